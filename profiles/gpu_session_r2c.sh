#!/bin/bash
cd "$(dirname "$0")/.."
python profiles/diag_tight.py cfg3i 512 1e-8 3000 15 2>&1 | tail -24
python profiles/diag_tight.py cfg5n10 512 1e-9 7400 60 2>&1 | tail -24
python profiles/diag_tight.py cfg4 128 1e-9 147600 90 2>&1 | grep -B1 -A6 "flagdiff [1-9]\|mismatch [2-9]" | head -60
python -m pytest tests/test_gpu_fullsize.py -m gpu -q 2>&1 | tail -3
PROF_B1=65536 PROF_B4=8192 ncu --set full --clock-control none --import-source on -k regex:k_resjac -c 1 -s 1 -o gpurun_out/k1_r2a -f python profiles/prof_driver.py > gpurun_out/r2e_ncu_k1.log 2>&1
PROF_B1=8192 PROF_B4=8192 ncu --set full --clock-control none --import-source on -k regex:k_simulate -c 1 -s 1 -o gpurun_out/k4_r2a -f python profiles/prof_driver.py > gpurun_out/r2e_ncu_k4.log 2>&1
for k in k1_r2a k4_r2a; do python profiles/ncu_extract.py gpurun_out/$k.ncu-rep > gpurun_out/${k}_ncu_summary.txt 2>/dev/null; done
cat gpurun_out/k1_r2a_ncu_summary.txt
