#!/bin/bash
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -q > gpurun_out/r2f_pytest.log 2>&1
tail -12 gpurun_out/r2f_pytest.log
PROF_B1=65536 PROF_B4=8192 ncu --set full --clock-control none --import-source on -k regex:k_resjac -c 1 -s 1 -o gpurun_out/k1_r2a -f python profiles/prof_driver.py > gpurun_out/r2e_ncu_k1.log 2>&1
PROF_B1=8192 PROF_B4=8192 ncu --set full --clock-control none --import-source on -k regex:k_simulate -c 1 -s 1 -o gpurun_out/k4_r2a -f python profiles/prof_driver.py > gpurun_out/r2e_ncu_k4.log 2>&1
for k in k1_r2a k4_r2a; do python profiles/ncu_extract.py gpurun_out/$k.ncu-rep > gpurun_out/${k}_ncu_summary.txt 2>/dev/null; done
cat gpurun_out/k1_r2a_ncu_summary.txt
tail -3 gpurun_out/r2e_ncu_k4.log
