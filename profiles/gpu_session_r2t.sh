#!/bin/bash
cd "$(dirname "$0")/.."
for v in n6 n7 n6g0; do PLB_LIB=$PWD/profiles/variants/libplb_$v.so python profiles/k4_probe.py 65536 iso 2>&1 | tail -1; done
PROF_B1=8192 PROF_B4=8192 ncu --set full --clock-control none --import-source on -k regex:k_simulate -c 1 -s 1 -o gpurun_out/k4_r2w8 -f python profiles/prof_driver.py > gpurun_out/r2w_ncu_k4.log 2>&1
python profiles/ncu_extract.py gpurun_out/k4_r2w8.ncu-rep | grep -E "stall|issue|inst_executed.sum|duration|icc|local"
