#!/bin/bash
# Session r3j: ncu --set full capture of K4 (iso) of the ONE_EVAL build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PROF_B1=8192 PROF_B4=8192 ncu --set full --clock-control none --import-source on -k regex:k_simulate -c 1 -s 1 -o gpurun_out/k4_r3j -f python profiles/prof_driver.py > gpurun_out/r3j_ncu_k4.log 2>&1
python profiles/ncu_extract.py gpurun_out/k4_r3j.ncu-rep > gpurun_out/k4_r3j_ncu_summary.txt 2>/dev/null
cat gpurun_out/k4_r3j_ncu_summary.txt | grep -E "stall|issue_active|icc|inst_executed.sum|duration|local"
ls -la gpurun_out/k4_r3j.ncu-rep
