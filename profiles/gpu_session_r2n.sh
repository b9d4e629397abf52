#!/bin/bash
cd "$(dirname "$0")/.."
compute-sanitizer --tool memcheck python profiles/sanitize_driver.py 2>&1 | tail -8 > gpurun_out/r2q_memcheck.txt; cat gpurun_out/r2q_memcheck.txt
SAN_FAMILIES=iso compute-sanitizer --tool racecheck python profiles/sanitize_driver.py 2>&1 > gpurun_out/r2q_racecheck_full.txt; grep -c "Warning\|WARN" gpurun_out/r2q_racecheck_full.txt; grep -E "ERROR|Error|RACECHECK SUMMARY|ok \[" gpurun_out/r2q_racecheck_full.txt | head; grep -o "plb_[a-z]*\.cuh:[0-9]*" gpurun_out/r2q_racecheck_full.txt | sort | uniq -c | sort -rn | head -12
SAN_FAMILIES=iso compute-sanitizer --tool synccheck python profiles/sanitize_driver.py 2>&1 | tail -4
