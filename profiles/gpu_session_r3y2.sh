#!/bin/bash
# Session r3y2: memcheck of three of the option-matrix families of the last build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SAN_FAMILIES=widemhc,thseimhc,wthlgm timeout 100 compute-sanitizer --tool memcheck python profiles/sanitize_driver.py 2>&1 | tail -7 > gpurun_out/r3y_memcheck_matrix.txt
cat gpurun_out/r3y_memcheck_matrix.txt
