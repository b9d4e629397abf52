#!/bin/bash
# Session r3y: memcheck of the wide SEI family of the last build (Mi / cpl in the global workspace, four groups per CTA)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SAN_FAMILIES=wsei timeout 80 compute-sanitizer --tool memcheck python profiles/sanitize_driver.py 2>&1 | tail -5 > gpurun_out/r3y_memcheck_wsei.txt
cat gpurun_out/r3y_memcheck_wsei.txt
