#!/bin/bash
# Session r3n2b (gpurun --gpus 2): the library's fan-out tests on the final build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 40 python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -3 > gpurun_out/r3n2b_pytest_multi.log
cat gpurun_out/r3n2b_pytest_multi.log
