#!/bin/bash
# wide thermal family: new tests, then the whole suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_wide.py -q -m gpu 2>&1 | tail -40 > gpurun_out/rv_wth.log
cat gpurun_out/rv_wth.log
python -m pytest tests -q -m gpu --deselect tests/test_gpu_wide.py 2>&1 | tail -8
