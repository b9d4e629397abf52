#!/bin/bash
# Session r3m: the rest of the option matrix (tests/test_gpu_matrix.py) + the suites of the families whose dispatch moved to the table
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_matrix.py -q -m gpu 2>&1 | grep -v "^  warn\|Warning" | tail -70 > gpurun_out/r3m_pytest_matrix.log
python -m pytest tests -q -m gpu --deselect tests/test_gpu_matrix.py 2>&1 | tail -6 > gpurun_out/r3m_pytest_rest.log
cat gpurun_out/r3m_pytest_matrix.log gpurun_out/r3m_pytest_rest.log
