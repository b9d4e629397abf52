#!/bin/bash
# Session r3b: Fickian_method = :spectral sibling builds against the oracle; the N_r builds again (same templates); perf of
# the new siblings (one segment, 16 384 systems)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_spectral.py -q -m gpu 2>&1 | tail -60 > gpurun_out/r3b_pytest_sp.log
python -m pytest tests/test_gpu_nr.py tests/test_gpu_parity.py tests/test_gpu_thermal.py tests/test_gpu_sei.py -q -m gpu 2>&1 | tail -8 > gpurun_out/r3b_pytest_rest.log
cat gpurun_out/r3b_pytest_sp.log; cat gpurun_out/r3b_pytest_rest.log
