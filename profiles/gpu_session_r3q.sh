#!/bin/bash
# Session r3q: thermal families with their factored blocks in the global workspace (6 / 5 / 3x / 2x systems per SM): probe + tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for f in thermal thsei wth wthsei lgmth thermal_r12 thermal_r14 thermal_sp; do python profiles/k4_probe.py 16384 $f 2>&1 | tail -1; done > gpurun_out/r3q_families.txt
python -m pytest tests/test_gpu_thermal.py tests/test_gpu_thsei.py tests/test_gpu_wide.py tests/test_gpu_matrix.py tests/test_gpu_lgm50.py tests/test_gpu_nr.py tests/test_gpu_spectral.py tests/test_gpu_mhc.py tests/test_gpu_tight.py -q -m gpu 2>&1 | tail -6 > gpurun_out/r3q_pytest.log
cut -c1-170 gpurun_out/r3q_families.txt; cat gpurun_out/r3q_pytest.log
