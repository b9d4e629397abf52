#!/bin/bash
# Session r3c: the spectral tests with the thermal bound fixed, throughput of the new sibling builds (one segment, 16 384
# systems), memcheck over them
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_spectral.py tests/test_gpu_nr.py -q -m gpu 2>&1 | tail -5 > gpurun_out/r3c_pytest.log
for f in iso iso_r12 iso_r14 iso_sp thermal thermal_r12 thermal_r14 thermal_sp sei sei_r12 sei_r14 sei_sp; do python profiles/k4_probe.py 16384 $f 2>&1 | tail -1; done > gpurun_out/r3c_families.txt
SAN_FAMILIES=iso12,th14,sei14,isosp,thsp,seisp compute-sanitizer --tool memcheck python profiles/sanitize_driver.py 2>&1 | tail -12 > gpurun_out/r3c_memcheck.txt
cat gpurun_out/r3c_pytest.log gpurun_out/r3c_families.txt gpurun_out/r3c_memcheck.txt
