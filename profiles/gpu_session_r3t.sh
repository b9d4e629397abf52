#!/bin/bash
# Session r3t: final geometry (factored blocks in the global workspace, history vectors back in shared memory): tests, families, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -5 > gpurun_out/r3t_pytest.log
for f in iso thermal sei wide wsei wth thsei wthsei mhc lgm lgmth iso_r12 iso_r14 iso_sp thermal_r12 thermal_r14 thermal_sp sei_r12 sei_r14 sei_sp; do python profiles/k4_probe.py 16384 $f 2>&1 | tail -1; done > gpurun_out/r3t_families.txt
python bench.py > gpurun_out/r3t_bench_cfg2.json 2> gpurun_out/r3t_bench_cfg2.err
cat gpurun_out/r3t_pytest.log; cut -c1-150 gpurun_out/r3t_families.txt; cut -c1-400 gpurun_out/r3t_bench_cfg2.json
