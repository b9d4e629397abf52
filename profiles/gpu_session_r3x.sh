#!/bin/bash
# Session r3x: configs[4] (wide SEI) bench line of the final build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py --workload cfg5 --steps 2 --warmup 3 --extra none > gpurun_out/r3x_bench_cfg5.json 2> gpurun_out/r3x_bench_cfg5.err
grep "^{" gpurun_out/r3x_bench_cfg5.json | cut -c1-300; tail -2 gpurun_out/r3x_bench_cfg5.err
