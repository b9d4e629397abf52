#!/bin/bash
# Session r3n8 (gpurun --gpus 8): the 8-GPU bench line of the final build (configs[2] at its named scale rides along)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r3n8_bench_cfg2_n8.json 2> gpurun_out/r3n8_bench.err
grep "^{" gpurun_out/r3n8_bench_cfg2_n8.json | cut -c1-300; tail -3 gpurun_out/r3n8_bench.err
