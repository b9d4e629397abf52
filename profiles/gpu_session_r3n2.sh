#!/bin/bash
# Session r3n2 (gpurun --gpus 2): the library's fan-out tests and the 2-GPU bench line of the final build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -4 > gpurun_out/r3n2_pytest_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r3n2_bench_cfg2_n2.json 2> gpurun_out/r3n2_bench.err
cat gpurun_out/r3n2_pytest_multi.log; grep "^{" gpurun_out/r3n2_bench_cfg2_n2.json | cut -c1-400; tail -3 gpurun_out/r3n2_bench.err
