#!/bin/bash
# 2 GPUs: the library's multi-GPU fan-out (plb_group_*), two handles on two devices, bench at N=2 (timed all-gather)
cd "$(dirname "$0")/.."
nvidia-smi -L
python -m pytest tests/test_gpu_multi.py tests/test_gpu_stops.py -m gpu -q -k "group or two_handles" 2>&1 | tail -5 | tee gpurun_out/r2r_pytest_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2r_bench_n2.json 2> gpurun_out/r2r_bench_n2.err; tail -2 gpurun_out/r2r_bench_n2.err; cut -c1-400 gpurun_out/r2r_bench_n2.json
