#!/bin/bash
# N GPUs: the default bench command as the driver launches it (headline + the config that rides along at this GPU count)
cd "$(dirname "$0")/.."
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2s_bench_n$N.json 2> gpurun_out/r2s_bench_n$N.err; tail -2 gpurun_out/r2s_bench_n$N.err; cut -c1-300 gpurun_out/r2s_bench_n$N.json
