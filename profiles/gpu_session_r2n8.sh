#!/bin/bash
# final build on N GPUs of one box (N from the environment), as the driver launches it; multi-GPU library tests first
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${NGPU:-8}
python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -2 > gpurun_out/rn_multi_$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/rn_bench_n$N.json 2> gpurun_out/rn_bench_n$N.err
tail -2 gpurun_out/rn_bench_n$N.err; cat gpurun_out/rn_multi_$N.log; grep "^{" gpurun_out/rn_bench_n$N.json | cut -c1-300
