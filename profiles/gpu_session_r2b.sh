#!/bin/bash
cd "$(dirname "$0")/.."
python profiles/bw_probe.py 2>&1 | tail -5
python profiles/diag_vmin.py 2>&1 | tail -30
python -m pytest tests/test_gpu_tight.py tests/test_gpu_parity.py -m gpu -q > gpurun_out/r2d_pytest.log 2>&1
tail -30 gpurun_out/r2d_pytest.log
python bench.py --steps 2 --warmup 1 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; tail -3 gpurun_out/r2d_bench.err; cat gpurun_out/r2d_bench.json
