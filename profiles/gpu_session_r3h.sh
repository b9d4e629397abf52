#!/bin/bash
# Session r3h: the ONE_EVAL default build: whole GPU suite, smoke, default bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -12 > gpurun_out/r3h_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3h_smoke.log 2>&1
python bench.py > gpurun_out/r3h_bench_cfg2.json 2> gpurun_out/r3h_bench_cfg2.err
cat gpurun_out/r3h_pytest.log; tail -2 gpurun_out/r3h_smoke.log; cat gpurun_out/r3h_bench_cfg2.json
