#!/bin/bash
# Session r3a (second sitting of round 2): the N_r = 12 / 14 sibling builds against the oracle, then the whole GPU suite
# and the default bench line of the restored build.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_nr.py -q -m gpu -x 2>&1 | tail -40 > gpurun_out/r3a_pytest_nr.log
python -m pytest tests -q -m gpu --durations=12 2>&1 | tail -30 > gpurun_out/r3a_pytest.log
python bench.py > gpurun_out/r3a_bench_cfg2.json 2> gpurun_out/r3a_bench_cfg2.err
cat gpurun_out/r3a_pytest_nr.log; tail -22 gpurun_out/r3a_pytest.log; cut -c1-600 gpurun_out/r3a_bench_cfg2.json
