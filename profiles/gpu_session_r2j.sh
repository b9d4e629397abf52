#!/bin/bash
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -q > gpurun_out/r2m_pytest.log 2>&1
tail -6 gpurun_out/r2m_pytest.log
python bench.py > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; tail -2 gpurun_out/r2m_bench.err; cut -c1-600 gpurun_out/r2m_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2m_bench_reference.json 2>/dev/null; cut -c1-300 gpurun_out/r2m_bench_reference.json
