#!/bin/bash
# thermal + SEI family, then everything
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_thsei.py -q -m gpu 2>&1 | grep -v "^$" | grep -E "^E  |passed|failed|FAILED|thsei solve|segment|Error" | head -60 > gpurun_out/ru_thsei.log
cat gpurun_out/ru_thsei.log
python -m pytest tests -q -m gpu --deselect tests/test_gpu_thsei.py 2>&1 | tail -8
python bench.py --no-cpu-baseline --extra none 2>/dev/null | cut -c1-200
python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline --extra none 2>/dev/null | cut -c1-200
