#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lgm50.py -q -m gpu 2>&1 | grep -v "^$" | grep -E "^E  |passed|failed|FAILED|LGM50|Error|^tests" | head -60
python bench.py --no-cpu-baseline --extra none 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel_ms'])"
