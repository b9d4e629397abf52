#!/bin/bash
# dc inputs, MHC in its own families, the failing wide-thermal test, headline + K1
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dc.py tests/test_gpu_mhc.py "tests/test_gpu_wide.py::test_simulate_parity" -q -m gpu 2>&1 | grep -v "^$" | grep -E "^E  |passed|failed|FAILED|Error|^wide" | head -60 > gpurun_out/rt_new.log
cat gpurun_out/rt_new.log
python bench.py --no-cpu-baseline --extra none 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel_ms'])"
