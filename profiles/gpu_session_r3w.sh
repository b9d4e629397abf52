#!/bin/bash
# Session r3w: final verification of the last build: whole GPU suite, the two SEI family probes, smoke
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -3 > gpurun_out/r3w_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 > gpurun_out/r3w_smoke.log
for f in sei wsei; do python profiles/k4_probe.py 16384 $f 2>&1 | tail -1; done > gpurun_out/r3w_families.txt
cat gpurun_out/r3w_pytest.log gpurun_out/r3w_smoke.log; cut -c1-150 gpurun_out/r3w_families.txt
