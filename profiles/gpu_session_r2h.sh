#!/bin/bash
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -q > gpurun_out/r2k_pytest.log 2>&1
tail -8 gpurun_out/r2k_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
