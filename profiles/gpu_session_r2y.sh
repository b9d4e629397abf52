#!/bin/bash
# rxn_MHC: new GPU tests first, then the whole GPU suite and the headline line (must not move)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_mhc.py -q -m gpu -x 2>&1 | tail -25 > gpurun_out/ry_mhc.log
python -m pytest tests -q -m gpu 2>&1 | tail -12 > gpurun_out/ry_pytest.log
python bench.py --no-cpu-baseline --extra none > gpurun_out/ry_bench.json 2> gpurun_out/ry_bench.err
cat gpurun_out/ry_mhc.log; cat gpurun_out/ry_pytest.log; cut -c1-300 gpurun_out/ry_bench.json
