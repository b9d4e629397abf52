#!/bin/bash
# round 2, session a: all GPU tests with the TMA-staged K1 and the lane-parallel parameter set-up; K1 A/B; K4 probe
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -q --durations=10 > gpurun_out/r2c_pytest.log 2>&1
tail -25 gpurun_out/r2c_pytest.log
for fam in iso sei; do
  python profiles/k1_probe.py 65536 $fam 2>&1 | tail -1
  PLB_K1_NO_TMA=1 python profiles/k1_probe.py 65536 $fam 2>&1 | tail -1
done
python profiles/k1_probe.py 65535 iso 2>&1 | tail -1
python profiles/k1_probe.py 32768 thermal 2>&1 | tail -1
python profiles/k4_probe.py 65536 iso 2>&1 | tail -1
