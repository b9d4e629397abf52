#!/bin/bash
# Session r3d: A/B of PLB_TICK_ONE_EVAL (every evaluation of a tick runs the residual+Jacobian instantiation)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for i in 1 2; do
python profiles/k4_probe.py 65536 iso 2>&1 | tail -1
PLB_LIB=$PWD/profiles/variants/libplb_oneeval.so python profiles/k4_probe.py 65536 iso 2>&1 | tail -1
done > gpurun_out/r3d_ab.txt
cat gpurun_out/r3d_ab.txt
