#!/bin/bash
# Session r3l: geometry / unroll knobs re-measured on top of ONE_EVAL (iso)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python profiles/k4_probe.py 65536 iso 2>&1 | tail -1 > gpurun_out/r3l_ab.txt
for v in w6g0 w7g1 unroll2; do PLB_LIB=$PWD/profiles/variants/libplb_$v.so timeout 300 python profiles/k4_probe.py 65536 iso 2>&1 | tail -1; done >> gpurun_out/r3l_ab.txt
cat gpurun_out/r3l_ab.txt | cut -c1-170
