#!/bin/bash
# Session r3f: with PLB_TICK_ONE_EVAL=1 as the default: a second barrier before the solve (syncsolve), or only that one (solveonly)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for f in iso thermal; do
python profiles/k4_probe.py 32768 $f 2>&1 | tail -1
PLB_LIB=$PWD/profiles/variants/libplb_syncsolve.so timeout 120 python profiles/k4_probe.py 32768 $f 2>&1 | tail -1
PLB_LIB=$PWD/profiles/variants/libplb_solveonly.so timeout 120 python profiles/k4_probe.py 32768 $f 2>&1 | tail -1
done > gpurun_out/r3f_ab.txt
cat gpurun_out/r3f_ab.txt | cut -c1-170
