#!/bin/bash
# Session r3g: PLB_TICK_SPLIT_LSETUP (an lsetup iteration takes two ticks) vs the ONE_EVAL default
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for f in iso thermal sei wsei; do
python profiles/k4_probe.py 32768 $f 2>&1 | tail -1
PLB_LIB=$PWD/profiles/variants/libplb_split.so timeout 120 python profiles/k4_probe.py 32768 $f 2>&1 | tail -1
done > gpurun_out/r3g_ab.txt
cat gpurun_out/r3g_ab.txt | cut -c1-170
