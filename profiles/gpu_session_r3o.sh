#!/bin/bash
# Session r3o: wide families: three two-warp groups in ONE CTA (one tick barrier for all of an SM's systems) vs three CTAs of one group
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for f in wide wsei; do
python profiles/k4_probe.py 16384 $f 2>&1 | tail -1
PLB_LIB=$PWD/profiles/variants/libplb_wide3x1.so timeout 300 python profiles/k4_probe.py 16384 $f 2>&1 | tail -1
done > gpurun_out/r3o_ab.txt
cat gpurun_out/r3o_ab.txt | cut -c1-170
