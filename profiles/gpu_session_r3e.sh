#!/bin/bash
# Session r3e: PLB_TICK_ONE_EVAL in every family (one segment, 16 384 systems), default build vs variant
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for f in iso thermal sei thsei wide wsei wth wthsei mhc lgm lgmth iso_r12 thermal_r14 iso_sp; do
python profiles/k4_probe.py 16384 $f 2>&1 | tail -1
PLB_LIB=$PWD/profiles/variants/libplb_oneeval.so python profiles/k4_probe.py 16384 $f 2>&1 | tail -1
done > gpurun_out/r3e_ab.txt
cat gpurun_out/r3e_ab.txt | cut -c1-170
