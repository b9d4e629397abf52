#!/bin/bash
# Session r3k: PLB_TICK_SYNC_STEP (one barrier per step attempt instead of one per evaluation) vs the default
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for f in iso thermal sei wsei; do
python profiles/k4_probe.py 32768 $f 2>&1 | tail -1
PLB_LIB=$PWD/profiles/variants/libplb_syncstep.so timeout 300 python profiles/k4_probe.py 32768 $f 2>&1 | tail -1
done > gpurun_out/r3k_ab.txt
cat gpurun_out/r3k_ab.txt | cut -c1-170
