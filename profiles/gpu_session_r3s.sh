#!/bin/bash
# Session r3s: thermal / wide families: phi_4 / phi_5 back in shared memory against one system more per SM
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
V=$PWD/profiles/variants
{
for f in thermal thsei wthsei wide; do python profiles/k4_probe.py 16384 $f 2>&1 | tail -1; done
for f in thermal thsei wthsei wide; do PLB_LIB=$V/libplb_X.so timeout 300 python profiles/k4_probe.py 16384 $f 2>&1 | tail -1; done
for f in thermal wide; do PLB_LIB=$V/libplb_Y.so timeout 300 python profiles/k4_probe.py 16384 $f 2>&1 | tail -1; done
} > gpurun_out/r3s_ab.txt
cut -c1-170 gpurun_out/r3s_ab.txt
