#!/bin/bash
# Session r3u: thermal families with pd / wT in the global workspace too: 7 or 8 systems per SM; iso with nine
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
V=$PWD/profiles/variants
{
for f in thermal thsei wth; do python profiles/k4_probe.py 16384 $f 2>&1 | tail -1; done
for f in thermal thsei wth; do PLB_LIB=$V/libplb_X.so timeout 300 python profiles/k4_probe.py 16384 $f 2>&1 | tail -1; done
PLB_LIB=$V/libplb_X.so timeout 300 python profiles/k4_probe.py 65536 iso 2>&1 | tail -1
PLB_LIB=$V/libplb_Y.so timeout 300 python profiles/k4_probe.py 16384 thermal 2>&1 | tail -1
} > gpurun_out/r3u_ab.txt
cut -c1-170 gpurun_out/r3u_ab.txt
