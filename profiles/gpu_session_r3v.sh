#!/bin/bash
# Session r3v: SEI families with Mi / cpl in the global workspace: wide SEI four groups per SM (phi_5 global) vs three (all on chip);
# SEI eight systems per SM (phi_5 global) vs seven (all on chip)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
V=$PWD/profiles/variants
{
for f in sei wsei; do python profiles/k4_probe.py 16384 $f 2>&1 | tail -1; done
for f in sei wsei; do PLB_LIB=$V/libplb_X.so timeout 300 python profiles/k4_probe.py 16384 $f 2>&1 | tail -1; done
} > gpurun_out/r3v_ab.txt
cut -c1-170 gpurun_out/r3v_ab.txt
