#!/bin/bash
# Session r3r: non-thermal families with their factored blocks in the global workspace: what to do with the freed shared memory
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
V=$PWD/profiles/variants
{
python profiles/k4_probe.py 65536 iso 2>&1 | tail -1
python profiles/k4_probe.py 16384 sei 2>&1 | tail -1
python profiles/k4_probe.py 16384 wide 2>&1 | tail -1
python profiles/k4_probe.py 16384 wsei 2>&1 | tail -1
[ -f $V/libplb_g0.so ] && PLB_LIB=$V/libplb_g0.so timeout 300 python profiles/k4_probe.py 65536 iso 2>&1 | tail -1
[ -f $V/libplb_g0.so ] && PLB_LIB=$V/libplb_g0.so timeout 300 python profiles/k4_probe.py 16384 wsei 2>&1 | tail -1
[ -f $V/libplb_sei8.so ] && PLB_LIB=$V/libplb_sei8.so timeout 300 python profiles/k4_probe.py 16384 sei 2>&1 | tail -1
[ -f $V/libplb_sei7g0.so ] && PLB_LIB=$V/libplb_sei7g0.so timeout 300 python profiles/k4_probe.py 16384 sei 2>&1 | tail -1
[ -f $V/libplb_wide4.so ] && PLB_LIB=$V/libplb_wide4.so timeout 300 python profiles/k4_probe.py 16384 wide 2>&1 | tail -1
} >> gpurun_out/r3r_ab.txt
cut -c1-170 gpurun_out/r3r_ab.txt
