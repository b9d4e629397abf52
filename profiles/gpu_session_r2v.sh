#!/bin/bash
cd "$(dirname "$0")/.."
PLB_LIB=$PWD/profiles/variants/libplb_g0w5.so python profiles/k4_probe.py 65536 sei 2>&1 | tail -1
PLB_LIB=$PWD/profiles/variants/libplb_g0w4.so python profiles/k4_probe.py 32768 thermal 2>&1 | tail -1
PLB_LIB=$PWD/profiles/variants/libplb_g0wide.so python profiles/k4_probe.py 32768 wsei 2>&1 | tail -1
