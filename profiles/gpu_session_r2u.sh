#!/bin/bash
cd "$(dirname "$0")/.."
PLB_LIB=$PWD/profiles/variants/libplb_n7g1.so python profiles/k4_probe.py 65536 iso 2>&1 | tail -1
