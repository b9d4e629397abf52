#!/bin/bash
cd "$(dirname "$0")/.."
for v in w4 w5; do PLB_LIB=$PWD/profiles/variants/libplb_$v.so python profiles/k4_probe.py 65536 iso 2>&1 | tail -1; done
python profiles/k4_probe.py 65536 iso 2>&1 | tail -1
