#!/bin/bash
cd "$(dirname "$0")/.."
python profiles/k1_probe.py 65536 sei 2>&1 | tail -1
PLB_LIB=$PWD/profiles/variants/libplb_seitma3.so python profiles/k1_probe.py 65536 sei 2>&1 | tail -1
PLB_LIB=$PWD/profiles/variants/libplb_seitma3.so python profiles/k1_probe.py 65536 iso 2>&1 | tail -1
