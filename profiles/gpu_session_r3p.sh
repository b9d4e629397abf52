#!/bin/bash
# Session r3p: wide thermal 2 groups x 1 CTA vs 1 x 2; wide SEI with its new default; the wide test files
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python profiles/k4_probe.py 16384 wth 2>&1 | tail -1 > gpurun_out/r3p_ab.txt
PLB_LIB=$PWD/profiles/variants/libplb_wth2x1.so timeout 300 python profiles/k4_probe.py 16384 wth 2>&1 | tail -1 >> gpurun_out/r3p_ab.txt
python profiles/k4_probe.py 16384 wsei 2>&1 | tail -1 >> gpurun_out/r3p_ab.txt
python -m pytest tests/test_gpu_wide.py tests/test_gpu_matrix.py tests/test_gpu_tight.py tests/test_gpu_ragged.py -q -m gpu 2>&1 | tail -4 >> gpurun_out/r3p_ab.txt
cat gpurun_out/r3p_ab.txt | cut -c1-170
