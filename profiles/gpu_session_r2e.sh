#!/bin/bash
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_tight.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q > gpurun_out/r2g_pytest.log 2>&1
tail -8 gpurun_out/r2g_pytest.log
for fam in iso sei thermal; do python profiles/k1_probe.py $([ $fam = thermal ] && echo 32768 || echo 65536) $fam 2>&1 | tail -1; done
python profiles/k4_probe.py 65536 iso 2>&1 | tail -1
for v in u2 u5; do PLB_LIB=$PWD/profiles/variants/libplb_$v.so python profiles/k4_probe.py 65536 iso 2>&1 | tail -1; done
