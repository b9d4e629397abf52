#!/bin/bash
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_thermal.py tests/test_gpu_sei.py tests/test_gpu_wide.py tests/test_gpu_ragged.py -m gpu -q 2>&1 | tail -4
for fam in iso sei thermal wsei; do
  B=65536; [ $fam = thermal ] && B=32768
  python profiles/k1_probe.py $B $fam 2>&1 | tail -1
done
