#!/bin/bash
cd "$(dirname "$0")/.."
python profiles/diag_maxit.py 1e-7 2>&1 | tail -16
python -m pytest tests/test_gpu_tight.py -m gpu -q -k "cfg4" 2>&1 | grep -E "^E .*Assert|passed|failed" | cut -c1-300
