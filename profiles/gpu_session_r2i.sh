#!/bin/bash
cd "$(dirname "$0")/.."
python profiles/diag_tight.py cfg3i 512 1e-7 3000 15 2>&1 | grep -v "^   gpu\|^   cpu\|first rows\|I gpu\|I cpu" | tail -24
python profiles/diag_flips.py 2>&1 | tail -8
