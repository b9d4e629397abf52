#!/bin/bash
cd "$(dirname "$0")/.."
python profiles/diag_tight.py cfg4 512 1e-9 147600 90 2>&1 | grep -A12 "FLAGDIFF\|rel [1-9].[0-9]*e-0[1-4]" | head -80
