#!/bin/bash
cd "$(dirname "$0")/.."
python profiles/k4_probe.py 65536 iso 2>&1 | tail -1
python profiles/k4_probe.py 65536 sei 2>&1 | tail -1
python profiles/k4_probe.py 32768 wsei 2>&1 | tail -1
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
