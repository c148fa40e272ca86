#!/bin/bash
for wl in c1; do python tools/prof_band.py 0 3 $wl 2>&1 | tail -2 | cut -c1-300; done
timeout 1100 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
