#!/bin/bash
for d in 1 2 4 0; do
  echo "=== PFBG_COLS2_DEBUG=$d"
  PFBG_COLS2_DEBUG=$d timeout 120 python -m pytest tests/test_gpu_cols2.py -x -q -m gpu -k "against_dft and 128" 2>&1 | grep -E "passed|failed|Error|error|assert" | head -6
done
