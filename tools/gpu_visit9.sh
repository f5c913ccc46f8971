#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for s in 0 3000 6000 9000 12000 15000 20000; do
  echo "== stagger $s"; SLSB_ATTN_STAGGER=$s timeout 120 python tools/attn_trace.py 2>&1 | head -12 | cut -c1-200
done
SLSB_ATTN_STAGGER=9000 timeout 300 python -m pytest tests/test_ops_gpu.py -q -m gpu --no-header -p no:cacheprovider -x -k attention 2>&1 | tail -2
