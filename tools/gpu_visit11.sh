#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
export SLSB_ATTN_SPLIT=1
for s in 0 2000 3000 3500 4000 5000; do
  SLSB_ATTN_STAGGER=$s timeout 120 python tools/attn_trace.py > gpurun_out/attn2_stagger_$s.log 2>&1
  echo "split stagger $s: $(head -1 gpurun_out/attn2_stagger_$s.log)"
done
head -18 gpurun_out/attn2_stagger_3500.log | cut -c1-200
SLSB_ATTN_STAGGER=3500 timeout 300 python -m pytest tests/test_ops_gpu.py -q -m gpu --no-header -p no:cacheprovider -k attention 2>&1 | tail -2
