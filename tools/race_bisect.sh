#!/bin/bash
# Bisect a bit-instability over the A/B switches: the same stress under each setting; the setting under which mismatches vanish names the kernel.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
REPS=${REPS:-12}
one() { echo "=== $*"; env "$@" timeout ${TMO:-300} python tools/stress_batch.py $REPS 2>&1 | grep -v Warning | grep -v WeightNorm | tail -n 12 | cut -c1-1800; }
for cfg in ${CONFIGS:-SLSB_X=0 SLSB_POOL_TMA=0 SLSB_RED_ADD_V1=1 SLSB_ATTN_IMPL=1 SLSB_NO_PDL=1 SLSB_NO_TAIL_SPLIT=1 SLSB_LN_STREAM=0 SLSB_GEMM_PAIR=0 SLSB_POS_V1=1 SLSB_LN_GEMM_V1=1}; do
  one $cfg
done
