#!/bin/bash
# A/B visit: one-kernel conv0 (SLSB_CONV0_V1=0) and split-row attention (SLSB_ATTN_SPLIT=1) against the shipped defaults.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name: $*"; timeout "${T:-600}" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n "${TAILN:-6}" gpurun_out/$name.log | cut -c1-400; }
T=300 TAILN=12 run conv0_ops python -m pytest tests/test_ops_gpu.py -q -m gpu --no-header -p no:cacheprovider -k conv0 -s
SLSB_ATTN_SPLIT=1 T=300 TAILN=6 run attn_ops python -m pytest tests/test_ops_gpu.py -q -m gpu --no-header -p no:cacheprovider -k attention
SLSB_ATTN_SPLIT=1 T=120 TAILN=24 run attn_trace_split python tools/attn_trace.py
T=120 TAILN=3 run attn_trace_base python tools/attn_trace.py
B="python bench.py --steps 20 --warmup 3 --legs none --no-cpu-baseline --sustained-steps 0"
T=600 TAILN=1 run bench_base $B
SLSB_CONV0_V1=0 T=600 TAILN=1 run bench_conv0 $B
SLSB_ATTN_SPLIT=1 T=600 TAILN=1 run bench_attn $B
SLSB_CONV0_V1=0 SLSB_ATTN_SPLIT=1 T=600 TAILN=1 run bench_both $B
SLSB_CONV0_V1=0 SLSB_ATTN_SPLIT=1 T=900 TAILN=4 run parity_both python -m pytest tests/test_parity_gpu.py -q -m gpu --no-header -p no:cacheprovider -x
for f in bench_base bench_conv0 bench_attn bench_both; do python - <<PY
import json
for l in open("gpurun_out/$f.log"):
    if l.startswith("{"):
        d = json.loads(l); r = d["roofline"]; o = r["other_kernels_ms_per_step"]
        print("$f", round(d["value"], 1), round(d["ms_per_step"], 3), "conv", round(o["conv_gemm"], 3), "attn", round(o["attention"], 3), "cold", round(d["cold_burst"]["value"], 1), d["clocks"]["sm_mhz"])
PY
done
