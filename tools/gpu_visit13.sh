#!/bin/bash
# A/B: conv1-6 with the A tile multicast across the CTA pair (SLSB_LN2_MCAST=1)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name: $*"; timeout "${T:-600}" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n "${TAILN:-6}" gpurun_out/$name.log | cut -c1-400; }
SLSB_LN2_MCAST=1 T=300 TAILN=4 run conv_ops_mc python -m pytest tests/test_ops_gpu.py -q -m gpu --no-header -p no:cacheprovider -k "conv"
B="python bench.py --steps 20 --warmup 3 --legs none --no-cpu-baseline --sustained-steps 0"
T=600 TAILN=1 run bench_base $B
SLSB_LN2_MCAST=1 T=600 TAILN=1 run bench_mc $B
SLSB_LN2_MCAST=1 T=900 TAILN=3 run parity_mc python -m pytest tests/test_parity_gpu.py -q -m gpu --no-header -p no:cacheprovider -x
for f in bench_base bench_mc; do python - <<PY
import json
for l in open("gpurun_out/$f.log"):
    if l.startswith("{"):
        d = json.loads(l); r = d["roofline"]; o = r["other_kernels_ms_per_step"]
        print("$f", round(d["value"], 1), round(d["ms_per_step"], 3), "conv", round(o["conv_gemm"], 3), "attn", round(o["attention"], 3), "cold", round(d["cold_burst"]["value"], 1), d["clocks"]["sm_mhz"])
PY
done
