#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name: $*"; timeout "${T:-600}" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n "${TAILN:-6}" gpurun_out/$name.log | cut -c1-400; }
T=300 TAILN=4 run attn_ops python -m pytest tests/test_ops_gpu.py -q -m gpu --no-header -p no:cacheprovider -x -k attention
T=120 TAILN=30 run attn_trace4 python tools/attn_trace.py
SLSB_ATTN_NW=2 T=120 TAILN=3 run attn_trace2 python tools/attn_trace.py
T=600 TAILN=1 run bench4 python bench.py --steps 20 --warmup 3 --legs none --no-cpu-baseline --sustained-steps 0
SLSB_ATTN_NW=2 T=600 TAILN=1 run bench2 python bench.py --steps 20 --warmup 3 --legs none --no-cpu-baseline --sustained-steps 0
for f in bench4 bench2; do python - <<PY
import json
for l in open("gpurun_out/$f.log"):
    if l.startswith("{"):
        d = json.loads(l); r = d["roofline"]
        print("$f", round(d["value"], 1), d["ms_per_step"], "attn", r["other_kernels_ms_per_step"]["attention"], d["clocks"]["sm_mhz"])
PY
done
