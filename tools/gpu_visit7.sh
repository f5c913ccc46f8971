#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name: $*"; timeout "${T:-600}" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n "${TAILN:-6}" gpurun_out/$name.log | cut -c1-400; }
T=600 TAILN=3 run ops python -m pytest tests/test_ops_gpu.py -q -m gpu --no-header -p no:cacheprovider -x
T=300 TAILN=20 run attn_trace python tools/attn_trace.py
T=600 TAILN=12 run flac python -m pytest tests/test_configs_gpu.py -q -m gpu --no-header -p no:cacheprovider -x -k "flac or pcm16 or sparse"
T=900 TAILN=1 run bench python bench.py --steps 20 --warmup 3 --legs heads,ingest --no-cpu-baseline --sustained-steps 0
T=1500 TAILN=3 run parity python -m pytest tests/test_parity_gpu.py -q -m gpu --no-header -p no:cacheprovider -s -x
