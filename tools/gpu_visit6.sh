#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name: $*"; timeout "${T:-600}" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n "${TAILN:-6}" gpurun_out/$name.log | cut -c1-400; }
T=600 TAILN=3 run ops python -m pytest tests/test_ops_gpu.py -q -m gpu --no-header -p no:cacheprovider -x
T=300 TAILN=22 run attn_trace python tools/attn_trace.py
T=300 TAILN=16 run gemm_bench python tools/gemm_bench.py
SLSB_RED_ADD_V1=1 T=300 TAILN=16 run gemm_bench_v1 python tools/gemm_bench.py
T=900 TAILN=1 run bench python bench.py --steps 20 --warmup 3 --legs heads --no-cpu-baseline --sustained-steps 0
T=1500 TAILN=3 run parity python -m pytest tests/test_parity_gpu.py -q -m gpu --no-header -p no:cacheprovider -s -x
grep -E "^\.?\[" gpurun_out/parity.log | cut -c1-200
T=1200 TAILN=3 run configs python -m pytest tests/test_configs_gpu.py tests/test_host.py -q -m gpu --no-header -p no:cacheprovider -x
