#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name: $*"; timeout "${T:-600}" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n "${TAILN:-6}" gpurun_out/$name.log; }
T=600 TAILN=25 run ops_tc python -m pytest tests/test_ops_gpu.py -q -m gpu --no-header -p no:cacheprovider -s -k "gemm_bf16_tc or conv_implicit"
T=600 TAILN=4 run ops_rest python -m pytest tests/test_ops_gpu.py -q -m gpu --no-header -p no:cacheprovider -k "not (gemm_bf16_tc or conv_implicit)"
T=1500 TAILN=4 run parity python -m pytest tests/test_parity_gpu.py -q -m gpu --no-header -p no:cacheprovider -s
grep -E "^\.?\[|max\|err" gpurun_out/parity.log | cut -c1-330
T=600 TAILN=14 run gemm_bench python tools/gemm_bench.py
T=900 TAILN=3 run bench python bench.py --steps 10 --warmup 3 --no-cpu-baseline
