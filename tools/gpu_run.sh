#!/bin/bash
# One GPU-box session: op tests (each file in its own process so a trap cannot poison the next), parity tests, bench.
# Everything is logged under gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name: $*"; timeout "${T:-600}" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n "${TAILN:-15}" gpurun_out/$name.log; }
T=900 run ops_gemm_tc python -m pytest tests/test_ops_gpu.py -q -m gpu -k "tc" -s --no-header -p no:cacheprovider
T=900 run ops_rest python -m pytest tests/test_ops_gpu.py -q -m gpu -k "not tc" -s --no-header -p no:cacheprovider
T=1500 run parity python -m pytest tests/test_parity_gpu.py -q -m gpu -s --no-header -p no:cacheprovider
T=600 run smoke python __graft_entry__.py smoke
T=900 run bench python bench.py --steps 10 --warmup 3
