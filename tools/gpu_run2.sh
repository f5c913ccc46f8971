#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name: $*"; timeout "${T:-600}" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n "${TAILN:-6}" gpurun_out/$name.log; }
T=900 run ops python -m pytest tests/test_ops_gpu.py -q -m gpu --no-header -p no:cacheprovider -x
T=900 run parity_varlen python -m pytest tests/test_parity_gpu.py -q -m gpu --no-header -p no:cacheprovider -k "variable_length or api_surface"
T=900 TAILN=3 run bench python bench.py --steps 10 --warmup 3 --no-cpu-baseline
bash tools/gpu_prof.sh
