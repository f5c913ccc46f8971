#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name: $*"; timeout "${T:-600}" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n "${TAILN:-6}" gpurun_out/$name.log; }
T=600 TAILN=30 run ops_a python -m pytest tests/test_ops_gpu.py -q -m gpu --no-header -p no:cacheprovider -s -k "attention or conv_ln_gelu or conv0_tensor"
T=600 TAILN=4 run ops_b python -m pytest tests/test_ops_gpu.py -q -m gpu --no-header -p no:cacheprovider -k "not (attention or conv_ln_gelu or conv0_tensor)"
T=1500 TAILN=4 run parity python -m pytest tests/test_parity_gpu.py -q -m gpu --no-header -p no:cacheprovider -s
grep -E "^\.?\[|max\|err" gpurun_out/parity.log | cut -c1-200
T=900 TAILN=3 run bench python bench.py --steps 10 --warmup 3 --no-cpu-baseline
CMD="python tools/prof_step.py --steps 1 --warmup 1"
$CMD > gpurun_out/prof_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
