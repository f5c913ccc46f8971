#!/bin/bash
# One GPU visit: unit + parity tests, the bench line, and (NCU_LIST=1) the ncu launch list of the same bench command.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name: $*"; timeout "${T:-600}" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n "${TAILN:-6}" gpurun_out/$name.log; }
T=600 TAILN=4 run ops python -m pytest tests/test_ops_gpu.py -q -m gpu --no-header -p no:cacheprovider -x
T=1500 TAILN=4 run parity python -m pytest tests/test_parity_gpu.py -q -m gpu --no-header -p no:cacheprovider -s -x
grep -E "^\.?\[|max\|err" gpurun_out/parity.log | cut -c1-160
T=900 TAILN=1 run bench python bench.py --steps 20 --warmup 3 ${BENCH_ARGS:-}
if [ "${AB_PDL:-0}" = "1" ]; then
  SLSB_NO_PDL=1 T=900 TAILN=1 run bench_nopdl python bench.py --steps 20 --warmup 3 --no-cpu-baseline --sustained-steps 0
fi
if [ "${NCU_LIST:-0}" = "1" ]; then
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sustained-steps 0"
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv $CMD > gpurun_out/ncu_list.log 2>&1
  echo "launch list rc=$?"
fi
