#!/bin/bash
# One GPU visit: unit + parity + config tests, the bench line (every leg), and (NCU_LIST=1) the ncu launch list of the bench's
# headline loop (legs off, no pre-warm: under ncu every launch is serialised, so only the launch mix / shares are read from it).
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name: $*"; timeout "${T:-600}" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n "${TAILN:-6}" gpurun_out/$name.log | cut -c1-400; }
if [ "${SKIP_TESTS:-0}" != "1" ]; then
T=600 TAILN=4 run ops python -m pytest tests/test_ops_gpu.py -q -m gpu --no-header -p no:cacheprovider -x
T=1500 TAILN=4 run parity python -m pytest tests/test_parity_gpu.py -q -m gpu --no-header -p no:cacheprovider -s -x
grep -E "^\.?\[|max\|err" gpurun_out/parity.log | cut -c1-160
T=1200 TAILN=4 run configs python -m pytest tests/test_configs_gpu.py tests/test_host.py -q -m gpu --no-header -p no:cacheprovider -x
fi
if [ "${SKIP_BENCH:-0}" != "1" ]; then
T=1200 TAILN=1 run bench python bench.py --steps 20 --warmup 3 ${BENCH_ARGS:-}
fi
if [ "${NCU_LIST:-0}" = "1" ]; then
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sustained-steps 0 --prewarm-seconds 0 --legs none"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv $CMD > gpurun_out/ncu_list.log 2>&1
  echo "launch list rc=$?"
fi
