#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): the N-rank tests, then the bench line under torchrun with the sharded DF-eval leg (strong scaling:
# 611 829 trials over N ranks, NCCL all_gather + rank-0 score.txt + EER inside the timed region, SHA-256 of the gathered scores).
set -u
N=${1:-2}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi -L | head -8
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_configs_gpu.py -q -m gpu --no-header -p no:cacheprovider -k "two_rank or own_device" > gpurun_out/multi_tests.log 2>&1
  echo "multi tests rc=$?"; tail -3 gpurun_out/multi_tests.log
fi
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 3 \
    --legs df_eval --no-cpu-baseline --sustained-steps 100 > gpurun_out/bench_${N}gpu.log 2>&1
echo "bench rc=$?"
grep '^{' gpurun_out/bench_${N}gpu.log | tail -1 | cut -c1-300
python - <<PY
import json
for l in open("gpurun_out/bench_${N}gpu.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print("N=$N value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "sustained", round(d.get("sustained", {}).get("value", 0), 1))
        print("df_eval", {k: d["df_eval"].get(k) for k in ("world", "seconds", "value", "device_seconds_scoring", "device_seconds_gather", "host_seconds_score_file", "eer", "sha256_scores", "collective")})
PY
