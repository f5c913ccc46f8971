#!/bin/bash
# Round-2 evidence visit: every GPU test, the full bench line (all legs), then the profiling pass (launch list + ncu --set full).
set -u
bash tools/gpu_round.sh
cp gpurun_out/bench.log gpurun_out/bench_full.log
bash tools/gpu_profile.sh
