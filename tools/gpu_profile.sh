#!/bin/bash
# Profiling visit (after the plain commands have exited 0 without ncu): launch list of the bench command + ncu --set full of
# the dominant kernel (CTA-pair tcgen05 encoder GEMM: one launch each of qkv / out_proj / fc1 / fc2), the conv pair kernel, LayerNorm, attention.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sustained-steps 0 --prewarm-seconds 0 --legs none"
$CMD > gpurun_out/bench_plain.log 2>&1 || { echo "plain bench failed"; tail -5 gpurun_out/bench_plain.log; exit 1; }
tail -1 gpurun_out/bench_plain.log | cut -c1-160
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
P="python tools/prof_step.py --steps 1 --warmup 1 --head sls"
$P > gpurun_out/prof_plain.log 2>&1 || { echo "plain prof_step failed"; exit 1; }
# launch order inside a step: ... LN, qkv GEMM, attention, out GEMM, LN, fc1, fc2 ... : skip the first step (warm-up) entirely
ncu --set full --clock-control none --import-source on -k regex:tc_gemm_pair_kernel -s 100 -c 4 -o gpurun_out/prof_gemm -f $P > gpurun_out/ncu_gemm.log 2>&1
echo "gemm capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"tc_gemm_ln2x_kernel|conv0_tc_kernel" -s 7 -c 3 -o gpurun_out/prof_ln2 -f $P > gpurun_out/ncu_ln2.log 2>&1
echo "ln2 capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"ln_stream_kernel|attn_tc_kernel" -s 80 -c 4 -o gpurun_out/prof_hbm -f $P > gpurun_out/ncu_hbm.log 2>&1
echo "hbm capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"sls_fuse_pool" -s 1 -c 1 -o gpurun_out/prof_pool -f $P > gpurun_out/ncu_pool.log 2>&1
echo "pool capture rc=$?"
# then, in the repo: python tools/ncu_summarize.py --round rNN  (writes profiles/)
# head kernels of the heads main.py builds (H-SAE / H-WIN): fused select+pool
PS="python tools/prof_step.py --steps 1 --warmup 1 --head sae"
$PS > gpurun_out/prof_sae_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"topk_pool_kernel|pool_finish_kernel" -s 2 -c 2 -o gpurun_out/prof_sae -f $PS > gpurun_out/ncu_sae.log 2>&1
echo "sae head capture rc=$?"
PW="python tools/prof_step.py --steps 1 --warmup 1 --head window"
$PW > gpurun_out/prof_win_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"window_select_kernel|window_vote_pool_kernel" -s 2 -c 2 -o gpurun_out/prof_win -f $PW > gpurun_out/ncu_win.log 2>&1
echo "window head capture rc=$?"
