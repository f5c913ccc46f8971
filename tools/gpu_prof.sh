#!/bin/bash
# ncu session: launch list of one bench-sized forward + full captures of the top kernels.  Logs under gpurun_out/.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
CMD="python tools/prof_step.py --steps 1 --warmup 1"
$CMD > gpurun_out/prof_plain.log 2>&1 || { echo "plain run failed"; tail gpurun_out/prof_plain.log; exit 1; }
tail -2 gpurun_out/prof_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches.csv
ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 113 -c 4 -o gpurun_out/prof_gemm -f $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "gemm capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:attn_tc_kernel -s 24 -c 1 -o gpurun_out/prof_attn -f $CMD > gpurun_out/ncu_attn.log 2>&1
echo "attn capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"conv0_kernel|ln_kernel|topk_threshold|mean_pool_kept" -s 16 -c 12 -o gpurun_out/prof_misc -f $CMD > gpurun_out/ncu_misc.log 2>&1
echo "misc capture rc=$?"
ls -la gpurun_out/
