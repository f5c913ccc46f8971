#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
CMD="python tools/prof_step.py --steps 1 --warmup 1"
$CMD > gpurun_out/prof_plain.log 2>&1 || { echo "plain run failed"; tail gpurun_out/prof_plain.log; exit 1; }
tail -1 gpurun_out/prof_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:tc_gemm_ln_kernel -s 7 -c 3 -o gpurun_out/prof_lngemm -f $CMD > gpurun_out/ncu_lngemm.log 2>&1
echo "lngemm capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:attn_tc_kernel -s 24 -c 1 -o gpurun_out/prof_attn -f $CMD > gpurun_out/ncu_attn.log 2>&1
echo "attn capture rc=$?"
