#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python tools/throttle_probe.py --chunks 25 --chunk 10 --mode device > gpurun_out/probe_device.log 2>&1; tail -26 gpurun_out/probe_device.log
timeout 300 python tools/throttle_probe.py --chunks 10 --chunk 10 --mode submit > gpurun_out/probe_submit.log 2>&1; tail -11 gpurun_out/probe_submit.log
timeout 300 python tools/throttle_probe.py --chunks 6 --chunk 10 --mode host > gpurun_out/probe_host.log 2>&1; tail -7 gpurun_out/probe_host.log
CMD="python tools/prof_step.py --steps 1 --warmup 1 --head sls"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc_kernel -s 24 -c 1 -o gpurun_out/prof_attn2 -f $CMD > gpurun_out/ncu_attn2.log 2>&1
echo "attn capture rc=$?"
