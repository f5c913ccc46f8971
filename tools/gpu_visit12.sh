#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python tools/gemm_bench.py > gpurun_out/gemm_bench.log 2>&1; cat gpurun_out/gemm_bench.log | grep -v Warn
SLSB_DEBUG_FLAGS=4 timeout 300 python tools/gemm_bench.py 2>&1 | grep -E "inplace" 
SLSB_DEBUG_FLAGS=1 timeout 300 python tools/gemm_bench.py 2>&1 | grep -E "^(qkv|fc1|out_bf16|fc2_bf16)"
SLSB_GEMM_TRACE=1 timeout 300 python tools/gemm_bench.py > gpurun_out/gemm_trace.log 2>&1
grep -A12 "== out_inplace" gpurun_out/gemm_trace.log | tail -9 | cut -c1-220
grep -A30 "== fc1" gpurun_out/gemm_trace.log | grep -A14 "gemm_trace" | tail -14 | cut -c1-220
