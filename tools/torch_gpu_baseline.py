"""Same-box GPU baseline: the oracle restatement of the reference path run as STOCK PYTORCH on the
B200 (cuDNN / cuBLAS eager kernels), fp32 and bf16 autocast, B = 64 x 64 600-sample clips.  Measurement tool only (it
imports oracle/, so it is neither part of the product nor of bench.py); numbers are recorded in profiles/ and DESIGN.md."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle.heads import OracleModel
from oracle.trunk import synth_clips

head = sys.argv[1] if len(sys.argv) > 1 else "sls"
B = 64
m = OracleModel(head=head).eval().cuda()
x = synth_clips(0, B).cuda()
res = {"head": head, "batch": B, "torch": torch.__version__}
for name, ctx in (("fp32", torch.autocast("cuda", enabled=False)), ("bf16_autocast", torch.autocast("cuda", dtype=torch.bfloat16))):
    with torch.no_grad(), ctx:
        for _ in range(3):
            m(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 10
        for _ in range(n):
            m(x)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
    res[name] = {"ms_per_step": ms, "utt_per_s": B / ms * 1e3}
print(json.dumps(res))
