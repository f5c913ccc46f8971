"""Short workload for ncu: W warm-up forwards + N profiled forwards of the bench configuration (B=64, bf16, H-SAE)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sls_b200

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--warmup", type=int, default=1)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--head", default="sae")
ap.add_argument("--layers", type=int, default=24)
a = ap.parse_args()
torch.manual_seed(0)
geo = sls_b200.TrunkGeometry(layers=a.layers)
if a.head == "sls":
    m = sls_b200.ModelSLS(None, "cuda", cp_path=None, geometry=geo)
    head = sls_b200.HEAD_SLS
else:
    m = (sls_b200.ModelWindowTopK if a.head == "window" else sls_b200.Model)(None, "cuda", cp_path=None, geometry=geo)
    head = sls_b200.HEAD_WINDOW if a.head == "window" else sls_b200.HEAD_SAE
m = m.to("cuda").eval()
eng = m.engine()
wav = eng.synth_clips(0, a.batch)
for _ in range(a.warmup):
    eng.forward(wav, head, sls_b200.PREC_BF16)
torch.cuda.synchronize()
l0 = eng.launch_count
for _ in range(a.steps):
    out = eng.forward(wav, head, sls_b200.PREC_BF16)
torch.cuda.synchronize()
print("ok", out[:2].tolist(), "launches/step", (eng.launch_count - l0) / a.steps)
