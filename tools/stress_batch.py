"""Race hunt 2: a clip's score must not depend on the batch around it.  Repeats full / pair / tail / permuted batches and reports
the first layer result that differs when a score does."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sls_b200

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
torch.manual_seed(1234)
m = sls_b200.ModelSLS(None, "cuda", cp_path=None).to("cuda").eval()
eng = m.engine()
wav = eng.synth_clips(0, 64)
T, D = 201, 1024
LAYERS = (0, 1, 2, 3, 5, 11, 23)
def run(x):
    with torch.no_grad():
        out = m(x).clone()
    B = x.shape[0]
    return out, [eng.get_tensor(f"layer_results.{i}", (B, T, D)) for i in LAYERS], eng.get_tensor("sls_weights", (B, 24))
bad = 0
for rep in range(reps):
    g = torch.Generator().manual_seed(rep)
    perm = torch.randperm(64, generator=g).to("cuda")
    full = run(wav)
    for name, idx in (("pair", torch.arange(0, 2, device="cuda")), ("tail", torch.arange(62, 64, device="cuda")), ("perm", perm),
                      ("mid5", torch.arange(30, 35, device="cuda"))):
        sub = run(wav[idx].contiguous())
        if not torch.equal(sub[0], full[0][idx]):
            bad += 1
            rows = (sub[0] != full[0][idx]).any(-1).nonzero().flatten().tolist()
            first = [l for l, a, b in zip(LAYERS, sub[1], full[1]) if not torch.equal(a, b[idx])]
            wdiff = not torch.equal(sub[2], full[2][idx])
            l0 = (sub[1][0] != full[1][0][idx]).nonzero()
            print(f"rep {rep} {name}: MISMATCH clips {rows[:8]} (of {len(rows)}) layers differing {first} sls_weights differ {wdiff} layer0 diffs {l0.shape[0]} {l0[:4].tolist()}", flush=True)
print(f"reps={reps} mismatches={bad} env={ {k: v for k, v in os.environ.items() if k.startswith('SLSB')} }")
