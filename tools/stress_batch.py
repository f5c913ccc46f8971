"""Race hunt: a clip's score must not depend on the batch around it, nor on the run.  Repeats full / permuted / sub-batches and,
when a score differs, reports the first layer result that differs and where (clip, frames, channels) - the row / column pattern
tells which kernel's tiling is involved."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sls_b200

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 24
torch.manual_seed(1234)
geo = sls_b200.TrunkGeometry(layers=nl)
m = sls_b200.ModelSLS(None, "cuda", cp_path=None, geometry=geo).to("cuda").eval()
eng = m.engine()
wav = eng.synth_clips(0, 64)
T, D = 201, 1024
LAYERS = tuple(range(nl))
def run(x):
    with torch.no_grad():
        out = m(x).clone()
    B = x.shape[0]
    return out, [eng.get_tensor(f"layer_results.{i}", (B, T, D)) for i in LAYERS], eng.get_tensor("sls_weights", (B, nl))
bad = 0
full0 = run(wav)
for rep in range(reps):
    g = torch.Generator().manual_seed(rep)
    perm = torch.randperm(64, generator=g).to("cuda")
    for name, idx in (("same", torch.arange(0, 64, device="cuda")), ("perm", perm)) + ((("pair", torch.arange(0, 2, device="cuda")),
                      ("mid5", torch.arange(30, 35, device="cuda"))) if rep % 16 == 0 else ()):
        sub = run(wav[idx].contiguous())
        if not torch.equal(sub[0], full0[0][idx]):
            bad += 1
            rows = (sub[0] != full0[0][idx]).any(-1).nonzero().flatten().tolist()
            first = [l for l, a, b in zip(LAYERS, sub[1], full0[1]) if not torch.equal(a, b[idx])]
            wdiff = not torch.equal(sub[2], full0[2][idx])
            msg = f"rep {rep} {name}: MISMATCH clips(pos) {rows[:8]} max|d|={float((sub[0] - full0[0][idx]).abs().max()):.2e} layers differing {first[:6]}{'...' if len(first) > 6 else ''} sls_weights differ {wdiff}"
            if first:
                l = first[0]
                d = (sub[1][l] != full0[1][l][idx]).nonzero()
                pos = sorted(set(d[:, 0].tolist()))
                fr = d[:, 1]; ch = d[:, 2]
                msg += (f" | layer {l}: {d.shape[0]} elements, clip positions {pos[:6]}, global rows {int(pos[0]) * T + int(fr.min())}..{int(pos[0]) * T + int(fr.max())} "
                        f"frames {int(fr.min())}..{int(fr.max())} ({len(set(fr.tolist()))} distinct), channels {int(ch.min())}..{int(ch.max())} ({len(set(ch.tolist()))} distinct), "
                        f"max|d|={float((sub[1][l] - full0[1][l][idx]).abs().max()):.3e}")
                # hypotheses about the corrupted segment (first differing clip / frame only)
                pc, fr0 = int(d[0, 0]), int(d[0, 1])
                c_lo, c_hi = int(ch.min()) // 64 * 64, int(ch.max()) // 64 * 64 + 64
                got = sub[1][l][pc, fr0, c_lo:c_hi]
                ref = full0[1][l][idx][pc, fr0, c_lo:c_hi]
                msg += f" | cols [{c_lo},{c_hi}) got[:6]={[round(float(v), 4) for v in got[:6]]} ref[:6]={[round(float(v), 4) for v in ref[:6]]}"
                if l > 0:
                    prev = full0[1][l - 1][idx][pc, fr0, c_lo:c_hi]
                    msg += f" prev_layer[:6]={[round(float(v), 4) for v in prev[:6]]} frac(got==prev)={float((got == prev).float().mean()):.2f}"
                    msg += f" |got-ref| vs |ref-prev| ratio median={float(((got - ref).abs() / ((ref - prev).abs() + 1e-9)).median()):.3f}"
                for ll in (l - 1, l):
                    if ll < 0:
                        continue
                    allrows = sub[1][ll][:, :, c_lo:c_hi].reshape(-1, c_hi - c_lo)
                    hit = ((allrows == got[None, :]).float().mean(-1) > 0.9).nonzero().flatten().tolist()
                    msg += f" rows of layer {ll} (this batch) equal to the got segment: {hit[:6]}"
            print(msg, flush=True)
print(f"layers={nl} reps={reps} mismatches={bad} of ~{reps * 2} env={ {k: v for k, v in os.environ.items() if k.startswith('SLSB')} }", flush=True)
