"""Debug aid: bf16 vs fp32 product path on long clips, layer by layer (which kernel diverges for T > 201?)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sls_b200

S = int(sys.argv[1]) if len(sys.argv) > 1 else 160000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
attn = int(sys.argv[3]) if len(sys.argv) > 3 else 0
layers = int(sys.argv[4]) if len(sys.argv) > 4 else 4
geo = sls_b200.TrunkGeometry(layers=layers)
torch.manual_seed(1)
ms = {}
for prec in ("fp32", "bf16"):
    torch.manual_seed(1)
    m = sls_b200.Model(None, "cuda", cp_path=None, precision=prec, geometry=geo)
    if prec == "bf16" and attn:
        base = m._engine_config
        def cfgf(base=base):
            c = base(); c.attn_impl = attn; return c
        m._engine_config = cfgf
    m.retain_intermediates = True
    ms[prec] = m.to("cuda").eval()
ms["bf16"].load_state_dict(ms["fp32"].state_dict())
eng0 = ms["fp32"].engine()
wav = eng0.synth_clips(500, B, S)
lens = torch.tensor([S - 7000 * i for i in range(B)], dtype=torch.int32, device="cuda")
T = eng0.frames(S)
res = {}
for prec, m in ms.items():
    with torch.no_grad():
        out = m(wav, return_sae_loss=False, sample_lengths=lens)
    eng = m.engine()
    res[prec] = {"out": out.cpu(), "x": eng.get_tensor("x", (B, T, 1024)).cpu(),
                 "layers": [eng.get_tensor(f"layer_results.{i}", (B, T, 1024)).cpu() for i in range(layers)]}
fl = [eng0.frames(int(n)) for n in lens.tolist()]
def rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))
print(f"S={S} B={B} T={T} attn={attn} frame_lens={fl} env={ {k: v for k, v in os.environ.items() if k.startswith('SLSB')} }")
print("out bf16", res["bf16"]["out"].tolist(), "fp32", res["fp32"]["out"].tolist())
for i in range(layers):
    a, b = res["bf16"]["layers"][i], res["fp32"]["layers"][i]
    per = [rel(a[j, :fl[j]], b[j, :fl[j]]) for j in range(B)]
    # where along T is the error?
    e = (a[0, :fl[0]] - b[0, :fl[0]]).norm(dim=-1) / (b[0, :fl[0]].norm(dim=-1) + 1e-30)
    worst = torch.topk(e, 5).indices.tolist()
    print(f"layer {i}: rel per utt {['%.3e' % p for p in per]} worst frames utt0 {worst} ({['%.2e' % float(e[w]) for w in worst]})")
print("x rel", [rel(res["bf16"]["x"][j, :fl[j]], res["fp32"]["x"][j, :fl[j]]) for j in range(B)])
