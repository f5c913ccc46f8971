"""Race hunt: the same full-size forward N times, every output compared bit for bit with the first run; on a mismatch the
layer-result snapshots are compared too, to find the first layer that differs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sls_b200

head = sys.argv[1] if len(sys.argv) > 1 else "sls"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 200
B = int(sys.argv[3]) if len(sys.argv) > 3 else 64
torch.manual_seed(1234)
m = (sls_b200.ModelSLS(None, "cuda", cp_path=None) if head == "sls" else sls_b200.Model(None, "cuda", cp_path=None)).to("cuda").eval()
if head != "sls":
    m.retain_intermediates = True
eng = m.engine()
wav = eng.synth_clips(0, B)
T, D, L = 201, 1024, 24
def run():
    with torch.no_grad():
        out = m(wav) if head == "sls" else m(wav, return_sae_loss=False)
    return out.clone(), [eng.get_tensor(f"layer_results.{i}", (B, T, D)) for i in (0, 1, 2, 5, 11, 23)], eng.get_tensor("x", (B, T, D))
ref = run()
bad = 0
for it in range(n):
    cur = run()
    if not torch.equal(cur[0], ref[0]) or not torch.equal(cur[2], ref[2]):
        bad += 1
        first = [i for i, (a, b) in zip((0, 1, 2, 5, 11, 23), zip(cur[1], ref[1])) if not torch.equal(a, b)]
        d = (cur[1][0] != ref[1][0]).nonzero()
        print(f"iter {it}: MISMATCH logprob_equal={torch.equal(cur[0], ref[0])} x_equal={torch.equal(cur[2], ref[2])} first differing layers {first} "
              f"layer0 diff count {d.shape[0]} sample {d[:5].tolist()}", flush=True)
print(f"head={head} B={B} iters={n} mismatches={bad} env={ {k: v for k, v in os.environ.items() if k.startswith('SLSB')} }")
