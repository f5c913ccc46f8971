"""Timeline of the persistent attention kernel (CTA 0): prints per-unit pipeline latencies in SM cycles."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
import sls_b200
lib = sls_b200.load_library()
B, T, H = 64, 201, 16
torch.manual_seed(0)
qkv = (torch.randn(B, T, 3 * H * 64, device="cuda") * 0.5).bfloat16()
out = torch.zeros(B, T, H * 64, device="cuda", dtype=torch.bfloat16)
trace = torch.zeros(64, 16, device="cuda", dtype=torch.int64)
P = lambda t: C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(3):
    assert lib.slsb_op_attention_trace(P(qkv), P(out), B, T, H, P(trace), st) == 0
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    lib.slsb_op_attention_trace(P(qkv), P(out), B, T, H, P(trace), st)
e1.record(); torch.cuda.synchronize()
print(f"kernel: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us per launch")
tr = trace.cpu()
t0 = int(tr[0, 9]) if int(tr[0, 9]) else int(tr[tr > 0].min())
names = ["S_iss", "S_com", "P_seen", "PV_com", "S_seen", "max_dn", "P_pub", "O_seen", "epi_dn", "st_free", "ld_land"]
print("unit " + " ".join(f"{n:>8s}" for n in names) + " | S_lat softmax(p1,p2) PVwake PV_lat epi period")
prev = {}
NW = int(os.environ.get("SLSB_ATTN_NW", "2"))
for u in range(24):
    r = [int(tr[u, e]) - t0 if int(tr[u, e]) else -1 for e in range(11)]
    s_lat = r[4] - r[1]; p1 = r[5] - r[4]; p2 = r[6] - r[5]; pvw = r[2] - r[6]; pv_lat = r[7] - r[3]; epi = r[8] - r[7]
    per = r[8] - prev.get(u % NW, r[8]); prev[u % NW] = r[8]
    print(f"{u:4d} " + " ".join(f"{v:8d}" for v in r) + f" | {s_lat:5d} {p1:5d} {p2:5d} {pvw:6d} {pv_lat:6d} {epi:5d} {per:6d}")
