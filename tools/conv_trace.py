"""Timeline of the conv + LN + GELU pair kernel (pair 0): SLSB_LN2_TRACE=1 python tools/conv_trace.py  (conv1 / conv2 shapes of the bench)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sls_b200
lib = sls_b200.load_library()
P = lambda t: C.c_void_p(t.data_ptr())
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
torch.manual_seed(0)
for name, B, Lin, k, s in (("conv1", 64, 12919, 3, 2), ("conv5", 64, 806, 2, 2)):
    Lout = (Lin - k) // s + 1
    x = torch.randn(B, Lin, 512, device="cuda").bfloat16()
    W = (torch.randn(512, k * 512, device="cuda") * 0.03).bfloat16()
    b, g, h = torch.randn(512, device="cuda") * 0.05, 1 + 0.1 * torch.randn(512, device="cuda"), 0.05 * torch.randn(512, device="cuda")
    out = torch.empty(B, Lout, 512, device="cuda", dtype=torch.bfloat16)
    print("==", name, file=sys.stderr, flush=True)
    for _ in range(2):
        assert lib.slsb_op_conv_ln_gelu(P(x), P(W), P(b), P(g), P(h), P(out), B, Lin, 512, k, s, st()) == 0
        torch.cuda.synchronize()
    if not os.environ.get("SLSB_LN2_TRACE"):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            lib.slsb_op_conv_ln_gelu(P(x), P(W), P(b), P(g), P(h), P(out), B, Lin, 512, k, s, st())
        e1.record(); torch.cuda.synchronize()
        print(f"{name}: {e0.elapsed_time(e1) * 100:.1f} us per launch  {2.0 * B * Lout * 512 * k * 512 / (e0.elapsed_time(e1) / 10) / 1e9:.0f} TFLOP/s")
