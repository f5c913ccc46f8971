"""Does replaying the forward as a CUDA graph beat the eager launch sequence?  (launch gaps between the ~185 kernels)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sls_b200
m = sls_b200.ModelSLS(None, "cuda", cp_path=None).to("cuda").eval()
eng = m.engine()
wav = eng.synth_clips(0, 64)
head, prec = sls_b200.HEAD_SLS, sls_b200.PREC_BF16
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3):
        out = eng.forward(wav, head, prec)
    torch.cuda.synchronize()
    def timeit(fn, n=20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    print("eager  ms/step:", timeit(lambda: eng.forward(wav, head, prec)))
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        out_g = eng.forward(wav, head, prec)
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    print("graph  ms/step:", timeit(g.replay))
    print("eager  ms/step:", timeit(lambda: eng.forward(wav, head, prec)))
    print("graph  ms/step:", timeit(g.replay))
    ref = eng.forward(wav, head, prec)
    g.replay(); torch.cuda.synchronize()
    print("bit-identical:", bool(torch.equal(ref, out_g)))
