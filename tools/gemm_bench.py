"""Times the encoder GEMM shapes of the bench workload through slsb_op_gemm with CUDA events (bf16 tcgen05 path)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sls_b200
lib = sls_b200.load_library()
P = lambda t: None if t is None else C.c_void_p(t.data_ptr())
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
M = 12864
shapes = [("qkv", 3072, 1024, 0, 1, False), ("out", 1024, 1024, 0, 0, True), ("out_inplace", 1024, 1024, 0, 0, "inplace"), ("fc2_inplace", 1024, 4096, 0, 0, "inplace"), ("out_bf16_nores", 1024, 1024, 0, 1, False),
          ("fc1", 4096, 1024, 1, 1, False), ("fc2", 1024, 4096, 0, 0, True), ("fc2_bf16_nores", 1024, 4096, 0, 1, False),
          ("sq4096", 4096, 4096, 0, 1, False)]
torch.manual_seed(0)
for name, N, K, act, obf, res in shapes:
    A = torch.randn(M, K, device="cuda").bfloat16()
    W = (torch.randn(N, K, device="cuda") * 0.03).bfloat16()
    b = torch.randn(N, device="cuda")
    R = torch.randn(M, N, device="cuda") if res else None
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16 if obf else torch.float32)
    if res == "inplace":        # residual == out: the reduce-add epilogue of the in-place stream
        out = R
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    if os.environ.get("SLSB_GEMM_TRACE"):           # timeline mode: two launches per shape (cold, warm), traces on stderr
        print(f"== {name}", file=sys.stderr, flush=True)
        for _ in range(2):
            lib.slsb_op_gemm(1, P(A), P(W), P(b), P(R), P(out), M, N, K, act, obf, st())
            torch.cuda.synchronize()
        continue
    for _ in range(3):
        lib.slsb_op_gemm(1, P(A), P(W), P(b), P(R), P(out), M, N, K, act, obf, st())
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lib.slsb_op_gemm(1, P(A), P(W), P(b), P(R), P(out), M, N, K, act, obf, st())
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    med = ts[len(ts) // 2]
    print(f"{name:16s} N={N:5d} K={K:5d} med={med*1e3:8.1f} us  min={ts[0]*1e3:8.1f} us  {2.0*M*N*K/med/1e9:8.1f} TFLOP/s (med)  flags={os.environ.get('SLSB_DEBUG_FLAGS','0')}")
if os.environ.get("SLSB_GEMM_TRACE"):
    sys.exit(0)
# cuBLAS reference for the same shapes (library baseline, not the product)
for name, N, K in [("qkv", 3072, 1024), ("out", 1024, 1024), ("fc1", 4096, 1024), ("fc2", 1024, 4096), ("sq4096", 4096, 4096)]:
    A = torch.randn(M, K, device="cuda").bfloat16(); W = torch.randn(N, K, device="cuda").bfloat16()
    for _ in range(3): torch.matmul(A, W.t())
    torch.cuda.synchronize(); ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(A, W.t()); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); med = ts[len(ts)//2]
    print(f"cublas {name:9s} N={N:5d} K={K:5d} med={med*1e3:8.1f} us {2.0*M*N*K/med/1e9:8.1f} TFLOP/s")
