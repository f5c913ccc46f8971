"""Kernel-level parity: every CUDA kernel, called through the C ABI, against a plain PyTorch fp32
reference of the same op on the same seeded inputs.  Tolerances are written next to each check."""
import math

import pytest
import torch
import torch.nn.functional as F

from helpers import P, ok, report, stream

pytestmark = pytest.mark.gpu

# the PyTorch references must be true fp32: cuDNN / cuBLAS default to TF32 for fp32 convs and matmuls on this GPU
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

FP32, BF16 = 0, 1
ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2


def _rand(shape, seed, scale=1.0, device="cuda"):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(device)


def _act(x, act):
    return F.gelu(x) if act == ACT_GELU else (F.relu(x) if act == ACT_RELU else x)


# ------------------------------------------------------------------------------------------ SIMT GEMM
@pytest.mark.parametrize("M,N,K,act,res", [(300, 200, 64, ACT_NONE, False), (257, 1024, 512, ACT_GELU, False),
                                           (1000, 512, 1024, ACT_NONE, True), (128, 4096, 1024, ACT_RELU, False)])
def test_gemm_fp32(lib, cuda, M, N, K, act, res):
    A, W, b = _rand((M, K), 1), _rand((N, K), 2, 0.05), _rand((N,), 3)
    R = _rand((M, N), 4) if res else None
    out = torch.empty(M, N, device=cuda)
    ok(lib, lib.slsb_op_gemm(FP32, P(A), P(W), P(b), P(R), P(out), M, N, K, act, 0, stream()), "gemm fp32")
    ref = _act(A.double() @ W.double().T + b.double(), act)
    if res:
        ref = ref + R.double()
    report(f"gemm_fp32 {M}x{N}x{K}", out, ref.float(), atol=2e-5, rtol=2e-5)   # fp32 accumulate over K <= 1024


# ------------------------------------------------------------------------------------------ tcgen05 GEMM
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 256, 256), (256, 512, 1024), (1000, 1024, 1024), (777, 3072, 1024),
                                   (640, 1024, 4096), (300, 384, 128), (300, 192, 512), (12864, 1024, 1024)])
def test_gemm_bf16_tc_bias(lib, cuda, M, N, K):
    A, W, b = _rand((M, K), 1).bfloat16(), _rand((N, K), 2, 0.05).bfloat16(), _rand((N,), 3)
    out = torch.empty(M, N, device=cuda, dtype=torch.bfloat16)
    ok(lib, lib.slsb_op_gemm(BF16, P(A), P(W), P(b), None, P(out), M, N, K, ACT_NONE, 1, stream()), "gemm tc")
    ref = A.float() @ W.float().T + b
    report(f"gemm_tc {M}x{N}x{K}", out, ref, atol=2e-3, rtol=8e-3)      # output rounded to bf16 (2^-8 rel)


@pytest.mark.parametrize("act,out_bf16,res", [(ACT_GELU, 1, False), (ACT_NONE, 0, True), (ACT_NONE, 0, False),
                                              (ACT_RELU, 0, False), (ACT_GELU, 0, True)])
def test_gemm_bf16_tc_epilogues(lib, cuda, act, out_bf16, res):
    M, N, K = 1000, 1024, 512
    A, W, b = _rand((M, K), 5).bfloat16(), _rand((N, K), 6, 0.05).bfloat16(), _rand((N,), 7)
    R = _rand((M, N), 8) if res else None
    out = torch.empty(M, N, device=cuda, dtype=torch.bfloat16 if out_bf16 else torch.float32)
    ok(lib, lib.slsb_op_gemm(BF16, P(A), P(W), P(b), P(R), P(out), M, N, K, act, out_bf16, stream()), "gemm tc epi")
    ref = _act(A.float() @ W.float().T + b, act)
    if res:
        ref = ref + R
    report(f"gemm_tc epi act={act} bf16={out_bf16} res={res}", out, ref, atol=2e-3 if out_bf16 else 2e-4, rtol=8e-3 if out_bf16 else 1e-4)


@pytest.mark.parametrize("M,N,K", [(1000, 1024, 512), (12864, 1024, 1024), (201, 1024, 4096), (300, 256, 64)])
def test_gemm_bf16_tc_in_place_residual(lib, cuda, M, N, K):
    """residual == out: the fp32 stream is advanced in place (TMA reduce-add epilogue); bit-stable, equals out-of-place."""
    A, W, b = _rand((M, K), 14).bfloat16(), _rand((N, K), 15, 0.05).bfloat16(), _rand((N,), 16)
    R = _rand((M, N), 17)
    inplace = R.clone()
    ok(lib, lib.slsb_op_gemm(BF16, P(A), P(W), P(b), P(inplace), P(inplace), M, N, K, ACT_NONE, 0, stream()), "gemm tc in place")
    report(f"gemm_tc in-place {M}x{N}x{K}", inplace, A.float() @ W.float().T + b + R, atol=2e-4 * (K / 512) ** 0.5, rtol=1e-4)
    separate = torch.empty_like(R)
    ok(lib, lib.slsb_op_gemm(BF16, P(A), P(W), P(b), P(R), P(separate), M, N, K, ACT_NONE, 0, stream()), "gemm tc out of place")
    assert torch.equal(inplace, separate)                      # same association: (acc + bias) + residual
    again = R.clone()
    ok(lib, lib.slsb_op_gemm(BF16, P(A), P(W), P(b), P(again), P(again), M, N, K, ACT_NONE, 0, stream()), "gemm tc in place")
    assert torch.equal(inplace, again)


@pytest.mark.parametrize("M,N,K,act", [(4864, 1024, 512, ACT_NONE), (6221, 1024, 256, ACT_GELU), (12864, 3072, 256, ACT_NONE),
                                       (12864, 4096, 128, ACT_GELU), (9000, 768 + 256, 192, ACT_RELU)])
def test_gemm_bf16_tc_tail_slices(lib, cuda, M, N, K, act):
    """Tile counts that leave a partial last round on 74 CTA pairs: its tiles are cut into 128/64-column slices.  Same result
    as the torch product, and bit-identical to the same rows computed by a launch small enough to have no partial round."""
    A, W, b = _rand((M, K), 31).bfloat16(), _rand((N, K), 32, 0.05).bfloat16(), _rand((N,), 33)
    out = torch.full((M, N), float("nan"), device=cuda, dtype=torch.bfloat16)
    ok(lib, lib.slsb_op_gemm(BF16, P(A), P(W), P(b), None, P(out), M, N, K, act, 1, stream()), "gemm tc tail")
    report(f"gemm_tc tail {M}x{N}x{K}", out, _act(A.float() @ W.float().T + b, act), atol=2e-3, rtol=8e-3)
    rows = slice(M - 300, M)                                   # the ragged end of the matrix: always inside the last round
    At = A[rows].contiguous()
    small = torch.empty(300, N, device=cuda, dtype=torch.bfloat16)
    ok(lib, lib.slsb_op_gemm(BF16, P(At), P(W), P(b), None, P(small), 300, N, K, act, 1, stream()), "gemm tc small")
    assert torch.equal(out[rows], small)


def test_gemm_bf16_tc_deterministic(lib, cuda):
    M, N, K = 2000, 1024, 1024
    A, W, b = _rand((M, K), 9).bfloat16(), _rand((N, K), 10, 0.05).bfloat16(), _rand((N,), 11)
    outs = []
    for _ in range(3):
        out = torch.empty(M, N, device=cuda, dtype=torch.bfloat16)
        ok(lib, lib.slsb_op_gemm(BF16, P(A), P(W), P(b), None, P(out), M, N, K, ACT_NONE, 1, stream()), "gemm tc")
        outs.append(out)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])   # bit-stable


@pytest.mark.parametrize("M,N,K,KS", [(64, 1024, 22848, 36), (2, 1024, 22848, 36), (100, 512, 640, 5), (130, 256, 1024, 16)])
def test_gemm_bf16_tc_split_k(lib, cuda, M, N, K, KS):
    """SLS fc1 path: partials[m][s][n] summed over the splits == the full product; splits are bit-stable."""
    A, W = _rand((M, K), 12).bfloat16(), _rand((N, K), 13, 0.05).bfloat16()
    part = torch.full((M, KS, N), float("nan"), device=cuda)
    ok(lib, lib.slsb_op_gemm_splitk(P(A), P(W), P(part), M, N, K, KS, stream()), "gemm split-k")
    ref = A.double() @ W.double().T
    report(f"gemm_tc split-k {M}x{N}x{K}/{KS}", part.double().sum(1), ref, atol=2e-3 * (K / 1024) ** 0.5, rtol=1e-4)
    per = -(-(K // 64) // KS) * 64
    report("gemm_tc split-k slab 1", part[:, 1], A[:, per:2 * per].double() @ W[:, per:2 * per].double().T, atol=1e-3, rtol=1e-4)
    part2 = torch.empty_like(part)
    ok(lib, lib.slsb_op_gemm_splitk(P(A), P(W), P(part2), M, N, K, KS, stream()), "gemm split-k")
    assert torch.equal(part, part2)


# ------------------------------------------------------------------------------------------ strided convs as implicit GEMM
@pytest.mark.parametrize("prec", [FP32, BF16])
@pytest.mark.parametrize("B,Lin,k,s", [(2, 1291, 3, 2), (3, 403, 2, 2), (1, 806, 2, 2), (2, 6459, 3, 2)])
def test_conv_implicit_gemm(lib, cuda, prec, B, Lin, k, s):
    C = N = 512
    dt = torch.bfloat16 if prec == BF16 else torch.float32
    x = _rand((B, Lin, C), 20).to(dt)
    w = _rand((N, C, k), 21, 1.0 / math.sqrt(C * k)).to(dt)
    b = _rand((N,), 22)
    Lout = (Lin - k) // s + 1
    wp = w.permute(0, 2, 1).reshape(N, k * C).contiguous()
    out = torch.empty(B, Lout, N, device=cuda, dtype=dt)
    ok(lib, lib.slsb_op_conv(prec, P(x), P(wp), P(b), P(out), B, Lin, C, N, k, s, stream()), "conv")
    ref = F.conv1d(x.float().transpose(1, 2), w.float(), b, stride=s).transpose(1, 2)
    report(f"conv prec={prec} B={B} Lin={Lin} k={k}", out, ref, atol=2e-3 if prec == BF16 else 2e-5, rtol=8e-3 if prec == BF16 else 2e-5)


# ------------------------------------------------------------------------------------------ positional conv
@pytest.mark.parametrize("prec", [FP32, BF16])
@pytest.mark.parametrize("B,T,lens", [(2, 201, None), (3, 97, [97, 50, 80]), (1, 256, None), (2, 257, None), (3, 374, [374, 300, 120]), (3, 499, None),
                                      (2, 512, None), (1, 700, None)])
def test_posconv(lib, cuda, prec, B, T, lens):
    D, K, G = 1024, 128, 16
    x = _rand((B, T, D), 30)
    w = _rand((D, D // G, K), 31, math.sqrt(4.0 / (K * D)) * 3)
    b = _rand((D,), 32, 0.05)
    wp = w.permute(0, 2, 1).reshape(D, K * (D // G)).contiguous()
    es = 2 if prec == BF16 else 4
    scratch = torch.zeros(B * (T + K) * D * es + (8 << 20), device=cuda, dtype=torch.uint8)
    lens_t = torch.tensor(lens, dtype=torch.int32, device=cuda) if lens else None
    out = torch.empty(B, T, D, device=cuda)
    wdev = wp.bfloat16() if prec == BF16 else wp
    ok(lib, lib.slsb_op_posconv(prec, P(x), P(wdev), P(b), P(out), P(scratch), B, T, D, K, P(lens_t), stream()), "posconv")
    xin = x.clone()
    if lens:
        for i, n in enumerate(lens):
            xin[i, n:] = 0
    if prec == BF16:
        xc, wc = xin.bfloat16().float(), w.bfloat16().float()
    else:
        xc, wc = xin, w
    y = F.conv1d(xc.transpose(1, 2), wc, b, padding=K // 2, groups=G)[:, :, :-1]
    ref = x + F.gelu(y).transpose(1, 2)
    if lens:   # only valid frames are defined
        for i, n in enumerate(lens):
            out[i, n:] = 0
            ref[i, n:] = 0
    report(f"posconv prec={prec}", out, ref, atol=3e-3 if prec == BF16 else 3e-5, rtol=1e-3 if prec == BF16 else 1e-5)


# ------------------------------------------------------------------------------------------ conv0 + LN + GELU
@pytest.mark.parametrize("out_bf16", [0, 1])
def test_conv0_ln_gelu(lib, cuda, out_bf16):
    B, S, C = 3, 4000, 512
    wav = _rand((B, S), 40)
    w, b = _rand((C, 10), 41, math.sqrt(2.0 / 10)), _rand((C,), 42, 0.05)
    g, be = 1 + _rand((C,), 43, 0.1), _rand((C,), 44, 0.05)
    L0 = (S - 10) // 5 + 1
    out = torch.empty(B, L0, C, device=cuda, dtype=torch.bfloat16 if out_bf16 else torch.float32)
    ok(lib, lib.slsb_op_conv0(out_bf16, P(wav), P(w), P(b), P(g), P(be), P(out), B, S, 0 if out_bf16 else 1, stream()), "conv0")
    y = F.conv1d(wav.unsqueeze(1), w.unsqueeze(1), b, stride=5).transpose(1, 2)
    ref = F.gelu(F.layer_norm(y, (C,), g, be, 1e-5))
    report(f"conv0 bf16={out_bf16}", out, ref, atol=1e-2 if out_bf16 else 2e-5, rtol=8e-3 if out_bf16 else 1e-5)


def test_conv0_tensor_core_hi_lo_split(lib, cuda):
    """conv0 + LN + GELU as ONE tcgen05 kernel; the raw audio and the taps are split into bf16 hi + lo parts so the
    products keep ~16 mantissa bits (x_hi w_hi + x_hi w_lo + x_lo w_hi)."""
    B, S, C = 3, 16000, 512
    wav = _rand((B, S), 45)
    w, b = _rand((C, 10), 46, math.sqrt(2.0 / 10)), _rand((C,), 47, 0.05)
    g, be = 1 + _rand((C,), 48, 0.1), _rand((C,), 49, 0.05)
    L0 = (S - 10) // 5 + 1
    out = torch.empty(B, L0, C, device=cuda, dtype=torch.bfloat16)
    scratch = torch.zeros(65536 + B * L0 * 128 + (8 << 20), device=cuda, dtype=torch.uint8)
    ok(lib, lib.slsb_op_conv0_tc(P(wav), P(w), P(b), P(g), P(be), P(out), P(scratch), B, S, stream()), "conv0 tc")
    y = F.conv1d(wav.unsqueeze(1), w.unsqueeze(1), b, stride=5).transpose(1, 2)
    ref = F.gelu(F.layer_norm(y, (C,), g, be, 1e-5))
    report("conv0_tc", out, ref, atol=1e-2, rtol=8e-3)           # bf16 output rounding dominates
    # against the bf16-rounded fp32 result the agreement is (almost) bit-level: the split loses only ~2^-17
    assert float((out.float() - ref.bfloat16().float()).abs().max()) <= 2 ** -6


@pytest.mark.parametrize("B,S,kind", [(3, 16000, "noise"), (1, 700, "noise"), (64, 64600, "noise"), (5, 12345, "dc"), (2, 4000, "silence")])
def test_conv0_fused_one_kernel(lib, cuda, B, S, kind):
    """conv0 as ONE kernel (csrc/conv0_tc.cu): A tiles built in shared memory, bias in spare K columns, LayerNorm statistics from
    the 11 x 11 Gram matrix of [w | b] instead of a pass over the 512 outputs, 16 epilogue warps.  Cases: ragged last tile, rows
    crossing utterance boundaries inside a tile, the bench shape, a DC offset 30x the signal (mean >> std: the E[y^2] - mean^2 form
    is at its worst) and digital silence (every row equals the bias vector)."""
    C = 512
    wav = _rand((B, S), 45)
    if kind == "dc":
        wav = 0.03 * wav + 0.9
    elif kind == "silence":
        wav = torch.zeros_like(wav)
    w, b = _rand((C, 10), 46, math.sqrt(2.0 / 10)), _rand((C,), 47, 0.05)
    g, be = 1 + _rand((C,), 48, 0.1), _rand((C,), 49, 0.05)
    L0 = (S - 10) // 5 + 1
    out = torch.full((B, L0, C), float("nan"), device=cuda, dtype=torch.bfloat16)
    guard = torch.full((4096,), 7.0, device=cuda, dtype=torch.bfloat16)      # allocated right after `out`: rows past the end must not be written
    scratch = torch.zeros(65536 + 4096, device=cuda, dtype=torch.uint8)
    ok(lib, lib.slsb_op_conv0_fused(P(wav), P(w), P(b), P(g), P(be), P(out), P(scratch), B, S, stream()), "conv0 fused")
    torch.cuda.synchronize()
    assert bool((guard == 7.0).all())
    y = F.conv1d(wav.double().unsqueeze(1), w.double().unsqueeze(1), b.double(), stride=5).transpose(1, 2)
    ref = F.gelu(F.layer_norm(y, (C,), g.double(), be.double(), 1e-5)).float()
    assert bool(torch.isfinite(out.float()).all())
    # dc: var is ~1e-3 of mean^2, so fp32 cancellation in E[y^2] - mean^2 costs ~3 digits of rstd (same form as the two-pass kernels)
    report(f"conv0_fused B={B} S={S} {kind}", out, ref, atol=3e-2 if kind == "dc" else 1e-2, rtol=8e-3)
    if kind == "noise":
        # against the bf16-rounded fp64 result: at most one bf16 ulp (2^-7 relative), i.e. a rounding flip of a value that sits on a
        # rounding boundary (the 423 M outputs of the bench shape always contain a few)
        rb = ref.bfloat16().float()
        assert bool(((out.float() - rb).abs() <= rb.abs() * 2.0 ** -7 + 1e-4).all())
        # and against the round-1 kernel (im2col + whole-row accumulator): same split arithmetic, statistics from another route
        old = torch.empty_like(out)
        scratch2 = torch.zeros(65536 + B * L0 * 128 + (1 << 20), device=cuda, dtype=torch.uint8)
        ok(lib, lib.slsb_op_conv0_tc(P(wav), P(w), P(b), P(g), P(be), P(old), P(scratch2), B, S, stream()), "conv0 tc")
        assert bool(((out.float() - old.float()).abs() <= old.float().abs() * 2.0 ** -7 + 1e-4).all())
        assert float((out != old).float().mean()) < 1e-2          # and the two kernels disagree (by that one ulp) on < 1 % of the outputs


@pytest.mark.parametrize("B,Lin,k,s", [(2, 1291, 3, 2), (3, 403, 2, 2), (2, 6459, 3, 2), (5, 130, 3, 2)])
def test_conv_ln_gelu_fused(lib, cuda, B, Lin, k, s):
    C = N = 512
    x = _rand((B, Lin, C), 23).bfloat16()
    w = _rand((N, C, k), 24, 1.0 / math.sqrt(C * k)).bfloat16()
    b = _rand((N,), 25, 0.3)
    g, be = 1 + _rand((C,), 26, 0.1), _rand((C,), 27, 0.05)
    Lout = (Lin - k) // s + 1
    wp = w.permute(0, 2, 1).reshape(N, k * C).contiguous()
    out = torch.empty(B, Lout, N, device=cuda, dtype=torch.bfloat16)
    ok(lib, lib.slsb_op_conv_ln_gelu(P(x), P(wp), P(b), P(g), P(be), P(out), B, Lin, C, k, s, stream()), "conv ln gelu")
    y = F.conv1d(x.float().transpose(1, 2), w.float(), b, stride=s).transpose(1, 2)
    ref = F.gelu(F.layer_norm(y, (C,), g, be, 1e-5))
    report(f"conv_ln_gelu B={B} Lin={Lin} k={k}", out, ref, atol=1e-2, rtol=8e-3)


@pytest.mark.parametrize("C", [512, 1024])
@pytest.mark.parametrize("in_bf16,out_bf16,gelu", [(0, 0, 0), (0, 1, 0), (1, 1, 1), (0, 0, 1)])
def test_layernorm(lib, cuda, C, in_bf16, out_bf16, gelu):
    rows = 1003
    x = _rand((rows, C), 50, 2.0) + 0.5
    x = x.bfloat16() if in_bf16 else x
    g, b = 1 + _rand((C,), 51, 0.1), _rand((C,), 52, 0.05)
    out = torch.empty(rows, C, device=cuda, dtype=torch.bfloat16 if out_bf16 else torch.float32)
    ok(lib, lib.slsb_op_layernorm(P(x), in_bf16, P(out), out_bf16, P(g), P(b), rows, C, gelu, 1, stream()), "ln")
    ref = F.layer_norm(x.float(), (C,), g, b, 1e-5)
    if gelu:
        ref = F.gelu(ref)
    report(f"ln C={C}", out, ref, atol=2e-2 if out_bf16 else 1e-5, rtol=8e-3 if out_bf16 else 1e-5)


@pytest.mark.parametrize("rows", [63, 64, 1003, 12864])
@pytest.mark.parametrize("taps", [False, True])
def test_layernorm_stream_taps(lib, cuda, rows, taps):
    """Encoder LayerNorm(1024) fp32 -> bf16 (bulk-copy fed persistent kernel from 64 rows up, one-warp-per-row kernel below):
    normalised rows vs torch, bf16 snapshot == the input rounded to bf16 (bit-exact), fc0 dots vs float64, bit-stable."""
    C = 1024
    x = _rand((rows, C), 53, 2.0) + 0.5
    g, b, dw = 1 + _rand((C,), 54, 0.1), _rand((C,), 55, 0.05), _rand((C,), 56, 0.05)
    outs = []
    for _ in range(2):
        out = torch.full((rows, C), float("nan"), device=cuda, dtype=torch.bfloat16)
        snap = torch.full((rows, C), float("nan"), device=cuda, dtype=torch.bfloat16)
        dots = torch.full((rows,), float("nan"), device=cuda)
        ok(lib, lib.slsb_op_layernorm_taps(P(x), P(out), P(g), P(b), P(dw) if taps else None, P(dots) if taps else None,
                                           P(snap) if taps else None, rows, C, stream()), "ln taps")
        outs.append((out, snap, dots))
    out, snap, dots = outs[0]
    report(f"ln stream rows={rows}", out, F.layer_norm(x, (C,), g, b, 1e-5), atol=2e-2, rtol=8e-3)
    if taps:
        assert torch.equal(snap, x.bfloat16())
        report("ln stream dots", dots.double(), x.double() @ dw.double(), atol=1e-4, rtol=1e-5)
    assert all(torch.equal(a, c) for a, c in zip(outs[0][:3 if taps else 1], outs[1][:3 if taps else 1]))


# ------------------------------------------------------------------------------------------ attention
def _attn_ref(qkv, B, T, H, lens):
    D = H * 64
    q, k, v = qkv.float().view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    s = q @ k.transpose(-1, -2)
    if lens is not None:
        mask = torch.arange(T, device=qkv.device)[None, :] >= torch.tensor(lens, device=qkv.device)[:, None]
        s = s.masked_fill(mask[:, None, None, :], float("-inf"))
    o = torch.softmax(s, -1) @ v
    return o.permute(0, 2, 1, 3).reshape(B, T, D)


@pytest.mark.parametrize("impl,bf16", [(1, 0), (1, 1), (2, 1), (3, 1)])
@pytest.mark.parametrize("B,T,lens", [(2, 201, None), (3, 137, [137, 60, 1]), (1, 256, None), (2, 49, None), (40, 201, None), (21, 100, None)])
def test_attention(lib, cuda, impl, bf16, B, T, lens):
    H = 16
    qkv = _rand((B, T, 3 * H * 64), 60, 0.5)
    qkv[..., :H * 64] *= 0.125 * 4      # q pre-scaled (keeps the softmax reasonably peaked)
    qkv = qkv.bfloat16() if bf16 else qkv
    out = torch.zeros(B, T, H * 64, device=cuda, dtype=qkv.dtype)
    lens_t = torch.tensor(lens, dtype=torch.int32, device=cuda) if lens else None
    ok(lib, lib.slsb_op_attention(impl, bf16, P(qkv), P(out), B, T, H, P(lens_t), stream()), "attention")
    ref = _attn_ref(qkv, B, T, H, lens)
    if lens:
        for i, n in enumerate(lens):
            out[i, n:] = 0
            ref[i, n:] = 0
    report(f"attention impl={impl} bf16={bf16} T={T}", out, ref, atol=1.5e-2 if bf16 else 2e-5, rtol=1e-2 if bf16 else 1e-5)


@pytest.mark.parametrize("T,first_big", [(201, 128), (201, 112), (256, 128), (160, 150)])
def test_attention_tc_rebase_path(lib, cuda, T, first_big):
    """Keys of the second key block dominate (scores jump by far more than 2^8): exercises the lazy online-softmax re-base
    of the persistent kernel (O_a rescaled in TMEM); plus rows where only SOME lanes of a warp re-base."""
    B, H = 3, 16
    qkv = _rand((B, T, 3 * H * 64), 62, 0.5)
    qkv[..., :H * 64] *= 0.5
    qkv[:, first_big:, H * 64:2 * H * 64] *= 12.0          # later keys 12x larger
    qkv[1, first_big:, H * 64:2 * H * 64] *= 0.02           # utterance 1: later keys tiny instead -> no re-base there
    qkv[2, ::3, :H * 64] *= 0.01                            # utterance 2: every third query nearly flat -> mixed warps
    qkv = qkv.bfloat16()
    out = torch.zeros(B, T, H * 64, device=cuda, dtype=torch.bfloat16)
    ok(lib, lib.slsb_op_attention(2, 1, P(qkv), P(out), B, T, H, None, stream()), "attention")
    report(f"attention re-base T={T}", out, _attn_ref(qkv, B, T, H, None), atol=2e-2, rtol=2e-2)


@pytest.mark.parametrize("B,T,lens", [(2, 257, None), (3, 288, [288, 257, 31]), (2, 300, [300, 193]), (2, 374, None), (3, 374, [374, 200, 150]),
                                      (1, 416, None), (2, 499, None), (5, 499, [499, 480, 353, 340, 1]), (1, 512, None), (37, 499, None)])
def test_attention_tc_key_blocks(lib, cuda, B, T, lens):
    """257 <= T <= 512 (5-10 s clips): the persistent tcgen05 kernel in its wide geometry (one 192 KB stage, up to four query
    tiles, key blocks of <= 192 columns: max round then exp round against the global row max); with and without padding masks,
    incl. blocks that are entirely padding, vs the fp32 torch reference and vs the CUDA-core kernel."""
    H = 16
    qkv = _rand((B, T, 3 * H * 64), 63, 0.5)
    qkv[..., :H * 64] *= 0.5
    qkv[0, T // 2:, H * 64:2 * H * 64] *= 6.0              # utterance 0: the row maxima live in the later key blocks
    qkv = qkv.bfloat16()
    lens_t = torch.tensor(lens, dtype=torch.int32, device=cuda) if lens else None
    outs = []
    for _ in range(2):
        out = torch.full((B, T, H * 64), float("nan"), device=cuda, dtype=torch.bfloat16)
        ok(lib, lib.slsb_op_attention(2, 1, P(qkv), P(out), B, T, H, P(lens_t), stream()), "attention wide")
        outs.append(out)
    assert torch.equal(outs[0], outs[1])                     # bit-stable
    out = outs[0].clone()
    simt = torch.zeros_like(out)
    ok(lib, lib.slsb_op_attention(1, 1, P(qkv), P(simt), B, T, H, P(lens_t), stream()), "attention simt")
    ref = _attn_ref(qkv, B, T, H, lens)
    if lens:
        for i, n in enumerate(lens):
            out[i, n:] = 0
            ref[i, n:] = 0
            simt[i, n:] = 0
    report(f"attention tc wide T={T}", out, ref, atol=1.5e-2, rtol=1e-2)
    report(f"attention tc wide vs simt T={T}", out, simt, atol=1.5e-2, rtol=1e-2)


@pytest.mark.parametrize("B,T,lens", [(5, 201, [201, 130, 64, 17, 1]), (64, 201, None), (3, 129, None), (7, 255, [255, 200, 129, 128, 100, 50, 2]),
                                      (19, 240, None), (9, 144, [144, 143, 97, 96, 95, 49, 48, 47, 16])])
def test_attention_tc_two_tiles(lib, cuda, B, T, lens):
    """129 <= T <= 256 (two query tiles per item, the 4-s clips of the headline config): odd item counts per CTA, padding masks
    that end inside / at the edge of / before a 16-key group, rows of the last tile past T, bit-stability over three launches into
    NaN-filled outputs, vs the fp32 torch reference and the CUDA-core kernel."""
    H = 16
    qkv = _rand((B, T, 3 * H * 64), 65, 0.5)
    qkv[..., :H * 64] *= 0.5
    qkv[0, T // 2:, H * 64:2 * H * 64] *= 6.0              # utterance 0: the row maxima live in the later key blocks
    qkv = qkv.bfloat16()
    lens_t = torch.tensor(lens, dtype=torch.int32, device=cuda) if lens else None
    outs = []
    for _ in range(3):
        out = torch.full((B, T, H * 64), float("nan"), device=cuda, dtype=torch.bfloat16)
        ok(lib, lib.slsb_op_attention(2, 1, P(qkv), P(out), B, T, H, P(lens_t), stream()), "attention two tiles")
        outs.append(out)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])      # bit-stable
    out = outs[0].clone()
    assert not torch.isnan(out.float()).any()
    simt = torch.zeros_like(out)
    ok(lib, lib.slsb_op_attention(1, 1, P(qkv), P(simt), B, T, H, P(lens_t), stream()), "attention simt")
    ref = _attn_ref(qkv, B, T, H, lens)
    if lens:
        for i, n in enumerate(lens):
            out[i, n:] = 0
            ref[i, n:] = 0
            simt[i, n:] = 0
    report(f"attention tc two tiles T={T}", out, ref, atol=1.5e-2, rtol=1e-2)
    report(f"attention tc two tiles vs simt T={T}", out, simt, atol=1.5e-2, rtol=1e-2)


def test_attention_tc_rejects_T_above_512(lib, cuda):
    B, T, H = 1, 513, 16
    qkv = _rand((B, T, 3 * H * 64), 64, 0.5).bfloat16()
    out = torch.zeros(B, T, H * 64, device=cuda, dtype=torch.bfloat16)
    assert lib.slsb_op_attention(2, 1, P(qkv), P(out), B, T, H, None, stream()) != 0
    assert b"512" in lib.slsb_last_error()


def test_attention_long_simt(lib, cuda):
    B, T, H = 1, 499, 16
    qkv = _rand((B, T, 3 * H * 64), 61, 0.5)
    out = torch.zeros(B, T, H * 64, device=cuda)
    ok(lib, lib.slsb_op_attention(1, 0, P(qkv), P(out), B, T, H, None, stream()), "attention long")
    report("attention simt T=499", out, _attn_ref(qkv, B, T, H, None), atol=2e-5, rtol=1e-5)


# ------------------------------------------------------------------------------------------ top-k select
@pytest.mark.parametrize("D,k", [(4096, 128), (1024, 64), (2048, 7)])
def test_topk_matches_canonical_rule(lib, cuda, D, k):
    rows = 300
    x = F.relu(_rand((rows, D), 70))
    x[5] = 0.0                       # all-zero row: k zeros kept, output all zero
    x[6, :] = 1.0                    # all-equal row: lowest k indices win
    x[7, : D // 2] = x[7, D // 2:]   # duplicated values -> real ties at the threshold
    x[8, 10:] = 0.0                  # fewer than k positives
    thr = torch.empty(rows, device=cuda)
    cut = torch.empty(rows, device=cuda, dtype=torch.int32)
    enc = torch.empty(rows, D, device=cuda)
    ok(lib, lib.slsb_op_topk(P(x), rows, D, k, P(thr), P(cut), P(enc), stream()), "topk")
    order = torch.sort(x, dim=-1, descending=True, stable=True).indices[:, :k]
    ref = torch.zeros_like(x).scatter_(-1, order, x.gather(-1, order))
    assert torch.equal(enc, ref)     # bit-exact selection, canonical lowest-index tie rule
    assert torch.equal(thr, x.gather(-1, order[:, -1:]).squeeze(-1))
    assert int((enc[6] != 0).nonzero().max()) == k - 1


# ------------------------------------------------------------------------------------------ fused select + pool
def _canonical_topk_mask(x, k):
    """Keep mask of the canonical rule: k largest per row, ties at the threshold go to the lowest indices."""
    order = torch.sort(x, dim=-1, descending=True, stable=True).indices[..., :k]
    return torch.zeros_like(x, dtype=torch.bool).scatter_(-1, order, True)


def _canonical_pool(kept_acts, lens, chunk=8):
    """Mean over the valid frames in the kernels' summation order: frames of a chunk in order, chunk partials in order
    (fp32 adds are exact IEEE operations, so a torch loop in the same order gives the same bits)."""
    B, T, D = kept_acts.shape
    total = torch.zeros(B, D, device=kept_acts.device)
    for c0 in range(0, T, chunk):
        part = torch.zeros(B, D, device=kept_acts.device)
        for t in range(c0, min(c0 + chunk, T)):
            valid = (t < lens).to(kept_acts.dtype)[:, None]
            part = part + kept_acts[:, t] * valid
        total = total + part
    return total / lens.to(torch.float32)[:, None]


@pytest.mark.parametrize("B,T,D,k,with_lens", [(3, 201, 4096, 128, False), (4, 50, 4096, 128, True), (2, 9, 1024, 64, True), (2, 33, 2048, 7, False)])
def test_topk_pool_fused_is_bit_exact(lib, cuda, B, T, D, k, with_lens):
    """H-SAE scoring path (model.py:68-79 + :245): per-frame top-k + masked mean pool in one kernel, activations read once."""
    x = F.relu(_rand((B, T, D), 71))
    x[0, 3] = 0.0                               # all-zero frame
    x[1, 2, :] = 0.5                            # all-equal frame: lowest k indices win
    x[0, 5, : D // 2] = x[0, 5, D // 2:]        # real ties at the threshold
    lens = torch.tensor([T - 3 * i for i in range(B)], dtype=torch.int32, device=cuda) if with_lens else None
    nc = lib.slsb_op_pool_chunks(T)
    assert nc == (T + 7) // 8
    thr = torch.empty(B * T, device=cuda); cut = torch.empty(B * T, device=cuda, dtype=torch.int32)
    partial = torch.empty(B, nc, D, device=cuda); pooled = torch.empty(B, D, device=cuda)
    ok(lib, lib.slsb_op_topk_pool(P(x), B, T, D, k, P(lens), P(thr), P(cut), P(partial), P(pooled), stream()), "topk_pool")
    mask = _canonical_topk_mask(x, k)
    full = torch.full((B,), T, device=cuda, dtype=torch.int32)
    ref = _canonical_pool(x * mask, lens if with_lens else full)
    assert torch.equal(pooled, ref)
    # thr / cut agree with the stand-alone selection kernel (what the dense / compact code consumers read)
    thr2 = torch.empty(B * T, device=cuda); cut2 = torch.empty(B * T, device=cuda, dtype=torch.int32)
    ok(lib, lib.slsb_op_topk(P(x), B * T, D, k, P(thr2), P(cut2), None, stream()), "topk")
    assert torch.equal(thr, thr2) and torch.equal(cut, cut2)
    # selection only (no pooling buffers)
    thr3 = torch.empty(B * T, device=cuda); cut3 = torch.empty(B * T, device=cuda, dtype=torch.int32)
    ok(lib, lib.slsb_op_topk_pool(P(x), B, T, D, k, None, P(thr3), P(cut3), None, None, stream()), "topk_pool select only")
    assert torch.equal(thr, thr3) and torch.equal(cut, cut3)


@pytest.mark.parametrize("B,T,D,k,window,ties", [(2, 201, 4096, 128, 8, False), (3, 37, 4096, 128, 8, True), (2, 20, 1024, 64, 4, True), (2, 11, 2048, 9, 2, False)])
def test_window_pool_fused_is_bit_exact(lib, cuda, B, T, D, k, window, ties):
    """H-WIN scoring path (model_window_topk.py:118-203): window sums -> per-window top-k -> votes -> per-frame top-k -> pooled kept
    activations, with the window sums and votes held in registers; against a torch restatement with the canonical tie rule."""
    x = F.relu(_rand((B, T, D), 72))
    if ties:
        x[0, :, 100:] = 0.0                     # fewer than k positive features: zeros tie at the threshold in every window
        x[1, 4] = 0.0
    stride = max(1, window // 2)
    nw = (T - window) // stride + 1
    sums = torch.zeros(B, nw, D, device=cuda)
    for j in range(window):                     # frame order, as the kernel and the reference's sum over the window
        sums = sums + torch.stack([x[:, w * stride + j] for w in range(nw)], 1)
    wmask_ref = _canonical_topk_mask(sums, k)
    votes_ref = torch.zeros(B, T, D, device=cuda)
    for w in range(nw):                         # ascending window order (:175-185)
        sl = slice(w * stride, w * stride + window)
        votes_ref[:, sl] = votes_ref[:, sl] + x[:, sl] * wmask_ref[:, w:w + 1]
    keep = _canonical_topk_mask(votes_ref, k)
    full = torch.full((B,), T, device=cuda, dtype=torch.int32)
    pooled_ref = _canonical_pool(x * keep, full)
    nc = lib.slsb_op_pool_chunks(T)
    wmask = torch.empty(B, nw, 256, device=cuda, dtype=torch.int32)
    for want_votes in (True, False):
        thr = torch.empty(B * T, device=cuda); cut = torch.empty(B * T, device=cuda, dtype=torch.int32)
        votes = torch.empty(B, T, D, device=cuda) if want_votes else None
        partial = torch.empty(B, nc, D, device=cuda); pooled = torch.empty(B, D, device=cuda)
        ok(lib, lib.slsb_op_window_pool(P(x), B, T, D, k, window, P(wmask), P(thr), P(cut), P(votes), P(partial), P(pooled), stream()), "window_pool")
        assert torch.equal(pooled, pooled_ref)
        if want_votes:
            assert torch.equal(votes, votes_ref)
            kth = torch.sort(votes_ref, dim=-1, descending=True, stable=True).values[..., k - 1].reshape(-1)
            assert torch.equal(thr, kth)
