"""Shared helpers for the GPU parity tests (call everything through the C ABI)."""
import ctypes as C

import torch


def P(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ok(lib, rc, what=""):
    if rc != 0:
        raise AssertionError(f"{what} failed rc={rc}: {lib.slsb_last_error().decode()}")
    torch.cuda.synchronize()


def report(name, got, ref, atol, rtol=0.0):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs()
    tol = atol + rtol * ref.abs()
    bad = err > tol
    mx = float(err.max())
    msg = f"{name}: max|err|={mx:.3e} ref_scale={float(ref.abs().mean()):.3e} mismatches={int(bad.sum())}/{bad.numel()}"
    if bad.any():
        idx = bad.nonzero()[:5].tolist()
        msg += f" first_bad={idx} got={[float(got[tuple(i)]) for i in idx]} ref={[float(ref[tuple(i)]) for i in idx]}"
    print(msg)
    assert not bad.any(), msg
    assert torch.isfinite(got).all(), name + ": non-finite output"
