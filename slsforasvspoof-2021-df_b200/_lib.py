"""ctypes binding of libslsb200.so (the C ABI declared in include/slsb200.h).

The product path has NO fallback: if the shared library is missing or the box has no sm_100 GPU,
every entry point raises.  Build it with ``python -c "import __graft_entry__ as g; g.build()"``
(or ``make -C slsforasvspoof-2021-df_b200/csrc``).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libslsb200.so")

HEAD_NONE, HEAD_SAE, HEAD_WINDOW, HEAD_SLS = 0, 1, 2, 3
HEAD_RETAIN = 0x100     # OR-ed into the head: keep layer results / SAE intermediates readable after the forward
PREC_FP32, PREC_BF16 = 0, 1
ATTN_AUTO, ATTN_SIMT, ATTN_TC, ATTN_TC_V1 = 0, 1, 2, 3


class SlsbError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [
        ("n_conv", C.c_int32), ("conv_dim", C.c_int32),
        ("conv_kernel", C.c_int32 * 8), ("conv_stride", C.c_int32 * 8),
        ("embed_dim", C.c_int32), ("ffn_dim", C.c_int32), ("n_heads", C.c_int32), ("n_layers", C.c_int32),
        ("pos_kernel", C.c_int32), ("pos_groups", C.c_int32),
        ("sae_dict", C.c_int32), ("sae_k", C.c_int32), ("sae_window", C.c_int32),
        ("cls_in", C.c_int32), ("cls_hidden", C.c_int32),
        ("sls_frames", C.c_int32), ("sls_hidden", C.c_int32),
        ("attn_impl", C.c_int32), ("reserved", C.c_int32 * 8),
    ]


_P = C.c_void_p
_SIGNATURES = {
    "slsb_abi_version": (C.c_int, []),
    "slsb_last_error": (C.c_char_p, []),
    "slsb_create": (C.c_int, [C.POINTER(Config), C.c_int, C.POINTER(_P)]),
    "slsb_destroy": (C.c_int, [_P]),
    "slsb_set_weight": (C.c_int, [_P, C.c_char_p, _P, C.c_int64, _P]),
    "slsb_finalize_weights": (C.c_int, [_P, _P]),
    "slsb_weight_numel": (C.c_int64, [_P, C.c_char_p]),
    "slsb_frames_for_samples": (C.c_int, [_P, C.c_int]),
    "slsb_forward": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "slsb_extract_feat": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "slsb_get_tensor": (C.c_int, [_P, C.c_char_p, _P, C.c_int64, _P]),
    "slsb_get_sparse": (C.c_int, [_P, _P, _P, _P, _P]),
    "slsb_sae_encode": (C.c_int, [_P, _P, C.c_int64, C.c_int, C.c_int, C.c_int, _P, _P]),
    "slsb_sae_decode": (C.c_int, [_P, _P, C.c_int64, C.c_int, _P, _P]),
    "slsb_sae_loss": (C.c_int, [_P, C.c_int, _P, _P]),
    "slsb_score_host": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "slsb_score_submit": (C.c_int64, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "slsb_score_wait": (C.c_int, [_P, C.c_int64]),
    "slsb_ingest_pcm16": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, _P, _P]),
    "slsb_score_pcm16_host": (C.c_int, [_P, _P, C.c_int64, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "slsb_synth_clips": (C.c_int, [_P, C.c_int64, C.c_int, C.c_int, _P]),
    "slsb_launch_count": (C.c_int64, [_P]),
    "slsb_profile_enable": (C.c_int, [_P, C.c_int]),
    "slsb_profile_read": (C.c_int, [_P, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "slsb_op_gemm": (C.c_int, [C.c_int, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "slsb_op_gemm_splitk": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "slsb_op_conv": (C.c_int, [C.c_int, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "slsb_op_conv_ln_gelu": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "slsb_op_conv0_tc": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int, _P]),
    "slsb_op_conv0_fused": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int, _P]),
    "slsb_op_posconv": (C.c_int, [C.c_int, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "slsb_op_conv0": (C.c_int, [C.c_int, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "slsb_op_layernorm": (C.c_int, [_P, C.c_int, _P, C.c_int, _P, _P, C.c_int64, C.c_int, C.c_int, C.c_int, _P]),
    "slsb_flac_decode": (C.c_int64, [_P, C.c_int64, C.c_int64, C.c_int, _P, C.c_int64, _P]),
    "slsb_flac_decode_mono16": (C.c_int64, [_P, C.c_int64, C.c_int64, C.c_int, _P, C.c_int64, _P]),
    "slsb_score_flac_host": (C.c_int, [_P, _P, C.c_int64, _P, C.c_int, C.c_int64, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "slsb_flac_scan": (C.c_int64, [_P, C.c_int64, C.c_int64, _P, _P, _P, _P, C.c_int64]),
    "slsb_flac_decode_frames": (C.c_int, [_P, _P, C.c_int, _P, _P, _P]),
    "slsb_flac_decode_frames_host": (C.c_int, [_P, _P, C.c_int, _P, _P]),
    "slsb_debug_pair_schedule": (C.c_int, [C.c_int, C.c_int, C.c_int, _P, C.c_int, _P]),
    "slsb_op_layernorm_taps": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_int64, C.c_int, _P]),
    "slsb_op_attention": (C.c_int, [C.c_int, C.c_int, _P, _P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "slsb_op_attention_trace": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "slsb_op_topk": (C.c_int, [_P, C.c_int64, C.c_int, C.c_int, _P, _P, _P, _P]),
    "slsb_op_topk_pool": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P, _P]),
    "slsb_op_window_pool": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P, _P, _P]),
    "slsb_op_pool_chunks": (C.c_int, [C.c_int]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """dlopen the library (works without a GPU; compute entry points then fail with a clear error)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SlsbError(
                f"{LIB_PATH} is missing: the CUDA extension is not built and there is no CPU fallback. "
                "Run `python -c \"import __graft_entry__ as g; g.build()\"`.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().slsb_last_error()
        raise SlsbError(f"{what or 'slsb call'} failed ({rc}): {msg.decode() if msg else '?'}")


def ptr(t) -> Optional[int]:
    """Raw data pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
