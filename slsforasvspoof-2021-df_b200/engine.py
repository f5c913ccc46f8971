"""Python handle around one ``slsb_engine`` (one per GPU per process).  Thin: every method is a
single C-ABI call with raw device pointers from torch tensors and torch's current stream."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _lib
from ._lib import (ATTN_AUTO, HEAD_NONE, HEAD_RETAIN, HEAD_SAE, HEAD_SLS, HEAD_WINDOW, PREC_BF16, PREC_FP32, Config, check, ptr,
                   stream_ptr)
from .weights import TrunkGeometry

PRECISIONS = {"fp32": PREC_FP32, "bf16": PREC_BF16}


def make_config(geo: TrunkGeometry, sae_dict=0, sae_k=128, sae_window=1, cls_in=0, cls_hidden=256, sls_frames=0,
                sls_hidden=1024, attn_impl=ATTN_AUTO) -> Config:
    cfg = Config()
    cfg.n_conv = len(geo.conv_layers)
    cfg.conv_dim = geo.conv_layers[-1][0]
    for i, (_, k, s) in enumerate(geo.conv_layers):
        cfg.conv_kernel[i], cfg.conv_stride[i] = k, s
    cfg.embed_dim, cfg.ffn_dim, cfg.n_heads, cfg.n_layers = geo.embed_dim, geo.ffn_dim, geo.heads, geo.layers
    cfg.pos_kernel, cfg.pos_groups = geo.conv_pos, geo.conv_pos_groups
    cfg.sae_dict, cfg.sae_k, cfg.sae_window = sae_dict, sae_k, sae_window
    cfg.cls_in, cfg.cls_hidden = cls_in, cls_hidden
    cfg.sls_frames, cfg.sls_hidden = sls_frames, sls_hidden
    cfg.attn_impl = attn_impl
    return cfg


class Engine:
    def __init__(self, cfg: Config, device: torch.device):
        if device.type != "cuda":
            raise _lib.SlsbError("the B200 scoring engine needs a CUDA device (no CPU fallback)")
        self.lib = _lib.load()
        self.device = device
        self.cfg = cfg
        self._h = C.c_void_p()
        idx = device.index if device.index is not None else torch.cuda.current_device()
        check(self.lib.slsb_create(C.byref(cfg), idx, C.byref(self._h)), "slsb_create")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.slsb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- weights ------------------------------------------------------------------------------
    def load_packed(self, packed: Dict[str, torch.Tensor]) -> None:
        st = stream_ptr(self.device)
        for name, t in packed.items():
            want = self.lib.slsb_weight_numel(self._h, name.encode())
            if want < 0:
                continue   # tensor this engine was not configured for (e.g. SLS weights on an SAE engine)
            t = t.to(device=self.device, dtype=torch.float32).contiguous()
            check(self.lib.slsb_set_weight(self._h, name.encode(), ptr(t), t.numel(), st), f"set_weight({name})")
        check(self.lib.slsb_finalize_weights(self._h, st), "finalize_weights")
        torch.cuda.current_stream(self.device).synchronize()   # packed temporaries may be freed by the caller

    # ---- hot path -----------------------------------------------------------------------------
    def frames(self, samples: int) -> int:
        return self.lib.slsb_frames_for_samples(self._h, samples)

    def forward(self, wav: torch.Tensor, head: int, precision: int, lens: Optional[torch.Tensor] = None, retain: bool = False) -> torch.Tensor:
        """``retain=True`` (``head | SLSB_HEAD_RETAIN``) keeps layer results / SAE activations / selection of this forward readable
        through ``get_tensor`` / ``get_sparse`` / ``sae_loss``; the plain scoring forward does not pay for them."""
        B, S = wav.shape
        out = torch.empty(B, 2, device=wav.device, dtype=torch.float32)
        check(self.lib.slsb_forward(self._h, ptr(wav), ptr(lens), B, S, head | (HEAD_RETAIN if retain else 0), precision, ptr(out),
                                    stream_ptr(wav.device)), "slsb_forward")
        return out

    def extract_feat(self, wav: torch.Tensor, precision: int, lens: Optional[torch.Tensor] = None) -> torch.Tensor:
        B, S = wav.shape
        x = torch.empty(B, self.frames(S), self.cfg.embed_dim, device=wav.device, dtype=torch.float32)
        check(self.lib.slsb_extract_feat(self._h, ptr(wav), ptr(lens), B, S, precision, ptr(x), stream_ptr(wav.device)), "slsb_extract_feat")
        return x

    def get_tensor(self, name: str, shape) -> torch.Tensor:
        t = torch.empty(*shape, device=self.device, dtype=torch.float32)
        check(self.lib.slsb_get_tensor(self._h, name.encode(), ptr(t), t.numel(), stream_ptr(self.device)), f"get_tensor({name})")
        return t

    def get_sparse(self, rows: int):
        """(indices int32 [rows, k], values fp32 [rows, k], counts int32 [rows]) of the last forward's SAE code."""
        k = self.cfg.sae_k
        idx = torch.empty(rows, k, device=self.device, dtype=torch.int32)
        val = torch.empty(rows, k, device=self.device, dtype=torch.float32)
        cnt = torch.empty(rows, device=self.device, dtype=torch.int32)
        check(self.lib.slsb_get_sparse(self._h, ptr(idx), ptr(val), ptr(cnt), stream_ptr(self.device)), "slsb_get_sparse")
        return idx, val, cnt

    def sae_encode(self, x: torch.Tensor, T: int, window: int, precision: int) -> torch.Tensor:
        rows = x.shape[0]
        out = torch.empty(rows, self.cfg.sae_dict, device=x.device, dtype=torch.float32)
        check(self.lib.slsb_sae_encode(self._h, ptr(x), rows, T, window, precision, ptr(out), stream_ptr(x.device)), "slsb_sae_encode")
        return out

    def sae_decode(self, enc: torch.Tensor, precision: int) -> torch.Tensor:
        rows = enc.shape[0]
        out = torch.empty(rows, self.cfg.embed_dim, device=enc.device, dtype=torch.float32)
        check(self.lib.slsb_sae_decode(self._h, ptr(enc), rows, precision, ptr(out), stream_ptr(enc.device)), "slsb_sae_decode")
        return out

    def sae_loss(self, precision: int) -> torch.Tensor:
        out = torch.empty(1, device=self.device, dtype=torch.float32)
        check(self.lib.slsb_sae_loss(self._h, precision, ptr(out), stream_ptr(self.device)), "slsb_sae_loss")
        return out[0]

    def score_host(self, wav_host: torch.Tensor, head: int, precision: int, lens_host: Optional[torch.Tensor] = None) -> torch.Tensor:
        """main.py:178-184 in one call: H2D, forward, exp(logp[:,1]), D2H (returns a CPU tensor)."""
        B, S = wav_host.shape
        scores = torch.empty(B, dtype=torch.float32, pin_memory=True)
        check(self.lib.slsb_score_host(self._h, ptr(wav_host), ptr(lens_host), B, S, head, precision, ptr(scores),
                                       stream_ptr(self.device)), "slsb_score_host")
        return scores

    def score_submit(self, wav_host: torch.Tensor, head: int, precision: int, lens_host: Optional[torch.Tensor] = None,
                     out: Optional[torch.Tensor] = None):
        """Pipelined variant: returns ``(ticket, scores)``; ``scores`` (pinned CPU tensor) is valid after
        ``score_wait(ticket)``.  The upload of this batch overlaps the forward of the previous submission."""
        B, S = wav_host.shape
        scores = out if out is not None else torch.empty(B, dtype=torch.float32, pin_memory=True)
        t = self.lib.slsb_score_submit(self._h, ptr(wav_host), ptr(lens_host), B, S, head, precision, ptr(scores), stream_ptr(self.device))
        if t < 0:
            check(-1, "slsb_score_submit")
        self._inflight = getattr(self, "_inflight", {})
        self._inflight[t] = (wav_host, lens_host, scores)          # keep the host buffers alive until waited
        return t, scores

    def score_wait(self, ticket: int = -1) -> None:
        check(self.lib.slsb_score_wait(self._h, ticket), "slsb_score_wait")
        infl = getattr(self, "_inflight", {})
        for k in [k for k in infl if ticket < 0 or k <= ticket]:
            del infl[k]

    def ingest_pcm16(self, pcm: torch.Tensor, offsets: torch.Tensor, lens: torch.Tensor, samples: int = 64600) -> torch.Tensor:
        """Device int16 PCM (clips back to back) -> fp32 [B, samples] with the reference's ``pad`` (truncate / tile-repeat)."""
        B = lens.numel()
        wav = torch.empty(B, samples, device=pcm.device, dtype=torch.float32)
        check(self.lib.slsb_ingest_pcm16(ptr(pcm), ptr(offsets), ptr(lens), B, samples, ptr(wav), stream_ptr(pcm.device)), "slsb_ingest_pcm16")
        return wav

    def score_pcm16_host(self, clips, head: int, precision: int, samples: int = 64600) -> torch.Tensor:
        """Scores a list of 1-D int16 host clips of any length (each is padded on the device like ``pad``)."""
        lens = torch.tensor([int(c.numel()) for c in clips], dtype=torch.int32)
        offsets = torch.zeros(len(clips), dtype=torch.int64)
        offsets[1:] = torch.cumsum(lens[:-1].to(torch.int64), 0)
        pcm = torch.cat([c.reshape(-1).to(torch.int16) for c in clips])
        return self.score_pcm16_arrays(pcm, offsets, lens, head, precision, samples)

    def score_pcm16_arrays(self, pcm: torch.Tensor, offsets: torch.Tensor, lens: torch.Tensor, head: int, precision: int,
                           samples: int = 64600) -> torch.Tensor:
        """``slsb_score_pcm16_host`` on host arrays: pcm int16 [total] (clips back to back), offsets int64 [B], lens int32 [B]."""
        if pcm.dtype != torch.int16 or offsets.dtype != torch.int64 or lens.dtype != torch.int32:
            raise TypeError("score_pcm16_arrays: pcm int16, offsets int64, lens int32")
        B = lens.numel()
        pcm = pcm.contiguous()
        pcm = pcm if pcm.is_pinned() else pcm.pin_memory()
        scores = torch.empty(B, dtype=torch.float32, pin_memory=True)
        check(self.lib.slsb_score_pcm16_host(self._h, ptr(pcm), pcm.numel(), ptr(offsets.contiguous()), ptr(lens.contiguous()), B, samples, head,
                                             precision, ptr(scores), stream_ptr(self.device)), "slsb_score_pcm16_host")
        return scores

    def score_flac_arrays(self, data: torch.Tensor, frames: torch.Tensor, total_samples: int, offsets: torch.Tensor, lens: torch.Tensor,
                          head: int, precision: int, samples: int = 64600):
        """``slsb_score_flac_host``: data uint8 [nbytes] (the clips' FLAC frames), frames uint8 [n_frames * 32] (packed
        ``slsb_flac_frame`` records), offsets int64 [B] / lens int32 [B] in samples of the decoded int16 buffer.  Returns
        ``(scores or None, status int32 [n_frames])`` - None when the device decoder refused a frame (host fallback)."""
        B = lens.numel()
        n_frames = frames.numel() // 32
        data = data if data.is_pinned() else data.pin_memory()
        scores = torch.empty(B, dtype=torch.float32, pin_memory=True)
        status = torch.empty(n_frames, dtype=torch.int32, pin_memory=True)
        rc = self.lib.slsb_score_flac_host(self._h, ptr(data), data.numel(), ptr(frames.contiguous()), n_frames, int(total_samples),
                                           ptr(offsets.contiguous()), ptr(lens.contiguous()), B, samples, head, precision, ptr(scores), ptr(status),
                                           stream_ptr(self.device))
        if rc == -2:
            return None, status
        check(rc, "slsb_score_flac_host")
        return scores, status

    def synth_clips(self, first_utt: int, count: int, samples: int = 64600) -> torch.Tensor:
        wav = torch.empty(count, samples, device=self.device, dtype=torch.float32)
        check(self.lib.slsb_synth_clips(ptr(wav), first_utt, count, samples, stream_ptr(self.device)), "slsb_synth_clips")
        return wav

    def profile(self, on: bool) -> None:
        check(self.lib.slsb_profile_enable(self._h, 1 if on else 0), "slsb_profile_enable")

    def profile_read(self, kind: int):
        ms, fl, n = C.c_double(), C.c_double(), C.c_int64()
        check(self.lib.slsb_profile_read(self._h, kind, C.byref(ms), C.byref(fl), C.byref(n)), "slsb_profile_read")
        return ms.value, fl.value, n.value

    @property
    def launch_count(self) -> int:
        return int(self.lib.slsb_launch_count(self._h))


__all__ = ["Engine", "make_config", "PRECISIONS", "HEAD_NONE", "HEAD_SAE", "HEAD_WINDOW", "HEAD_SLS", "HEAD_RETAIN", "PREC_FP32", "PREC_BF16"]
