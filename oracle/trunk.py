"""ORACLE (test infrastructure only) -- CPU fp32 restatement of the XLS-R 300M trunk.

This file is the *checker*, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  The product path (``slsforasvspoof-2021-df_b200``) never does.

What it restates
----------------
The arithmetic of the scoring path lives in fairseq @ a54021305d6b3c4c5959ac9395135f63202db8f1,
which is NOT shipped with the reference checkout (``.MISSING_LARGE_BLOBS:2``).  The only
in-tree statement of the dataflow is the vendored, un-imported copy
``/root/reference/wav2vec/wav2vec2.py``; every function below cites the lines it follows.
Missing fairseq modules are restated from their published semantics:

* ``Fp32LayerNorm``  = ``F.layer_norm(x.float(), ...)``           (wav2vec2.py:803-813)
* ``TransposeLast``  = ``x.transpose(-2, -1)``
* ``SamePad(128)``   = drop the last frame when the kernel is even  (wav2vec2.py:875)
* ``MultiheadAttention`` = separate biased q/k/v/out projections, q scaled by d**-0.5
  after projection, fp32 softmax, key-padding -> -inf            (wav2vec2.py:1009-1014, :1046-1052)
* ``get_activation_fn('gelu')`` = exact erf GELU in fp32

Parity pin
----------
The reference has no golden vectors for this path, so the trunk
is pinned against an *independent* implementation that is installed here:
``transformers.Wav2Vec2Model`` (``do_stable_layer_norm=True``), see
``oracle/make_golden.py::crosscheck_hf`` and ``tests/test_oracle.py``.  The heads are
pinned against the reference's own ``model.py`` / ``model_window_topk.py`` /
``model_backup.py`` executed verbatim through ``oracle/fairseq_stub.py``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class TrunkConfig:
    """XLS-R 300M hyper-parameters (they live in the absent checkpoint's cfg; corroborated by
    torchaudio.models.wav2vec2_xlsr_300m and the 315 M parameter count)."""

    conv_layers: List[Tuple[int, int, int]] = field(
        default_factory=lambda: [(512, 10, 5)] + [(512, 3, 2)] * 4 + [(512, 2, 2)] * 2
    )  # wav2vec2.py:97-99
    embed_dim: int = 1024
    ffn_dim: int = 4096
    heads: int = 16
    layers: int = 24
    conv_pos: int = 128
    conv_pos_groups: int = 16
    required_seq_len_multiple: int = 1  # wav2vec2.py:241-246 (vendored default)
    final_dim: int = 768      # pre-training-only heads, kept for state_dict key parity
    latent_vars: int = 320
    latent_groups: int = 2


def conv_out_length(n: int, cfg: TrunkConfig) -> int:
    """wav2vec2.py:523-538 (_get_feat_extract_output_lengths)."""
    for _, k, s in cfg.conv_layers:
        n = (n - k) // s + 1
    return n


class _Transpose(nn.Module):
    def forward(self, x):
        return x.transpose(-2, -1)


class _Fp32LayerNorm(nn.LayerNorm):
    def forward(self, x):
        out = F.layer_norm(
            x.float(), self.normalized_shape,
            self.weight.float() if self.weight is not None else None,
            self.bias.float() if self.bias is not None else None, self.eps)
        return out.type_as(x)


class ConvFeatureExtractor(nn.Module):
    """wav2vec2.py:773-851, mode='layer_norm', conv_bias=True."""

    def __init__(self, cfg: TrunkConfig):
        super().__init__()
        self.conv_layers = nn.ModuleList()
        in_d = 1
        for dim, k, s in cfg.conv_layers:
            conv = nn.Conv1d(in_d, dim, k, stride=s, bias=True)
            nn.init.kaiming_normal_(conv.weight)  # wav2vec2.py:796
            self.conv_layers.append(nn.Sequential(
                conv, nn.Dropout(0.0),
                nn.Sequential(_Transpose(), _Fp32LayerNorm(dim, elementwise_affine=True), _Transpose()),
                nn.GELU()))  # wav2vec2.py:803-813
            in_d = dim

    def forward(self, x):
        x = x.unsqueeze(1)  # wav2vec2.py:846
        for conv in self.conv_layers:
            x = conv(x)
        return x


class _SelfAttention(nn.Module):
    """fairseq MultiheadAttention (self-attention, eval) restated; ctor wav2vec2.py:1009-1014."""

    def __init__(self, dim: int, heads: int):
        super().__init__()
        self.heads, self.head_dim = heads, dim // heads
        self.scaling = self.head_dim ** -0.5
        self.k_proj = nn.Linear(dim, dim)
        self.v_proj = nn.Linear(dim, dim)
        self.q_proj = nn.Linear(dim, dim)
        self.out_proj = nn.Linear(dim, dim)

    def forward(self, x, key_padding_mask=None):
        T, B, C = x.shape
        q = self.q_proj(x) * self.scaling
        k = self.k_proj(x)
        v = self.v_proj(x)
        q = q.contiguous().view(T, B * self.heads, self.head_dim).transpose(0, 1)
        k = k.contiguous().view(T, B * self.heads, self.head_dim).transpose(0, 1)
        v = v.contiguous().view(T, B * self.heads, self.head_dim).transpose(0, 1)
        w = torch.bmm(q, k.transpose(1, 2))
        if key_padding_mask is not None:
            w = w.view(B, self.heads, T, T)
            w = w.masked_fill(key_padding_mask.unsqueeze(1).unsqueeze(2).to(torch.bool), float("-inf"))
            w = w.view(B * self.heads, T, T)
        w = F.softmax(w.float(), dim=-1).type_as(w)
        a = torch.bmm(w, v)
        a = a.transpose(0, 1).contiguous().view(T, B, C)
        return self.out_proj(a), None


class EncoderLayer(nn.Module):
    """wav2vec2.py:983-1083, layer_norm_first=True branch (:1044-1062)."""

    def __init__(self, cfg: TrunkConfig):
        super().__init__()
        self.self_attn = _SelfAttention(cfg.embed_dim, cfg.heads)
        self.self_attn_layer_norm = nn.LayerNorm(cfg.embed_dim)
        self.fc1 = nn.Linear(cfg.embed_dim, cfg.ffn_dim)
        self.fc2 = nn.Linear(cfg.ffn_dim, cfg.embed_dim)
        self.final_layer_norm = nn.LayerNorm(cfg.embed_dim)

    def forward(self, x, self_attn_padding_mask=None):
        residual = x
        x = self.self_attn_layer_norm(x)
        x, attn = self.self_attn(x, key_padding_mask=self_attn_padding_mask)
        x = residual + x
        residual = x
        x = self.final_layer_norm(x)
        x = F.gelu(self.fc1(x).float()).type_as(x)
        x = self.fc2(x)
        x = residual + x
        return x, attn


def _pad_to_multiple(x, multiple, dim=-1, value=0):
    """wav2vec/utils.py:18-29."""
    if x is None:
        return None, 0
    tsz = x.size(dim)
    m = tsz / multiple
    remainder = math.ceil(m) * multiple - tsz
    if m.is_integer():
        return x, 0
    pad_offset = (0,) * (-1 - dim) * 2
    return F.pad(x, (*pad_offset, 0, remainder), value=value), remainder


class Encoder(nn.Module):
    """wav2vec2.py:854-972."""

    def __init__(self, cfg: TrunkConfig):
        super().__init__()
        self.cfg = cfg
        D = cfg.embed_dim
        conv = nn.Conv1d(D, D, kernel_size=cfg.conv_pos, padding=cfg.conv_pos // 2, groups=cfg.conv_pos_groups)
        std = math.sqrt(4.0 / (cfg.conv_pos * D))  # wav2vec2.py:870
        nn.init.normal_(conv.weight, mean=0, std=std)
        nn.init.constant_(conv.bias, 0)
        # wav2vec2.py:874: weight_norm(name="weight", dim=2) -> params weight_g [1,1,K], weight_v
        conv = torch.nn.utils.weight_norm(conv, name="weight", dim=2)
        self.pos_conv = nn.Sequential(conv, nn.Identity(), nn.GELU())  # [1] stands in for SamePad
        self.layers = nn.ModuleList([EncoderLayer(cfg) for _ in range(cfg.layers)])
        self.layer_norm = nn.LayerNorm(D)
        for m in self.modules():  # init_bert_params, wav2vec2.py:899
            if isinstance(m, nn.Linear):
                m.weight.data.normal_(mean=0.0, std=0.02)
                if m.bias is not None:
                    m.bias.data.zero_()

    def _pos(self, x):
        y = self.pos_conv[0](x.transpose(1, 2))
        if self.cfg.conv_pos % 2 == 0:
            y = y[:, :, :-1]  # SamePad, wav2vec2.py:875
        return self.pos_conv[2](y).transpose(1, 2)

    def forward(self, x, padding_mask=None):
        # extract_features, wav2vec2.py:910-972
        if padding_mask is not None:
            x = x.masked_fill(padding_mask.unsqueeze(-1), 0.0)  # index_put(x, padding_mask, 0)
        x = x + self._pos(x)
        x, pad_length = _pad_to_multiple(x, self.cfg.required_seq_len_multiple, dim=-2, value=0)
        if pad_length > 0 and padding_mask is None:
            padding_mask = x.new_zeros((x.size(0), x.size(1)), dtype=torch.bool)
            padding_mask[:, -pad_length:] = True
        else:
            padding_mask, _ = _pad_to_multiple(padding_mask, self.cfg.required_seq_len_multiple, dim=-1, value=True)
        x = x.transpose(0, 1)
        layer_results = []
        for layer in self.layers:
            x, z = layer(x, self_attn_padding_mask=padding_mask)
            layer_results.append((x[:-pad_length] if pad_length > 0 else x, z))  # :958, raw residual stream
        x = x.transpose(0, 1)
        if pad_length > 0:
            x = x[:, :-pad_length]
        x = self.layer_norm(x)  # wav2vec2.py:905-906
        return x, layer_results


class Wav2Vec2Trunk(nn.Module):
    """Eval-only Wav2Vec2Model (wav2vec2.py:256-362 ctor, :540-647 forward) with fairseq parameter names."""

    def __init__(self, cfg: Optional[TrunkConfig] = None):
        super().__init__()
        self.cfg = cfg = cfg or TrunkConfig()
        embed = cfg.conv_layers[-1][0]
        self.feature_extractor = ConvFeatureExtractor(cfg)
        self.post_extract_proj = nn.Linear(embed, cfg.embed_dim)
        self.mask_emb = nn.Parameter(torch.zeros(cfg.embed_dim).uniform_())
        self.encoder = Encoder(cfg)
        self.layer_norm = nn.LayerNorm(embed)
        # pre-training-only heads (state_dict key parity with xlsr2_300m.pt; unused in eval)
        self.final_proj = nn.Linear(cfg.embed_dim, cfg.final_dim)
        self.project_q = nn.Linear(cfg.final_dim, cfg.final_dim)

    def feat_lengths(self, input_lengths: torch.Tensor) -> torch.Tensor:
        out = input_lengths.clone()
        for _, k, s in self.cfg.conv_layers:
            out = torch.floor((out - k) / s + 1)
        return out.to(torch.long)

    def forward(self, source, padding_mask=None, mask=False, features_only=True, layer=None):
        assert not mask and features_only and layer is None, "oracle restates the eval scoring call only"
        assert not self.training, "oracle must run in eval mode (SURVEY.md section 0 trap)"
        features = self.feature_extractor(source)          # :554
        features = features.transpose(1, 2)
        features = self.layer_norm(features)               # :563-564
        unmasked = features
        if padding_mask is not None and padding_mask.any():  # :567-586
            input_lengths = (1 - padding_mask.long()).sum(-1)
            output_lengths = self.feat_lengths(input_lengths)
            pm = torch.zeros(features.shape[:2], dtype=features.dtype, device=features.device)
            pm[(torch.arange(pm.shape[0], device=pm.device), output_lengths - 1)] = 1
            padding_mask = (1 - pm.flip([-1]).cumsum(-1).flip([-1])).bool()
        else:
            padding_mask = None
        features = self.post_extract_proj(features)        # :595-596
        x, layer_results = self.encoder(features, padding_mask=padding_mask)  # :635
        return {"x": x, "padding_mask": padding_mask, "features": unmasked, "layer_results": layer_results}


# --------------------------------------------------------------------------------------------
# Deterministic, machine-independent parameter / clip synthesis (integer hash -> float, no libm)
# --------------------------------------------------------------------------------------------
import zlib

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(z: np.ndarray) -> np.ndarray:
    z = (z + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))


def hash_normal(key: int, n: int, offset: int = 0) -> np.ndarray:
    """n pseudo-normal float32 (Irwin-Hall of 4x16-bit uniforms, unit variance), bit-exact everywhere.
    The same function is implemented on the device in csrc/synth.cu (integer arithmetic, one rounding)."""
    with np.errstate(over="ignore"):
        idx = np.arange(offset, offset + n, dtype=np.uint64)
        h = _splitmix64(idx ^ _splitmix64(np.uint64(key) * np.uint64(0xD6E8FEB86659FD93) & _M64))
    s = ((h & np.uint64(0xFFFF)) + ((h >> np.uint64(16)) & np.uint64(0xFFFF))
         + ((h >> np.uint64(32)) & np.uint64(0xFFFF)) + (h >> np.uint64(48)))  # in [0, 4*65535]
    f = s.astype(np.float32) - np.float32(131070.0)  # exact
    # var of one 16-bit uniform = (65536^2-1)/12 ; 4 of them
    return f * np.float32(1.0 / math.sqrt(4 * (65536.0 ** 2 - 1) / 12.0))


def synth_clips(first_utt: int, count: int, samples: int = 64600) -> torch.Tensor:
    """Synthetic clips keyed by utterance index (any rank can regenerate any clip; SURVEY.md section 8d cfg 3)."""
    out = np.empty((count, samples), dtype=np.float32)
    for i in range(count):
        out[i] = hash_normal(0x5EED0000 + first_utt + i, samples)
    return torch.from_numpy(out)


@torch.no_grad()
def seeded_init_(module: nn.Module, seed: int = 1234) -> nn.Module:
    """Overwrite every parameter/buffer with a hash-seeded value of the scale the fairseq
    initialisers would give (wav2vec2.py:796, :870-872, :899), *plus* non-trivial biases and
    LayerNorm affines so a dropped bias/affine cannot hide behind a zero init."""
    names = dict(module.named_parameters())
    names.update({k: v for k, v in module.named_buffers() if v.dtype.is_floating_point})
    for name, p in sorted(names.items()):
        n = p.numel()
        # keyed by the parameter NAME, so the trunk gets the same weights under every head
        z = torch.from_numpy(hash_normal(seed * 100003 + zlib.crc32(name.encode()), n)).view(p.shape)
        leaf = name.split(".")[-1]
        if leaf == "weight_g":
            continue  # set below from weight_v
        if p.dim() == 1:
            if leaf == "weight":  # LayerNorm / BatchNorm gain
                p.copy_(1.0 + 0.1 * z)
            elif leaf == "running_var":
                p.copy_(1.0 + 0.1 * z.abs())
            else:  # biases, b_dec, mask_emb, running_mean
                p.copy_(0.05 * z)
        elif leaf == "weight_v":
            cfgk, D = p.shape[2], p.shape[0]
            p.copy_(z * math.sqrt(4.0 / (cfgk * D)))
        elif p.dim() == 3:  # conv weights: kaiming_normal_, fan_in = in*k
            p.copy_(z * math.sqrt(2.0 / (p.shape[1] * p.shape[2])))
        elif p.dim() == 2:
            if "sae." in name or name.startswith("sae"):
                p.copy_(z / math.sqrt(p.shape[1]) if leaf == "weight" else z)
            elif p.shape[1] > 8192:   # SLS fc1 (22847 -> 1024): keep pre-activations O(1)
                p.copy_(0.1 * z / math.sqrt(p.shape[1]))
            else:
                p.copy_(0.02 * z)   # init_bert_params
        else:
            p.copy_(0.02 * z)
    for name, p in names.items():
        if name.endswith("weight_g"):
            v = names[name[:-1] + "v"]
            p.copy_(v.norm(dim=(0, 1), keepdim=True) * (1.0 + 0.1 * torch.from_numpy(
                hash_normal(seed * 100003 + 7777, p.numel())).view(p.shape)))
    return module
