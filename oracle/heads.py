"""ORACLE (test infrastructure only) -- CPU fp32 restatement of the three classifier heads that sit
on the XLS-R trunk, so the checker can run on the GPU box where ``/root/reference`` does not exist.

* H-SAE  : ``/root/reference/model.py:42-104`` (AutoEncoderTopK), ``:144-260`` (Model.forward)
* H-WIN  : ``/root/reference/model_window_topk.py:68-203`` (encode/_window_topk), ``:324-393``
* H-SLS  : ``/root/reference/model_backup.py:186-202`` (getAttenF, the only surviving piece) + the upstream
           SLS classifier (QiShanZhang/SLSforASVspoof-2021-DF ``model.py``; NOT in this tree --
           "parity unpinned" for everything after getAttenF).

Pinned (tests/test_oracle.py, oracle/make_golden.py) against the reference's own files run verbatim
through ``oracle/fairseq_stub.py`` while ``/root/reference`` is present; the outputs are committed
under ``tests/golden/``.

Canonical top-k tie rule (declared deviation): the reference uses ``torch.topk(sorted=False)``
whose choice among equal values is implementation-defined.  Oracle and CUDA
kernels both keep the LOWEST feature index among equal values.  For H-SAE this never changes the
result unless two positive activations are bit-equal; for H-WIN it decides frames with fewer than
k non-zero votes.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .trunk import TrunkConfig, Wav2Vec2Trunk


def canonical_topk_mask(x: torch.Tensor, k: int) -> torch.Tensor:
    """0/1 mask of the k largest entries of the last dim; ties -> lowest index wins."""
    order = torch.sort(x, dim=-1, descending=True, stable=True).indices[..., :k]
    return torch.zeros_like(x).scatter_(-1, order, 1.0)


class SAETopK(nn.Module):
    """model.py:42-104 / model_window_topk.py:40-116."""

    def __init__(self, activation_dim: int, dict_size: int, k: int, window_size: int = 1):
        super().__init__()
        self.window_size = window_size
        self.register_buffer("k", torch.tensor(k, dtype=torch.int))
        self.decoder = nn.Linear(dict_size, activation_dim, bias=False)
        self.decoder.weight.data = self.decoder.weight.data / torch.norm(self.decoder.weight.data, dim=0, keepdim=True)
        self.encoder = nn.Linear(activation_dim, dict_size)
        self.encoder.weight.data = self.decoder.weight.T.clone()
        self.encoder.bias.data.zero_()
        self.b_dec = nn.Parameter(torch.zeros(activation_dim))

    def pre_topk(self, x):
        return F.relu(self.encoder(x - self.b_dec))  # model.py:70

    def encode(self, x: torch.Tensor, temporal_dim: Optional[int] = None):
        k = int(self.k)
        shape = x.shape
        if x.dim() == 3:
            B, T, _ = x.shape
        elif temporal_dim is not None:
            B, T = x.shape[0] // temporal_dim, temporal_dim
        else:
            B = T = None
        acts = self.pre_topk(x.reshape(-1, shape[-1]))
        if self.window_size == 1 or T is None:
            out = acts * canonical_topk_mask(acts, k)  # model.py:73-77 (values scattered into zeros)
        else:
            out = window_topk(acts.reshape(B, T, -1), k, self.window_size).reshape(B * T, -1)
        return out.reshape(*shape[:-1], -1)

    def decode(self, x):
        return self.decoder(x) + self.b_dec

    def forward(self, x):
        e = self.encode(x)
        return self.decode(e), e


def window_topk(x: torch.Tensor, k: int, window_size: int) -> torch.Tensor:
    """model_window_topk.py:118-203 without the Python loops; stride = window//2 (:133)."""
    B, T, D = x.shape
    stride = max(1, window_size // 2)
    if stride >= T:
        raise ValueError("sequence shorter than the window stride is not exercised by any reference caller")
    nw = (T - window_size) // stride + 1                      # :141
    xw = x.unfold(1, window_size, stride)                     # [B, nw, D, w]   (:153)
    sums = xw.sum(dim=-1)                                     # :158  (sum order: j = 0..w-1)
    mask_w = canonical_topk_mask(sums, k)                     # :161-165
    votes = torch.zeros_like(x)
    for i in range(nw):                                       # :175-185  (ascending window order)
        s = i * stride
        votes[:, s:s + window_size] += x[:, s:s + window_size] * mask_w[:, i:i + 1]
    return x * canonical_topk_mask(votes, k)                  # :188-197


class OracleModel(nn.Module):
    """Same ctor/forward surface and state_dict names as the reference ``Model`` (model.py:144-260)."""

    def __init__(self, args=None, device="cpu", cp_path="xlsr2_300m.pt", use_sae=True, use_sparse_features=True,
                 sae_dict_size=4096, sae_k=128, sae_window_size=1, sae_weight=0.1, head="sae",
                 trunk_cfg: Optional[TrunkConfig] = None):
        super().__init__()
        self.head = head
        self.use_sae, self.use_sparse_features, self.sae_weight = use_sae, use_sparse_features, sae_weight
        self.ssl_model = nn.Module()
        self.ssl_model.model = Wav2Vec2Trunk(trunk_cfg)
        self.ssl_model.out_dim = D = self.ssl_model.model.cfg.embed_dim
        if head in ("sae", "window"):
            self.sae = SAETopK(D, sae_dict_size, sae_k, sae_window_size if head == "window" else 1)
            in_dim = sae_dict_size if use_sparse_features else D
            self.pool = nn.AdaptiveAvgPool1d(1)
            self.classifier = nn.Sequential(nn.LayerNorm(in_dim), nn.Linear(in_dim, 256), nn.ReLU(),
                                            nn.Dropout(0.3), nn.Linear(256, 2))   # model.py:183-189
        elif head == "sls":
            self.first_bn = nn.BatchNorm2d(num_features=1)
            self.selu = nn.SELU()
            self.fc0 = nn.Linear(D, 1)
            self.sig = nn.Sigmoid()
            self.sls_in = None  # set on first forward from T: ((T//3) * (D//3)) = 22847 for T=201, D=1024
            T = 201
            self.fc1 = nn.Linear((T // 3) * (D // 3), 1024)
            self.fc3 = nn.Linear(1024, 2)
        else:
            raise ValueError(head)

    def trunk(self, x, padding_mask=None):
        if x.ndim == 3:
            x = x[:, :, 0]       # model.py:134-137
        return self.ssl_model.model(x, padding_mask=padding_mask, mask=False, features_only=True)

    def forward(self, x, return_sae_loss=False, padding_mask=None, taps: Optional[dict] = None):
        res = self.trunk(x, padding_mask)
        if taps is not None:
            taps["x"] = res["x"]
            taps["layer_results"] = [lr[0] for lr in res["layer_results"]]
        if self.head == "sls":
            return self._sls(res["layer_results"])
        x_ssl = res["x"]
        B, T, C = x_ssl.shape
        enc = self.sae.encode(x_ssl, temporal_dim=T)                     # model.py:221 / window:349
        if taps is not None:
            taps["encoded"] = enc
        sae_loss = F.mse_loss(self.sae.decode(enc.reshape(B * T, -1)), x_ssl.reshape(B * T, C)) if return_sae_loss else None
        feats = enc if self.use_sparse_features else self.sae.decode(enc.reshape(B * T, -1)).reshape(B, T, C)
        if res["padding_mask"] is not None:
            # extension (BASELINE config 4): mean over VALID frames only; reference callers never pad
            valid = (~res["padding_mask"]).unsqueeze(-1).to(feats.dtype)
            pooled = (feats * valid).sum(1) / valid.sum(1)
        else:
            pooled = self.pool(feats.transpose(1, 2)).squeeze(-1)         # model.py:245
        if taps is not None:
            taps["pooled"] = pooled
        out = F.log_softmax(self.classifier(pooled), dim=-1)              # model.py:246-247
        return (out, sae_loss) if return_sae_loss else out

    def _sls(self, layer_results):
        # getAttenF, model_backup.py:186-202
        pooled, full = [], []
        for lr in layer_results:
            y = lr[0].transpose(0, 1).transpose(1, 2)
            pooled.append(F.adaptive_avg_pool1d(y, 1).transpose(1, 2))
            xx = lr[0].transpose(0, 1)
            full.append(xx.reshape(xx.size(0), -1, xx.size(1), xx.size(2)))
        y0 = torch.cat(pooled, dim=1)
        fullfeature = torch.cat(full, dim=1)
        # upstream SLS classifier (not in this tree)
        y0 = self.sig(self.fc0(y0))
        y0 = y0.view(y0.shape[0], y0.shape[1], y0.shape[2], -1)
        fullfeature = (fullfeature * y0).sum(1).unsqueeze(1)
        x = self.selu(self.first_bn(fullfeature))
        x = F.max_pool2d(x, (3, 3))
        x = torch.flatten(x, 1)
        x = self.selu(self.fc1(x))
        x = self.selu(self.fc3(x))
        return F.log_softmax(x, dim=1)
