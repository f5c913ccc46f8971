"""Drop-in ``Model`` / ``SSLModel`` / ``AutoEncoderTopK`` for the reference's scoring path.

Same constructor, forward arity, attribute names and state_dict keys as
``/root/reference/model.py:42-260`` (H-SAE, what ``main.py --is_eval`` runs) and
``/root/reference/model_window_topk.py:40-393`` (H-WIN, ``--use_window_topk``); the arithmetic runs in
``libslsb200.so`` (hand-written sm_100a CUDA behind the C ABI of ``include/slsb200.h``).
There is no CPU path: calling forward without the built extension or without a B200 raises.
"""
from __future__ import annotations

import os
import weakref
from typing import Optional

import torch
import torch.nn as nn

from . import _lib
from .engine import Engine, make_config, PRECISIONS, HEAD_NONE, HEAD_SAE, HEAD_WINDOW, HEAD_SLS
from .weights import TrunkGeometry, TrunkParams, pack_state_dict


def _read_fairseq_checkpoint(cp_path: str):
    """A fairseq ``xlsr2_300m.pt`` without fairseq: the tensors of ckpt['model'].
    (reference: fairseq.checkpoint_utils.load_model_ensemble_and_task, model.py:113-115.)"""
    if not os.path.exists(cp_path):
        raise RuntimeError(
            f"Could not load SSL checkpoint '{cp_path}' (file not found). Pass cp_path=None for a random-init trunk "
            "(synthetic benchmarks / parity tests).")
    from .weights import load_checkpoint_tensors
    return load_checkpoint_tensors(cp_path)        # allowlist unpickler (weights.py): only tensors + plain containers resolve, every other global becomes an inert stub


def _load_fairseq_checkpoint(trunk: TrunkParams, cp_path: str, sd=None) -> None:
    sd = _read_fairseq_checkpoint(cp_path) if sd is None else sd
    missing, unexpected = trunk.load_state_dict(sd, strict=False)
    hard = [k for k in missing if not k.startswith(("quantizer", "project_q", "final_proj"))]
    if hard:
        raise RuntimeError(f"SSL checkpoint '{cp_path}' lacks trunk tensors: {hard[:5]} ...")


class _EngineOwner:
    """Mixin: lazily builds the engine on the parameters' device and keeps its weight arena in sync."""

    _engine: Optional[Engine] = None
    _weights_dirty: bool = True

    def _engine_config(self):  # pragma: no cover - overridden
        raise NotImplementedError

    def _sls_kp(self) -> int:
        return 0

    def _apply(self, fn, *a, **k):           # .to() / .cuda() / .float() move parameters -> repack
        self._weights_dirty = True
        return super()._apply(fn, *a, **k)

    def _mark_dirty(self, *_):
        self._weights_dirty = True

    def refresh_weights(self) -> None:
        """Re-pack parameters into the engine (call after modifying parameters in place)."""
        self._weights_dirty = True

    def engine(self) -> Engine:
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise _lib.SlsbError("Model parameters are on %s: the B200 path needs model.to('cuda') (no CPU fallback)" % dev)
        if getattr(self, "_is_replica", False):
            raise _lib.SlsbError("nn.DataParallel replicas are not supported: run one process per GPU "
                                 "(CUDA_VISIBLE_DEVICES / torchrun), see INTEGRATION.md")
        if self._engine is None or self._engine.device != dev:
            if self._engine is not None:
                self._engine.close()
            self._engine = Engine(self._engine_config(), dev)
            self._weights_dirty = True
        if self._weights_dirty:
            with torch.no_grad():
                packed = pack_state_dict(self.state_dict(), self.ssl_model.model.geo, sls_kp=self._sls_kp())
                self._engine.load_packed(packed)
            self._weights_dirty = False
        return self._engine

    def _prec(self) -> int:
        return PRECISIONS[os.environ.get("SLSB_PRECISION", self.precision)]


def _prep_wav(x: torch.Tensor) -> torch.Tensor:
    if x.ndim == 3:
        x = x[:, :, 0]            # model.py:134-137
    if x.ndim != 2:
        raise ValueError(f"expected audio of shape [B, S] or [B, S, 1], got {tuple(x.shape)}")
    return x.to(torch.float32).contiguous()


class AutoEncoderTopK(nn.Module):
    """Parameter holder + engine-backed encode/decode (model.py:42-104; window variant model_window_topk.py:40-216)."""

    def __init__(self, activation_dim: int, dict_size: int, k: int, window_size: int = 1):
        super().__init__()
        self.activation_dim, self.dict_size, self.window_size = activation_dim, dict_size, window_size
        assert isinstance(k, int) and k > 0, f"k={k} must be a positive integer"
        self.register_buffer("k", torch.tensor(k, dtype=torch.int))
        self.decoder = nn.Linear(dict_size, activation_dim, bias=False)
        self.decoder.weight.data = self.decoder.weight.data / torch.norm(self.decoder.weight.data, dim=0, keepdim=True)
        self.encoder = nn.Linear(activation_dim, dict_size)
        self.encoder.weight.data = self.decoder.weight.T.clone()
        self.encoder.bias.data.zero_()
        self.b_dec = nn.Parameter(torch.zeros(activation_dim))
        self._owner = None

    def _model(self):
        m = self._owner() if self._owner is not None else None
        if m is None:
            raise _lib.SlsbError("AutoEncoderTopK runs inside a Model (it shares the Model's engine)")
        return m

    def encode(self, x: torch.Tensor, temporal_dim: Optional[int] = None) -> torch.Tensor:
        m = self._model()
        shape = x.shape
        if x.dim() == 3:
            T = shape[1]
        elif temporal_dim is not None:
            T = temporal_dim
        else:
            T = 0
        window = self.window_size if T > 0 else 1     # flat input without temporal_dim: per-row top-k (window:89-95)
        flat = x.reshape(-1, shape[-1]).to(torch.float32).contiguous()
        out = m.engine().sae_encode(flat, T if T > 0 else 1, window, m._prec())
        return out.reshape(*shape[:-1], -1)

    def decode(self, x: torch.Tensor) -> torch.Tensor:
        m = self._model()
        flat = x.reshape(-1, x.shape[-1]).to(torch.float32).contiguous()
        return m.engine().sae_decode(flat, m._prec()).reshape(*x.shape[:-1], -1)

    def forward(self, x: torch.Tensor):
        encoded = self.encode(x)
        return self.decode(encoded), encoded


class SSLModel(nn.Module):
    """model.py:106-141.  ``self.model`` holds the XLS-R parameters under fairseq names."""

    def __init__(self, device, cp_path: Optional[str] = "xlsr2_300m.pt", geometry: Optional[TrunkGeometry] = None):
        super().__init__()
        sd = _read_fairseq_checkpoint(cp_path) if cp_path is not None else None
        if geometry is None and sd is not None:
            geometry = TrunkGeometry.from_state_dict(sd)      # fairseq builds the model from the checkpoint, not from defaults
        self.model = TrunkParams(geometry)
        if sd is not None:
            _load_fairseq_checkpoint(self.model, cp_path, sd)
        self.device = device
        self.out_dim = self.model.geo.embed_dim
        self._owner = None

    def extract_feat(self, input_data: torch.Tensor) -> torch.Tensor:
        m = self._owner() if self._owner is not None else None
        if m is None:
            raise _lib.SlsbError("SSLModel runs inside a Model (it shares the Model's engine)")
        return m.engine().extract_feat(_prep_wav(input_data), m._prec())


class Model(_EngineOwner, nn.Module):
    """Audio deepfake detector: XLS-R trunk + TopK SAE + mean-pool + MLP (model.py:144-260).

    ``sae_window_size > 1`` selects the window top-k variant (model_window_topk.py); extra keyword
    ``precision`` ('bf16' default, 'fp32' = CUDA-core verification mode) is the only addition.
    """

    def __init__(self, args, device, cp_path: Optional[str] = "xlsr2_300m.pt", use_sae: bool = True,
                 use_sparse_features: bool = True, sae_dict_size: int = 4096, sae_k: int = 128,
                 sae_weight: float = 0.1, sae_window_size: int = 1, precision: str = "bf16",
                 geometry: Optional[TrunkGeometry] = None):
        super().__init__()
        self.device = device
        self.use_sae, self.use_sparse_features, self.sae_weight = use_sae, use_sparse_features, sae_weight
        self.precision = precision
        self.ssl_model = SSLModel(device=device, cp_path=cp_path, geometry=geometry)
        D = self.ssl_model.out_dim
        if self.use_sae:
            self.sae = AutoEncoderTopK(activation_dim=D, dict_size=sae_dict_size, k=sae_k, window_size=sae_window_size)
            input_dim = sae_dict_size if use_sparse_features else D
        else:
            input_dim = D
        self.pool = nn.AdaptiveAvgPool1d(1)
        self.classifier = nn.Sequential(nn.LayerNorm(input_dim), nn.Linear(input_dim, 256), nn.ReLU(), nn.Dropout(0.3),
                                        nn.Linear(256, 2))
        self.last_sparse_features = None
        self.last_feature_indices = None
        # True: every forward keeps its layer results / SAE activations readable through the engine (get_tensor, last_sparse_code);
        # forward sets it for the call when the caller asks for sae_loss / interpretability, which need them (model.py:224-240)
        self.retain_intermediates = False
        self._input_dim = input_dim
        self._bind()
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._mark_dirty())

    def _bind(self):
        ref = weakref.ref(self)
        self.ssl_model._owner = ref
        if self.use_sae:
            self.sae._owner = ref

    def _engine_config(self):
        geo = self.ssl_model.model.geo
        if self.use_sae:
            return make_config(geo, sae_dict=self.sae.dict_size, sae_k=int(self.sae.k), sae_window=self.sae.window_size,
                               cls_in=self._input_dim)
        return make_config(geo, sae_dict=0, cls_in=self._input_dim)

    def _head(self) -> int:
        return HEAD_WINDOW if (self.use_sae and self.sae.window_size > 1) else HEAD_SAE

    def forward(self, input_data: torch.Tensor, return_sae_loss: bool = True, return_interpretability: bool = False,
                sample_lengths: Optional[torch.Tensor] = None):
        """Returns log-probs [B, 2] (class 1 = bonafide) with the reference's arity rule (model.py:253-260).
        ``sample_lengths`` (int32 [B], optional) enables the padding-mask path of wav2vec2.py:567-586."""
        eng = self.engine()
        wav = _prep_wav(input_data)
        lens = None if sample_lengths is None else sample_lengths.to(device=wav.device, dtype=torch.int32).contiguous()
        retain = self.retain_intermediates or (self.use_sae and (return_sae_loss or return_interpretability))
        output = eng.forward(wav, self._head(), self._prec(), lens, retain=retain)
        sae_loss, interp = None, None
        if self.use_sae and return_sae_loss:
            sae_loss = eng.sae_loss(self._prec())                                       # model.py:224-225
        if self.use_sae and return_interpretability:
            B, T = wav.shape[0], eng.frames(wav.shape[1])
            self.last_sparse_features = eng.get_tensor("encoded", (B, T, self.sae.dict_size))   # model.py:236-240
            self.last_feature_indices = self.last_sparse_features > 0
            interp = self.get_interpretability_info(None)
        if return_interpretability:
            return (output, sae_loss, interp) if return_sae_loss else (output, interp)
        if return_sae_loss:
            return output, sae_loss
        return output

    def last_sparse_code(self, batch: int, samples: int):
        """Compact form of the last forward's SAE code: ``(indices int32 [B, T, k], values fp32 [B, T, k], counts int32
        [B, T])``, indices ascending, unused slots -1.  ``last_sparse_features`` (dense, model.py:236-240) is its scatter;
        the analysis scripts can consume this instead of the 4096-wide dense tensor."""
        eng = self.engine()
        T = eng.frames(samples)
        idx, val, cnt = eng.get_sparse(batch * T)
        k = idx.shape[-1]
        return idx.view(batch, T, k), val.view(batch, T, k), cnt.view(batch, T)

    def get_interpretability_info(self, pooled_features):
        """model.py:262-293 (offline analysis surface; plain torch ops on the engine's sparse features)."""
        if self.last_sparse_features is None:
            return None
        f = self.last_sparse_features
        avg = f.mean(dim=1)
        vals, idx = avg.topk(k=min(20, f.shape[-1]), dim=-1)
        active = (f > 0).float()
        return {"avg_activation": avg, "top20_features": idx, "top20_values": vals, "sparsity": active.mean(dim=[1, 2]),
                "activation_freq": active.mean(dim=1), "sparse_features": f}

    def compute_total_loss(self, classification_loss: torch.Tensor, sae_loss: torch.Tensor = None):
        if sae_loss is None or not self.use_sae:
            return classification_loss
        return classification_loss + (self.sae_weight * sae_loss)


class ModelWindowTopK(Model):
    """model_window_topk.py:271-393: same surface with ``sae_window_size=8`` by default."""

    def __init__(self, args, device, cp_path: Optional[str] = "xlsr2_300m.pt", use_sae: bool = True,
                 use_sparse_features: bool = True, sae_dict_size: int = 4096, sae_k: int = 128, sae_window_size: int = 8,
                 sae_weight: float = 0.1, precision: str = "bf16", geometry: Optional[TrunkGeometry] = None):
        super().__init__(args, device, cp_path=cp_path, use_sae=use_sae, use_sparse_features=use_sparse_features,
                         sae_dict_size=sae_dict_size, sae_k=sae_k, sae_weight=sae_weight, sae_window_size=sae_window_size,
                         precision=precision, geometry=geometry)


class ModelSLS(_EngineOwner, nn.Module):
    """The SLS layer-attention head the north star names: getAttenF (model_backup.py:186-202) + the upstream
    classifier (fc0 / sigmoid weighting / BN / SELU / 3x3 max-pool / fc1 / fc3 / log-softmax)."""

    def __init__(self, args, device, cp_path: Optional[str] = "xlsr2_300m.pt", precision: str = "bf16",
                 geometry: Optional[TrunkGeometry] = None, frames: int = 201):
        super().__init__()
        self.device = device
        self.precision = precision
        self.ssl_model = SSLModel(device=device, cp_path=cp_path, geometry=geometry)
        D = self.ssl_model.out_dim
        self.first_bn = nn.BatchNorm2d(num_features=1)
        self.selu = nn.SELU(inplace=True)
        self.fc0 = nn.Linear(D, 1)
        self.sig = nn.Sigmoid()
        self.fc1 = nn.Linear((frames // 3) * (D // 3), 1024)      # 22847 for T = 201
        self.fc3 = nn.Linear(1024, 2)
        self.logsoftmax = nn.LogSoftmax(dim=1)
        self._frames = frames
        self.ssl_model._owner = weakref.ref(self)
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._mark_dirty())

    def _engine_config(self):
        return make_config(self.ssl_model.model.geo, sls_frames=self._frames, sls_hidden=self.fc1.out_features)

    def _sls_kp(self) -> int:
        q = 64 * 17          # matches csrc/engine.cu: a multiple of the fp32 17-way split and of the 64-wide k-blocks
        return (self.fc1.in_features + q - 1) // q * q

    def _head(self) -> int:
        return HEAD_SLS

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.engine().forward(_prep_wav(x), HEAD_SLS, self._prec())

    def layer_results(self, batch: int, samples: int):
        """[(x_i [T, B, C], None)] of the last forward, the tuple list fairseq returns (wav2vec2.py:958)."""
        eng = self.engine()
        T, D = eng.frames(samples), self.ssl_model.out_dim
        return [(eng.get_tensor(f"layer_results.{i}", (batch, T, D)).transpose(0, 1), None)
                for i in range(self.ssl_model.model.geo.layers)]


def getAttenF(layerResult):
    """model_backup.py:186-202 on a layer_results list (offline helper; the hot path fuses this on the device)."""
    pooled = [lr[0].transpose(0, 1).mean(dim=1, keepdim=True) for lr in layerResult]
    full = [lr[0].transpose(0, 1).unsqueeze(1) for lr in layerResult]
    return torch.cat(pooled, dim=1), torch.cat(full, dim=1)
