"""Parameter holders with fairseq / reference state_dict names, and the packer that turns a state_dict
into the engine's tensor table (names documented in INTEGRATION.md).

The holders never run a torch forward: they exist so that ``Model`` is an ``nn.Module`` whose
``parameters()``, ``state_dict()`` and ``load_state_dict(strict=True)`` behave like the reference's
(``/root/reference/main.py:518-521, :586-592``; key names per ``wav2vec/wav2vec2.py:264-362, :803-813,
:862-875, :1009-1028`` and ``model.py:53-66, :183-189``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import torch
import torch.nn as nn


@dataclass
class TrunkGeometry:
    """XLS-R 300M (wav2vec2-large, extractor_mode=layer_norm, conv_bias, layer_norm_first)."""
    conv_layers: List[Tuple[int, int, int]] = field(
        default_factory=lambda: [(512, 10, 5)] + [(512, 3, 2)] * 4 + [(512, 2, 2)] * 2)   # wav2vec2.py:97-99
    embed_dim: int = 1024
    ffn_dim: int = 4096
    heads: int = 16
    layers: int = 24
    conv_pos: int = 128
    conv_pos_groups: int = 16
    final_dim: int = 768
    latent_vars: int = 320
    latent_groups: int = 2

    def frames(self, samples: int) -> int:
        for _, k, s in self.conv_layers:        # wav2vec2.py:523-538
            samples = max(0, (samples - k) // s + 1)
        return samples

    @classmethod
    def from_state_dict(cls, sd: Dict[str, torch.Tensor], prefix: str = "") -> "TrunkGeometry":
        """The trunk's shape as the checkpoint's own tensors give it.  fairseq rebuilds the model from the checkpoint's ``cfg``
        (``load_model_ensemble_and_task``, model.py:113-115); the tensors carry the same information: encoder depth = number
        of ``encoder.layers.<i>``, widths from ``post_extract_proj`` / ``fc1``, conv stack from ``conv_layers.<i>.0.weight``
        (strides are not stored in tensors: XLS-R's 5,2,2,2,2,2,2 are kept, wav2vec2.py:97-99)."""
        g = cls()
        idx = [int(k[len(prefix) + len("encoder.layers."):].split(".", 1)[0]) for k in sd if k.startswith(prefix + "encoder.layers.")]
        if idx:
            g.layers = max(idx) + 1
        if prefix + "post_extract_proj.weight" in sd:
            g.embed_dim = int(sd[prefix + "post_extract_proj.weight"].shape[0])
            g.heads = g.embed_dim // 64
        if prefix + "encoder.layers.0.fc1.weight" in sd:
            g.ffn_dim = int(sd[prefix + "encoder.layers.0.fc1.weight"].shape[0])
        convs = []
        for i, (dim, k, st) in enumerate(g.conv_layers):
            w = sd.get(prefix + f"feature_extractor.conv_layers.{i}.0.weight")
            convs.append((int(w.shape[0]), int(w.shape[2]), st) if w is not None else (dim, k, st))
        g.conv_layers = convs
        return g


class _Quantizer(nn.Module):
    """Pre-training-only GumbelVectorQuantizer parameters (kept for strict state_dict loading)."""

    def __init__(self, geo: TrunkGeometry):
        super().__init__()
        embed = geo.conv_layers[-1][0]
        n = geo.latent_vars * geo.latent_groups
        self.vars = nn.Parameter(torch.zeros(1, n, geo.final_dim // geo.latent_groups))
        self.weight_proj = nn.Linear(embed, n)


class _AttnParams(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.k_proj = nn.Linear(dim, dim)
        self.v_proj = nn.Linear(dim, dim)
        self.q_proj = nn.Linear(dim, dim)
        self.out_proj = nn.Linear(dim, dim)


class _LayerParams(nn.Module):
    def __init__(self, geo: TrunkGeometry):
        super().__init__()
        self.self_attn = _AttnParams(geo.embed_dim)
        self.self_attn_layer_norm = nn.LayerNorm(geo.embed_dim)
        self.fc1 = nn.Linear(geo.embed_dim, geo.ffn_dim)
        self.fc2 = nn.Linear(geo.ffn_dim, geo.embed_dim)
        self.final_layer_norm = nn.LayerNorm(geo.embed_dim)


class _EncoderParams(nn.Module):
    def __init__(self, geo: TrunkGeometry):
        super().__init__()
        D = geo.embed_dim
        conv = nn.Conv1d(D, D, kernel_size=geo.conv_pos, padding=geo.conv_pos // 2, groups=geo.conv_pos_groups)
        nn.init.normal_(conv.weight, mean=0, std=math.sqrt(4.0 / (geo.conv_pos * D)))   # wav2vec2.py:870-872
        nn.init.constant_(conv.bias, 0)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            conv = torch.nn.utils.weight_norm(conv, name="weight", dim=2)             # -> weight_g, weight_v
        self.pos_conv = nn.Sequential(conv, nn.Identity(), nn.GELU())
        self.layers = nn.ModuleList([_LayerParams(geo) for _ in range(geo.layers)])
        self.layer_norm = nn.LayerNorm(D)
        for m in self.modules():                                                       # init_bert_params, :899
            if isinstance(m, nn.Linear):
                m.weight.data.normal_(mean=0.0, std=0.02)
                m.bias.data.zero_()


class TrunkParams(nn.Module):
    """Stands where fairseq's ``Wav2Vec2Model`` stands in the reference (``ssl_model.model``)."""

    def __init__(self, geo: TrunkGeometry = None):
        super().__init__()
        self.geo = geo = geo or TrunkGeometry()
        embed = geo.conv_layers[-1][0]
        fe = nn.Module()
        fe.conv_layers = nn.ModuleList()
        in_d = 1
        for dim, k, s in geo.conv_layers:
            conv = nn.Conv1d(in_d, dim, k, stride=s, bias=True)
            nn.init.kaiming_normal_(conv.weight)                                       # wav2vec2.py:796
            fe.conv_layers.append(nn.Sequential(
                conv, nn.Dropout(0.0), nn.Sequential(nn.Identity(), nn.LayerNorm(dim), nn.Identity()), nn.GELU()))
            in_d = dim
        self.feature_extractor = fe
        self.post_extract_proj = nn.Linear(embed, geo.embed_dim)
        self.mask_emb = nn.Parameter(torch.zeros(geo.embed_dim).uniform_())
        self.encoder = _EncoderParams(geo)
        self.layer_norm = nn.LayerNorm(embed)
        self.quantizer = _Quantizer(geo)
        self.project_q = nn.Linear(geo.final_dim, geo.final_dim)
        self.final_proj = nn.Linear(geo.embed_dim, geo.final_dim)

    def forward(self, *a, **k):   # pragma: no cover - the arithmetic lives in libslsb200.so
        raise RuntimeError("TrunkParams holds parameters only; call SSLModel.extract_feat / Model.forward")


def pack_state_dict(sd: Dict[str, torch.Tensor], geo: TrunkGeometry, prefix: str = "ssl_model.model.",
                    sls_kp: int = 0) -> Dict[str, torch.Tensor]:
    """state_dict (reference names) -> engine tensor table (fp32, contiguous, same device as the inputs).

    Folds done here, once per weight load (not on the hot path):
      * conv weights [out, in, k] -> [out, k*in] (tap-major, channels-last implicit GEMM);
      * pos_conv weight-norm  w = g * v / ||v||_(out,in)  (dim=2, wav2vec2.py:874) -> [out, k*64];
      * q/k/v fused into one [3D, D] matrix, q (weight and bias) pre-scaled by head_dim**-0.5 (exact: power of two).
    """
    def g(name):
        return sd[prefix + name].detach().float()

    out: Dict[str, torch.Tensor] = {}
    for i, (dim, k, s) in enumerate(geo.conv_layers):
        w = g(f"feature_extractor.conv_layers.{i}.0.weight")
        if i == 0:
            out["conv0.w"] = w.reshape(dim, k)
        else:
            out[f"conv{i}.w"] = w.permute(0, 2, 1).reshape(dim, -1)
        out[f"conv{i}.b"] = g(f"feature_extractor.conv_layers.{i}.0.bias")
        out[f"conv{i}.ln.w"] = g(f"feature_extractor.conv_layers.{i}.2.1.weight")
        out[f"conv{i}.ln.b"] = g(f"feature_extractor.conv_layers.{i}.2.1.bias")
    out["feat_ln.w"], out["feat_ln.b"] = g("layer_norm.weight"), g("layer_norm.bias")
    out["proj.w"], out["proj.b"] = g("post_extract_proj.weight"), g("post_extract_proj.bias")
    v, gg = g("encoder.pos_conv.0.weight_v"), g("encoder.pos_conv.0.weight_g")
    w = v * (gg / v.norm(dim=(0, 1), keepdim=True))
    out["pos.w"] = w.permute(0, 2, 1).reshape(w.shape[0], -1)
    out["pos.b"] = g("encoder.pos_conv.0.bias")
    scale = (geo.embed_dim // geo.heads) ** -0.5
    for l in range(geo.layers):
        p = f"encoder.layers.{l}."
        out[f"L{l}.ln1.w"], out[f"L{l}.ln1.b"] = g(p + "self_attn_layer_norm.weight"), g(p + "self_attn_layer_norm.bias")
        out[f"L{l}.qkv.w"] = torch.cat([g(p + "self_attn.q_proj.weight") * scale, g(p + "self_attn.k_proj.weight"),
                                        g(p + "self_attn.v_proj.weight")], 0)
        out[f"L{l}.qkv.b"] = torch.cat([g(p + "self_attn.q_proj.bias") * scale, g(p + "self_attn.k_proj.bias"),
                                        g(p + "self_attn.v_proj.bias")], 0)
        out[f"L{l}.out.w"], out[f"L{l}.out.b"] = g(p + "self_attn.out_proj.weight"), g(p + "self_attn.out_proj.bias")
        out[f"L{l}.ln2.w"], out[f"L{l}.ln2.b"] = g(p + "final_layer_norm.weight"), g(p + "final_layer_norm.bias")
        out[f"L{l}.fc1.w"], out[f"L{l}.fc1.b"] = g(p + "fc1.weight"), g(p + "fc1.bias")
        out[f"L{l}.fc2.w"], out[f"L{l}.fc2.b"] = g(p + "fc2.weight"), g(p + "fc2.bias")
    out["enc_ln.w"], out["enc_ln.b"] = g("encoder.layer_norm.weight"), g("encoder.layer_norm.bias")

    def h(name):
        return sd[name].detach().float()

    if "sae.encoder.weight" in sd:
        out["sae.enc.w"], out["sae.enc.b"] = h("sae.encoder.weight"), h("sae.encoder.bias")
        out["sae.b_dec"], out["sae.dec.w"] = h("sae.b_dec"), h("sae.decoder.weight")
    if "classifier.0.weight" in sd:
        out["cls.ln.w"], out["cls.ln.b"] = h("classifier.0.weight"), h("classifier.0.bias")
        out["cls.fc1.w"], out["cls.fc1.b"] = h("classifier.1.weight"), h("classifier.1.bias")
        out["cls.fc2.w"], out["cls.fc2.b"] = h("classifier.4.weight"), h("classifier.4.bias")
    if "fc0.weight" in sd:
        out["sls.fc0.w"], out["sls.fc0.b"] = h("fc0.weight").reshape(-1), h("fc0.bias")
        out["sls.bn"] = torch.stack([h("first_bn.weight")[0], h("first_bn.bias")[0],
                                     h("first_bn.running_mean")[0], h("first_bn.running_var")[0]])
        w1 = h("fc1.weight")
        pad = sls_kp - w1.shape[1]
        out["sls.fc1.w"] = torch.nn.functional.pad(w1, (0, pad)) if pad > 0 else w1
        out["sls.fc1.b"] = h("fc1.bias")
        out["sls.fc3.w"], out["sls.fc3.b"] = h("fc3.weight"), h("fc3.bias")
    return {k: v.contiguous() for k, v in out.items()}


# ---------------------------------------------------------------------------------------------------------------------
# checkpoints without fairseq (main.py:531-592, model.py:113-115)
# ---------------------------------------------------------------------------------------------------------------------
# Exact (module, name) pairs a checkpoint may resolve.  Whole modules are NOT trusted: ``builtins.eval``, ``builtins.getattr``,
# ``torch.utils.cpp_extension.load``, ``torch.storage._load_from_bytes``, ``functools.partial`` ... are all reachable through
# "safe-looking" roots, so anything that is not listed here (or is not a torch dtype / legacy storage class / numpy scalar
# type, checked by type below) is replaced by an inert stub and is never called.
_SAFE_GLOBALS = frozenset({
    ("collections", "OrderedDict"), ("collections", "defaultdict"),
    ("torch._utils", "_rebuild_tensor_v2"), ("torch._utils", "_rebuild_tensor"), ("torch._utils", "_rebuild_parameter"),
    ("torch._utils", "_rebuild_parameter_with_state"), ("torch._tensor", "_rebuild_from_type_v2"),
    ("torch", "Size"), ("torch", "device"), ("torch", "Tensor"), ("torch.nn.parameter", "Parameter"),
    ("torch.storage", "UntypedStorage"), ("torch.storage", "TypedStorage"),
    ("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"),
    ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar"), ("numpy", "ndarray"), ("numpy", "dtype"),
    ("argparse", "Namespace"), ("_codecs", "encode"), ("copyreg", "_reconstructor"),
    ("builtins", "set"), ("builtins", "frozenset"), ("builtins", "dict"), ("builtins", "list"), ("builtins", "tuple"),
    ("builtins", "int"), ("builtins", "float"), ("builtins", "complex"), ("builtins", "bool"), ("builtins", "str"),
    ("builtins", "bytes"), ("builtins", "bytearray"), ("builtins", "slice"), ("builtins", "range"), ("builtins", "object"),
})


def _is_safe_global(module: str, name: str) -> bool:
    if (module, name) in _SAFE_GLOBALS:
        return True
    if module == "torch" and "." not in name:
        obj = getattr(torch, name, None)
        if isinstance(obj, torch.dtype):                                   # torch.float32, torch.bfloat16, ...
            return True
        if name.endswith("Storage") and isinstance(obj, type):           # legacy typed storages named in persistent ids
            return True
    if module == "numpy" and "." not in name:
        import numpy as np
        obj = getattr(np, name, None)
        return isinstance(obj, type) and issubclass(obj, np.generic)       # numpy scalar types (np.float64, ...)
    return False


def _stub_class(module: str, name: str):
    """Stand-in for everything outside the allowlist (omegaconf configs, fairseq task / criterion objects, and equally any
    function a hostile file names): constructible / callable with anything, absorbs any pickled state, runs no code of the
    named object - the named module is never even imported."""
    def _init(self, *a, **k):
        self.__dict__["_args"] = (a, k)

    def _setstate(self, state):
        self.__dict__["_state"] = state

    return type(name, (), {"__module__": module, "__init__": _init, "__setstate__": _setstate, "__call__": lambda self, *a, **k: self,
                           "__getattr__": lambda self, k: None, "__reduce__": lambda self: (dict, ()),
                           "__setitem__": lambda self, k, v: None, "append": lambda self, v: None, "extend": lambda self, v: None,
                           "add": lambda self, v: None, "update": lambda self, *a, **k: None})


class _TensorOnlyUnpickler:
    """``pickle_module`` for ``torch.load``: tensors and plain containers load normally; a global is resolved only when its exact
    (module, name) is on ``_SAFE_GLOBALS`` (or it is a torch dtype / storage class / numpy scalar type); every other global
    becomes an inert stub.  So a fairseq ``xlsr2_300m.pt`` (``cfg`` = omegaconf objects) or a training checkpoint written by
    ``main.py`` loads on a box that has neither fairseq nor omegaconf, and a crafted file cannot reach ``eval`` / ``exec`` /
    ``os.system`` / ``torch.utils.cpp_extension.load`` through it (tests/test_host.py::test_checkpoint_unpickler_*)."""
    import pickle as _pickle
    __name__ = "pickle"

    class Unpickler(_pickle.Unpickler):
        def find_class(self, module, name):
            if _is_safe_global(module, name):
                return super().find_class(module, name)
            return _stub_class(module, name)

    @classmethod
    def load(cls, f, **kw):
        return cls.Unpickler(f, **kw).load()


def load_checkpoint_tensors(path: str) -> Dict[str, torch.Tensor]:
    """Returns the state_dict stored in ``path``: a fairseq checkpoint (``ckpt['model']``), a ``main.py`` checkpoint
    (a bare state_dict, keys possibly prefixed with ``module.``) or ``{'state_dict': ...}``.  No fairseq import."""
    ckpt = torch.load(path, map_location="cpu", weights_only=False, pickle_module=_TensorOnlyUnpickler)
    if isinstance(ckpt, dict):
        for key in ("model_state_dict", "state_dict", "model"):          # main.py:531-536 _get_state_dict order
            if key in ckpt and isinstance(ckpt[key], dict):
                ckpt = ckpt[key]
                break
    if not isinstance(ckpt, dict) or not any(torch.is_tensor(v) for v in ckpt.values()):
        raise RuntimeError(f"'{path}' holds no state_dict")
    return {k: v for k, v in ckpt.items() if torch.is_tensor(v)}


def fix_module_prefix(state_dict: Dict[str, torch.Tensor], model_is_wrapped: bool) -> Dict[str, torch.Tensor]:
    """main.py:542-560: add / strip the ``module.`` prefix so the keys match the model's DataParallel wrapping."""
    if not state_dict:
        return state_dict
    has = all(isinstance(k, str) and k.startswith("module.") for k in state_dict)
    if model_is_wrapped and not has:
        return {"module." + k: v for k, v in state_dict.items()}
    if not model_is_wrapped and has:
        return {k[len("module."):]: v for k, v in state_dict.items()}
    return state_dict


def load_model_checkpoint(model: nn.Module, path: str):
    """main.py:576-592 without fairseq: locate the state_dict, fix the prefix, strict load with the reference's
    non-strict fallback.  Returns ``load_state_dict``'s result."""
    sd = fix_module_prefix(load_checkpoint_tensors(path), isinstance(model, nn.DataParallel))
    try:
        return model.load_state_dict(sd, strict=True)
    except RuntimeError as e:
        print("Strict load failed; retrying with strict=False.\n    Reason: {}".format(str(e)[:300]))
        return model.load_state_dict(sd, strict=False)
