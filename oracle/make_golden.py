"""ORACLE (test infrastructure only) -- pins the oracle and mints the committed golden fixtures.

Run in the BUILD container (needs /root/reference and ``transformers``):

    python -m oracle.make_golden            # writes tests/golden/*.npz, prints the pin report

Pins
----
1. ``crosscheck_hf``       : oracle trunk vs ``transformers.Wav2Vec2Model(do_stable_layer_norm=True)``
                             (independent implementation of fairseq's wav2vec2).
2. ``crosscheck_reference``: oracle heads vs the reference's OWN ``model.py`` / ``model_window_topk.py`` /
                             ``model_backup.py::getAttenF`` executed verbatim through ``fairseq_stub``.
3. fixtures                : seeded weights (``seeded_init_``, seed 1234) + ``synth_clips`` -> log-probs and
                             strided taps of the intermediate tensors, full XLS-R-300M size.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

from .trunk import TrunkConfig, Wav2Vec2Trunk, seeded_init_, synth_clips
from .heads import OracleModel, SAETopK, window_topk, canonical_topk_mask

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def hf_model_from_oracle(trunk: Wav2Vec2Trunk):
    from transformers import Wav2Vec2Config, Wav2Vec2Model
    c = trunk.cfg
    hc = Wav2Vec2Config(
        hidden_size=c.embed_dim, num_hidden_layers=c.layers, num_attention_heads=c.heads,
        intermediate_size=c.ffn_dim, feat_extract_norm="layer", conv_bias=True, do_stable_layer_norm=True,
        conv_dim=[d for d, _, _ in c.conv_layers], conv_kernel=[k for _, k, _ in c.conv_layers],
        conv_stride=[s for _, _, s in c.conv_layers],
        num_conv_pos_embeddings=c.conv_pos, num_conv_pos_embedding_groups=c.conv_pos_groups,
        hidden_dropout=0.0, attention_dropout=0.0, activation_dropout=0.0, feat_proj_dropout=0.0,
        layerdrop=0.0, final_dropout=0.0, mask_time_prob=0.0, mask_feature_prob=0.0,
        hidden_act="gelu", attn_implementation="eager")
    hf = Wav2Vec2Model(hc).eval()
    sd = trunk.state_dict()
    m = {}
    for i in range(len(c.conv_layers)):
        for p in ("weight", "bias"):
            m[f"feature_extractor.conv_layers.{i}.conv.{p}"] = sd[f"feature_extractor.conv_layers.{i}.0.{p}"]
            m[f"feature_extractor.conv_layers.{i}.layer_norm.{p}"] = sd[f"feature_extractor.conv_layers.{i}.2.1.{p}"]
    for p in ("weight", "bias"):
        m[f"feature_projection.layer_norm.{p}"] = sd[f"layer_norm.{p}"]
        m[f"feature_projection.projection.{p}"] = sd[f"post_extract_proj.{p}"]
        m[f"encoder.layer_norm.{p}"] = sd[f"encoder.layer_norm.{p}"]
    m["encoder.pos_conv_embed.conv.bias"] = sd["encoder.pos_conv.0.bias"]
    m["encoder.pos_conv_embed.conv.parametrizations.weight.original0"] = sd["encoder.pos_conv.0.weight_g"]
    m["encoder.pos_conv_embed.conv.parametrizations.weight.original1"] = sd["encoder.pos_conv.0.weight_v"]
    for l in range(c.layers):
        s, d = f"encoder.layers.{l}.", f"encoder.layers.{l}."
        for p in ("weight", "bias"):
            for proj in ("q_proj", "k_proj", "v_proj", "out_proj"):
                m[d + f"attention.{proj}.{p}"] = sd[s + f"self_attn.{proj}.{p}"]
            m[d + f"layer_norm.{p}"] = sd[s + f"self_attn_layer_norm.{p}"]
            m[d + f"feed_forward.intermediate_dense.{p}"] = sd[s + f"fc1.{p}"]
            m[d + f"feed_forward.output_dense.{p}"] = sd[s + f"fc2.{p}"]
            m[d + f"final_layer_norm.{p}"] = sd[s + f"final_layer_norm.{p}"]
    missing, unexpected = hf.load_state_dict(m, strict=False)
    missing = [k for k in missing if "masked_spec_embed" not in k]
    assert not missing and not unexpected, (missing, unexpected)
    return hf


@torch.no_grad()
def crosscheck_hf(cfg: TrunkConfig, batch: int = 2, samples: int = 64600) -> dict:
    trunk = seeded_init_(Wav2Vec2Trunk(cfg), 1234).eval()
    hf = hf_model_from_oracle(trunk)
    x = synth_clips(0, batch, samples)
    o = trunk(x)
    h = hf(x, output_hidden_states=True)
    rep = {"x_maxabs": float((o["x"] - h.last_hidden_state).abs().max())}
    errs = []
    for i in range(cfg.layers - 1):   # hidden_states[i+1] = raw output of layer i for i < L-1
        errs.append(float((o["layer_results"][i][0].transpose(0, 1) - h.hidden_states[i + 1]).abs().max()))
    rep["layer_maxabs"] = max(errs) if errs else 0.0
    rep["x_scale"] = float(o["x"].abs().mean())
    return rep


@torch.no_grad()
def crosscheck_reference(cfg: TrunkConfig, batch: int = 2, samples: int = 64600) -> dict:
    from . import fairseq_stub
    rep = {}
    x = synth_clips(0, batch, samples)
    # ---- H-SAE: /root/reference/model.py verbatim --------------------------------------------
    ref_model = fairseq_stub.import_reference("model")
    fairseq_stub.set_next_trunk(cfg=cfg)
    ref = ref_model.Model(None, "cpu").eval()
    mine = OracleModel(head="sae", trunk_cfg=cfg).eval()
    seeded_init_(mine, 1234)
    ref.load_state_dict(mine.state_dict(), strict=True)   # same names, same shapes -> strict
    a = ref(x, return_sae_loss=False)
    b = mine(x)
    rep["sae_logprob_maxabs"] = float((a - b).abs().max())
    a2, loss_a = ref(x, return_sae_loss=True)
    b2, loss_b = mine(x, return_sae_loss=True)
    rep["sae_loss_abs"] = abs(float(loss_a) - float(loss_b))
    # ---- H-WIN: /root/reference/model_window_topk.py verbatim --------------------------------
    ref_win = fairseq_stub.import_reference("model_window_topk")
    fairseq_stub.set_next_trunk(cfg=cfg)
    refw = ref_win.Model(None, "cpu").eval()
    minew = OracleModel(head="window", sae_window_size=8, trunk_cfg=cfg).eval()
    seeded_init_(minew, 1234)
    refw.load_state_dict(minew.state_dict(), strict=True)
    rep["win_logprob_maxabs_vs_ref_impl_defined_ties"] = float((refw(x, return_sae_loss=False) - minew(x)).abs().max())
    # tie-free input (all activations > 0, T = 8 + 4n so every frame is covered): reference is deterministic
    g = torch.Generator().manual_seed(7)
    acts = torch.rand(2, 32, 512, generator=g) + 0.01
    sae_ref = ref_win.AutoEncoderTopK(128, 512, k=64, window_size=8)   # test_overlapping_windows.py:17-27 shape
    rep["win_topk_tiefree_maxabs"] = float((sae_ref._window_topk(acts, 64, 8) - window_topk(acts, 64, 8)).abs().max())
    # ---- getAttenF: /root/reference/model_backup.py:186-202 verbatim --------------------------
    ref_bk = fairseq_stub.import_reference("model_backup")
    sls = OracleModel(head="sls", trunk_cfg=cfg).eval()
    seeded_init_(sls, 1234)
    lr = sls.trunk(x)["layer_results"]
    y_ref, full_ref = ref_bk.getAttenF(lr)
    w = torch.sigmoid(sls.fc0(y_ref))
    fused_ref = (full_ref * w.view(w.shape[0], w.shape[1], w.shape[2], -1)).sum(1)
    # same quantity through the oracle head's own code path
    taps = {}
    out = sls(x, taps=taps)
    stack = torch.stack([t.transpose(0, 1) for t in taps["layer_results"]], 1)
    w2 = torch.sigmoid(sls.fc0(stack.mean(2)))
    rep["sls_getAttenF_fused_maxabs"] = float(((stack * w2.unsqueeze(-1)).sum(1) - fused_ref).abs().max())
    rep["sls_logprob_finite"] = bool(torch.isfinite(out).all())
    return rep


@torch.no_grad()
def make_fixture(head: str, batch: int = 2, seed: int = 1234) -> dict:
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    m = OracleModel(head=head, sae_window_size=8).eval()
    seeded_init_(m, seed)
    x = synth_clips(0, batch)
    taps = {}
    t0 = time.time()
    out = m(x, taps=taps)
    dt = time.time() - t0
    fx = {"logprob": out.numpy(), "seed": np.int64(seed), "batch": np.int64(batch), "oracle_seconds": np.float64(dt),
          "x_tap": taps["x"][:, ::25, ::64].contiguous().numpy()}
    for i in (0, 5, 11, 17, 23):
        fx[f"layer{i}_tap"] = taps["layer_results"][i].transpose(0, 1)[:, ::25, ::64].contiguous().numpy()
    if "pooled" in taps:
        fx["pooled_tap"] = taps["pooled"][:, ::16].contiguous().numpy()
        fx["nnz_per_frame"] = (taps["encoded"] > 0).sum(-1).numpy().astype(np.int32)
    return fx


def make_eer_fixture() -> dict:
    """EER known-answer cases: produced by the REFERENCE's eval_metrics_DF.compute_eer where /root/reference is mounted
    (asserted equal to oracle.eer), else by oracle.eer alone.  float32 scores so the fixture is exact on any host."""
    from oracle.eer import compute_eer
    ref = None
    if os.path.isdir("/root/reference"):
        sys.path.insert(0, "/root/reference")
        import eval_metrics_DF as ref
    rs = np.random.RandomState(7)
    cases = {}
    for name, (nt, nn, q) in {"gauss": (500, 4000, 0), "ties": (300, 900, 8), "tiny": (3, 5, 0), "separable": (50, 70, 0)}.items():
        t, n = rs.randn(nt) + 1.0, rs.randn(nn)
        if name == "separable":
            t += 10
        if q:
            t, n = np.round(t * q) / q, np.round(n * q) / q
        t, n = t.astype(np.float32), n.astype(np.float32)
        e = compute_eer(t.astype(np.float64), n.astype(np.float64))
        if ref is not None:
            e_ref = ref.compute_eer(t.astype(np.float64), n.astype(np.float64))
            assert e_ref[0] == e[0] and float(e_ref[1]) == e[1], (name, e_ref, e)
        cases[name + "_t"], cases[name + "_n"], cases[name + "_eer"] = t, n, np.array(e)
    return cases


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "eer_cases.npz"), **make_eer_fixture())
    small = TrunkConfig(layers=2)
    print("[pin] HF cross-check (2 layers):", crosscheck_hf(small))
    print("[pin] HF cross-check (24 layers):", crosscheck_hf(TrunkConfig(), batch=1))
    if os.path.isdir("/root/reference"):
        print("[pin] reference heads (2-layer trunk):", crosscheck_reference(small))
    for head in ("sae", "window", "sls"):
        fx = make_fixture(head)
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"xlsr300m_{head}_b2.npz"), **fx)
        print(f"[fixture] {head}: logprob={fx['logprob'].tolist()} ({fx['oracle_seconds']:.2f}s)")


if __name__ == "__main__":
    sys.exit(main())
