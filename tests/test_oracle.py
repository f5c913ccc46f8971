"""CPU tests of the oracle itself: against the committed golden fixtures, against an independent wav2vec2
implementation (HF transformers) and, when /root/reference is present, against the reference's own head code."""
import os

import numpy as np
import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_synth_clips_known_answer():
    from oracle.trunk import synth_clips
    x = synth_clips(0, 2, 64600)
    assert x.shape == (2, 64600) and x.dtype == torch.float32
    # known-answer values of the integer-hash generator (machine independent by construction)
    assert abs(float(x.std()) - 1.0) < 0.02 and abs(float(x.mean())) < 0.02
    again = synth_clips(1, 1, 64600)
    assert torch.equal(again[0], x[1])


@pytest.mark.parametrize("head", ["sae", "window", "sls"])
def test_oracle_reproduces_golden(head):
    from oracle.heads import OracleModel
    from oracle.trunk import seeded_init_, synth_clips
    fx = np.load(os.path.join(GOLDEN, f"xlsr300m_{head}_b2.npz"))
    m = OracleModel(head=head, sae_window_size=8).eval()
    seeded_init_(m, int(fx["seed"]))
    taps = {}
    with torch.no_grad():
        out = m(synth_clips(0, int(fx["batch"])), taps=taps)
    assert np.abs(out.numpy() - fx["logprob"]).max() <= 2e-5       # fp32, different BLAS blocking across hosts
    assert np.abs(taps["x"][:, ::25, ::64].numpy() - fx["x_tap"]).max() <= 2e-4
    for i in (0, 5, 11, 17, 23):
        got = taps["layer_results"][i].transpose(0, 1)[:, ::25, ::64].numpy()
        assert np.abs(got - fx[f"layer{i}_tap"]).max() <= 2e-3 * max(1.0, np.abs(fx[f"layer{i}_tap"]).max())


def test_oracle_trunk_matches_hf_wav2vec2():
    from oracle.make_golden import crosscheck_hf
    from oracle.trunk import TrunkConfig
    rep = crosscheck_hf(TrunkConfig(layers=2), batch=1, samples=16000)
    assert rep["x_maxabs"] <= 1e-5 and rep["layer_maxabs"] <= 1e-5, rep


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference checkout not present on this box")
def test_oracle_heads_match_reference_code():
    from oracle.make_golden import crosscheck_reference
    from oracle.trunk import TrunkConfig
    rep = crosscheck_reference(TrunkConfig(layers=1), batch=2, samples=64600)
    assert rep["sae_logprob_maxabs"] <= 1e-6 and rep["sae_loss_abs"] <= 1e-6, rep
    assert rep["win_topk_tiefree_maxabs"] == 0.0, rep
    assert rep["win_logprob_maxabs_vs_ref_impl_defined_ties"] <= 5e-3, rep     # reference ties are implementation-defined
    assert rep["sls_getAttenF_fused_maxabs"] <= 1e-5 and rep["sls_logprob_finite"], rep


def test_canonical_topk_and_window_rule():
    from oracle.heads import canonical_topk_mask, window_topk
    x = torch.tensor([[1.0, 3.0, 3.0, 0.0, 3.0, 2.0]])
    assert canonical_topk_mask(x, 2).tolist() == [[0, 1, 1, 0, 0, 0]]      # ties -> lowest index
    g = torch.Generator().manual_seed(0)
    a = torch.relu(torch.randn(2, 21, 64, generator=g))
    out = window_topk(a, 8, 8)
    assert out.shape == a.shape and int((out > 0).sum(-1).max()) <= 8
    assert torch.equal(out * (out > 0), out) and torch.all((out == 0) | (out == a))


def test_eer_matches_reference_code_and_golden():
    """oracle.eer vs the reference's own eval_metrics_DF.compute_eer (where mounted) and the committed known answers."""
    import os, sys
    from oracle.eer import compute_eer
    fx = np.load(os.path.join(os.path.dirname(__file__), "golden", "eer_cases.npz"))
    names = sorted({k[:-2] for k in fx.files if k.endswith("_t")})
    assert names == ["gauss", "separable", "ties", "tiny"]
    ref = None
    if os.path.isdir("/root/reference"):
        sys.path.insert(0, "/root/reference")
        import eval_metrics_DF as ref
    for n in names:
        t, s = fx[n + "_t"].astype(np.float64), fx[n + "_n"].astype(np.float64)
        e = compute_eer(t, s)
        assert e[0] == fx[n + "_eer"][0] and e[1] == fx[n + "_eer"][1], n
        if ref is not None:
            e_ref = ref.compute_eer(t, s)
            assert float(e_ref[0]) == e[0] and float(e_ref[1]) == e[1], n
    assert fx["separable_eer"][0] == 0.0
