"""End-to-end parity of the CUDA scoring path against the CPU oracle and the committed golden fixtures.

Gates: log-probs within 1e-4 abs in fp32 mode, 2e-2 abs in bf16 mode, identical argmax;
intermediate tensors are compared too, because random-init logits are weakly sensitive.
Everything is called through the product API (``Model.forward`` -> C ABI); the oracle is only the checker.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TOL = {"fp32": 1e-4, "bf16": 2e-2}


def _oracle(head):
    from oracle.heads import OracleModel
    from oracle.trunk import seeded_init_
    m = OracleModel(head=head, sae_window_size=8).eval()
    seeded_init_(m, 1234)
    return m


def _product(sls, head, oracle_model, precision):
    if head == "sls":
        m = sls.ModelSLS(None, "cuda", cp_path=None, precision=precision)
    elif head == "window":
        m = sls.ModelWindowTopK(None, "cuda", cp_path=None, precision=precision)
    else:
        m = sls.Model(None, "cuda", cp_path=None, precision=precision)
    missing, unexpected = m.load_state_dict(oracle_model.state_dict(), strict=False)
    assert not unexpected and all(k.startswith("ssl_model.model.quantizer") for k in missing), (missing, unexpected)
    return m.to("cuda").eval()


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.fixture(scope="module")
def clips():
    from oracle.trunk import synth_clips
    return synth_clips(0, 2)


@pytest.fixture(scope="module", params=["sae", "window", "sls"])
def case(request, sls, cuda, clips):
    head = request.param
    om = _oracle(head)
    taps = {}
    with torch.no_grad():
        ref = om(clips, taps=taps)
    return {"head": head, "oracle": om, "ref": ref, "taps": taps}


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_logprob_parity_and_taps(sls, cuda, clips, case, precision):
    head, ref, taps = case["head"], case["ref"], case["taps"]
    m = _product(sls, head, case["oracle"], precision)
    x = clips.to(cuda)
    with torch.no_grad():
        if head != "sls":
            fast = m(x, return_sae_loss=False).cpu()       # scoring forward: nothing retained, fused select + pool
            m.retain_intermediates = True                  # the taps below read layer results / SAE codes of the forward
        out = m(x) if head == "sls" else m(x, return_sae_loss=False)
    out = out.cpu()
    if head != "sls":
        assert torch.equal(out, fast)                      # retaining intermediates does not change a single bit of the score
    eng = m.engine()
    B, T, D = 2, 201, 1024
    layer_rel = []
    for i in (0, 5, 11, 17, 23):
        got = eng.get_tensor(f"layer_results.{i}", (B, T, D)).cpu()
        layer_rel.append((i, _rel(got, taps["layer_results"][i].transpose(0, 1))))
    x_rel = _rel(eng.get_tensor("x", (B, T, D)).cpu(), taps["x"])
    err = float((out - ref).abs().max())
    print(f"[{head}/{precision}] logprob max|err|={err:.3e} out={out.tolist()} ref={ref.tolist()} x_rel={x_rel:.3e} layer_rel={layer_rel}")
    assert torch.isfinite(out).all()
    assert err <= TOL[precision], f"{head}/{precision}: {err} > {TOL[precision]}"
    assert torch.equal(out.argmax(-1), ref.argmax(-1))
    lim = 5e-6 if precision == "fp32" else 1.5e-2              # 2 x the observed 2.3e-6 / 7.5e-3 (profiles/r02 parity logs)
    assert x_rel <= lim and all(r <= lim for _, r in layer_rel), (x_rel, layer_rel)
    # committed golden fixture (minted in the build container by oracle/make_golden.py)
    fx = np.load(os.path.join(GOLDEN, f"xlsr300m_{head}_b2.npz"))
    assert float(np.abs(out.numpy() - fx["logprob"]).max()) <= TOL[precision]
    got_tap = eng.get_tensor("x", (B, T, D)).cpu()[:, ::25, ::64].numpy()
    tap_err = float(np.abs(got_tap - fx["x_tap"]).max())
    print(f"[{head}/{precision}] x_tap max|err|={tap_err:.3e}")
    assert tap_err <= (2e-5 if precision == "fp32" else 4e-2)          # observed 6.7e-6 / 1.84e-2 on every box of round 2
    if head != "sls" and precision == "fp32":
        pooled = eng.get_tensor("pooled", (B, 4096)).cpu()
        assert float((pooled - taps["pooled"]).abs().max()) <= 1e-4
        enc = eng.get_tensor("encoded", (B, T, 4096)).cpu()
        nnz = (enc > 0).sum(-1)
        nnz_diff = int((nnz != (taps["encoded"] > 0).sum(-1)).sum())
        print(f"[{head}/{precision}] frames whose kept count differs from the oracle: {nnz_diff}; pooled max|err|={float((pooled - taps['pooled']).abs().max()):.3e}")
        assert nnz_diff == 0                                                # selection agrees frame by frame (observed 0: the fp32 path is bit-deterministic)
        if head == "sae":
            assert int(nnz.max()) <= 128


def test_api_surface_and_bit_stability(sls, cuda, clips):
    om = _oracle("sae")
    m = _product(sls, "sae", om, "bf16")
    x = clips.to(cuda)
    with torch.no_grad():
        a = m(x, return_sae_loss=False)
        b = m(x, return_sae_loss=False)
        assert torch.equal(a, b)                                            # bit-stable run to run
        x4 = torch.cat([x, x.flip(0)], 0)
        c = m(x4, return_sae_loss=False)
        assert torch.equal(c[:2], a) and torch.equal(c[2:], a.flip(0))      # batch composition does not change a score
        out, loss = m(x)                                                    # default arity: 2-tuple (model.py:258-259)
        _, loss_ref = om(clips, return_sae_loss=True)
        assert abs(float(loss) - float(loss_ref)) <= 2e-2 * max(1.0, abs(float(loss_ref)))
        out3 = m(x, return_sae_loss=True, return_interpretability=True)
        assert len(out3) == 3 and set(out3[2]) >= {"avg_activation", "top20_features", "sparsity", "sparse_features"}
        assert m.last_sparse_features.shape == (2, 201, 4096)
        feats = m.ssl_model.extract_feat(x.unsqueeze(-1))                   # [B, S, 1] accepted (model.py:134-137)
        assert feats.shape == (2, 201, 1024) and m.ssl_model.out_dim == 1024
        enc = m.sae.encode(feats.reshape(-1, 1024))
        assert torch.equal(enc.reshape(2, 201, -1), m.last_sparse_features)  # same kernels, same bits as the forward's
        assert int((enc > 0).sum(-1).max()) <= 128
        rec = m.sae.decode(enc)
        assert rec.shape == (402, 1024)
        # end-to-end host API equals the device path
        scores = m.engine().score_host(clips.pin_memory(), sls.HEAD_SAE, sls.PREC_BF16)
        assert torch.equal(scores, torch.exp(a[:, 1]).cpu())
    assert isinstance(m, torch.nn.Module) and sum(p.numel() for p in m.parameters()) > 3e8
    assert m.compute_total_loss(torch.tensor(1.0), torch.tensor(2.0)) == pytest.approx(1.2)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_variable_length_padding_mask(sls, cuda, precision):
    """BASELINE config 4: right-padded clips + key-padding mask (wav2vec2.py:567-586), vs the oracle run with
    ``padding_mask`` and vs the oracle run on the un-padded clip alone."""
    from oracle.trunk import synth_clips
    om = _oracle("sae")
    m = _product(sls, "sae", om, precision)
    lens = [40000, 16000, 31234]
    S = 40000
    x = synth_clips(10, 3, S)
    pm = torch.zeros(3, S, dtype=torch.bool)
    for i, n in enumerate(lens):
        x[i, n:] = 0
        pm[i, n:] = True
    with torch.no_grad():
        ref = om(x, padding_mask=pm)
        alone = om(x[1:2, :lens[1]])
        out = m(x.to(cuda), return_sae_loss=False, sample_lengths=torch.tensor(lens)).cpu()
    print(f"[varlen/{precision}] out={out.tolist()} ref={ref.tolist()} alone={alone.tolist()}")
    assert float((ref[1] - alone[0]).abs().max()) <= 1e-4          # the oracle's own consistency
    err = float((out - ref).abs().max())
    assert err <= TOL[precision]
    # identical argmax wherever the oracle's own margin is not a tie at this precision (margin > 2 x observed error);
    # utterance 0 of this batch has an oracle margin of 1.5e-3, below bf16 resolution of the path
    margin = (ref[:, 0] - ref[:, 1]).abs()
    decided = margin > 2 * err
    assert torch.equal(out.argmax(-1)[decided], ref.argmax(-1)[decided]) and int(decided.sum()) >= 2


LONG_LENS = [80000, 120000, 160000, 96000, 140000, 110000, 155000, 88000, 159999, 81234, 131072, 100000]   # 5 ... 10 s


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_long_clips_5_to_10_s_parity(sls, cuda, precision):
    """BASELINE config 4 at its stated upper range: 24-layer model, clips of 5, 7.5 and 10 s (T = 249 ... 499 frames) and
    lengths in between, bucketed by frame count, right-padded inside a bucket with a key-padding mask (wav2vec2.py:567-586).
    bf16: attention runs on the tcgen05 kernel in its wide geometry (T > 256); fp32: CUDA-core verification path.
    Gate: |log-prob - oracle(clip alone, un-padded)| <= 1e-4 fp32 / 2e-2 bf16, identical argmax on every margin-decided clip
    (>= 8 of the 12 must be decided: margin > 2 x the observed error)."""
    from oracle.trunk import synth_clips
    om = _oracle("sae")
    m = _product(sls, "sae", om, precision)
    eng = m.engine()
    clips = [synth_clips(300 + i, 1, n)[0] for i, n in enumerate(LONG_LENS)]
    assert eng.frames(80000) == 249 and eng.frames(120000) == 374 and eng.frames(160000) == 499
    with torch.no_grad():
        ref = torch.cat([om(c[None]) for c in clips])
    got = torch.empty_like(ref)
    batches = sls.bucket_by_frames(LONG_LENS, eng.frames, bucket_frames=64, max_batch=4)
    assert sorted(i for b in batches for i in b) == list(range(len(clips)))
    for batch in batches:
        S = LONG_LENS[batch[0]]
        wav = torch.zeros(len(batch), S)
        for j, i in enumerate(batch):
            wav[j, :LONG_LENS[i]] = clips[i]
        with torch.no_grad():
            out = m(wav.to(cuda), return_sae_loss=False, sample_lengths=torch.tensor([LONG_LENS[i] for i in batch]))
        got[torch.tensor(batch)] = out.cpu()
    err = float((got - ref).abs().max())
    margin = (ref[:, 0] - ref[:, 1]).abs()
    decided = margin > 2 * err
    print(f"[long clips/{precision}] max|logprob err|={err:.3e} decided={int(decided.sum())}/{len(clips)} margins={margin.tolist()}")
    assert torch.isfinite(got).all() and err <= TOL[precision]
    assert int(decided.sum()) >= 8 and torch.equal(got.argmax(-1)[decided], ref.argmax(-1)[decided])


def test_score_file_roundtrip(sls, cuda, tmp_path):
    om = _oracle("sae")
    m = _product(sls, "sae", om, "bf16")
    ds = sls.SyntheticEvalSet(7)
    path = str(tmp_path / "scores.txt")
    sls.produce_evaluation_file(ds, m, "cuda", path, batch_size=4)
    rows = [l.split(" ") for l in open(path).read().splitlines()]
    assert len(rows) == 7 and all(len(r) == 2 for r in rows)
    assert [r[0] for r in rows] == [f"SYN_{i:07d}" for i in range(7)]
    dev = sls.score_synthetic_shard(m, 0, 7, batch=4).cpu()
    assert np.allclose([float(r[1]) for r in rows], dev.numpy(), rtol=0, atol=0)   # host-buffer path == device-resident path
