"""BASELINE configs 3-5 as parity-test cases on the GPU (small sizes + size-independent properties):
config 3 utterance-sharded scoring is bit-identical for every sharding, config 4 length-bucketed variable-length clips
match the oracle run one utterance at a time, config 5 the window-TopK head under sharding; plus the device EER."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _small(sls, head, precision="bf16", layers=2):
    from oracle.heads import OracleModel
    from oracle.trunk import TrunkConfig, seeded_init_
    om = OracleModel(head=head, trunk_cfg=TrunkConfig(layers=layers), sae_window_size=8).eval()
    seeded_init_(om, 4321)
    cls = {"sls": sls.ModelSLS, "sae": sls.Model, "window": sls.ModelWindowTopK}[head]
    m = cls(None, "cuda", cp_path=None, precision=precision, geometry=sls.TrunkGeometry(layers=layers))
    m.load_state_dict(om.state_dict(), strict=False)
    return om, m.to("cuda").eval()


@pytest.mark.parametrize("head", ["sls", "window"])
def test_sharded_scoring_is_bit_identical_for_every_sharding(sls, cuda, head):
    """configs 3 / 5: scores keyed by utterance index do not depend on world size, shard boundaries or batch size."""
    _, m = _small(sls, head)
    n = 37
    whole = sls.score_synthetic_shard(m, 0, n, batch=16)
    for world in (2, 3, 8):
        parts = [sls.score_synthetic_shard(m, *sls.shard_range(n, r, world), batch=5) for r in range(world)]
        assert torch.equal(torch.cat(parts), whole), (head, world)
    host = torch.stack([torch.from_numpy(sls.synth_clip_host(i)) for i in (0, 17, 36)])
    dev = torch.cat([m.engine().synth_clips(i, 1) for i in (0, 17, 36)]).cpu()
    assert torch.equal(host, dev)                                      # any rank can regenerate any clip, bit-exactly
    assert torch.isfinite(whole).all() and float(whole.min()) > 0 and float(whole.max()) < 1


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 2e-2)])
def test_variable_length_buckets_match_per_utterance_oracle(sls, cuda, precision, tol):
    """config 4: clips of 1-4 s, bucketed by frame count, right-padded + masked, vs the oracle on each clip alone."""
    from oracle.trunk import synth_clips
    om, m = _small(sls, "sae", precision)
    lens = [16000, 16400, 23456, 33333, 40000, 47999, 52000, 64000, 64600]
    clips = [synth_clips(100 + i, 1, n)[0] for i, n in enumerate(lens)]
    got = sls.score_variable_length(m, clips, bucket_frames=64, max_batch=4)
    with torch.no_grad():
        ref = torch.stack([torch.exp(om(c[None])[0, 1]) for c in clips])
    err = float((got - ref).abs().max())
    print(f"[varlen buckets/{precision}] max|score err|={err:.3e} got={got.tolist()} ref={ref.tolist()}")
    assert err <= tol
    again = sls.score_variable_length(m, clips[::-1], bucket_frames=64, max_batch=4)
    assert torch.equal(again.flip(0), got)                             # order / batch composition does not change a score


def test_eer_on_device_matches_oracle(sls, cuda):
    from oracle.eer import compute_eer
    g = torch.Generator().manual_seed(5)
    scores = torch.rand(20011, generator=g)
    labels = torch.rand(20011, generator=g) < 0.1
    scores[labels] += 0.15
    got = sls.compute_eer(scores.to(cuda), labels.to(cuda))
    want = compute_eer(scores[labels].double().numpy(), scores[~labels].double().numpy())
    assert got == want


def test_score_sharded_tool_single_gpu(sls, cuda, tmp_path):
    out = str(tmp_path / "score.txt")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "score_sharded.py"), "--utts", "70", "--layers", "2", "--batch", "32",
                        "--verify", "--out", out], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    rep = json.loads(r.stdout.strip().splitlines()[-1])
    assert rep["verify_bit_identical"] is True and rep["n"] == 70 and 0.0 <= rep["eer"] <= 1.0
    utts, vals = sls.read_score_file(out)
    assert len(utts) == 70 and utts[69] == "SYN_0000069" and np.isfinite(vals).all()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs on one node")
def test_two_rank_nccl_gather_equals_single_rank(sls, cuda, tmp_path):
    """config 3 on 2 GPUs: torchrun, NCCL all-gather; the gathered vector equals the 1-rank run bit for bit."""
    tool = os.path.join(ROOT, "tools", "score_sharded.py")
    one = subprocess.run([sys.executable, tool, "--utts", "50", "--layers", "2", "--batch", "16"], capture_output=True, text=True, timeout=600)
    two = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", tool, "--utts", "50", "--layers", "2", "--batch", "16", "--verify"],
                         capture_output=True, text=True, timeout=900)
    assert one.returncode == 0 and two.returncode == 0, one.stderr + two.stderr
    a = json.loads(one.stdout.strip().splitlines()[-1])
    b = json.loads([l for l in two.stdout.strip().splitlines() if l.startswith("{")][-1])
    assert a["checksum"] == b["checksum"] and a["first"] == b["first"] and b["verify_bit_identical"] is True and b["world"] == 2


def test_pcm16_ingest_is_bit_exact_with_reference_pad(sls, cuda):
    """next row N2: int16 -> float32 / 32768 + pad() on the device == numpy astype/divide + the reference's pad, bit for bit
    (ragged clips: shorter than, equal to and longer than 64 600 samples, odd lengths, a 1-sample clip)."""
    rs = np.random.RandomState(11)
    lens = [1, 7, 16000, 32301, 64599, 64600, 64601, 100003]
    clips = [rs.randint(-32768, 32768, size=n).astype(np.int16) for n in lens]
    pcm = torch.from_numpy(np.concatenate(clips)).to(cuda)
    off = torch.tensor(np.concatenate(([0], np.cumsum(lens)[:-1])), dtype=torch.int64, device=cuda)
    ln = torch.tensor(lens, dtype=torch.int32, device=cuda)
    _, m = _small(sls, "sae")
    eng = m.engine()
    for S in (64600, 1000, 3):
        got = eng.ingest_pcm16(pcm, off, ln, S).cpu().numpy()
        want = np.stack([sls.pad_clip(c.astype(np.float32) / np.float32(32768.0), S) for c in clips])
        assert got.dtype == np.float32 and np.array_equal(got, want), S
    # end to end from int16 host clips == the float path on the same padded clips
    host16 = [torch.from_numpy(c) for c in clips[2:6]]
    a = eng.score_pcm16_host(host16, sls.HEAD_SAE, sls.PREC_BF16)
    wav = torch.from_numpy(np.stack([sls.pad_clip(c.astype(np.float32) / np.float32(32768.0), 64600) for c in clips[2:6]])).pin_memory()
    b = eng.score_host(wav, sls.HEAD_SAE, sls.PREC_BF16)
    assert torch.equal(a, b)
    with pytest.raises(sls.SlsbError):
        eng.score_pcm16_host([torch.zeros(0, dtype=torch.int16)], sls.HEAD_SAE, sls.PREC_BF16)


@pytest.mark.parametrize("head", ["sae", "window"])
def test_sparse_code_is_the_compact_form_of_the_dense_code(sls, cuda, head):
    """next row N4: (indices, values) emitted by slsb_get_sparse scatter back to exactly the dense encoded tensor."""
    from oracle.trunk import synth_clips
    _, m = _small(sls, head, "fp32")
    x = synth_clips(3, 2).to(cuda)
    with torch.no_grad():
        out = m(x, return_sae_loss=False, return_interpretability=True)
    dense = m.last_sparse_features
    idx, val, cnt = m.last_sparse_code(2, x.shape[1])
    assert idx.shape == (2, 201, 128) and idx.dtype == torch.int32
    valid = idx >= 0
    assert torch.equal(valid.sum(-1).to(torch.int32), cnt) and int(cnt.max()) <= 128
    scat = torch.zeros_like(dense)
    scat.view(-1, dense.shape[-1]).scatter_add_(1, idx.view(-1, 128).clamp(min=0).long(), (val * valid).view(-1, 128))
    assert torch.equal(scat, dense)
    srt = torch.where(valid, idx, torch.full_like(idx, 1 << 30))
    assert bool((srt[..., 1:] >= srt[..., :-1]).all())                          # ascending feature index


def test_full_size_batch64_properties(sls, cuda):
    """BASELINE config 2 at its full size (24 layers, B = 64 x 64 600 samples, bf16, SLS head), through size-independent
    properties: the first two clips reproduce the committed golden log-probs (oracle, 2e-2), every clip's score is bit-identical
    to what it gets in a batch of 2 and under a permutation of the batch, and the scores are proper probabilities."""
    from oracle.heads import OracleModel
    from oracle.trunk import seeded_init_
    om = OracleModel(head="sls").eval()
    seeded_init_(om, 1234)
    m = sls.ModelSLS(None, "cuda", cp_path=None, precision="bf16")
    m.load_state_dict(om.state_dict(), strict=False)
    m = m.to("cuda").eval()
    eng = m.engine()
    wav = eng.synth_clips(0, 64)
    with torch.no_grad():
        full = m(wav)
        pair = m(wav[:2].contiguous())
        tail = m(wav[62:].contiguous())
        perm = torch.randperm(64, generator=torch.Generator().manual_seed(3)).to(cuda)
        shuffled = m(wav[perm].contiguous())
    fx = np.load(os.path.join(ROOT, "tests", "golden", "xlsr300m_sls_b2.npz"))
    gerr = float(np.abs(full[:2].cpu().numpy() - fx["logprob"]).max())
    assert gerr <= 2e-2, f"golden log-probs: {gerr}"
    assert torch.equal(full[:2], pair), f"batch of 2 differs from the same clips in the batch of 64: {full[:2].tolist()} vs {pair.tolist()}"
    assert torch.equal(full[62:], tail), f"tail pair differs: {full[62:].tolist()} vs {tail.tolist()}"
    moved = (shuffled != full[perm]).any(-1).nonzero().flatten().tolist()
    assert not moved, f"permuted batch: clips at positions {moved} changed, max |d| = {float((shuffled - full[perm]).abs().max()):.3e}"
    p = torch.exp(full)
    assert torch.isfinite(full).all() and float((p.sum(-1) - 1).abs().max()) < 1e-5


def test_edge_cases_empty_short_and_minimal_clips(sls, cuda):
    """Empty batch / clips shorter than the conv stack's receptive field fail loudly (the reference dies inside conv1d
    there); the shortest legal clips (1, 2, 9 frames) and odd batch sizes match the oracle; [B, S, 1] input is accepted
    (model.py:130-134)."""
    from oracle.trunk import synth_clips
    om, m = _small(sls, "sae", "fp32")
    with pytest.raises(sls.SlsbError, match="empty batch"):
        m(torch.zeros(0, 64600, device=cuda), return_sae_loss=False)
    with pytest.raises(sls.SlsbError, match="too short"):
        m(torch.zeros(2, 399, device=cuda), return_sae_loss=False)
    for samples, batch in ((400, 1), (720, 3), (3200, 5)):             # 1, 2 and 9 frames
        clips = synth_clips(500, batch, samples)
        with torch.no_grad():
            ref = om(clips)
            got = m(clips.to(cuda), return_sae_loss=False).cpu()
            got3 = m(clips.to(cuda).unsqueeze(-1), return_sae_loss=False).cpu()
        err = float((got - ref).abs().max())
        print(f"[minimal clips] S={samples} B={batch} frames={m.engine().frames(samples)} max|err|={err:.2e}")
        assert err <= 1e-4 and torch.equal(got, got3)
    _, mb = _small(sls, "sae", "bf16")
    clips = synth_clips(600, 7, 3200)
    with torch.no_grad():
        err = float((mb(clips.to(cuda), return_sae_loss=False).cpu() - om(clips)).abs().max())
    assert err <= 2e-2


def test_score_pcm_shard_equals_float_host_path(sls, cuda, tmp_path):
    """next row N2 end to end: WAV files -> PCM shard -> score_pcm_shard (2 bytes / sample uploaded, pad() on the device) gives
    bit for bit the scores of the reference-shaped path (host float32 conversion + pad, then slsb_score_host), for any batch
    size and for a rank's sub-range of the shard."""
    rs = np.random.RandomState(3)
    lens = [16000, 64600, 70001, 33000, 5, 64599, 129200]
    clips = [(rs.randn(n) * 3000).clip(-32768, 32767).astype(np.int16) for n in lens]
    paths = []
    for i, c in enumerate(clips):
        paths.append(str(tmp_path / f"c{i}.wav"))
        sls.write_wav_pcm16(paths[-1], c)
    sls.wav_files_to_shard(str(tmp_path / "shard"), [f"u{i}" for i in range(len(clips))], paths, workers=2)
    shard = sls.PcmShard(str(tmp_path / "shard"))
    _, m = _small(sls, "sls")
    got = sls.score_pcm_shard(m, shard, batch=3)
    wav = torch.from_numpy(np.stack([sls.pad_clip(c.astype(np.float32) / np.float32(32768.0), 64600) for c in clips])).pin_memory()
    want = m.engine().score_host(wav, sls.HEAD_SLS, sls.PREC_BF16)
    assert torch.equal(got, want.cpu())
    assert torch.equal(sls.score_pcm_shard(m, shard, batch=64), got)
    assert torch.equal(sls.score_pcm_shard(m, shard, batch=2, lo=2, hi=6), got[2:6])
    assert torch.isfinite(got).all() and float(got.min()) > 0 and float(got.max()) < 1


def test_score_files_tool_end_to_end(sls, cuda, tmp_path):
    """A1 -> A18 on files: FLAC corpus + trial list -> tools/score_files.py -> score.txt in protocol order with the scores of the
    in-process path (same seed-1234 random-init 2-layer model), and an EER against a trial_metadata-style key file."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import flac_enc
    rs = np.random.RandomState(4)
    utts = [f"DF_E_{3000000 + i}" for i in range(6)]
    clips = [(rs.randn(n) * 2500).astype(np.int16) for n in (64600, 70000, 8000, 32000, 64601, 500)]
    os.makedirs(tmp_path / "flac")
    for u, c in zip(utts, clips):
        with open(tmp_path / "flac" / f"{u}.flac", "wb") as f:
            f.write(flac_enc.encode(c.astype(np.int64), kind="fixed2", porder=2, rate=16000))
    (tmp_path / "trials.txt").write_text("\n".join(utts) + "\n")
    (tmp_path / "keys.txt").write_text("".join(f"LA_00{i} {u} nocodec asvspoof A14 {'bonafide' if i % 2 else 'spoof'} notrim eval\n"
                                               for i, u in enumerate(utts)))
    out = tmp_path / "score.txt"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "score_files.py"), "--protocol", str(tmp_path / "trials.txt"),
                        "--audio-dir", str(tmp_path), "--head", "sls", "--layers", "2", "--batch", "4", "--out", str(out),
                        "--keys", str(tmp_path / "keys.txt")], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, RANK="0", WORLD_SIZE="1", LOCAL_RANK="0"))
    assert r.returncode == 0, r.stderr[-2000:]
    rec = json.loads(r.stdout.strip().splitlines()[-1])
    ids, scores = sls.read_score_file(str(out))
    assert ids == utts and rec["trials"] == 6 and rec["scored_with_keys"] == 6 and 0.0 <= rec["eer"] <= 1.0
    torch.manual_seed(1234)
    m = sls.ModelSLS(None, cuda, cp_path=None, precision="bf16", geometry=sls.TrunkGeometry(layers=2)).to(cuda).eval()
    want = m.engine().score_pcm16_host([torch.from_numpy(c) for c in clips], sls.HEAD_SLS, sls.PREC_BF16)
    assert np.array_equal(scores.astype(np.float32), want.numpy())           # repr(float) round-trips float32 exactly


def _tiny_checkpoints(sls, tmp_path, layers=2, window=False):
    """A fairseq-style trunk checkpoint (cfg + model tensors) and a main.py-style trained checkpoint (module.-prefixed state_dict)."""
    from oracle.heads import OracleModel
    from oracle.trunk import TrunkConfig, seeded_init_
    om = OracleModel(head="window" if window else "sae", trunk_cfg=TrunkConfig(layers=layers), sae_window_size=8).eval()
    seeded_init_(om, 777)
    sd = om.state_dict()
    cp = str(tmp_path / "xlsr_like.pt")
    torch.save({"cfg": {"model": {"encoder_layers": layers}}, "model": {k[len("ssl_model.model."):]: v for k, v in sd.items()
                                                                    if k.startswith("ssl_model.model.")}}, cp)
    best = str(tmp_path / "best.pth")
    torch.save({"module." + k: v for k, v in sd.items()}, best)
    return om, cp, best


@pytest.mark.parametrize("window", [False, True])
def test_main_eval_replay_writes_the_reference_score_file(sls, cuda, tmp_path, window):
    """VERDICT r1 missing #5 (GPU half): the evaluation branch of main.py (:630-653) replayed step by step by tools/main_eval.py
    - same flags, Model(...) keyword arguments, nn.DataParallel wrap, checkpoint load, genSpoof_list, Dataset_ASVspoof2021_eval
    over FLAC files, stale-file removal, produce_evaluation_file with batch 20 - ends in a score file whose rows are
    "{utt} {repr(float)}", in protocol order, equal to exp(oracle log-prob[:, 1]) within the bf16 gate and bit-equal to a direct
    forward of the same model.  (The reference's own main.py runs verbatim on the shim up to the first forward in
    tests/test_host.py::test_reference_main_py_runs_on_the_shim; this box has no /root/reference.)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import flac_enc
    om, cp, best = _tiny_checkpoints(sls, tmp_path, window=window)
    rs = np.random.RandomState(21)
    n = 23                                                             # two batches of 20 + 3: the last one is ragged
    utts = [f"DF_E_{2000011 + 3 * i}" for i in range(n)]
    lens = rs.randint(20000, 90000, size=n)
    clips = [(3000 * np.sin(np.arange(m) * (0.01 + 0.002 * i)) + rs.randn(m) * 500).astype(np.int16) for i, m in enumerate(lens)]
    os.makedirs(tmp_path / "db" / "flac")
    for u, c in zip(utts, clips):
        (tmp_path / "db" / "flac" / f"{u}.flac").write_bytes(flac_enc.encode(c.astype(np.int64), kind="fixed2", porder=2, rate=16000))
    (tmp_path / "trl.txt").write_text("".join(u + "\n" for u in utts))
    out = tmp_path / "scores" / "scores_DF.txt"
    os.makedirs(out.parent)
    out.write_text("stale line\n")
    cmd = [sys.executable, os.path.join(ROOT, "tools", "main_eval.py"), "--is_eval", "--track", "DF", "--cp_path", cp, "--model_path", best,
           "--database_path", str(tmp_path / "db"), "--protocols_path", str(tmp_path / "trl.txt"), "--eval_output", str(out),
           "--batch_size", "14", "--num_epochs", "100"] + (["--use_window_topk"] if window else [])
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "Total parameters:" in r.stdout and f"Scores saved to: {out}" in r.stdout
    rows = [ln.split(" ") for ln in out.read_text().splitlines()]
    assert [r_[0] for r_ in rows] == utts and all(len(r_) == 2 for r_ in rows)
    got = np.array([float(r_[1]) for r_ in rows])
    assert all(repr(float(np.float32(v))) == r_[1] or repr(v) == r_[1] for v, r_ in zip(got, rows))       # Python float repr, main.py:190-192
    x = torch.from_numpy(np.stack([sls.pad_clip(c.astype(np.float32) / np.float32(32768.0), 64600) for c in clips]))
    with torch.no_grad():
        ref = torch.exp(om(x)[:, 1]).numpy()
    assert np.abs(got - ref).max() <= 2e-2, np.abs(got - ref).max()
    cls = sls.ModelWindowTopK if window else sls.Model
    m = cls(None, "cuda", cp_path=cp)
    sls.load_model_checkpoint(m, best)
    m = m.to("cuda").eval()
    with torch.no_grad():
        direct = torch.cat([torch.exp(m(x[i:i + 20].to(cuda), return_sae_loss=False)[:, 1]) for i in (0, 20)]).cpu().numpy()
    assert np.array_equal(got.astype(np.float32), direct)


def test_checkpoint_roundtrip_on_the_gpu(sls, cuda, tmp_path):
    """VERDICT r1 missing #6 / row N3: load_model_checkpoint(path) -> forward equals the forward of the same state_dict loaded in
    memory, bit for bit: bare state_dict, ``module.``-prefixed (main.py:542-560), wrapped in {'model_state_dict': ...} (:531-536),
    and the strict -> non-strict fallback when a key is missing (:586-592); the trunk checkpoint (model.py:113-115) likewise."""
    from oracle.trunk import synth_clips
    om, cp, best = _tiny_checkpoints(sls, tmp_path)
    sd = om.state_dict()
    x = synth_clips(40, 3).to(cuda)
    mem = sls.Model(None, "cuda", cp_path=None, geometry=sls.TrunkGeometry(layers=2))
    mem.load_state_dict(sd, strict=False)
    mem = mem.to(cuda).eval()
    with torch.no_grad():
        want = mem(x, return_sae_loss=False)
        ref = om(x.cpu())
    assert float((want.cpu() - ref).abs().max()) <= 2e-2
    files = {"bare": sd, "prefixed": {"module." + k: v for k, v in sd.items()},
             "full": {"model_state_dict": sd, "optimizer_state_dict": {"state": {}}, "epoch": 7, "best_val_eer": 1.5}}
    for name, obj in files.items():
        p = str(tmp_path / f"{name}.pth")
        torch.save(obj, p)
        for wrapped in (False, True):
            m = sls.Model(None, "cuda", cp_path=cp)                     # geometry + trunk from the fairseq-style file
            target = torch.nn.DataParallel(m) if wrapped else m
            res = sls.load_model_checkpoint(target, p)
            assert not res.unexpected_keys and all(k.replace("module.", "").startswith("ssl_model.model.quantizer") or "project_q" in k or "final_proj" in k
                                                   or "mask_emb" in k for k in res.missing_keys), (name, res)
            m = m.to(cuda).eval()
            with torch.no_grad():
                got = (target if wrapped else m)(x, return_sae_loss=False)
            assert torch.equal(got, want), (name, wrapped)
    # trunk-only load (cp_path) gives the trunk of the trained model: identical features
    m = sls.Model(None, "cuda", cp_path=cp).to(cuda).eval()
    with torch.no_grad():
        assert torch.equal(m.ssl_model.extract_feat(x), mem.ssl_model.extract_feat(x))
    # a checkpoint that lacks a classifier tensor: strict load fails, the non-strict retry keeps the model's own value
    part = {k: v for k, v in sd.items() if k != "classifier.4.bias"}
    p = str(tmp_path / "partial.pth")
    torch.save(part, p)
    m = sls.Model(None, "cuda", cp_path=cp)
    keep = m.classifier[4].bias.detach().clone()
    res = sls.load_model_checkpoint(m, p)
    assert "classifier.4.bias" in res.missing_keys and torch.equal(m.classifier[4].bias.detach(), keep)
    mem.classifier[4].bias.data.copy_(keep)
    mem.refresh_weights()
    m = m.to(cuda).eval()
    with torch.no_grad():
        assert torch.equal(m(x, return_sae_loss=False), mem(x, return_sae_loss=False))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs on one node")
def test_engine_runs_on_its_own_device_and_restores_the_callers(sls, cuda):
    """an engine on cuda:1 while torch's current device is cuda:0 - every C-ABI entry switches to the engine's device
    and hands the caller's device back; same bits as the engine on cuda:0."""
    _, m0 = _small(sls, "sae")
    x = m0.engine().synth_clips(0, 2)
    with torch.no_grad():
        a = m0(x, return_sae_loss=False)
    assert torch.cuda.current_device() == 0
    _, m1 = _small(sls, "sae")
    m1 = m1.to("cuda:1")
    with torch.no_grad():
        b = m1(x.to("cuda:1"), return_sae_loss=False)
    assert torch.cuda.current_device() == 0 and b.device.index == 1
    assert torch.equal(a.cpu(), b.cpu())
    with torch.no_grad():
        assert torch.equal(m0(x, return_sae_loss=False), a)             # the cuda:0 engine still works afterwards


def test_device_flac_decoder_is_bit_exact_and_scores_like_the_host_path(sls, cuda, tmp_path):
    """next row N2 on the device: one GPU thread per FLAC frame (csrc/flac_gpu.cu) == the host decoder (csrc/flac_decode.cpp), sample for
    sample, over every subframe type / Rice setting / block size; FLAC files -> scores with the decode on the device == the host-decode
    pipeline bit for bit; a batch with a stream the device decoder does not take falls back to the host decoder."""
    import ctypes as C
    import flac_enc
    from helpers import P, stream
    lib = sls.load_library()
    rs = np.random.RandomState(12)
    scans, refs = [], []
    kinds = ["verbatim", "fixed0", "fixed1", "fixed2", "fixed3", "fixed4", "lpc1", "lpc4", "lpc8", "lpc12", "lpc13", "lpc32"]
    for i in range(48):
        n = int(rs.randint(1, 70000))
        amp = float(rs.choice([3, 300, 12000]))
        x = np.clip(amp * np.sin(np.arange(n) * rs.uniform(0.001, 1.0)) + rs.randn(n) * amp * rs.uniform(0, 0.5), -32768, 32767).astype(np.int64)
        if i == 0:
            x = np.where((np.arange(n) // 50) % 2 == 0, 32767, -32768)                        # full-scale square wave: long Rice code words
        if i == 1:
            x = (x >> 3) << 3                                                              # wasted bits
        data = flac_enc.encode(x, kind=kinds[i % len(kinds)], porder=int(rs.randint(0, 5)), method=int(rs.randint(0, 2)),
                               blocksize=int(rs.choice([192, 576, 1024, 4096, 4608, 333])), escape_partitions=(1,) if i % 7 == 3 else ())
        scans.append(sls.scan_flac_bytes(data, 64600, sample_rate=None))
        refs.append(sls.decode_flac_bytes(data, 64600, sample_rate=None))
        assert np.array_equal(refs[-1], x[:64600].astype(np.int16))
    d, fr, tot, off, lens = sls.pack_flac_batch(scans, 64600)
    host_pcm, host_st = sls.decode_flac_frames_host(d, fr, tot)
    bytes_dev = torch.zeros(d.size + 8, dtype=torch.uint8, device=cuda)
    bytes_dev[:d.size] = torch.from_numpy(d).to(cuda)
    frames_dev = torch.from_numpy(fr.view(np.uint8).reshape(-1).copy()).to(cuda)
    pcm_dev = torch.zeros(tot, dtype=torch.int16, device=cuda)
    st_dev = torch.zeros(len(fr), dtype=torch.int32, device=cuda)
    assert lib.slsb_flac_decode_frames(P(bytes_dev), P(frames_dev), len(fr), P(pcm_dev), P(st_dev), stream()) == 0
    torch.cuda.synchronize()
    assert np.array_equal(st_dev.cpu().numpy(), host_st) and (host_st > 0).all()
    got = pcm_dev.cpu().numpy()
    assert np.array_equal(got, host_pcm)
    for o, n, r in zip(off.tolist(), lens.tolist(), refs):
        assert np.array_equal(got[o:o + n], r)
    # files -> scores: decode on the device == decode on the host, bit for bit; a stereo file makes its batch take the host path
    _, m = _small(sls, "sae")
    paths = []
    clips = [(rs.randn(n) * 2000).astype(np.int64) for n in (70000, 12345, 64600, 300, 40000, 64601, 9, 20000, 33333)]
    for i, c in enumerate(clips):
        paths.append(str(tmp_path / f"c{i}.flac"))
        with open(paths[-1], "wb") as f:
            f.write(flac_enc.encode(c, kind="lpc8" if i % 2 else "fixed2", porder=2, rate=16000))
    ref_scores = sls.score_audio_files(m, paths, batch=4, workers=2)
    stats = {}
    dev_scores = sls.score_flac_files_device(m, paths, batch=4, workers=2, stats=stats)
    assert torch.equal(dev_scores, ref_scores) and stats["device_batches"] == 3 and stats["host_batches"] == 0
    assert 0 < stats["flac_bytes"] < stats["pcm_bytes"]
    stereo = str(tmp_path / "st.flac")
    with open(stereo, "wb") as f:
        f.write(flac_enc.encode(np.stack([clips[1], clips[1] // 2], 1), kind="fixed2", stereo=10, rate=16000))
    mixed = paths[:3] + [stereo] + paths[3:5]
    stats = {}
    mixed_scores = sls.score_flac_files_device(m, mixed, batch=4, workers=2, stats=stats)
    assert torch.equal(mixed_scores, sls.score_audio_files(m, mixed, batch=4, workers=2)) and stats["host_batches"] == 1 and stats["device_batches"] == 1
