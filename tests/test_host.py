"""CPU tests of the host side: C-ABI library loads and exports what include/slsb200.h declares, the drop-in
``Model`` has the reference's surface and state_dict keys, packing shapes, score-file format, sharding and the
world_size-2 gather (gloo).  No compute entry point is called here (there is no GPU in the build container)."""
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(sls, lib):
    header = open(os.path.join(ROOT, "include", "slsb200.h")).read()
    declared = set(re.findall(r"\b(slsb_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(sls.EXPORTED_SYMBOLS), declared ^ set(sls.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.slsb_abi_version() == 1


def test_no_cpu_fallback(sls):
    m = sls.Model(None, "cpu", cp_path=None, geometry=sls.TrunkGeometry(layers=1))
    with pytest.raises(sls.SlsbError):
        m(torch.zeros(1, 64600))
    if not torch.cuda.is_available():
        import ctypes as C
        from importlib import import_module
        cfg = sls.make_config(sls.TrunkGeometry(layers=1))
        h = C.c_void_p()
        rc = sls.load_library().slsb_create(C.byref(cfg), 0, C.byref(h))
        assert rc != 0 and b"no CUDA device" in sls.load_library().slsb_last_error()


def test_missing_checkpoint_raises_like_reference(sls):
    with pytest.raises(RuntimeError):
        sls.Model(None, "cpu", cp_path="/nonexistent/xlsr2_300m.pt", geometry=sls.TrunkGeometry(layers=1))


def test_state_dict_keys_match_reference_naming(sls):
    m = sls.Model(None, "cpu", cp_path=None, geometry=sls.TrunkGeometry(layers=2))
    keys = set(m.state_dict())
    for k in ["ssl_model.model.feature_extractor.conv_layers.0.0.weight", "ssl_model.model.feature_extractor.conv_layers.6.2.1.bias",
              "ssl_model.model.post_extract_proj.weight", "ssl_model.model.mask_emb", "ssl_model.model.layer_norm.weight",
              "ssl_model.model.encoder.pos_conv.0.weight_g", "ssl_model.model.encoder.pos_conv.0.weight_v",
              "ssl_model.model.encoder.pos_conv.0.bias", "ssl_model.model.encoder.layers.1.self_attn.q_proj.weight",
              "ssl_model.model.encoder.layers.1.self_attn.out_proj.bias", "ssl_model.model.encoder.layers.0.fc1.weight",
              "ssl_model.model.encoder.layers.0.final_layer_norm.weight", "ssl_model.model.encoder.layer_norm.bias",
              "sae.k", "sae.encoder.weight", "sae.encoder.bias", "sae.decoder.weight", "sae.b_dec",
              "classifier.0.weight", "classifier.1.weight", "classifier.4.bias"]:
        assert k in keys, k
    assert m.state_dict()["sae.encoder.weight"].shape == (4096, 1024)
    assert m.state_dict()["ssl_model.model.encoder.pos_conv.0.weight_g"].shape == (1, 1, 128)
    # oracle (== reference naming, verified against /root/reference/model.py by strict load in oracle/make_golden.py)
    from oracle.heads import OracleModel
    from oracle.trunk import TrunkConfig
    o = OracleModel(head="sae", trunk_cfg=TrunkConfig(layers=2))
    missing, unexpected = m.load_state_dict(o.state_dict(), strict=False)
    assert not unexpected and all("quantizer" in k for k in missing)
    # module. prefix round trip (main.py:542-560)
    wrapped = torch.nn.DataParallel(m) if torch.cuda.is_available() else None
    sd = {"module." + k: v for k, v in m.state_dict().items()}
    assert all(k.startswith("module.ssl_model") or k.startswith("module.sae") or k.startswith("module.classifier") for k in sd)


def test_pack_shapes_and_folds(sls):
    geo = sls.TrunkGeometry(layers=1)
    m = sls.Model(None, "cpu", cp_path=None, geometry=geo)
    packed = sls.pack_state_dict(m.state_dict(), geo)
    assert packed["conv0.w"].shape == (512, 10) and packed["conv1.w"].shape == (512, 1536) and packed["conv6.w"].shape == (512, 1024)
    assert packed["pos.w"].shape == (1024, 128 * 64) and packed["L0.qkv.w"].shape == (3072, 1024)
    sd = m.state_dict()
    # conv weights are tap-major: packed[o, t*C + c] == w[o, c, t]
    w = sd["ssl_model.model.feature_extractor.conv_layers.2.0.weight"]
    assert torch.equal(packed["conv2.w"][7, 2 * 512 + 5], w[7, 5, 2])
    # q rows pre-scaled by 64**-0.5, k/v untouched
    q = sd["ssl_model.model.encoder.layers.0.self_attn.q_proj.weight"]
    assert torch.equal(packed["L0.qkv.w"][:1024], q * 0.125)
    assert torch.equal(packed["L0.qkv.w"][1024:2048], sd["ssl_model.model.encoder.layers.0.self_attn.k_proj.weight"])
    # weight-norm fold equals torch's own weight_norm forward
    conv = m.ssl_model.model.encoder.pos_conv[0]
    x = torch.randn(1, 1024, 50)
    ref_w = torch._weight_norm(conv.weight_v, conv.weight_g, 2)
    assert torch.allclose(packed["pos.w"].reshape(1024, 128, 64).permute(0, 2, 1), ref_w, atol=1e-7)
    s = sls.ModelSLS(None, "cpu", cp_path=None, geometry=geo)
    ps = sls.pack_state_dict(s.state_dict(), geo, sls_kp=s._sls_kp())
    assert s.fc1.in_features == 22847 and ps["sls.fc1.w"].shape == (1024, 22848) and ps["sls.bn"].shape == (4,)
    assert float(ps["sls.fc1.w"][:, 22847:].abs().max()) == 0.0


def test_frames_formula(sls):
    geo = sls.TrunkGeometry()
    assert geo.frames(64600) == 201 and geo.frames(16000) == 49 and geo.frames(160000) == 499
    assert geo.frames(400) == 1 and geo.frames(399) == 0 and geo.frames(0) == 0      # receptive field of the conv stack = 400 samples


def test_pad_clip_matches_reference_semantics(sls):
    x = np.arange(5, dtype=np.float32)
    assert sls.pad_clip(x, 12).tolist() == [0, 1, 2, 3, 4, 0, 1, 2, 3, 4, 0, 1]      # data_utils_SSL.py:58-65 tile-repeat
    assert sls.pad_clip(np.arange(20, dtype=np.float32), 12).tolist() == list(range(12))  # truncate head
    assert sls.pad_clip(x, 5).tolist() == x.tolist()
    if os.path.isdir("/root/reference"):
        import importlib.util, types
        src = open("/root/reference/data_utils_SSL.py").read()
        fn = src[src.index("def pad("):src.index("class Dataset_ASVspoof2019_train")]
        ns = {"np": np}
        exec(fn, ns)                                                                  # the reference's own pad()
        for n in (1, 7, 64599, 64600, 70000):
            y = np.random.RandomState(n).randn(n).astype(np.float32)
            assert np.array_equal(ns["pad"](y, 64600), sls.pad_clip(y, 64600))


def test_score_file_format(sls, tmp_path):
    p = str(tmp_path / "s.txt")
    sls.write_score_file(p, ["LA_E_1", "LA_E_2"], [0.5, 1.25e-07])
    assert open(p).read() == "LA_E_1 0.5\nLA_E_2 1.25e-07\n"                      # main.py:190-192: python float repr
    sls.write_score_file(p, ["a"], [0.123456789], fmt="6f")
    assert open(p).read() == "a 0.123457\n"
    import pandas
    sls.write_score_file(p, ["u1", "u2", "u3"], [0.1, 0.2, 0.3])
    df = pandas.read_csv(p, sep=" ", header=None, skipinitialspace=True)           # evaluate_2021_DF.py:24
    assert df.shape == (3, 2)


def test_compute_eer_equals_oracle_on_golden(sls):
    """scoring.compute_eer (torch, runs on the scores' device) == the reference algorithm, bit for bit, ties included."""
    from oracle.eer import compute_eer as eer_ref
    fx = np.load(os.path.join(ROOT, "tests", "golden", "eer_cases.npz"))
    for n in ("gauss", "ties", "tiny", "separable"):
        t, s = fx[n + "_t"], fx[n + "_n"]
        perm = np.random.RandomState(1).permutation(t.size + s.size)            # protocol order is arbitrary
        scores = torch.from_numpy(np.concatenate((t, s))[perm])
        labels = torch.from_numpy((np.arange(t.size + s.size) < t.size)[perm])
        got = sls.compute_eer(scores, labels)
        # targets-first STABLE order inside each class is what the reference sees after its own boolean masks
        want = eer_ref(np.concatenate((t, s))[perm][labels.numpy()].astype(np.float64), np.concatenate((t, s))[perm][~labels.numpy()].astype(np.float64))
        assert got == want, (n, got, want)
    with pytest.raises(ValueError):
        sls.compute_eer(torch.zeros(3), torch.ones(3, dtype=torch.bool))


def test_bucket_by_frames_partitions_and_bounds(sls):
    """BASELINE config 4 host logic: every index once, one 64-frame bucket per batch, longest first, <= max_batch."""
    rs = np.random.RandomState(3)
    lens = rs.randint(16000, 160001, size=500).tolist()
    batches = sls.bucket_by_frames(lens, _frames, bucket_frames=64, max_batch=32)
    flat = [i for b in batches for i in b]
    assert sorted(flat) == list(range(500))
    for b in batches:
        assert 1 <= len(b) <= 32
        fr = [_frames(lens[i]) for i in b]
        assert len({(f - 1) // 64 for f in fr}) == 1
        assert [lens[i] for i in b] == sorted((lens[i] for i in b), reverse=True)
    assert sls.bucket_by_frames([], _frames) == []


def test_batch_by_length_partitions_and_bounds(sls):
    """Length-sorted batching (config 4): every index once, longest first inside and across batches, padding of a batch below
    max_pad_frames, size <= max_batch (up to 4 x for short clips when min_rows asks for it), one ragged batch at most per closing rule."""
    rs = np.random.RandomState(5)
    lens = rs.randint(16000, 160001, size=700).tolist()
    for min_rows in (None, 12864):
        batches = sls.batch_by_length(lens, _frames, max_pad_frames=64, max_batch=32, min_rows=min_rows)
        flat = [i for b in batches for i in b]
        assert sorted(flat) == list(range(700))
        assert [lens[i] for i in flat] == sorted(lens, reverse=True)
        for b in batches:
            fr = [_frames(lens[i]) for i in b]
            assert fr[0] - fr[-1] < 64
            cap = 32 if not min_rows else max(32, min(128, min_rows // fr[0]))
            assert 1 <= len(b) <= cap
        if min_rows:
            assert max(len(b) for b in batches) > 32                       # one-second clips really get the larger batches
    assert sls.batch_by_length([], _frames) == []
    # three clips of very different lengths are not padded into one batch
    assert [len(b) for b in sls.batch_by_length([160000, 16000, 80000], _frames)] == [1, 1, 1]


def _frames(n):
    for k, s in [(10, 5), (3, 2), (3, 2), (3, 2), (3, 2), (2, 2), (2, 2)]:
        n = (n - k) // s + 1
    return n


def test_checkpoint_loads_without_fairseq_or_omegaconf(sls, tmp_path):
    """next row N3: a fairseq-style checkpoint whose ``cfg`` pickles classes of modules that are NOT importable here
    (omegaconf / fairseq stand-ins) still yields its tensors; ``module.``-prefixed main.py checkpoints load into Model."""
    import pickle, types
    fake = types.ModuleType("omegaconf_standin.dictconfig")
    class DictConfig:                                   # pickled by reference: module + qualname
        def __init__(self, d): self.d = d
    DictConfig.__module__, DictConfig.__qualname__ = "omegaconf_standin.dictconfig", "DictConfig"
    fake.DictConfig = DictConfig
    sys.modules["omegaconf_standin"] = types.ModuleType("omegaconf_standin")
    sys.modules["omegaconf_standin.dictconfig"] = fake
    geo = sls.TrunkGeometry(layers=1)
    trunk = sls.TrunkParams(geo)
    sd = {k: torch.randn_like(v) for k, v in trunk.state_dict().items()}
    path = str(tmp_path / "xlsr_like.pt")
    torch.save({"args": None, "cfg": DictConfig({"model": {"encoder_layers": 1}}), "model": sd, "extra_state": {"epoch": 3}}, path)
    del sys.modules["omegaconf_standin"], sys.modules["omegaconf_standin.dictconfig"]
    with pytest.raises(Exception):
        torch.load(path, map_location="cpu", weights_only=False)                      # what a plain load does without the module
    got = sls.load_checkpoint_tensors(path)
    assert set(got) == set(sd) and all(torch.equal(got[k], sd[k]) for k in sd)
    ssl = sls.SSLModel("cpu", cp_path=path, geometry=geo)                              # model.py:113-115 replacement
    assert torch.equal(ssl.model.state_dict()["post_extract_proj.weight"], sd["post_extract_proj.weight"])
    # main.py:542-560 checkpoints: bare state_dict with the DataParallel prefix
    m = sls.Model(None, "cpu", cp_path=None, geometry=geo)
    full = {"module." + k: torch.randn_like(v) if v.is_floating_point() else v.clone() for k, v in m.state_dict().items()}
    p2 = str(tmp_path / "best.pth")
    torch.save(full, p2)
    res = sls.load_model_checkpoint(m, p2)                                            # un-wrapped model: prefix stripped
    assert not res.missing_keys and not res.unexpected_keys
    assert torch.equal(m.state_dict()["classifier.1.weight"], full["module.classifier.1.weight"])
    dp = torch.nn.DataParallel(sls.Model(None, "cpu", cp_path=None, geometry=geo))    # main.py:518 wrapping: prefix kept
    res = sls.load_model_checkpoint(dp, p2)
    assert not res.missing_keys and torch.equal(dp.module.state_dict()["sae.b_dec"], full["module.sae.b_dec"])
    assert sls.fix_module_prefix({"a": 1}, True) == {"module.a": 1} and sls.fix_module_prefix({"module.a": 1}, False) == {"a": 1}
    with pytest.raises(RuntimeError):
        torch.save({"note": "no tensors"}, p2); sls.load_checkpoint_tensors(p2)


class _Boom:
    """Pickles as a call of an arbitrary global: ``__reduce__`` -> (callable, args)."""
    def __init__(self, fn, args): self.fn, self.args = fn, args
    def __reduce__(self): return self.fn, self.args


def test_checkpoint_unpickler_resolves_only_the_allowlist(sls, tmp_path, monkeypatch):
    """the loader must not execute what a crafted checkpoint names.  Each payload would run code under a plain
    ``torch.load(weights_only=False)`` (and under a root-module allowlist: all of them live below builtins / os / torch /
    functools); here every one becomes an inert stub, the marker file is never written, and the tensors still load."""
    import functools, os as _os, subprocess
    from importlib import import_module
    weights = import_module("slsforasvspoof-2021-df_b200.weights")
    marker = tmp_path / "pwned"
    cmd = f"touch {marker}"
    payloads = {
        "os.system": _Boom(_os.system, (cmd,)),
        "builtins.eval": _Boom(eval, (f"__import__('os').system({cmd!r})",)),
        "builtins.exec": _Boom(exec, (f"import os; os.system({cmd!r})",)),
        "builtins.getattr": _Boom(getattr, ("abc", "upper")),
        "builtins.__import__": _Boom(__import__, ("os",)),
        "subprocess.check_output": _Boom(subprocess.check_output, (["touch", str(marker)],)),
        "functools.partial": _Boom(functools.partial, (_os.system, cmd)),
        "torch.load": _Boom(torch.load, (str(marker),)),
        "torch.storage._load_from_bytes": _Boom(torch.storage._load_from_bytes, (b"x",)),
        "torch.hub.load": _Boom(torch.hub.load, ("a/b", "c")),
    }
    good = {"w": torch.arange(6.0).reshape(2, 3), "n": torch.tensor([1, 2], dtype=torch.int64), "h": torch.ones(2, dtype=torch.bfloat16)}
    for name, boom in payloads.items():
        path = str(tmp_path / "evil.pt")
        torch.save({"model": good, "cfg": boom, "extra": [boom, {"k": boom}]}, path)
        got = sls.load_checkpoint_tensors(path)
        assert not marker.exists(), name
        assert set(got) == set(good) and all(torch.equal(got[k], good[k]) for k in good), name
    # the allowlist is exact pairs, not module roots
    assert weights._is_safe_global("collections", "OrderedDict") and weights._is_safe_global("torch", "float32")
    assert weights._is_safe_global("torch", "FloatStorage") and weights._is_safe_global("numpy", "float64")
    for mod, nm in [("builtins", "eval"), ("builtins", "exec"), ("builtins", "getattr"), ("builtins", "__import__"), ("os", "system"),
                    ("torch.utils.cpp_extension", "load"), ("torch", "load"), ("torch.storage", "_load_from_bytes"), ("functools", "partial"),
                    ("copyreg", "__newobj__x"), ("numpy", "load"), ("torch", "hub"), ("torch", "ops"), ("numpy", "ndarray.tofile")]:
        assert not weights._is_safe_global(mod, nm), (mod, nm)
    # numpy payloads that fairseq checkpoints really carry (optimizer history, extra_state) still load
    path = str(tmp_path / "np.pt")
    torch.save({"model": good, "extra_state": {"best": np.float64(0.25), "hist": np.arange(4), "ns": __import__("argparse").Namespace(lr=[1e-3])}}, path)
    assert set(sls.load_checkpoint_tensors(path)) == set(good)


def test_checkpoint_unpickler_survives_corrupted_files(sls, tmp_path):
    """Mutation fuzz (like the FLAC decoder's): a byte-flipped or truncated checkpoint either still yields tensors or raises
    an ordinary exception - no crash, no hang, nothing executed."""
    import zipfile
    geo = sls.TrunkGeometry(layers=1)
    sd = {k: v for k, v in list(sls.TrunkParams(geo).state_dict().items())[:6]}
    src = tmp_path / "ok.pt"
    torch.save({"model": sd, "cfg": {"a": [1, 2, 3]}}, str(src))
    blob = src.read_bytes()
    with zipfile.ZipFile(str(src)) as z:
        pkl = [n for n in z.namelist() if n.endswith("data.pkl")][0]
        info = z.getinfo(pkl)
    lo = blob.find(b"data.pkl") + 8                                          # mutate inside / around the pickle stream
    rs = np.random.RandomState(3)
    outcomes = {"ok": 0, "raised": 0}
    for trial in range(120):
        b = bytearray(blob)
        if trial % 4 == 3:
            b = b[:rs.randint(16, len(b))]
        else:
            for _ in range(rs.randint(1, 4)):
                b[lo + rs.randint(0, max(1, info.file_size + 64))] ^= 1 << rs.randint(0, 8)
        pth = tmp_path / "mut.pt"
        pth.write_bytes(bytes(b))
        try:
            got = sls.load_checkpoint_tensors(str(pth))
            assert all(torch.is_tensor(v) for v in got.values())
            outcomes["ok"] += 1
        except Exception:
            outcomes["raised"] += 1
    assert outcomes["ok"] + outcomes["raised"] == 120 and outcomes["raised"] > 0


_MAIN_PY_LAUNCHER = r'''
import os, runpy, sys, types
sys.path.insert(0, {root!r})
import sls_b200
# the shim of INTEGRATION.md: the reference imports `model`, `model_window_topk`, `data_utils_SSL`; they resolve to this package
m = types.ModuleType("model"); m.Model = sls_b200.Model; sys.modules["model"] = m
w = types.ModuleType("model_window_topk"); w.Model = sls_b200.ModelWindowTopK; sys.modules["model_window_topk"] = w
d = types.ModuleType("data_utils_SSL")
d.genSpoof_list, d.Dataset_ASVspoof2021_eval, d.Dataset_in_the_wild_eval = sls_b200.genSpoof_list, sls_b200.Dataset_ASVspoof2021_eval, sls_b200.Dataset_in_the_wild_eval
d.Dataset_ASVspoof2019_train = type("Dataset_ASVspoof2019_train", (), {{}})          # training only: never touched by --is_eval
sys.modules["data_utils_SSL"] = d
tb = types.ModuleType("tensorboardX"); tb.SummaryWriter = type("SummaryWriter", (), {{"__init__": lambda self, *a, **k: None}})
sys.modules["tensorboardX"] = tb                                                    # not installed here; training only
sys.path.insert(1, {ref!r})                                                         # core_scripts.startup_config: the reference's own
sys.argv = ["main.py"] + {argv!r}
runpy.run_path(os.path.join({ref!r}, "main.py"), run_name="__main__")
'''


@pytest.mark.skipif(not os.path.exists("/root/reference/main.py"), reason="the reference tree is only mounted in the build container")
@pytest.mark.parametrize("window", [False, True])
def test_reference_main_py_runs_on_the_shim(sls, tmp_path, window):
    """VERDICT r1 missing #5: the reference's OWN main.py, unmodified, executed with --is_eval on top of the shim modules: argument
    parsing, Model(...) with main.py's keyword arguments (:493-517), nn.DataParallel(model).to(device) (:518), the parameter count
    (:520-521), the checkpoint branch (:531-592: torch.load, _get_state_dict, _fix_module_prefix, strict load), genSpoof_list,
    Dataset_ASVspoof2021_eval, the 6-worker DataLoader and the first model(batch_x, return_sae_loss=False) call (:158-181).
    This container has no GPU, so that first forward must end in this package's "no CPU fallback" error and nothing else;
    the same sequence runs to the score file on the GPU box (tests/test_configs_gpu.py::test_main_eval_replay_*)."""
    import flac_enc
    geo = sls.TrunkGeometry(layers=1)
    torch.manual_seed(5)
    trunk = sls.TrunkParams(geo)
    cp = str(tmp_path / "xlsr_like.pt")
    torch.save({"cfg": {"model": {"encoder_layers": 1}}, "model": trunk.state_dict()}, cp)
    cls = sls.ModelWindowTopK if window else sls.Model
    trained = torch.nn.DataParallel(cls(None, "cpu", cp_path=cp))
    assert trained.module.ssl_model.model.geo.layers == 1                              # geometry comes from the checkpoint, as with fairseq
    with torch.no_grad():
        trained.module.classifier[4].bias.add_(0.25)
    best = str(tmp_path / "best.pth")
    torch.save(trained.state_dict(), best)                                            # what main.py's training loop writes (module.-prefixed)
    rs = np.random.RandomState(2)
    utts = [f"DF_E_{2000011 + i}" for i in range(3)]
    os.makedirs(tmp_path / "db" / "flac")
    for u in utts:
        (tmp_path / "db" / "flac" / f"{u}.flac").write_bytes(flac_enc.encode((rs.randn(20000) * 3000).astype(np.int64), kind="fixed2", rate=16000))
    (tmp_path / "trl.txt").write_text("".join(u + "\n" for u in utts))
    out = tmp_path / "scores" / "scores_DF.txt"
    os.makedirs(out.parent)
    out.write_text("stale line\n")
    argv = ["--is_eval", "--track", "DF", "--cp_path", cp, "--model_path", best, "--database_path", str(tmp_path / "db"),
            "--protocols_path", str(tmp_path / "trl.txt"), "--eval_output", str(out)] + (["--use_window_topk"] if window else [])
    script = tmp_path / "launch.py"
    script.write_text(_MAIN_PY_LAUNCHER.format(root=ROOT, ref="/root/reference", argv=argv))
    r = subprocess.run([sys.executable, str(script)], cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
    log = r.stdout + r.stderr
    assert "Total parameters:" in r.stdout and "LOADING CHECKPOINT" in r.stdout and "Loaded weights (fresh optimizer/epoch)" in r.stdout, log[-3000:]
    assert ("Using Window-based TopK" if window else "Using Per-Timestep TopK") in r.stdout
    assert "Strict load failed" not in log                                            # strict=True succeeded: every key matched
    assert r.returncode != 0 and "no CPU fallback" in log, log[-3000:]                 # reached the first forward, and only then stopped
    assert not out.exists() or out.read_text() == ""                                  # the stale score file was removed (main.py:646-647)


def test_conv0_layernorm_statistics_from_the_gram_matrix():
    """The identity csrc/conv0_tc.cu is built on (host arithmetic only, fp32 like the kernel): with a = [x_0..x_9, 1] and
    W' = [w | b], the LayerNorm statistics of the 512 conv0 outputs y_n = a . W'_n are mean = a . s / 512 and
    E[y^2] = a^T G a / 512 (s = column sums, G = W'^T W', both built in double once per weight load).  Cases: noise, a DC offset
    30x the signal, digital silence, a bias with a large common component (mean >> std: the cancellation case)."""
    rs = np.random.RandomState(7)
    C, k = 512, 10
    for kind in ("noise", "dc", "silence", "bias_offset"):
        w = (rs.randn(C, k) * np.sqrt(2.0 / k)).astype(np.float32)
        b = (rs.randn(C) * 0.05 + (5.0 if kind == "bias_offset" else 0.0)).astype(np.float32)
        x = rs.randn(4000, k).astype(np.float32)
        if kind == "dc":
            x = (0.03 * x + 0.9).astype(np.float32)
        elif kind == "silence":
            x[:] = 0
        wp = np.concatenate([w, b[:, None]], 1).astype(np.float64)
        s32 = (wp.sum(0) / C).astype(np.float32)
        g32 = (wp.T @ wp / C).astype(np.float32)
        a = np.concatenate([x, np.ones((len(x), 1), np.float32)], 1)
        mean = np.zeros(len(x), np.float32)
        e2 = np.zeros(len(x), np.float32)
        for i in range(k + 1):                      # the kernel's order: row i of G against a, then a_i times that, fp32 throughout
            t = np.zeros(len(x), np.float32)
            for j in range(k + 1):
                t = (g32[i, j] * a[:, j] + t).astype(np.float32)
            e2 = (a[:, i] * t + e2).astype(np.float32)
            mean = (a[:, i] * s32[i] + mean).astype(np.float32)
        var = np.maximum(e2 - mean * mean, 0).astype(np.float32)
        rstd = 1.0 / np.sqrt(var + np.float32(1e-5))
        y = a.astype(np.float64) @ wp.T             # direct: the 512 outputs, statistics over them in double
        mean_ref, var_ref = y.mean(1), y.var(1)
        rstd_ref = 1.0 / np.sqrt(var_ref + 1e-5)
        assert np.abs(mean - mean_ref).max() <= 1e-6, kind                      # observed <= 3.2e-7
        # what the epilogue applies is (y - mean) * rstd: compare the normalised outputs, not var itself
        z = (y - mean[:, None].astype(np.float64)) * rstd[:, None].astype(np.float64)
        z_ref = (y - mean_ref[:, None]) * rstd_ref[:, None]
        assert np.abs(z - z_ref).max() <= (5e-5 if kind == "bias_offset" else 2e-6), (kind, float(np.abs(z - z_ref).max()))   # observed 1.7e-5 / 6.7e-7


def test_conv0_hi_lo_split_arithmetic():
    """The operand layout csrc/conv0_tc.cu builds in shared memory, restated in numpy: A row = [x_hi | x_hi | x_lo | 1 1 0..],
    W row = [w_hi | w_lo | w_hi | b_hi b_lo 0..] (bf16 parts, fp32 accumulation), so that A . W = x_hi w_hi + x_hi w_lo + x_lo w_hi
    + b_hi + b_lo.  Against the exact fp32 conv + bias the dropped x_lo w_lo term and the bf16 rounding of the lo parts leave a
    relative error of ~2^-16 of the operand magnitudes - three orders of magnitude below the bf16 rounding of the output."""
    bf = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).bfloat16().float().numpy()
    rs = np.random.RandomState(11)
    k, C = 10, 512
    x = rs.randn(3000, k).astype(np.float32) * 0.3
    w = (rs.randn(C, k) * np.sqrt(2.0 / k)).astype(np.float32)
    b = (rs.randn(C) * 0.05).astype(np.float32)
    x_hi = bf(x); x_lo = bf(x - x_hi)
    w_hi = bf(w); w_lo = bf(w - w_hi)
    b_hi = bf(b); b_lo = bf(b - b_hi)
    A = np.zeros((len(x), 64), np.float32)
    A[:, 0:k] = x_hi; A[:, 16:16 + k] = x_hi; A[:, 32:32 + k] = x_lo; A[:, 48] = 1.0; A[:, 49] = 1.0
    W = np.zeros((C, 64), np.float32)
    W[:, 0:k] = w_hi; W[:, 16:16 + k] = w_lo; W[:, 32:32 + k] = w_hi; W[:, 48] = b_hi; W[:, 49] = b_lo
    assert np.array_equal(bf(A), A) and np.array_equal(bf(W), W)          # every operand element is exactly representable in bf16
    got = A.astype(np.float64) @ W.astype(np.float64).T                   # the tensor core accumulates exact bf16 products in fp32
    ref = x.astype(np.float64) @ w.astype(np.float64).T + b.astype(np.float64)
    scale = (np.abs(x).astype(np.float64) @ np.abs(w).astype(np.float64).T + np.abs(b))
    rel = np.abs(got - ref) / scale
    assert rel.max() <= 2.0 ** -15, float(rel.max())                      # observed ~1e-5; bf16 output rounding is 2^-9
    plain = bf(x).astype(np.float64) @ bf(w).astype(np.float64).T + bf(b).astype(np.float64)
    assert (np.abs(plain - ref) / scale).max() > 50 * rel.max()           # what a single-bf16 operand GEMM would give


def test_shard_ranges_cover_exactly(sls):
    for n, w in [(611829, 8), (10, 4), (3, 8), (0, 2)]:
        r = [sls.shard_range(n, k, w) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n and all(r[i][1] == r[i + 1][0] for i in range(w - 1))
        assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


_GLOO_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import sls_b200
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank, n = dist.get_rank(), 11
lo, hi = sls_b200.shard_range(n, rank, 2)
local = torch.arange(lo, hi, dtype=torch.float32) * 0.5       # stands in for this rank's scores
full = sls_b200.gather_scores(local, n, rank, 2)
assert torch.equal(full, torch.arange(n, dtype=torch.float32) * 0.5), full
if rank == 0:
    sls_b200.write_score_file(sys.argv[4], ["u%d" % i for i in range(n)], full.tolist())
dist.barrier(); dist.destroy_process_group()
'''


def test_world_size_2_gather_gloo(tmp_path):
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER)
    out = str(tmp_path / "scores.txt")
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(port), str(r), out]) for r in range(2)]
    assert all(p.wait(timeout=240) == 0 for p in procs)
    lines = open(out).read().splitlines()
    assert len(lines) == 11 and lines[3] == "u3 1.5"


@pytest.mark.parametrize("M,N", [(12864, 3072), (12864, 4096), (12864, 1024), (4864, 1024), (6221, 1024), (300, 256), (201, 1024),
                                 (256 * 74, 256), (256 * 75, 256), (1, 256), (64 * 499, 2048)])
@pytest.mark.parametrize("pairs", [74, 66, 1, 7])
def test_pair_gemm_schedule_covers_every_output_once(lib, M, N, pairs):
    """Host replay of the CTA-pair GEMM's static schedule (full rounds in snake order + a column-sliced partial last round):
    every 256-row x N output block is produced exactly once, slices are 256 / 128 / 64 columns wide and aligned, and no pair
    works more than one (sliced) item longer than another."""
    import ctypes as C
    cap = 4 * ((M + 255) // 256) * (N // 256) + 8
    items = (C.c_int32 * (5 * cap))()
    split = C.c_int32(0)
    n = lib.slsb_debug_pair_schedule(M, N, pairs, items, cap, C.byref(split))
    assert 0 < n <= cap and split.value in (1, 2, 4)
    rows = np.ctypeslib.as_array(items).reshape(cap, 5)[:n]
    m_tiles = (M + 255) // 256
    cover = np.zeros((m_tiles, N // 64), dtype=np.int32)
    for pair, rnd, row0, col0, cols in rows:
        assert cols in (256, 128, 64) and col0 % cols == 0 and row0 % 256 == 0 and col0 + cols <= N
        cover[row0 // 256, col0 // 64:(col0 + cols) // 64] += 1
    assert (cover == 1).all()
    used = min(pairs, m_tiles * (N // 256))
    cost = np.zeros(used)
    for pair, rnd, row0, col0, cols in rows:
        cost[pair] += cols / 256.0
    full = (m_tiles * (N // 256)) // used
    assert cost.min() >= full and cost.max() <= full + 1
    if (m_tiles * (N // 256)) % used:
        assert cost.max() - full <= max(0.25, -(-((m_tiles * (N // 256)) % used * split.value) // used) / split.value) + 1e-9
    # rounds are consecutive per pair (the device roles walk `it` upwards until the schedule ends)
    for p in range(used):
        r = sorted(int(x[1]) for x in rows if x[0] == p)
        assert r == list(range(len(r)))


def test_wav_decode_and_pcm_shard_roundtrip(sls, tmp_path):
    """Ingest host side (next row N2): 16-bit WAV -> int16, stereo -> mono mean, wrong rate / width rejected; a PCM shard
    returns exactly the clips it was built from, and batch() hands out the (pcm, offsets, lens) triple of the C ABI with
    long clips truncated to their head (pad() keeps x[:max_len], data_utils_SSL.py:60-61)."""
    import wave
    rs = np.random.RandomState(5)
    lens = [1, 900, 16000, 64600, 70001, 33]
    clips = [rs.randint(-32768, 32768, size=n).astype(np.int16) for n in lens]
    paths = []
    for i, c in enumerate(clips):
        p = str(tmp_path / f"u{i}.wav")
        sls.write_wav_pcm16(p, c)
        paths.append(p)
        assert np.array_equal(sls.read_wav_pcm16(p), c)
    assert all(np.array_equal(a, b) for a, b in zip(sls.decode_wav_files(paths, workers=3), clips))
    st = str(tmp_path / "stereo.wav")
    l, r = np.array([100, -3, 32767, -32768, 1], np.int16), np.array([101, -4, 32767, -32768, 2], np.int16)
    with wave.open(st, "wb") as w:
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(16000)
        w.writeframes(np.stack([l, r], 1).astype("<i2").tobytes())
    assert sls.read_wav_pcm16(st).tolist() == [101, -4, 32767, -32768, 2]        # mean, halves rounded away from zero
    bad = str(tmp_path / "bad.wav")
    sls.write_wav_pcm16(bad, clips[1], sample_rate=8000)
    with pytest.raises(sls.AudioFormatError):
        sls.read_wav_pcm16(bad)
    with wave.open(bad, "wb") as w:
        w.setnchannels(1); w.setsampwidth(1); w.setframerate(16000); w.writeframes(b"\\x00\\x01")
    with pytest.raises(sls.AudioFormatError):
        sls.read_wav_pcm16(bad)

    d = str(tmp_path / "shard")
    ids = [f"DF_E_{i:07d}" for i in range(len(clips))]
    sls.wav_files_to_shard(d, ids, paths, workers=2)
    sh = sls.PcmShard(d)
    assert len(sh) == len(clips) and sh.utt_ids == ids
    assert all(np.array_equal(sh.clip(i), c) for i, c in enumerate(clips))
    pcm, off, ln = sh.batch(1, 4)
    assert pcm.dtype == np.int16 and off.dtype == np.int64 and ln.dtype == np.int32
    assert ln.tolist() == lens[1:4] and off.tolist() == [0, 900, 16900] and np.array_equal(pcm, np.concatenate(clips[1:4]))
    pcm, off, ln = sh.batch(2, 6, max_samples=64600)
    assert ln.tolist() == [16000, 64600, 64600, 33] and off.tolist() == [0, 16000, 80600, 145200]
    for j, i in enumerate(range(2, 6)):
        assert np.array_equal(pcm[off[j]:off[j] + ln[j]], clips[i][:64600])
    # what the device computes from such a triple == the reference's host path on the same clip
    want = sls.pad_clip(clips[4].astype(np.float32) / np.float32(32768.0), 64600)
    assert np.array_equal(sls.pad_clip(pcm[off[2]:off[2] + ln[2]].astype(np.float32) / np.float32(32768.0), 64600), want)
    with pytest.raises(sls.AudioFormatError):
        sls.write_pcm_shard(str(tmp_path / "e"), ["a"], [np.zeros(0, np.int16)])
    np.save(os.path.join(d, "offsets.npy"), np.array([0, 5, 5], np.int64))
    with pytest.raises(sls.AudioFormatError):
        sls.PcmShard(d)


# ---------------------------------------------------------------------------------------------------------------- FLAC
# The three worked examples of RFC 9639 (appendix D): complete files with their STREAMINFO MD5 - the third-party known
# answers of the native decoder (frame CRC-8 / CRC-16 and the MD5 are verified by the decoder itself on every call).
RFC9639_EXAMPLES = {
    "d1_verbatim_wasted_bits": (
        "664c6143 80000022 1000 1000 00000f 00000f 0ac442f0 00000001 3e84b41807dc690307586a3dad1a2e0f"
        "fff869180000bf 0358fd 03128b aa9a", 44100, 2, 16, [[25588, 10416]]),
    "d2_fixed_rice_side_right_two_frames": (
        "664c6143 00000022 0010 0010 000017 000044 0ac442f0 00000013 d5b0564975e98b8d8b930422757b8103"
        "03000012 0000000000000000 0000000000000000 0010"
        "0400003a 20000000 7265666572656e6365206c6962464c414320312e332e33203230313930383034 01000000 0e000000 5449544c453dd7a9d79cd795d79d"
        "81000006 000000000000"
        "fff86998000f99 1208670162 3d1442998f5df70d6fe00c17caeb21000ee7a77a24a1590c1217b603097b784faa9a33d285e070ad5b1b4851b4010d99d2cd1a68f1e6 b810"
        "fff869180102a4 02c382c40bc14a 03ee48dd03b67c 1330", 44100, 2, 16,
        [[10372, 6070], [18041, 10545], [14942, 8743], [17876, 10449], [15627, 9143], [17899, 10463], [16242, 9502], [18077, 10569],
         [16824, 9840], [18263, 10680], [17295, 10113], [-14418, -8428], [-15201, -8895], [-14508, -8476], [-15195, -8896],
         [-14818, -8653], [-15486, -9072], [-15349, -8958], [-16054, -9410]]),
    "d3_lpc_8bit": (
        "664c6143 80000022 1000 1000 00001f 00001f 07d00070 00000018 f8f9e396f5cbcfc6dc807f9977906b32"
        "fff868020017e9 44004f6f313d1047d227cb6d090831452bdc28222280 57a3", 32000, 1, 8,
        [[v] for v in (0, 79, 111, 78, 8, -61, -90, -68, -13, 42, 67, 53, 13, -27, -46, -38, -12, 14, 24, 19, 6, -4, -5, 0)]),
}


def _flac_decode(lib, data, max_samples=0, verify=1):
    import ctypes as C
    buf = np.frombuffer(data, dtype=np.uint8)
    info = (C.c_int32 * 6)()
    cap = 1 << 21
    out = np.empty(cap, dtype=np.int32)
    n = lib.slsb_flac_decode(buf.ctypes.data, buf.size, max_samples, verify, out.ctypes.data, cap, info)
    ch = max(int(info[1]), 1)
    return int(n), list(info), (out[:n * ch].reshape(n, ch).copy() if n > 0 else None)


@pytest.mark.parametrize("name", sorted(RFC9639_EXAMPLES))
def test_flac_decoder_rfc9639_examples(lib, name):
    hexs, rate, ch, bps, want = RFC9639_EXAMPLES[name]
    data = bytes.fromhex(hexs.replace(" ", ""))
    n, info, pcm = _flac_decode(lib, data)
    assert n == len(want) and info[:4] == [rate, ch, bps, len(want)] and info[4] == 1        # MD5 of STREAMINFO verified
    assert pcm.tolist() == want
    bad = bytearray(data)
    bad[-5] ^= 0x10                                                                            # a flipped payload bit
    assert _flac_decode(lib, bytes(bad))[0] in (-6, -3, -4, -7)
    assert _flac_decode(lib, data[:-3])[0] == -3
    assert _flac_decode(lib, b"RIFF" + data[4:])[0] == -2


def test_flac_decoder_every_syntax_element_against_test_encoder(sls, lib):
    """CONSTANT / VERBATIM / FIXED 0-4 / LPC up to order 32, Rice methods 0 and 1, partition orders, escape partitions, wasted
    bits, all four stereo modes, odd block sizes, multi-byte frame numbers: decode(encode(x)) == x with CRCs and MD5 verified;
    early stop returns the head; corrupted streams are rejected."""
    import flac_enc
    rs = np.random.RandomState(0)
    t = np.arange(20000)
    mono = (8000 * np.sin(t * 0.05) + rs.randn(t.size) * 300).astype(np.int64)
    stereo = np.stack([mono, (mono * 0.8 + rs.randn(t.size) * 100).astype(np.int64)], 1)
    cases = [dict(kind="verbatim"), dict(kind="fixed0"), dict(kind="fixed1", porder=1), dict(kind="fixed2", porder=3),
             dict(kind="fixed3", porder=4, method=1), dict(kind="fixed4", porder=2, escape_partitions=(1, 3)),
             dict(kind="lpc8", porder=2), dict(kind="lpc32", porder=1, blocksize=1152), dict(kind="fixed2", blocksize=777),
             dict(kind="fixed2", blocksize=16, extra_metadata=False), dict(kind="lpc4", with_md5=False)]
    for kw in cases:
        x = mono[:3000] if kw.get("blocksize") == 16 else mono                 # 188 frames: two-byte coded frame numbers
        n, info, pcm = _flac_decode(lib, flac_enc.encode(x, **kw))
        assert n == len(x) and np.array_equal(pcm[:, 0], x), kw
        assert info[4] == (0 if kw.get("with_md5") is False else 1), kw
    for mode in (None, 8, 9, 10):
        n, info, pcm = _flac_decode(lib, flac_enc.encode(stereo, kind="lpc6", stereo=mode, porder=3))
        assert n == len(stereo) and info[4] == 1 and np.array_equal(pcm, stereo), mode
    # extremes: full-scale square wave (constant sub-blocks, large residuals), silence, wasted bits, 24-bit and 8-bit samples
    sq = np.where((t // 50) % 2 == 0, 32767, -32768)
    for x, kw in ((sq, dict(kind="fixed1", porder=2)), (np.zeros(5000, np.int64), dict(kind="constant")),
                  ((mono >> 3) << 3, dict(kind="fixed2")), (mono * 200, dict(kind="lpc8", bps=24)), (mono >> 8, dict(kind="fixed2", bps=8))):
        n, info, pcm = _flac_decode(lib, flac_enc.encode(x, **kw))
        assert n == len(x) and info[4] == 1 and np.array_equal(pcm[:, 0], x), kw
    wide = rs.randint(-2 ** 28, 2 ** 28, size=3000).astype(np.int64) << 3                      # 32-bit samples: independent channels decode,
    n, info, pcm = _flac_decode(lib, flac_enc.encode(wide, bps=32, kind="lpc4", blocksize=1024))   # a 33-bit side channel is refused
    assert n == len(wide) and info[2] == 32 and info[4] == 1 and np.array_equal(pcm[:, 0], wide)
    assert _flac_decode(lib, flac_enc.encode(np.stack([wide, wide + 5], 1), bps=32, kind="verbatim", stereo=10, blocksize=512))[0] == -9
    data = flac_enc.encode(mono, kind="lpc8", porder=2)
    tagged = b"ID3\x04\x00\x00" + bytes([0, 0, 1, 5]) + bytes(133) + data + b"TAG" + bytes(125)       # ID3v2 in front (133 bytes), ID3v1 behind
    n, info, pcm = _flac_decode(lib, tagged)
    assert n == len(mono) and info[4] == 1 and np.array_equal(pcm[:, 0], mono)
    n, info, pcm = _flac_decode(lib, data, max_samples=6460)
    assert n == 6460 and info[4] == 0 and np.array_equal(pcm[:, 0], mono[:6460])            # early stop: head only, MD5 not checked
    bad = bytearray(data)
    bad[len(bad) // 2] ^= 0x01
    assert _flac_decode(lib, bytes(bad))[0] < 0
    md5_bad = bytearray(data)
    md5_bad[4 + 4 + 18] ^= 0xFF                                                                # STREAMINFO MD5 field
    assert _flac_decode(lib, bytes(md5_bad))[0] == -8 and _flac_decode(lib, bytes(md5_bad), verify=0)[0] == len(mono)


def test_flac_unknown_length_stream_with_early_stop_and_threads(sls, lib):
    """(1) STREAMINFO total == 0 ("unknown") + max_samples: the digest covers only the decoded head, so a non-zero
    stored MD5 must NOT raise a spurious MD5 error - neither at a block boundary nor inside a block; the complete decode of the
    same stream still verifies it.  (2) first use of the decoder from many threads at once (CRC tables built race-free)."""
    import concurrent.futures as cf
    import flac_enc
    rs = np.random.RandomState(1)
    x = (6000 * np.sin(np.arange(20000) * 0.03) + rs.randn(20000) * 200).astype(np.int64)
    data = bytearray(flac_enc.encode(x, kind="lpc8", porder=2, blocksize=4096))
    si = 8                                                                # "fLaC" + 4-byte block header, then STREAMINFO
    data[si + 13] &= 0xF0
    data[si + 14:si + 18] = bytes(4)                                      # 36-bit total-samples field -> 0
    data = bytes(data)
    for cut in (4096, 8192, 5000, 1, 19999):
        n, info, pcm = _flac_decode(lib, data, max_samples=cut)
        assert n == cut and info[3] == 0 and info[4] == 0 and np.array_equal(pcm[:, 0], x[:cut]), cut
    n, info, pcm = _flac_decode(lib, data)                                 # decoded to the end: MD5 verified
    assert n == len(x) and info[4] == 1 and np.array_equal(pcm[:, 0], x)
    n, info, pcm = _flac_decode(lib, data, max_samples=30000)              # limit beyond the end: still the whole stream
    assert n == len(x) and info[4] == 1
    bad = bytearray(data)
    bad[si + 18 + 3] ^= 0x55                                               # wrong stored MD5: caught on a complete decode only
    assert _flac_decode(lib, bytes(bad))[0] == -8 and _flac_decode(lib, bytes(bad), max_samples=5000)[0] == 5000
    with cf.ThreadPoolExecutor(max_workers=16) as ex:
        outs = list(ex.map(lambda i: sls.decode_flac_bytes(data, 6000 + i, sample_rate=None), range(64)))
    assert all(np.array_equal(o, x[:6000 + i].astype(np.int16)) for i, o in enumerate(outs))


def test_flac_files_to_shard_through_the_ingest_api(sls, tmp_path):
    """read_flac_pcm16 / decode_audio_files / audio_files_to_shard: mono and stereo 16 kHz FLAC + a WAV, heads of 64 600 samples."""
    import flac_enc
    rs = np.random.RandomState(1)
    clips = [(rs.randn(n) * 2000).astype(np.int16) for n in (70000, 12345, 64600)]
    paths = []
    for i, c in enumerate(clips[:2]):
        paths.append(str(tmp_path / f"a{i}.flac"))
        with open(paths[-1], "wb") as f:
            f.write(flac_enc.encode(c.astype(np.int64), kind="fixed2", porder=2, rate=16000))
    paths.append(str(tmp_path / "a2.wav"))
    sls.write_wav_pcm16(paths[-1], clips[2])
    assert np.array_equal(sls.read_flac_pcm16(paths[0]), clips[0])
    assert np.array_equal(sls.read_flac_pcm16(paths[0], max_samples=64600), clips[0][:64600])
    got = sls.decode_audio_files(paths, workers=3, max_samples=64600)
    assert all(np.array_equal(g, c[:64600]) for g, c in zip(got, clips))
    st = np.stack([clips[1].astype(np.int64), clips[1].astype(np.int64) + 3], 1)
    mono = sls.decode_flac_bytes(flac_enc.encode(st, kind="fixed1", stereo=10))
    tot = st.sum(1)
    assert np.array_equal(mono, (np.sign(tot) * ((2 * np.abs(tot) + 2) // 4)).astype(np.int16))           # (2x + 3) / 2, halves away from zero
    with pytest.raises(sls.AudioFormatError, match="16000 Hz"):
        sls.decode_flac_bytes(flac_enc.encode(clips[1].astype(np.int64), rate=44100))
    with pytest.raises(sls.AudioFormatError, match="16-bit"):
        sls.decode_flac_bytes(flac_enc.encode(clips[1].astype(np.int64) >> 8, bps=8))
    with pytest.raises(sls.AudioFormatError, match="not a FLAC"):
        sls.decode_flac_bytes(b"OggS" + bytes(100))
    sls.audio_files_to_shard(str(tmp_path / "sh"), ["a", "b", "c"], paths, workers=2)
    sh = sls.PcmShard(str(tmp_path / "sh"))
    assert [len(sh.clip(i)) for i in range(3)] == [64600, 12345, 64600] and np.array_equal(sh.clip(1), clips[1])


def test_score_files_tool_decode_stage(sls, tmp_path):
    """tools/score_files.py --shard-only (no GPU): trial list + <dir>/flac/<utt>.flac -> the rank's PCM shard, in protocol order,
    heads of 64 600 samples; two ranks split the list like shard_range."""
    import flac_enc
    rs = np.random.RandomState(2)
    utts = [f"DF_E_{2000011 + i}" for i in range(5)]
    clips = [(rs.randn(n) * 1500).astype(np.int16) for n in (70000, 3000, 64600, 16000, 99)]
    os.makedirs(tmp_path / "flac")
    for u, c in zip(utts, clips):
        with open(tmp_path / "flac" / f"{u}.flac", "wb") as f:
            f.write(flac_enc.encode(c.astype(np.int64), kind="fixed2", porder=1, rate=16000))
    (tmp_path / "trials.txt").write_text("\n".join(utts) + "\n")
    tool = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "score_files.py")
    for world in (1, 2):
        for rank in range(world):
            env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
            r = subprocess.run([sys.executable, tool, "--protocol", str(tmp_path / "trials.txt"), "--audio-dir", str(tmp_path), "--shard-only",
                                "--out", str(tmp_path / f"score{world}.txt"), "--workers", "2"], capture_output=True, text=True, env=env, timeout=300)
            assert r.returncode == 0, r.stderr[-2000:]
            rec = json.loads(r.stdout.strip().splitlines()[-1])
            lo, hi = sls.shard_range(len(utts), rank, world)
            sh = sls.PcmShard(rec["shard"])
            assert rec["trials"] == hi - lo and sh.utt_ids == utts[lo:hi]
            assert all(np.array_equal(sh.clip(i - lo), clips[i][:64600]) for i in range(lo, hi))
    r = subprocess.run([sys.executable, tool, "--protocol", str(tmp_path / "trials.txt"), "--audio-dir", str(tmp_path / "nowhere"), "--shard-only",
                        "--out", str(tmp_path / "x.txt")], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "missing" in r.stderr


def test_flac_decoder_survives_corrupted_streams(lib):
    """Untrusted input: bit flips, byte substitutions, truncations and absurd STREAMINFO fields either decode (when the damage
    missed everything that is checked) or return an error code - never crash, never write past the output buffer."""
    import flac_enc
    rs = np.random.RandomState(7)
    x = (rs.randn(9000) * 4000).astype(np.int64)
    streams = [flac_enc.encode(x, kind="lpc8", porder=3, blocksize=1024), flac_enc.encode(np.stack([x, x // 2], 1), kind="fixed3", stereo=10, porder=2),
               bytes.fromhex(RFC9639_EXAMPLES["d2_fixed_rice_side_right_two_frames"][0].replace(" ", ""))]
    guard = 64
    ok = err = 0
    for data in streams:
        for trial in range(150):
            bad = bytearray(data)
            kind = trial % 3
            if kind == 0:
                for _ in range(1 + trial % 4):
                    bad[rs.randint(4, len(bad))] ^= 1 << rs.randint(8)
            elif kind == 1:
                pos = rs.randint(4, len(bad))
                bad[pos:pos + rs.randint(1, 9)] = bytes(rs.randint(0, 256, size=rs.randint(1, 9)).tolist())
            else:
                bad = bad[:rs.randint(5, len(bad))]
            buf = np.frombuffer(bytes(bad), dtype=np.uint8)
            out = np.full(len(x) + guard, 12345, dtype=np.int16)
            n = lib.slsb_flac_decode_mono16(buf.ctypes.data, buf.size, 0, 1, out.ctypes.data, len(x), None)
            assert n <= len(x) and (out[len(x):] == 12345).all()
            ok, err = ok + (n >= 0), err + (n < 0)
    assert err > 300 and ok + err == 450                       # almost every mutation is caught by a CRC, the MD5 or a syntax check
    huge = bytearray(streams[0])
    huge[4 + 4 + 13] |= 0x0F                                   # total samples: top bits of the 36-bit field
    huge[4 + 4 + 14:4 + 4 + 18] = b"\xff\xff\xff\xff"
    buf = np.frombuffer(bytes(huge), dtype=np.uint8)
    info = (__import__("ctypes").c_int32 * 6)()
    out32 = np.empty(1 << 16, dtype=np.int32)
    assert lib.slsb_flac_decode(buf.ctypes.data, buf.size, 0, 1, out32.ctypes.data, out32.size, info) < 0


def test_eval_datasets_mirror_the_reference(sls, tmp_path):
    """data_utils_SSL.py surface on the native decoders: genSpoof_list's three branches, Dataset_ASVspoof2021_eval
    (<base>/flac/<utt>.flac) and Dataset_in_the_wild_eval (<base><file>) return (float32 [64600], utt_id) equal to the
    reference's int16 / 32768 -> pad(); they collate through a DataLoader like main.py:161-165 uses them."""
    import flac_enc
    rs = np.random.RandomState(9)
    utts = ["DF_E_2000011", "DF_E_2000013", "DF_E_2000024"]
    clips = [(rs.randn(n) * 3000).astype(np.int16) for n in (80000, 64600, 20001)]
    os.makedirs(tmp_path / "flac")
    for u, c in zip(utts, clips):
        (tmp_path / "flac" / f"{u}.flac").write_bytes(flac_enc.encode(c.astype(np.int64), kind="lpc8", porder=3, rate=16000))
    (tmp_path / "eval.txt").write_text("".join(u + "\n" for u in utts))
    (tmp_path / "train.txt").write_text("LA_0079 LA_T_1138215 - - bonafide\nLA_0079 LA_T_1271820 - A01 spoof\n")
    assert sls.genSpoof_list(str(tmp_path / "eval.txt"), is_train=False, is_eval=True) == utts
    labels, keys = sls.genSpoof_list(str(tmp_path / "train.txt"), is_train=True, is_eval=False)
    assert keys == ["LA_T_1138215", "LA_T_1271820"] and labels == {"LA_T_1138215": 1, "LA_T_1271820": 0}
    assert sls.genSpoof_list(str(tmp_path / "train.txt")) == (labels, keys)
    ds = sls.Dataset_ASVspoof2021_eval(utts, str(tmp_path))
    assert len(ds) == 3
    for i, c in enumerate(clips):
        x, u = ds[i]
        assert u == utts[i] and x.dtype == torch.float32 and x.shape == (64600,)
        assert np.array_equal(x.numpy(), sls.pad(c.astype(np.float32) / np.float32(32768.0)))
    xb, ub = next(iter(torch.utils.data.DataLoader(ds, batch_size=3, shuffle=False, drop_last=False)))
    assert xb.shape == (3, 64600) and list(ub) == utts
    sls.write_wav_pcm16(str(tmp_path / "wild_7.wav"), clips[2])
    x, u = sls.Dataset_in_the_wild_eval(["wild_7.wav"], str(tmp_path) + "/")[0]
    assert u == "wild_7.wav" and np.array_equal(x.numpy(), sls.pad(clips[2].astype(np.float32) / np.float32(32768.0)))
    # multi-channel input is down-mixed like librosa.load(mono=True): per-channel float32, THEN the float mean - an odd
    # channel sum keeps its half LSB (the int16 scorer path rounds it); stereo WAV and stereo FLAC, early stop included
    import wave
    st = np.stack([clips[2].astype(np.int64), (clips[2].astype(np.int64) // 3) | 1], 1)               # plenty of odd sums
    want = np.mean(st.astype(np.float32) / np.float32(32768.0), axis=1, dtype=np.float32)              # librosa.to_mono
    with wave.open(str(tmp_path / "wild_st.wav"), "wb") as w:
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(16000)
        w.writeframes(st.astype("<i2").tobytes())
    (tmp_path / "wild_st.flac").write_bytes(flac_enc.encode(st, kind="fixed2", stereo=10, rate=16000))
    for name in ("wild_st.wav", "wild_st.flac"):
        x, _ = sls.Dataset_in_the_wild_eval([name], str(tmp_path) + "/")[0]
        assert np.array_equal(x.numpy(), sls.pad(want)), name
        assert np.array_equal(sls.read_audio_float32(str(tmp_path / name), max_samples=777), want[:777])
        assert np.abs(sls.read_audio_pcm16(str(tmp_path / name)).astype(np.float32) / 32768 - want).max() == pytest.approx(0.5 / 32768)
    # formats the reference would resample / convert are refused with ONE exception type (never guessed): 8 kHz, 24-bit, junk
    with wave.open(str(tmp_path / "w8k.wav"), "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(8000); w.writeframes(bytes(200))
    with wave.open(str(tmp_path / "w24.wav"), "wb") as w:
        w.setnchannels(1); w.setsampwidth(3); w.setframerate(16000); w.writeframes(bytes(300))
    (tmp_path / "junk.wav").write_bytes(b"RIFF\x10\x00\x00\x00WAVEfmt \x28\x00\x00\x00\xfe\xff" + bytes(60))       # WAVE_FORMAT_EXTENSIBLE stub
    for name in ("w8k.wav", "w24.wav", "junk.wav"):
        with pytest.raises(sls.AudioFormatError):
            sls.Dataset_in_the_wild_eval([name], str(tmp_path) + "/")[0]
        with pytest.raises(sls.AudioFormatError):
            sls.read_wav_pcm16(str(tmp_path / name))


def test_evaluate_2021_DF_tool_matches_oracle_eer(sls, tmp_path):
    """tools/evaluate_2021_DF.py: the reference's CLI contract (3 arguments, trial-count / column checks, phase filter,
    "eer: %.2f") with the EER of the pinned oracle restatement of eval_metrics_DF.compute_eer."""
    from oracle.eer import compute_eer as oracle_eer
    rs = np.random.RandomState(13)
    n = 400
    utts = [f"DF_E_{4000000 + i}" for i in range(n)]
    bona = rs.rand(n) < 0.2
    phase = np.where(rs.rand(n) < 0.7, "eval", "progress")
    scores = np.where(bona, rs.beta(5, 2, n), rs.beta(2, 5, n))
    scores[::17] = 0.5                                                  # ties
    os.makedirs(tmp_path / "keys" / "CM")
    (tmp_path / "keys" / "CM" / "trial_metadata.txt").write_text("".join(
        f"LA_0009 {u} nocodec asvspoof A14 {'bonafide' if b else 'spoof'} notrim {p} traditional_vocoder - - - -\n" for u, b, p in zip(utts, bona, phase)))
    sls.write_score_file(str(tmp_path / "score.txt"), utts, scores.tolist())
    tool = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "evaluate_2021_DF.py")
    for ph in ("eval", "progress"):
        r = subprocess.run([sys.executable, tool, str(tmp_path / "score.txt"), str(tmp_path / "keys"), ph], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-1500:]
        m = phase == ph
        want = oracle_eer(scores[m & bona], scores[m & ~bona])[0]
        assert r.stdout.strip().splitlines()[-1] == "eer: %.2f" % (100 * want)
    sls.write_score_file(str(tmp_path / "short.txt"), utts[:-1], scores[:-1].tolist())
    r = subprocess.run([sys.executable, tool, str(tmp_path / "short.txt"), str(tmp_path / "keys"), "eval"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 1 and "CHECK: submission has 399 of 400 expected trials." in r.stdout
    r = subprocess.run([sys.executable, tool, str(tmp_path / "score.txt"), str(tmp_path / "keys"), "dev"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 1 and "phase must be" in r.stdout
    r = subprocess.run([sys.executable, tool, str(tmp_path / "score.txt")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 1 and "invalid input arguments" in r.stdout


def test_c_abi_is_plain_c_and_callable_from_a_c_host(sls, tmp_path):
    """The boundary is a C ABI: include/slsb200.h compiles as strict C99 (no C++ / torch types) and a C program linked against
    libslsb200.so calls it - ABI version, the FLAC decoder on RFC 9639's first example, the GEMM schedule replay, and
    slsb_create failing loudly without a GPU (no CPU fallback)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "host.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "slsb200.h"
int main(void) {
    static const unsigned char ex1[] = {0x66,0x4c,0x61,0x43,0x80,0x00,0x00,0x22,0x10,0x00,0x10,0x00,0x00,0x00,0x0f,0x00,0x00,0x0f,0x0a,0xc4,0x42,0xf0,
        0x00,0x00,0x00,0x01,0x3e,0x84,0xb4,0x18,0x07,0xdc,0x69,0x03,0x07,0x58,0x6a,0x3d,0xad,0x1a,0x2e,0x0f,0xff,0xf8,0x69,0x18,0x00,0x00,0xbf,
        0x03,0x58,0xfd,0x03,0x12,0x8b,0xaa,0x9a};
    int32_t pcm[8], info[6], items[5 * 4096], split = 0;
    int64_t n = slsb_flac_decode(ex1, (int64_t)sizeof ex1, 0, 1, pcm, 8, info);
    int k = slsb_debug_pair_schedule(12864, 3072, 74, items, 4096, &split);
    slsb_config cfg;
    memset(&cfg, 0, sizeof cfg);
    printf("abi=%d flac=%lld L=%d R=%d md5=%d items=%d split=%d\n", slsb_abi_version(), (long long)n, pcm[0], pcm[1], info[4], k, split);
    return 0;
}
''')
    exe = tmp_path / "host"
    libdir = os.path.dirname(sls.LIB_PATH)
    cc = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(root, "include"), str(src), "-o", str(exe),
                         "-L", libdir, "-l:" + os.path.basename(sls.LIB_PATH), "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip() == "abi=1 flac=1 L=25588 R=10416 md5=1 items=632 split=2"       # 8 x 74 whole tiles + 20 tiles x 2 slices


def test_score_pcm_shard_prefetch_keeps_order_and_buffers_apart(sls, tmp_path):
    """Host logic of score_pcm_shard with a stand-in engine (no GPU): every batch reaches the scorer exactly once, in order, with
    the (pcm, offsets, lens) of its own clips even though the next batch is being staged concurrently into the other buffer."""
    import time as _time
    rs = np.random.RandomState(21)
    lens = rs.randint(50, 400, size=23).tolist()
    clips = [rs.randint(-32768, 32768, size=n).astype(np.int16) for n in lens]
    sls.write_pcm_shard(str(tmp_path / "s"), [f"u{i}" for i in range(len(clips))], clips)
    shard = sls.PcmShard(str(tmp_path / "s"))
    calls = []

    class FakeEngine:
        def score_pcm16_arrays(self, pcm, off, ln, head, prec, samples):
            _time.sleep(0.01)                                   # let the prefetch of the next batch overlap this "forward"
            got = [pcm[int(o):int(o) + int(n)].numpy().copy() for o, n in zip(off, ln)]
            calls.append(got)
            return torch.tensor([float(g.astype(np.int64).sum()) for g in got])

    class FakeModel:
        def engine(self): return FakeEngine()
        def _head(self): return 3
        def _prec(self): return 1

    for batch, lo, hi, cut in ((4, 0, None, 300), (5, 3, 19, 100), (64, 0, None, 64600), (1, 20, 23, 64600)):
        calls.clear()
        out = sls.score_pcm_shard(FakeModel(), shard, batch=batch, samples=cut, lo=lo, hi=hi)
        h = len(clips) if hi is None else hi
        want = [clips[i][:cut] for i in range(lo, h)]
        flat = [g for c in calls for g in c]
        assert len(flat) == len(want) and all(np.array_equal(a, b) for a, b in zip(flat, want))
        assert out.tolist() == [float(w.astype(np.int64).sum()) for w in want]
        assert [len(c) for c in calls] == [min(batch, h - a) for a in range(lo, h, batch)]
    assert sls.score_pcm_shard(FakeModel(), shard, lo=5, hi=5).numel() == 0


def test_library_sass_is_blackwell_native(sls):
    """Build guard (cuobjdump, no GPU): the hot kernels carry tcgen05 / TMEM / TMA SASS (UTC*MMA, LDTM, UTMALDG, UTMASTG,
    UTMAREDG, UBLKCP) and no kernel falls back to the legacy mma.sync path (HMMA)."""
    import shutil
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "sass_evidence.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-1500:]
    rows = {ln.split('",')[0].strip('"'): dict(zip(r.stdout.splitlines()[0].split(",")[1:], map(int, ln.split('",')[1].split(","))))
            for ln in r.stdout.splitlines()[1:]}
    pair, attn, ln2, lns = rows["tc_gemm_pair_kernel"], rows["attn_tc_kernel<false>"], rows["tc_gemm_ln2x_kernel<1>"], rows["ln_stream_kernel<true>"]
    assert pair["UTCHMMA"] > 0 and pair["LDTM"] > 0 and pair["UTMALDG"] > 0 and pair["UTMASTG"] > 0 and pair["UTMAREDG"] > 0
    assert attn["UTCHMMA"] > 0 and attn["STTM"] > 0 and attn["UTMALDG"] > 0           # P written back into tensor memory
    wide = rows["attn_tc_kernel<true>"]                                                # T in (256, 512]: same tcgen05 structure
    assert wide["UTCHMMA"] > 0 and wide["STTM"] > 0 and wide["UTMALDG"] > 0
    assert ln2["UTCHMMA"] > 0 and ln2["UTMASTG"] > 0 and lns["UBLKCP"] > 0
    c0 = rows["conv0_tc_kernel"]                                                       # conv0: audio -> tcgen05 -> LN + GELU -> TMA store
    assert c0["UTCHMMA"] > 0 and c0["LDTM"] > 0 and c0["UTMASTG"] > 0 and c0["MUFU.TANH"] > 0
    assert all(v["HMMA"] == 0 for v in rows.values())


def test_score_audio_files_pipeline_with_stand_in_engine(sls, tmp_path):
    """score_audio_files (decode pool running ahead of the scorer, no shard on disk) hands every file to the scorer once, in
    order, as the head of its decoded clip; a stand-in engine replaces the GPU."""
    import flac_enc
    rs = np.random.RandomState(23)
    clips = [(rs.randn(n) * 2000).astype(np.int16) for n in rs.randint(200, 3000, size=19)]
    paths = []
    for i, c in enumerate(clips):
        if i % 3 == 0:
            paths.append(str(tmp_path / f"f{i}.wav"))
            sls.write_wav_pcm16(paths[-1], c)
        else:
            paths.append(str(tmp_path / f"f{i}.flac"))
            (tmp_path / f"f{i}.flac").write_bytes(flac_enc.encode(c.astype(np.int64), kind="fixed2", blocksize=512, rate=16000))
    seen = []

    class FakeEngine:
        def score_pcm16_arrays(self, pcm, off, ln, head, prec, samples):
            got = [pcm[int(o):int(o) + int(n)].numpy().copy() for o, n in zip(off, ln)]
            seen.extend(got)
            return torch.tensor([float(g.astype(np.int64).sum()) for g in got])

    class FakeModel:
        def engine(self): return FakeEngine()
        def _head(self): return 1
        def _prec(self): return 1

    for batch, cut, workers, ahead in ((4, 1000, 3, 2), (64, 64600, 1, 1), (1, 500, 6, 8)):
        seen.clear()
        out = sls.score_audio_files(FakeModel(), paths, batch=batch, samples=cut, workers=workers, ahead=ahead)
        want = [c[:cut] for c in clips]
        assert len(seen) == len(want) and all(np.array_equal(a, b) for a, b in zip(seen, want))
        assert out.tolist() == [float(w.astype(np.int64).sum()) for w in want]
    assert sls.score_audio_files(FakeModel(), [], batch=4).numel() == 0
    with pytest.raises(FileNotFoundError):
        sls.score_audio_files(FakeModel(), paths[:3] + [str(tmp_path / "missing.flac")], batch=2)


def test_bench_reference_arm_contract(tmp_path):
    """`bench.py --impl reference` (CPU oracle port timed on the host cores): one JSON line with the GPU arm's metric / unit /
    workload, `impl`, `cpu_baseline`, and an `e2e` equal to the line's own value; under a multi-rank launch only rank 0 works and
    prints, the other ranks exit 0 silently."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--head", "sae"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=root)
    assert r.returncode == 0, r.stderr[-1500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "utterances_per_second" and d["unit"] == "utt/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["e2e"] == {"value": d["value"], "unit": "utt/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert "batch=64 x 64600-sample clips per GPU (BASELINE config 2)" in d["config"]["workload"] and d["config"]["head"] == "sae"
    r1 = subprocess.run(cmd + ["--gpus", "2"], capture_output=True, text=True, timeout=300, cwd=root,
                        env=dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"))
    assert r1.returncode == 0 and not [ln for ln in r1.stdout.splitlines() if ln.startswith("{")]


def test_compute_eer_matches_oracle_on_random_cases_with_ties(sls):
    """Property test of scoring.compute_eer (torch, any device) against the pinned oracle restatement of
    eval_metrics_DF.compute_eer: heavy ties (quantised scores), tiny and lopsided class counts - EER and threshold equal exactly."""
    from hypothesis import given, settings, strategies as st
    from oracle.eer import compute_eer as oracle_eer

    @settings(max_examples=60, deadline=None)
    @given(st.integers(1, 40), st.integers(1, 40), st.integers(2, 12), st.integers(0, 2 ** 31 - 1))
    def check(n_bona, n_spoof, levels, seed):
        rs = np.random.RandomState(seed)
        bona = np.round(rs.beta(4, 2, n_bona) * levels) / levels
        spoof = np.round(rs.beta(2, 4, n_spoof) * levels) / levels
        scores = np.concatenate([bona, spoof])
        lab = np.concatenate([np.ones(n_bona, bool), np.zeros(n_spoof, bool)])
        perm = rs.permutation(scores.size)                         # protocol order mixes the classes
        got = sls.compute_eer(torch.from_numpy(scores[perm]), torch.from_numpy(lab[perm]))
        want = oracle_eer(scores[perm][lab[perm]], scores[perm][~lab[perm]])
        assert got[0] == float(want[0]) and got[1] == float(want[1])

    check()


def test_flac_decoder_roundtrip_property(lib):
    """decode(encode(x)) == x with the MD5 verified, over random signals, block sizes, predictors, Rice settings, stereo modes and
    bit depths (test encoder: tests/flac_enc.py)."""
    import flac_enc
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.integers(0, 2 ** 31 - 1), st.sampled_from([8, 12, 16, 20, 24]), st.integers(1, 1500),
           st.sampled_from([16, 64, 192, 255, 256, 576, 1000, 1152, 4096]),
           st.sampled_from(["verbatim", "constant", "fixed0", "fixed1", "fixed2", "fixed3", "fixed4", "lpc1", "lpc2", "lpc5", "lpc8", "lpc12", "lpc20", "lpc32"]),
           st.integers(0, 4), st.integers(0, 1), st.sampled_from([None, None, 8, 9, 10]), st.sampled_from(["noise", "sine", "silence", "dc", "full", "sparse"]))
    def check(seed, bps, n, blocksize, kind, porder, method, stereo, signal):
        rs = np.random.RandomState(seed)
        hi = (1 << (bps - 1)) - 1
        t = np.arange(n)

        def make():
            if signal == "noise":
                return rs.randint(-hi // 4, hi // 4 + 1, size=n)
            if signal == "sine":
                return (0.6 * hi * np.sin(t * rs.uniform(0.01, 0.5)) + rs.randn(n) * max(1, hi / 500)).astype(np.int64)
            if signal == "silence":
                return np.zeros(n, np.int64)
            if signal == "dc":
                return np.full(n, rs.randint(-hi, hi + 1), np.int64)
            if signal == "full":
                return np.where(rs.rand(n) < 0.5, hi, -hi - 1).astype(np.int64)
            return (rs.randint(-hi, hi + 1, size=n) * (rs.rand(n) < 0.05)).astype(np.int64) << rs.randint(0, 3)
        x = make() if stereo is None else np.stack([make(), make()], 1)
        x = np.clip(x, -hi - 1, hi)
        esc = tuple(i for i in range(1 << porder) if rs.rand() < 0.2)
        data = flac_enc.encode(x, bps=bps, rate=16000, blocksize=blocksize, kind=kind, stereo=stereo, porder=porder, method=method,
                               escape_partitions=esc)
        n_dec, info, pcm = _flac_decode(lib, data)
        assert n_dec == n and info[:3] == [16000, 1 if stereo is None else 2, bps] and info[4] == 1
        assert np.array_equal(pcm if stereo is not None else pcm[:, 0], x)

    check()


# ---------------------------------------------------------------------------------------------- device FLAC decoder: host side
def _frame_core(sls, data, max_samples=64600):
    """scan -> frame jobs -> the frame core compiled for the host (csrc/flac_frame.h, the code the GPU runs one thread per frame)."""
    sc = sls.scan_flac_bytes(data, max_samples, sample_rate=None)
    d, fr, tot, off, lens = sls.pack_flac_batch([sc], max_samples)
    pcm, st = sls.decode_flac_frames_host(d, fr, tot)
    return sc, pcm, st, fr


def test_flac_frame_core_matches_host_decoder_on_every_syntax_element(sls, lib):
    """The frame core of the device decoder (flac_frame.h: streamed Rice decode + predictor restore, word-wise bit reader) equals
    flac_decode.cpp sample for sample on CONSTANT / VERBATIM / FIXED 0-4 / LPC 1-32, both Rice methods, partition orders 0-4,
    escape partitions, wasted bits, odd block sizes, two-byte frame numbers - for whole clips and for heads that end inside a frame."""
    import flac_enc
    rs = np.random.RandomState(5)
    t = np.arange(20000)
    mono = (8000 * np.sin(t * 0.05) + rs.randn(t.size) * 300).astype(np.int64)
    sq = np.where((t // 50) % 2 == 0, 32767, -32768)
    cases = [(mono, dict(kind="verbatim")), (mono, dict(kind="fixed0")), (mono, dict(kind="fixed1", porder=1)), (mono, dict(kind="fixed2", porder=3)),
             (mono, dict(kind="fixed3", porder=4, method=1)), (mono, dict(kind="fixed4", porder=2, escape_partitions=(1, 3))),
             (mono, dict(kind="lpc8", porder=2)), (mono, dict(kind="lpc12", porder=3)), (mono, dict(kind="lpc13", porder=1)),
             (mono, dict(kind="lpc32", porder=1, blocksize=1152)), (mono, dict(kind="fixed2", blocksize=777)),
             (mono[:3000], dict(kind="fixed2", blocksize=16, extra_metadata=False)), (mono, dict(kind="lpc4", with_md5=False)),
             (sq, dict(kind="fixed1", porder=2)), (np.zeros(5000, np.int64), dict(kind="constant")), ((mono >> 3) << 3, dict(kind="fixed2")),
             (mono, dict(kind="lpc1", blocksize=192)), (mono, dict(kind="lpc5", blocksize=4608, porder=2))]
    for x, kw in cases:
        data = flac_enc.encode(x, **kw)
        for cut in (64600, 5000, 1):
            ref = sls.decode_flac_bytes(data, cut, sample_rate=None)
            sc, pcm, st, fr = _frame_core(sls, data, cut)
            assert sc.device_ok and (st > 0).all(), (kw, cut, st)
            assert st.tolist() == sc.blocks.tolist()                      # the status of a decoded frame is its block size
            assert np.array_equal(pcm, ref) and np.array_equal(ref, x[:cut].astype(np.int16)), (kw, cut)
            assert int(fr["keep"].sum()) == ref.size


def test_flac_scan_frame_table_and_rejections(sls, lib):
    """slsb_flac_scan: frame boundaries are exact (offsets chain, the lengths cover the audio part of the file), the RFC 9639 example
    files scan and decode, streams the device decoder does not take are flagged, damaged streams are refused."""
    import flac_enc
    rs = np.random.RandomState(6)
    x = (5000 * np.sin(np.arange(30000) * 0.02) + rs.randn(30000) * 500).astype(np.int64)
    data = flac_enc.encode(x, kind="lpc8", porder=3, blocksize=1024)
    sc = sls.scan_flac_bytes(data, None, sample_rate=None)
    assert len(sc.off) == 30 and sc.samples == 30000 and sc.total == 30000 and sc.blocks.tolist() == [1024] * 29 + [30000 - 29 * 1024]
    assert np.array_equal(sc.off[1:], sc.off[:-1] + sc.len[:-1]) and int(sc.off[-1] + sc.len[-1]) == len(data)
    head = sls.scan_flac_bytes(data, 2500, sample_rate=None)                  # the head needs 3 frames
    assert len(head.off) == 3 and head.samples == 2500
    # a sync code planted inside a frame (0xFF 0xF8 ...) must not split it: raw 16-bit samples that spell sync codes
    evil = np.tile(np.array([-8, -1, -7, -2], dtype=np.int64), 3000)          # 0xFFF8 0xFFFF 0xFFF9 0xFFFE as verbatim samples
    data2 = flac_enc.encode(evil, kind="verbatim", blocksize=4096)
    sc2, pcm2, st2, _ = _frame_core(sls, data2)
    assert len(sc2.off) == 3 and (st2 > 0).all() and np.array_equal(pcm2, evil.astype(np.int16))
    # RFC 9639 appendix D: example 1 (mono... if 16-bit mono) decodes through the frame core; stereo examples are flagged, not decoded
    for name, (hexs, rate, ch, bps, want) in RFC9639_EXAMPLES.items():
        raw = bytes.fromhex(hexs.replace(" ", ""))
        s3 = sls.scan_flac_bytes(raw, None, sample_rate=None)
        assert (s3.rate, s3.channels, s3.bps) == (rate, ch, bps) and s3.samples == len(want), name
        assert s3.device_ok == (ch == 1 and bps == 16), name
        if s3.device_ok:
            _, pcm3, st3, _ = _frame_core(sls, raw, None)
            assert (st3 > 0).all() and pcm3.tolist() == [w[0] for w in want], name
    stereo = flac_enc.encode(np.stack([x, x // 2], 1), kind="fixed2", stereo=10)
    assert not sls.scan_flac_bytes(stereo, 64600, sample_rate=None).device_ok
    assert not sls.scan_flac_bytes(flac_enc.encode(x >> 8, bps=8), 64600, sample_rate=None).device_ok
    with pytest.raises(sls.AudioFormatError):
        sls.scan_flac_bytes(flac_enc.encode(x, rate=44100), 64600)             # wrong rate, as the host path
    bad = bytearray(data)
    bad[len(bad) // 2] ^= 0x04                                                 # payload damage: some frame's CRC-16 no longer matches
    with pytest.raises(sls.AudioFormatError):
        sls.scan_flac_bytes(bytes(bad), None, sample_rate=None)
    with pytest.raises(sls.AudioFormatError):
        sls.scan_flac_bytes(data[:len(data) // 2], None, sample_rate=None)     # truncated
    # the frame core refuses what it is not built for instead of guessing (status -9), the caller then decodes on the host
    d, fr, tot, off, lens = sls.pack_flac_batch([sls.scan_flac_bytes(stereo, 64600, sample_rate=None)], 64600)
    _, st = sls.decode_flac_frames_host(d, fr, tot)
    assert (st == -9).all()


def test_flac_frame_core_property_random_streams(sls):
    """decode_frames(scan(encode(x))) == x over random signals, predictors, Rice settings and block sizes; several clips per batch."""
    import flac_enc
    rs = np.random.RandomState(7)
    scans, refs = [], []
    for i in range(24):
        n = int(rs.randint(1, 9000))
        amp = float(rs.choice([3, 300, 12000]))
        x = np.clip(amp * np.sin(np.arange(n) * rs.uniform(0.001, 1.0)) + rs.randn(n) * amp * rs.uniform(0, 0.5), -32768, 32767).astype(np.int64)
        kind = str(rs.choice(["verbatim", "constant" if False else "fixed0", "fixed1", "fixed2", "fixed3", "fixed4", "lpc2", "lpc8", "lpc12", "lpc20"]))
        data = flac_enc.encode(x, kind=kind, porder=int(rs.randint(0, 5)), method=int(rs.randint(0, 2)), blocksize=int(rs.choice([192, 576, 1024, 4096, 333])))
        cut = int(rs.choice([64600, max(1, n // 2), max(1, n - 1)]))
        scans.append(sls.scan_flac_bytes(data, cut, sample_rate=None))
        refs.append(x[:cut].astype(np.int16))
    d, fr, tot, off, lens = sls.pack_flac_batch(scans, 64600)
    pcm, st = sls.decode_flac_frames_host(d, fr, tot)
    assert (st > 0).all() and tot == sum(r.size for r in refs)
    for o, n, r in zip(off.tolist(), lens.tolist(), refs):
        assert n == r.size and np.array_equal(pcm[o:o + n], r)
