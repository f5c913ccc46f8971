"""ORACLE (test infrastructure only) -- a stand-in ``fairseq`` package so that the reference's
``model.py`` / ``model_window_topk.py`` / ``model_backup.py`` import and run VERBATIM in the
build container (``import fairseq`` fails there: the source zip was stripped,
``/root/reference/.MISSING_LARGE_BLOBS:2``).

``fairseq.checkpoint_utils.load_model_ensemble_and_task([cp_path], strict=False)`` is the only
fairseq entry point those files use (``model.py:113-115``); the stub returns the restated trunk
(``oracle/trunk.py``).  Used only by ``oracle/make_golden.py`` and ``tests/test_oracle.py`` when
``/root/reference`` exists; it cannot travel to the GPU box, hence the committed fixtures.
"""
from __future__ import annotations

import sys
import types

from .trunk import TrunkConfig, Wav2Vec2Trunk

_next_trunk = {"cfg": None, "instance": None}


def set_next_trunk(cfg: TrunkConfig = None, instance: Wav2Vec2Trunk = None) -> None:
    """Choose what the next ``load_model_ensemble_and_task`` call hands back."""
    _next_trunk["cfg"], _next_trunk["instance"] = cfg, instance


def _load_model_ensemble_and_task(paths, strict=False, **_):
    trunk = _next_trunk["instance"] or Wav2Vec2Trunk(_next_trunk["cfg"])
    trunk.eval()
    return [trunk], None, None


def install() -> None:
    if "fairseq" in sys.modules and getattr(sys.modules["fairseq"], "__sls_oracle_stub__", False):
        return
    pkg = types.ModuleType("fairseq")
    pkg.__sls_oracle_stub__ = True
    cu = types.ModuleType("fairseq.checkpoint_utils")
    cu.load_model_ensemble_and_task = _load_model_ensemble_and_task
    pkg.checkpoint_utils = cu
    sys.modules["fairseq"] = pkg
    sys.modules["fairseq.checkpoint_utils"] = cu


def import_reference(module_name: str, ref_root: str = "/root/reference"):
    """Import a reference module verbatim (read-only tree; nothing is copied)."""
    import importlib
    install()
    if ref_root not in sys.path:
        sys.path.insert(0, ref_root)
    return importlib.import_module(module_name)
