"""Import alias for the package directory ``slsforasvspoof-2021-df_b200/`` (hyphens are not importable
with the ``import`` statement).  ``import sls_b200`` gives the package itself."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("slsforasvspoof-2021-df_b200")
sys.modules[__name__] = _pkg
