"""Import shim: lets the reference's own files import `torch_geometric` names and get the B200 implementations.

Put this directory on ``sys.path`` (``PYTHONPATH=/root/repo/compat:/root/repo``) and
``/root/reference/utils/models.py`` (``from torch_geometric.nn import GATConv, GATv2Conv``, line 11) and
``/root/reference/5_train_SpotV2Net.py`` (``from torch_geometric.loader import DataLoader``, line 11) import unchanged.
Only the names the reference's hot path touches exist; everything else of PyG is deliberately absent (SURVEY.md §8b,
"Import-compat note").  This is NOT PyG and must never shadow a real installation: it refuses to load if one is found.
"""
import importlib.util as _u
import os as _os
import sys as _sys

__version__ = "2.3.0+spotv2net_b200.shim"

_here = _os.path.dirname(_os.path.abspath(__file__))
for _p in _sys.path:
    _cand = _os.path.join(_p or ".", "torch_geometric", "__init__.py")
    if _os.path.exists(_cand) and _os.path.dirname(_os.path.abspath(_cand)) != _here:
        raise ImportError(f"a real torch_geometric is installed at {_cand}; remove spotv2net_b200's compat/ shim from sys.path")
del _u
