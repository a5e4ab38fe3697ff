"""spotv2net_b200 — the SpotV2Net GAT hot path on B200 (sm_100a).

Public surface (mirrors what the reference imports and calls):

* :class:`GATConv`   — PyG 2.3.0 ``GATConv`` API (utils/models.py:11,87-113,146)
* :class:`GATModel`  — the SpotV2Net model (utils/models.py:61-152)
* :class:`WindowDataset`, :class:`WindowLoader`, :class:`SpotBatch` — device-side collation
  replacing CovarianceLaggedDataset / CovarianceLaggedMultiOutputDataset + PyG DataLoader (utils/dataset.py:160-412)
* :func:`evaluate`, :func:`attention_weights` — the evaluation / attention-export loops of 6_results.ipynb

Everything computes in libspotv2_gat.so (include/spotv2_gat.h); importing the package does not
load it, the first call does, and a missing library is an error, never a fallback.
"""
from ._lib import SpotV2Error, load as load_library          # noqa: F401
from .gat_conv import GATConv, Topology, WindowSource, topology_from_edge_index   # noqa: F401
from .models import GATModel, gat_layer_plan                 # noqa: F401
from .data import SpotBatch, WindowDataset, WindowLoader, complete_graph_edge_index, batched_topology  # noqa: F401
from .infer import attention_weights, evaluate               # noqa: F401

__all__ = ["GATConv", "GATModel", "WindowDataset", "WindowLoader", "SpotBatch", "SpotV2Error",
           "Topology", "WindowSource", "topology_from_edge_index", "complete_graph_edge_index", "gat_layer_plan",
           "load_library", "batched_topology", "evaluate", "attention_weights"]
