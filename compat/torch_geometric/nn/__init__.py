"""`torch_geometric.nn` names used by /root/reference/utils/models.py:11."""
from spotv2net_b200 import GATConv  # noqa: F401


class GATv2Conv:        # imported by the reference but never constructed (utils/models.py builds GATConv only)
    def __init__(self, *a, **k):
        raise NotImplementedError("GATv2Conv is outside the scope of spotv2net_b200 (the reference never instantiates it)")
