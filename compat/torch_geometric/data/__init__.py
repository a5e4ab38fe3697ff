"""`torch_geometric.data` names imported by /root/reference/utils/dataset.py:8 (the dataset classes themselves need h5py
and natsort and are replaced by spotv2net_b200.WindowDataset)."""
from types import SimpleNamespace


class Data(SimpleNamespace):
    """Attribute bag with PyG's constructor spelling: Data(x=..., edge_index=..., edge_attr=..., y_x=...)."""

    def to(self, device):
        import torch
        for k, v in list(vars(self).items()):
            if torch.is_tensor(v):
                setattr(self, k, v.to(device))
        return self


class InMemoryDataset:
    def __init__(self, *a, **k):
        raise NotImplementedError("use spotv2net_b200.WindowDataset: the matrices stay resident in HBM and batches are "
                                  "collated on the device")
