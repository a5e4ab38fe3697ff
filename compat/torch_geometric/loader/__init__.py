"""`torch_geometric.loader.DataLoader` as /root/reference/5_train_SpotV2Net.py:90-91 uses it: (dataset, batch_size, shuffle)."""
from spotv2net_b200 import WindowLoader


class DataLoader(WindowLoader):
    """Accepts a ``spotv2net_b200.WindowDataset`` (or a slice of one); collation runs on the device."""

    def __init__(self, dataset, batch_size: int = 1, shuffle: bool = False, **kwargs):
        if not hasattr(dataset, "collate"):
            raise TypeError("this shim collates spotv2net_b200.WindowDataset objects (device-side); build the dataset with "
                            "WindowDataset(vol, volvol, seq_length) instead of CovarianceLaggedDataset")
        super().__init__(dataset, batch_size=batch_size, shuffle=shuffle, generator=kwargs.get("generator"))
