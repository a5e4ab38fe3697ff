"""Graph-batch collation on the device.

Replaces, for the default (fully connected, single-output) configuration, the reference's
``CovarianceLaggedDataset`` + PyG ``DataLoader`` / ``Batch.from_data_list`` pair
(/root/reference/utils/dataset.py:160-289; 5_train_SpotV2Net.py:66-91,142-143): instead of
materialising every windowed graph on the host, caching it with ``torch.save`` and concatenating
samples one by one in Python, the two ``[T, N, N]`` matrix stacks stay resident in HBM and each
batch is gathered straight into the reference's tensor layouts by one CUDA kernel launch pair
(``spotv2_collate_windows``).  A batch quacks like a PyG ``Batch`` (``x``, ``edge_index``,
``edge_attr``, ``y_x``, ``batch``, ``ptr``, ``num_graphs``, ``to()``), so ``GATModel.forward``
and the reference's training loop consume it unchanged, and it carries the validated topology
so no per-batch edge_index check is needed.
"""
from __future__ import annotations

from typing import Iterator, Optional, Sequence

import numpy as np
import torch
from torch import Tensor

from . import _lib
from ._lib import SpotV2Error, check, ptr, stream_ptr
from .gat_conv import Topology, WindowSource, topology_from_edge_index

REFERENCE_DROP_FIRST = 8357          # hard-coded cut at utils/dataset.py:288


def complete_graph_edge_index(N: int) -> Tensor:
    """Local edge_index [2, N(N-1)] in the reference's order (utils/dataset.py:216-226):
    upper-triangle pairs (r < c, row-major) as r -> c, then the same pairs as c -> r."""
    iu = torch.triu_indices(N, N, offset=1)
    return torch.cat([iu, iu.flip(0)], dim=1)


class SpotBatch:
    """Duck-typed PyG ``Batch`` for identical complete graphs."""

    def __init__(self, x, edge_index, edge_attr, y_x, num_graphs, nodes_per_graph, spot_topology=None, spot_windows=None,
                 spot_x16=None):
        self.x, self.edge_index, self.edge_attr, self.y_x = x, edge_index, edge_attr, y_x
        self.num_graphs, self.nodes_per_graph, self.spot_topology = num_graphs, nodes_per_graph, spot_topology
        self.spot_windows = spot_windows        # WindowSource: lets the GAT layers read the windows instead of edge_attr
        # (pair [2, B*N, ld16] fp16, scale block [8]): x as the projection GEMMs' operand pair, emitted by the collation.
        # It also rides on the tensor itself (x._spot_pair, with the tensor's version), which is where GATConv looks.
        self.spot_x16 = spot_x16

    @property
    def batch(self) -> Tensor:
        return torch.arange(self.num_graphs, device=self.x.device).repeat_interleave(self.nodes_per_graph)

    @property
    def ptr(self) -> Tensor:
        return torch.arange(self.num_graphs + 1, device=self.x.device) * self.nodes_per_graph

    def to(self, device, non_blocking: bool = False):
        device = torch.device(device)
        if device == self.x.device:
            return self
        if self.edge_attr is None:
            raise SpotV2Error("a structured batch (no materialised edge_attr) lives on its dataset's device")
        moved = [t.to(device, non_blocking=non_blocking) for t in (self.x, self.edge_index, self.edge_attr, self.y_x)]
        return SpotBatch(*moved, self.num_graphs, self.nodes_per_graph, None)


_EDGE_CACHE: dict = {}


def batched_topology(B: int, N: int, device) -> tuple[Tensor, Topology]:
    """edge_index [2, B*N*(N-1)] for B copies of the complete graph plus its validated table (cached)."""
    key = (B, N, torch.device(device))
    hit = _EDGE_CACHE.get(key)
    if hit is None:
        local = complete_graph_edge_index(N).to(device)
        offs = (torch.arange(B, device=device) * N).view(1, B, 1)
        ei = (local.view(2, 1, -1) + offs).reshape(2, -1).contiguous()
        hit = (ei, topology_from_edge_index(ei, B * N, N))
        if len(_EDGE_CACHE) > 16:
            _EDGE_CACHE.clear()
        _EDGE_CACHE[key] = hit
    return hit


def load_matrix_stack(path: str) -> np.ndarray:
    """[T, N, N] float64 stack from the reference's H5 layout (keys '0'..'T-1', one N x N matrix
    each; 3_create_matrix_dataset.py:215-222) or from a .npy file."""
    if str(path).endswith(".npy"):
        return np.load(path)
    try:
        import h5py  # not in this image; present wherever the reference runs
    except ImportError as e:
        raise SpotV2Error(f"reading {path} needs h5py, which is not installed; pass arrays or .npy instead") from e
    with h5py.File(path, "r") as f:
        keys = sorted(f.keys(), key=int)
        return np.stack([np.asarray(f[k]) for k in keys])


class WindowDataset:
    """Device-resident equivalent of ``CovarianceLaggedDataset`` (utils/dataset.py:160-289).

    Sample ``k`` is the window starting at ``t0 = k + drop_first``: ``x[i, c*L + t] = vol[t0+t, i, c]``,
    ``edge_attr[e, k*L + t]`` = (volvol[t0+t, r, c], volvol[.., src, src], volvol[.., dst, dst]),
    ``y_x[i] = vol[t0+L, i, i]``.

    ``future_steps=K`` gives ``CovarianceLaggedMultiOutputDataset`` (utils/dataset.py:293-412; chosen by
    5_train_SpotV2Net.py:66-76 when ``output_node_channels > 1``): same x / edges, targets
    ``y_x[i*K + k] = vol[t0+L+k, i, i]``, and ``K - 1`` fewer samples.
    """

    def __init__(self, vol, volvol, seq_length: int, device="cuda", drop_first: int = REFERENCE_DROP_FIRST,
                 future_steps: Optional[int] = None, structured: bool = False):
        vol = torch.as_tensor(np.asarray(vol) if not torch.is_tensor(vol) else vol)
        volvol = torch.as_tensor(np.asarray(volvol) if not torch.is_tensor(volvol) else volvol)
        if vol.shape != volvol.shape or vol.dim() != 3 or vol.shape[1] != vol.shape[2]:
            raise SpotV2Error("vol and volvol must both be [T, N, N]")
        self.T, self.N = int(vol.shape[0]), int(vol.shape[1])
        self.L = int(seq_length)
        self.drop_first = int(drop_first)
        # structured=True: batches carry window references (SpotBatch.spot_windows) instead of the materialised
        # [B*N*(N-1), 3L] edge_attr (edge_attr is None); GATModel / GATConv then read the [L, N, N] windows directly
        self.structured = bool(structured)
        self.future_steps = None if future_steps is None else int(future_steps)
        if self.future_steps is not None and self.future_steps < 1:
            raise SpotV2Error("future_steps must be >= 1")
        self._tail = self.L + (self.future_steps - 1 if self.future_steps else 0)
        if self.T - self._tail - self.drop_first <= 0:
            raise SpotV2Error(f"T={self.T} too short for seq_length={self.L} and drop_first={self.drop_first}")
        self.device = torch.device(device)
        self.vol = vol.to(self.device, torch.float32).contiguous()          # numpy float64 -> fp32, as torch.tensor(.., dtype=float)
        self.volvol = volvol.to(self.device, torch.float32).contiguous()
        self.num_node_features = self.N * self.L
        self.num_edge_features = 3 * self.L
        # per-matrix sums of the co-volatility stack: batch statistics of edge_attr (GATModel standardize=True) without a pass
        # over the edges (WindowSource.edge_stats)
        vv64 = self.volvol.double()
        up, dg = vv64.triu(1), vv64.diagonal(dim1=1, dim2=2)
        self.mat_stats = torch.stack([up.sum((1, 2)), (up * up).sum((1, 2)), dg.sum(1), (dg * dg).sum(1)], dim=1).contiguous()
        del vv64, up, dg
        # x values are entries of vol: the fp16 operand pair of every batch shares one power-of-two scale, fixed here
        self.x_scale = torch.empty(8, device=self.device, dtype=torch.float32)
        lib = _lib.load()
        with torch.cuda.device(self.device):
            check(lib.spotv2_stack_scale(ptr(self.vol), self.vol.numel(), ptr(self.x_scale), stream_ptr(self.device)),
                  "spotv2_stack_scale")

    def __len__(self) -> int:
        return self.T - self._tail - self.drop_first

    def collate(self, indices: Sequence[int] | Tensor) -> SpotBatch:
        idx = torch.as_tensor(indices, dtype=torch.int64)
        if idx.numel() == 0:
            raise SpotV2Error("empty batch")
        if int(idx.min()) < 0 or int(idx.max()) >= len(self):
            raise IndexError("sample index out of range")
        B, N, L = int(idx.numel()), self.N, self.L
        t0 = (idx + self.drop_first).to(torch.int32).to(self.device)
        dev = self.device
        x = torch.empty(B * N, N * L, device=dev, dtype=torch.float32)
        ea = None if self.structured else torch.empty(B * N * (N - 1), 3 * L, device=dev, dtype=torch.float32)
        y = torch.empty(B * N, device=dev, dtype=torch.float32)
        lib = _lib.load()
        ld16 = lib.spotv2_gat_ld16(N * L)
        x16 = torch.empty(2, B * N, ld16, device=dev, dtype=torch.float16)
        with torch.cuda.device(dev):
            check(lib.spotv2_collate_windows_pair(ptr(self.vol), ptr(self.volvol), self.T, N, L, ptr(t0), B, ptr(x), ptr(x16[0]),
                                                  ptr(x16[1]), ld16, ptr(self.x_scale), ptr(ea), ptr(y), stream_ptr(dev)),
                  "spotv2_collate_windows_pair")
        x._spot_pair = (x16, self.x_scale, x._version)
        if self.future_steps is not None:        # [B, K, N] gather of the next K diagonals -> [B*N*K], node major
            K = self.future_steps
            t = (t0.to(torch.int64) + L).view(B, 1) + torch.arange(K, device=dev).view(1, K)
            y = self.vol.diagonal(dim1=1, dim2=2)[t].permute(0, 2, 1).reshape(-1).contiguous()
        ei, topo = batched_topology(B, N, dev)
        # window references ride only on structured batches: with a materialised edge_attr the layers must use it
        src = WindowSource(self.volvol, t0, L, checked=True, mat_stats=self.mat_stats)
        if ea is not None:
            ea._spot_stats_src = (src, ea._version)       # where GATModel finds the statistics of this very tensor
        return SpotBatch(x, ei, ea, y, B, N, topo, src if self.structured else None, (x16, self.x_scale))

    def __getitem__(self, k):
        if isinstance(k, slice):
            return _Subset(self, range(*k.indices(len(self))))
        return self.collate([int(k)])


class _Subset:
    def __init__(self, ds: WindowDataset, rng: range):
        self.ds, self.rng = ds, rng

    def __len__(self):
        return len(self.rng)

    def collate(self, indices):
        base = torch.as_tensor(indices, dtype=torch.int64)
        return self.ds.collate(base * self.rng.step + self.rng.start)


class WindowLoader:
    """``DataLoader(dataset, batch_size, shuffle)`` stand-in (5_train_SpotV2Net.py:90-91)."""

    def __init__(self, dataset, batch_size: int = 1, shuffle: bool = False, generator: Optional[torch.Generator] = None):
        self.dataset, self.batch_size, self.shuffle, self.generator = dataset, int(batch_size), shuffle, generator

    def __len__(self) -> int:
        return (len(self.dataset) + self.batch_size - 1) // self.batch_size

    def __iter__(self) -> Iterator[SpotBatch]:
        n = len(self.dataset)
        order = torch.randperm(n, generator=self.generator) if self.shuffle else torch.arange(n)
        for s in range(0, n, self.batch_size):
            yield self.dataset.collate(order[s:s + self.batch_size])
