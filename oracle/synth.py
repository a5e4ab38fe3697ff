"""Synthetic standardised data in the reference's layouts (CPU, per-sample loops).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Restates, sample by sample
the way the reference does it, the graph construction of
``CovarianceLaggedDataset.process`` (/root/reference/utils/dataset.py:182-289)
and PyG 2.3.0's ``Batch.from_data_list`` collation
([PyG] data/collate.py; call site /root/reference/5_train_SpotV2Net.py:90,142).
The TAQ data is not available (/root/reference/README.md:66-70), so the H5
contents are replaced by symmetric N(0,1) matrices of the same shape
(SURVEY.md §8d).
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Optional, Sequence

import numpy as np
import torch


def synthetic_matrices(T: int, N: int, seed: int = 1234):
    """Two stacks [T, N, N] float64, symmetric, ~N(0,1): stand-ins for the
    keys '0'..'T-1' of vols_mats_taq_standardized.h5 / volvols_mats_taq_standardized.h5
    (/root/reference/3_create_matrix_dataset.py:215-222, 4_standardize_data.py:55-77)."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(2):
        a = torch.randn(T, N, N, generator=g, dtype=torch.float64)
        out.append(((a + a.transpose(1, 2)) / np.sqrt(2.0)).numpy())
    return out[0], out[1]


def complete_graph_edge_index(N: int) -> torch.Tensor:
    """utils/dataset.py:216-226: argwhere(triu(k=1)) pairs as (src=r, dst=c),
    followed by the same pair list swapped."""
    mask = np.triu(np.ones((N, N)), k=1) > 0
    upper = torch.tensor(np.argwhere(mask), dtype=torch.long).t().contiguous()
    lower = upper[[1, 0], :]
    return torch.cat([upper, lower], dim=1)


def window_sample(M_vol: np.ndarray, M_vv: np.ndarray, t0: int, L: int, future_steps: Optional[int] = None):
    """One Data(x, edge_index, edge_attr, y_x) exactly as utils/dataset.py:200-282
    builds it (per-lag objects stacked on a new last dim, then flattened).
    future_steps = K: the multi-output variant (CovarianceLaggedMultiOutputDataset, utils/dataset.py:332-405):
    y_x [N*K], y_x[i*K + k] = diag(vol[t0 + L + k])[i] (the last lag's [N, K] target block, flattened)."""
    N = M_vol.shape[1]
    xs, eas = [], []
    edge_index = complete_graph_edge_index(N)
    mask = np.triu(np.ones((N, N)), k=1) > 0
    y = None
    for j in range(L):
        cov = M_vol[t0 + j]
        covol = M_vv[t0 + j]
        adj = covol.copy()
        np.fill_diagonal(adj, 0)
        variances = torch.tensor(np.diag(covol), dtype=torch.float)
        cov_e = torch.tensor(adj[mask], dtype=torch.float)
        cov_e = torch.cat([cov_e, cov_e])
        eas.append(torch.stack([cov_e, variances[edge_index[0]], variances[edge_index[1]]], dim=1))
        xs.append(torch.tensor(cov, dtype=torch.float))
        if future_steps is None:
            y = torch.tensor(np.diag(M_vol[t0 + j + 1]), dtype=torch.float)
        else:                                                       # utils/dataset.py:380-386
            y = torch.stack([torch.tensor(np.diag(M_vol[t0 + j + k + 1]), dtype=torch.float)
                             for k in range(future_steps)], dim=1).reshape(-1)
    x = torch.stack(xs, dim=2).reshape(N, -1)
    edge_attr = torch.stack(eas, dim=2).reshape(edge_index.shape[1], -1)
    return SimpleNamespace(x=x, edge_index=edge_index, edge_attr=edge_attr, y_x=y)


class Batch(SimpleNamespace):
    """Duck-typed stand-in for a PyG ``Batch`` (attributes the reference reads:
    utils/models.py:138-140, 5_train_SpotV2Net.py:143-154)."""

    def to(self, device):
        for k, v in list(vars(self).items()):
            if torch.is_tensor(v):
                setattr(self, k, v.to(device))
        return self


def collate(samples: Sequence[SimpleNamespace]) -> Batch:
    """[PyG] data/collate.py: cat x / edge_attr / y_x on dim 0, cat edge_index
    on dim 1 with cumulative node offsets, build batch and ptr."""
    xs, eis, eas, ys, bvec, ptr = [], [], [], [], [], [0]
    off = 0
    for b, s in enumerate(samples):
        n = s.x.shape[0]
        xs.append(s.x)
        eis.append(s.edge_index + off)
        eas.append(s.edge_attr)
        ys.append(s.y_x)
        bvec.append(torch.full((n,), b, dtype=torch.long))
        off += n
        ptr.append(off)
    return Batch(x=torch.cat(xs, 0), edge_index=torch.cat(eis, 1), edge_attr=torch.cat(eas, 0),
                 y_x=torch.cat(ys, 0), batch=torch.cat(bvec), ptr=torch.tensor(ptr),
                 num_graphs=len(samples))


def make_batch(M_vol, M_vv, t0s: Sequence[int], L: int, future_steps: Optional[int] = None) -> Batch:
    return collate([window_sample(M_vol, M_vv, int(t), L, future_steps) for t in t0s])


def random_complete_batch(B: int, N: int, F_in: int, Fe: int, seed: int = 0, dtype=torch.float32) -> Batch:
    """Unstructured i.i.d. N(0,1) features on the reference's complete-graph
    topology (generic operator contract)."""
    g = torch.Generator().manual_seed(seed)
    ei = complete_graph_edge_index(N)
    R = ei.shape[1]
    edge_index = torch.cat([ei + b * N for b in range(B)], dim=1)
    return Batch(x=torch.randn(B * N, F_in, generator=g, dtype=dtype),
                 edge_index=edge_index,
                 edge_attr=torch.randn(B * R, Fe, generator=g, dtype=dtype),
                 y_x=torch.randn(B * N, generator=g, dtype=dtype),
                 batch=torch.arange(B).repeat_interleave(N),
                 ptr=torch.arange(B + 1) * N, num_graphs=B)
