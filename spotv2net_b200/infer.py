"""Evaluation / analysis helpers: what /root/reference/6_results.ipynb does with a trained model.

* ``evaluate``: the batched no-grad loop of the notebook's "Evaluate on the train / validation / test set" cells
  (code lines 294-330, 332-372, 402-440 of the notebook's concatenated source): de-standardised predictions
  ``model(data) * std + mean``, de-standardised targets, the naive benchmark ``data.x[:, 0]`` and the mean batch MSE.
* ``attention_weights``: the notebook's modified ``GATModel.forward`` (its lines 173-190) calls every layer with
  ``return_attention_weights=True`` and keeps the per-layer ``(edge_index, alpha)``; here that is
  ``model.collect_attention = True`` (see ``GATModel``), wrapped for one batch.
All arithmetic runs in the library's CUDA kernels; nothing here falls back to the CPU.
"""
from __future__ import annotations

from typing import Iterable

import torch


@torch.no_grad()
def evaluate(model, loader: Iterable, mean: float = 0.0, std: float = 1.0) -> dict:
    """Returns ``{"preds", "actual", "naive", "mse"}``: 1-D tensors concatenated over the loader's batches (reshape with
    ``.view(-1, N)`` as the notebook does) and the mean of the per-batch MSE losses."""
    was_training = model.training
    model.eval()
    crit = torch.nn.MSELoss()
    preds, actual, naive, total, nb = [], [], [], 0.0, 0
    for data in loader:
        y_hat = model(data) * std + mean
        y = data.y_x * std + mean
        preds.append(y_hat)
        actual.append(y)
        naive.append(data.x[:, 0])
        total += crit(y_hat, y).item()
        nb += 1
    model.train(was_training)
    return {"preds": torch.cat(preds), "actual": torch.cat(actual), "naive": torch.cat(naive), "mse": total / max(nb, 1)}


@torch.no_grad()
def attention_weights(model, data) -> list:
    """Per-layer ``(edge_index_with_self_loops [2, E'], alpha [E', H])`` of one batch, PyG ordering (real edges in batch
    order, then one self loop per node)."""
    was_training, old = model.training, getattr(model, "collect_attention", False)
    model.eval()
    model.collect_attention = True
    try:
        model(data)
        return list(model.attention_weights)
    finally:
        model.collect_attention = old
        model.train(was_training)
