"""Training harness: the caller of the hot path, restated from the reference's ``train()``
(/root/reference/5_train_SpotV2Net.py:23-203) on top of the B200 model and the device-side collation.

Same hyper-parameter keys (``config/GNN_param.yaml``), same split (first ``split_proportion`` of the
samples train, the rest test, in index order), same loss / optimiser choices, same "save the state_dict
when the test loss improves by ``tolerance``" rule, same artefact names.  Defects of the fork are routed
around, not reproduced (the ``pdb.set_trace()``, the undefined ``mae``/``rmse`` print).

With ``torch.distributed`` initialised the loop is data parallel over graph snapshots: every rank takes
its shard of each global batch and the flat gradient arena is all-reduced once per step (``dp.py``).
"""
from __future__ import annotations

import math
import os
import sys
from typing import Optional

import numpy as np
import torch
import torch.distributed as dist
import yaml

from .data import WindowDataset, WindowLoader, load_matrix_stack
from .dp import FlatGradArena, shard_snapshots
from .gat_conv import WindowSource
from .models import GATModel


def _dist_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def build_model(p: dict, num_node_features: int, num_edge_features: int) -> GATModel:
    """5_train_SpotV2Net.py:100-110."""
    if p.get("modeltype", "gat") != "gat":
        raise ValueError("only modeltype 'gat' is in scope (RecurrentGCN is dead code in the reference)")
    return GATModel(num_node_features=num_node_features, num_edge_features=num_edge_features,
                    num_heads=p["num_heads"], output_node_channels=p["output_node_channels"],
                    dim_hidden_layers=p["dim_hidden_layers"], dropout_att=p["dropout_att"], dropout=p["dropout"],
                    activation=p["activation"], concat_heads=p["concat_heads"], negative_slope=p["negative_slope"],
                    standardize=p["standardize"])


def build_optimizer(p: dict, params):
    """5_train_SpotV2Net.py:127-134 (unknown name: print + exit, like the reference)."""
    name = p["optimizer"]
    if name == "adam":
        return torch.optim.Adam(params, lr=p["learning_rate"])
    if name == "adamw":
        return torch.optim.AdamW(params, lr=p["learning_rate"])
    if name == "rmsprop":
        return torch.optim.RMSprop(params, lr=p["learning_rate"])
    print("Choose an available optimizer")
    sys.exit()


def _scaled(batch, scale):
    """5_train_SpotV2Net.py:145-147 / 173-175: x, edge_attr and the targets times ``scale_up``.  A structured batch
    has no edge_attr to scale: its edge features are the vol-of-vol windows, so the window stack is scaled instead."""
    if scale:
        batch.x = batch.x * scale
        if batch.edge_attr is not None:
            batch.edge_attr = batch.edge_attr * scale
        win = getattr(batch, "spot_windows", None)
        if win is not None:
            ms = None if win.mat_stats is None else win.mat_stats * torch.tensor([scale, scale * scale, scale, scale * scale],
                                                                                dtype=torch.float64, device=win.mat_stats.device)
            batch.spot_windows = WindowSource(win.volvol * scale, win.t0, win.L, win.checked, ms)
        batch.y_x = batch.y_x * scale
    return batch


def train(seed: Optional[int] = None, trial=None, p: Optional[dict] = None, *, config_path: str = "config/GNN_param.yaml",
          vol=None, volvol=None, device="cuda", output_root: str = "output", drop_first: Optional[int] = None,
          verbose: bool = True):
    """Returns (train_losses, test_losses).  ``vol`` / ``volvol``: optional [T, N, N] arrays replacing the H5
    files named in the config (the TAQ data is not public)."""
    if p is None:
        with open(config_path, "r") as f:
            p = yaml.safe_load(f)
        p["seed"] = seed
    elif seed is not None:
        p = dict(p, seed=seed)
    rank, world = _dist_info()
    if trial is not None:
        folder = os.path.join(output_root, "{}_{}".format(p["modelname"], "optuna"), str(trial.number))
    else:
        folder = os.path.join(output_root, "{}_{}".format(p["modelname"], p["seq_length"]))
    if rank == 0:
        os.makedirs(folder, exist_ok=True)
        with open(os.path.join(folder, "GNN_param.yaml"), "w") as f:
            yaml.dump(p, f)

    torch.manual_seed(p["seed"])
    np.random.seed(p["seed"])
    torch.cuda.manual_seed_all(p["seed"])

    if vol is None:
        vol, volvol = load_matrix_stack(p["volfile"]), load_matrix_stack(p["volvolfile"])
    kw = {} if drop_first is None else {"drop_first": drop_first}
    if p["output_node_channels"] > 1:          # 5_train_SpotV2Net.py:66-76: the multi-output dataset, K = output channels
        kw["future_steps"] = p["output_node_channels"]
    dataset = WindowDataset(vol, volvol, seq_length=p["seq_length"], device=device, **kw)
    train_size = int(p["split_proportion"] * len(dataset))
    train_set, test_set = dataset[:train_size], dataset[train_size:]
    gen = torch.Generator().manual_seed(int(p["seed"]))          # same shuffle on every rank
    train_loader = WindowLoader(train_set, batch_size=p["batch_size"], shuffle=True, generator=gen)
    test_loader = WindowLoader(test_set, batch_size=p["batch_size"], shuffle=False)

    model = build_model(p, dataset.num_node_features, dataset.num_edge_features).to(device)
    criterion = torch.nn.MSELoss()
    optimizer = build_optimizer(p, model.parameters())
    arena = FlatGradArena(model.parameters()) if world > 1 else None

    def run_batch(loader_set, order, training):
        idx = order
        if world > 1 and training and idx.numel() % world == 0:
            idx = shard_snapshots(idx, rank, world)
        batch = _scaled(loader_set.collate(idx), p.get("scale_up"))
        y_hat = model(batch)
        return criterion(y_hat, batch.y_x)

    train_losses, test_losses = [], []
    prev_test = float("inf")
    for epoch in range(p["num_epochs"]):
        model.train()
        total = torch.zeros((), device=device, dtype=torch.float64)      # summed on the device: no sync per step
        n_train = len(train_set)
        order = torch.randperm(n_train, generator=gen)
        steps = 0
        for s in range(0, n_train, p["batch_size"]):
            loss = run_batch(train_set, order[s:s + p["batch_size"]], True)
            if arena is not None:
                arena.zero()
            else:
                optimizer.zero_grad()
            loss.backward()
            if arena is not None:
                arena.all_reduce()
            optimizer.step()
            total += loss.detach().double()
            steps += 1
        if world > 1:
            # a rank's loss is the mean over ITS shard of the global batch; the reported figure is the mean over ranks
            # (equal shards), so that it equals the single-process number (batches that did not split evenly were run
            # whole by every rank: the mean over ranks is then that same value)
            dist.all_reduce(total)
            total /= world
        avg_train = total.item() / max(steps, 1)
        train_losses.append(avg_train)

        model.eval()
        tot_test, n_test_batches = 0.0, 0
        with torch.no_grad():
            for batch in test_loader:
                batch = _scaled(batch, p.get("scale_up"))
                tot_test += criterion(model(batch), batch.y_x).item()
                n_test_batches += 1
        avg_test = tot_test / max(n_test_batches, 1)
        test_losses.append(avg_test)

        if rank == 0 and (epoch == 0 or avg_test + float(p["tolerance"]) < prev_test):
            prev_test = avg_test
            torch.save(model.state_dict(), os.path.join(folder, "{}_weights_seed_{}.pth".format(p["modelname"], p["seed"])))
        if verbose and rank == 0:
            print(f"Epoch: {epoch + 1}/{p['num_epochs']}, Train Loss: {avg_train:.10f}, Test Loss: {avg_test:.10f}, "
                  f"Train RMSE: {math.sqrt(avg_train):.10f}, Test RMSE: {math.sqrt(avg_test):.10f}")

    if rank == 0:
        np.save(os.path.join(folder, "train_losses_seed_{}.npy".format(p["seed"])), np.array(train_losses))
        np.save(os.path.join(folder, "test_losses_seed_{}.npy".format(p["seed"])), np.array(test_losses))
    return train_losses, test_losses
