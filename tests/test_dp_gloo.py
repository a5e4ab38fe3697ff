"""N > 1 path on CPU: world size 2 over gloo.  Gradients all-reduced through the flat arena must equal
the single-process gradient on the concatenated batch (distributed tests)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pyg_gat, synth
from spotv2net_b200.dp import FlatGradArena, shard_snapshots

CFG = dict(num_node_features=18, num_edge_features=9, num_heads=2, output_node_channels=1, dim_hidden_layers=[8, 4],
           concat_heads=True)
N, L, B = 6, 3, 8


def _batch(idx):
    vol, vv = synth.synthetic_matrices(L + B + 2, N, seed=5)
    return synth.make_batch(vol, vv, [int(i) for i in idx], L)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    model = pyg_gat.OracleGATModel(**CFG)
    arena = FlatGradArena(model.parameters())
    bt = _batch(shard_snapshots(torch.arange(B), rank, world))
    arena.zero()
    torch.nn.functional.mse_loss(model(bt), bt.y_x).backward()
    assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(arena.params, arena.views))
    arena.all_reduce()
    if rank == 0:
        torch.save(arena.flat.clone(), out)
    dist.destroy_process_group()


def test_flat_arena_allreduce_matches_single_process(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "flat.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    flat = torch.load(out)
    torch.manual_seed(0)
    model = pyg_gat.OracleGATModel(**CFG)
    bt = _batch(torch.arange(B))
    torch.nn.functional.mse_loss(model(bt), bt.y_x).backward()
    ref = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    assert flat.shape == ref.shape
    assert (flat - ref).abs().max() <= 1e-5 * ref.abs().max()


def test_shard_snapshots_is_a_partition():
    idx = torch.arange(12)
    parts = [shard_snapshots(idx, r, 4) for r in range(4)]
    assert sorted(torch.cat(parts).tolist()) == idx.tolist()
    with pytest.raises(ValueError):
        shard_snapshots(torch.arange(10), 0, 4)


def test_arena_survives_zero_grad_set_to_none():
    lin = torch.nn.Linear(3, 2)
    arena = FlatGradArena(lin.parameters())
    lin(torch.ones(1, 3)).sum().backward()
    g = arena.flat.clone()
    lin.zero_grad(set_to_none=True)
    arena.zero()
    lin(torch.ones(1, 3)).sum().backward()
    assert torch.equal(arena.flat, g) and arena.nbytes == 4 * 8
