"""Worker of tests/test_dp_nccl.py (launched by torch.distributed.run, one rank per GPU, NCCL).

Two checks of the N > 1 path on real hardware (SURVEY.md:279):
 1. spotv2net_b200.GATModel + FlatGradArena: every rank runs its shard of ONE global batch through the CUDA hot path,
    the flat arena is all-reduced (average) and must equal the single-GPU gradient on the concatenated batch;
 2. spotv2net_b200.train.train with torch.distributed initialised (its DP branch) must reproduce the single-process
    loss curve and final weights on the same global batches.
Rank 0 writes the errors as JSON to argv[1]; the test asserts on them.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

import spotv2net_b200 as sv
from spotv2net_b200.dp import FlatGradArena, shard_snapshots
from spotv2net_b200.train import train
from oracle import synth


def relerr(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


def main():
    out_path = sys.argv[1]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    res = {"world": world}

    # ---- 1. one global batch, default layer geometry in the first layer -------------------------------------------
    N, L, B = 30, 42, 16 * world
    vol, vv = synth.synthetic_matrices(L + B + 2, N, seed=21)
    kw = dict(num_node_features=N * L, num_edge_features=3 * L, num_heads=6, output_node_channels=1, dim_hidden_layers=[500])
    ds = sv.WindowDataset(vol, vv, seq_length=L, device=dev, drop_first=0)
    torch.manual_seed(0)
    model = sv.GATModel(**kw).to(dev)
    arena = FlatGradArena(model.parameters())
    bt = ds.collate(shard_snapshots(torch.arange(B), rank, world))
    arena.zero()
    torch.nn.functional.mse_loss(model(bt), bt.y_x).backward()
    assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(arena.params, arena.views))
    arena.all_reduce()
    torch.cuda.synchronize()
    torch.manual_seed(0)
    single = sv.GATModel(**kw).to(dev)
    full = ds.collate(torch.arange(B))
    torch.nn.functional.mse_loss(single(full), full.y_x).backward()
    ref = torch.cat([p.grad.reshape(-1) for p in single.parameters()])
    res["arena_vs_single_gpu"] = relerr(arena.flat, ref)
    per = {}
    off = 0
    for (k, p) in single.named_parameters():
        per[k] = relerr(arena.flat[off:off + p.numel()], p.grad.reshape(-1))
        off += p.numel()
    res["per_parameter"] = per
    # every rank holds the same reduced arena
    g = [torch.empty_like(arena.flat) for _ in range(world)]
    dist.all_gather(g, arena.flat)
    res["ranks_identical"] = all(torch.equal(g[0], t) for t in g)

    # ---- 2. train(): DP branch against a single-process run -------------------------------------------------------
    N2, L2, T2 = 30, 3, 46
    vol2, vv2 = synth.synthetic_matrices(T2, N2, seed=77)
    p = dict(modelname="t", modeltype="gat", seq_length=L2, batch_size=8, dim_hidden_layers=[16], output_node_channels=1,
             num_heads=3, concat_heads=True, activation="relu", optimizer="adam", learning_rate=1e-3, negative_slope=0.2,
             dropout_att=0.0, dropout=0.0, standardize=False, num_epochs=2, tolerance=1e-9, split_proportion=0.8,
             scale_up=None, seed=5)
    root = os.path.join(os.path.dirname(out_path), "dp")
    tr_dp, te_dp = train(p=dict(p), vol=vol2, volvol=vv2, device=dev, output_root=root, drop_first=2, verbose=False)
    dist.barrier()
    if rank == 0:
        sd_dp = torch.load(os.path.join(root, "t_3", "t_weights_seed_5.pth"))
        # single process: same code with the process group hidden
        import spotv2net_b200.train as tmod
        saved = tmod._dist_info
        tmod._dist_info = lambda: (0, 1)
        try:
            root1 = os.path.join(os.path.dirname(out_path), "single")
            tr_1, te_1 = train(p=dict(p), vol=vol2, volvol=vv2, device=dev, output_root=root1, drop_first=2, verbose=False)
        finally:
            tmod._dist_info = saved
        sd_1 = torch.load(os.path.join(root1, "t_3", "t_weights_seed_5.pth"))
        res["train_loss_rel"] = max(abs(a - b) / abs(b) for a, b in zip(tr_dp + te_dp, tr_1 + te_1))
        res["train_weights_rel"] = max(relerr(sd_dp[k], sd_1[k]) for k in sd_1)
        with open(out_path, "w") as f:
            json.dump(res, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
