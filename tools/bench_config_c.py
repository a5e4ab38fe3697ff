#!/usr/bin/env python
"""BASELINE config C: wide multi-head variant (8 heads, hidden [256, 256], concat -> layer 0: 1260 -> 8x256 concat,
layer 1: 2048 -> 256 head mean; utils/models.py:90-104), batch 4096 x 30-node graphs, forward + backward through
the public GATModel API.  Timed in both precisions: "fp32" (1e-5 parity, 3 fp16 products per projection) and "half"
(one fp16 tensor-core product, fp32 accumulate - the reduced-precision variant the config names).  One JSON line."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spotv2net_b200 as sv  # noqa: E402


def main(B=4096, steps=5, warmup=3):
    dev = torch.device("cuda", 0)
    N, L = 30, 42
    g = torch.Generator(device=dev).manual_seed(1234)
    T = B + L + 1
    mats = []
    for _ in range(2):
        a = torch.randn(T, N, N, device=dev, generator=g)
        mats.append((a + a.transpose(1, 2)) / 2 ** 0.5)
    ds = sv.WindowDataset(mats[0], mats[1], seq_length=L, device=dev, drop_first=0)
    bt = ds.collate(torch.arange(B))
    torch.manual_seed(0)
    model = sv.GATModel(N * L, 3 * L, 8, 1, dim_hidden_layers=[256, 256], concat_heads=True).to(dev)
    res = {}
    for prec in ("fp32", "half"):
        model.set_precision(prec)

        def step():
            model.zero_grad(set_to_none=True)
            loss = torch.nn.functional.mse_loss(model(bt), bt.y_x)
            loss.backward()
            return loss
        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        res[prec] = {"ms_per_step": ms, "graphs_per_s": B / (ms * 1e-3), "loss": loss.item()}
    print(json.dumps({"workload": "BASELINE configs[2]: 8 heads, hidden [256,256], concat, two GAT layers + Linear + MSE, "
                                  "batch 4096, GATModel forward + autograd backward", "steps": steps, "warmup": warmup,
                      "results": res}))


if __name__ == "__main__":
    main()
