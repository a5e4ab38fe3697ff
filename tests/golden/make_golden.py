"""Generates tests/golden/*.npz from the fp64 edge-list oracle (oracle/pyg_gat.py).

The reference ships no golden vectors and PyG cannot be imported in this image
(SURVEY.md §8c), so these fixtures are produced by the PyG-order restatement in
float64 and pin (a) the oracle against accidental edits and (b) the CUDA path on
the GPU box, where /root/reference and this script's inputs do not exist.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyg_gat, synth  # noqa: E402

CASES = {
    # name: (B, N, L or (F, Fe), H, C, concat, slope, structured, weight_scale)
    "default_mean":   dict(B=3, N=30, L=4, H=6, C=20, concat=False, slope=0.2, structured=True, wscale=1.0),
    "concat_heads":   dict(B=2, N=30, L=3, H=4, C=16, concat=True, slope=0.2, structured=True, wscale=1.0),
    "single_head":    dict(B=2, N=30, F=24, Fe=6, H=1, C=10, concat=True, slope=0.05, structured=False, wscale=1.0),
    "odd_channels":   dict(B=4, N=7, F=9, Fe=5, H=3, C=5, concat=False, slope=0.5, structured=False, wscale=2.0),
    "stress_weights": dict(B=2, N=12, F=16, Fe=8, H=7, C=12, concat=False, slope=0.8, structured=False, wscale=4.0),
}


def build_case(name, cfg, seed):
    torch.manual_seed(seed)
    B, N, H, C = cfg["B"], cfg["N"], cfg["H"], cfg["C"]
    if cfg["structured"]:
        L = cfg["L"]
        vol, vv = synth.synthetic_matrices(L + B + 2, N, seed=1234 + seed)
        batch = synth.make_batch(vol, vv, list(range(B)), L)
        batch.x, batch.edge_attr = batch.x.double(), batch.edge_attr.double()
    else:
        batch = synth.random_complete_batch(B, N, cfg["F"], cfg["Fe"], seed=seed, dtype=torch.float32)
        batch.x, batch.edge_attr = batch.x.double(), batch.edge_attr.double()
    Fin, Fe = batch.x.shape[1], batch.edge_attr.shape[1]
    layer = pyg_gat.OracleGATConv(Fin, C, heads=H, concat=cfg["concat"], negative_slope=cfg["slope"], edge_dim=Fe).double()
    with torch.no_grad():
        for p in layer.parameters():
            p.mul_(cfg["wscale"])
        layer.bias.normal_()
        for p in layer.parameters():            # make every stored value exactly representable in fp32
            p.copy_(p.float().double())
    out, (ei2, alpha) = layer(batch.x, batch.edge_index, batch.edge_attr, return_attention_weights=True)
    dout = torch.randn(out.shape, dtype=torch.float32).double()
    out.backward(dout)
    f32 = lambda t: t.detach().to(torch.float32).numpy()
    f64 = lambda t: t.detach().numpy()
    return dict(
        meta=np.array([B, N, Fin, Fe, H, C, int(cfg["concat"])], dtype=np.int64), slope=np.float64(cfg["slope"]),
        x=f32(batch.x), edge_index=batch.edge_index.numpy(), edge_attr=f32(batch.edge_attr), dout=f32(dout),
        lin_weight=f32(layer.lin_src.weight), att_src=f32(layer.att_src), att_dst=f32(layer.att_dst),
        lin_edge_weight=f32(layer.lin_edge.weight), att_edge=f32(layer.att_edge), bias=f32(layer.bias),
        # float64 results; every input and weight above is exactly representable in float32
        out=f64(out), alpha=f64(alpha), edge_index_with_loops=ei2.numpy(),
        g_lin_weight=f64(layer.lin_src.weight.grad), g_att_src=f64(layer.att_src.grad),
        g_att_dst=f64(layer.att_dst.grad), g_lin_edge_weight=f64(layer.lin_edge.weight.grad),
        g_att_edge=f64(layer.att_edge.grad), g_bias=f64(layer.bias.grad))


if __name__ == "__main__":
    for k, (name, cfg) in enumerate(CASES.items()):
        data = build_case(name, cfg, seed=k)
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **data)
        print(name, {a: v.shape for a, v in data.items() if hasattr(v, "shape") and v.ndim}, os.path.getsize(path))
