"""GPU box: gradients of one default-geometry layer in p_format 1 against p_format 0 (same inputs), several batch sizes."""
import sys, torch
sys.path.insert(0, ".")
import spotv2net_b200 as sv
from spotv2net_b200 import gat_conv
dev = torch.device("cuda", 0)
N, L, H, C_ = 30, 42, 6, 500
Fin, Fe = N * L, 3 * L
rel = lambda a, b: ((a - b).abs().max() / b.abs().max()).item()
for B in [int(a) for a in sys.argv[1:]] or [8, 64, 592, 4096]:
    torch.manual_seed(11)
    layer = sv.GATConv(Fin, C_, heads=H, concat=False, edge_dim=Fe).to(dev)
    g = torch.Generator(device=dev).manual_seed(3)
    x = torch.randn(B * N, Fin, device=dev, generator=g, requires_grad=True)
    ea = torch.randn(B * N * (N - 1), Fe, device=dev, generator=g)
    dout = torch.randn(B * N, C_, device=dev, generator=g)
    ei, _ = sv.batched_topology(B, N, dev)
    res = {}
    for pf in (0, 1):
        gat_conv.P_FORMAT = pf
        layer.zero_grad(); x.grad = None
        out = layer(x, ei, ea)
        out.backward(dout)
        res[pf] = dict(out=out.detach().clone(), dx=x.grad.clone(), **{k: p.grad.clone() for k, p in layer.named_parameters()})
    print(B, {k: f"{rel(res[1][k], res[0][k]):.1e}" for k in res[0]})
    if B > 592:
        e = (res[1]["dx"] - res[0]["dx"]).abs().view(B, -1).amax(1) / res[0]["dx"].abs().max()
        bad = (e > 1e-5).nonzero().flatten().tolist()
        print("   graphs with dx error > 1e-5:", len(bad), bad[:40])
