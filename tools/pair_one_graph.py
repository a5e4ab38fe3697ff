"""GPU box: isolate the graph of the B = 1184 batch whose p_format 1 gradient is off; compare dP | ds | dd element by element."""
import sys, torch
sys.path.insert(0, ".")
import spotv2net_b200 as sv
from spotv2net_b200 import gat_conv
dev = torch.device("cuda", 0)
N, L, H, C_ = 30, 42, 6, 500
Fin, Fe = N * L, 3 * L
B = 1184
torch.manual_seed(11)
layer = sv.GATConv(Fin, C_, heads=H, concat=False, edge_dim=Fe).to(dev)
g = torch.Generator(device=dev).manual_seed(3)
X = torch.randn(B * N, Fin, device=dev, generator=g)
EA = torch.randn(B * N * (N - 1), Fe, device=dev, generator=g)
DO = torch.randn(B * N, C_, device=dev, generator=g)
graphs = [868]
idx = torch.tensor(graphs, device=dev)
x = X.view(B, N, Fin)[idx].reshape(-1, Fin).clone().requires_grad_()
ea = EA.view(B, N * (N - 1), Fe)[idx].reshape(-1, Fe).contiguous()
dout = DO.view(B, N, C_)[idx].reshape(-1, C_).contiguous()
ei, _ = sv.batched_topology(len(graphs), N, dev)
res = {}
for pf in (0, 1):
    gat_conv.P_FORMAT = pf
    gat_conv._KEEP = {}
    layer.zero_grad(); x.grad = None
    layer(x, ei, ea).backward(dout)
    k = gat_conv._KEEP
    hp, blk = k["head_pitch"], k["dp_blk"]
    full = k["dP16"][0].float() + k["dP16"][1].float()
    dP = torch.stack([full[:, h * hp:h * hp + C_] for h in range(H)], 1) * blk[2]          # [N, H, C]
    sd = full[:, H * hp:H * hp + 2 * H] * blk[3]                                           # [N, 2H]
    res[pf] = (dP, sd)
dP0, sd0 = res[0]
dP1, sd1 = res[1]
e = (dP1 - dP0).abs() / dP0.abs().max()
print("dP worst", e.max().item(), "rows/heads with error > 1e-5:", (e.amax(2) > 1e-5).nonzero().tolist()[:20])
es = (sd1 - sd0).abs() / sd0.abs().max()
print("ds|dd worst", es.max().item(), (es > 1e-5).nonzero().tolist()[:20])
print("ds|dd values at bad entries", [(i, j, sd0[i, j].item(), sd1[i, j].item()) for i, j in (es > 1e-5).nonzero().tolist()[:8]])
