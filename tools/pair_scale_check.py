"""GPU box: p_format 1 against p_format 0 and the fp64 oracle on a small concat layer with inputs scaled by 1 and by 100."""
import sys, torch
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import spotv2net_b200 as sv
from spotv2net_b200 import gat_conv
from oracle import pyg_gat, synth
dev = torch.device("cuda", 0)
N, L, H, C_, B = 30, 3, 3, 16, 8
Fin, Fe = N * L, 3 * L
rel = lambda a, b: ((a.double().cpu() - b.double().cpu()).abs().max() / b.double().abs().max()).item()
vol, vv = synth.synthetic_matrices(L + B + 2, N, seed=77)
bt = synth.make_batch(vol, vv, list(range(B)), L)
for scale in (1.0, 100.0):
    torch.manual_seed(5)
    ref = pyg_gat.OracleGATConv(Fin, C_, heads=H, concat=True, edge_dim=Fe).double()
    layer = sv.GATConv(Fin, C_, heads=H, concat=True, edge_dim=Fe)
    layer.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
    layer.to(dev)
    x64 = (bt.x * scale).double().requires_grad_()
    ea64 = (bt.edge_attr * scale).double()
    g = torch.Generator().manual_seed(1)
    dout = torch.randn(B * N, H * C_, generator=g)
    o64 = ref(x64, bt.edge_index, ea64)
    o64.backward(dout.double())
    g64 = dict(out=o64.detach(), dx=x64.grad, **{k: p.grad for k, p in ref.named_parameters()})
    for pf in (0, 1):
        gat_conv.P_FORMAT = pf
        layer.zero_grad()
        x = (bt.x * scale).to(dev).requires_grad_()
        out = layer(x, bt.edge_index.to(dev), (bt.edge_attr * scale).to(dev))
        out.backward(dout.to(dev))
        mine = dict(out=out.detach(), dx=x.grad, **{k: p.grad for k, p in layer.named_parameters()})
        print(f"scale {scale:5.0f} p_format {pf}:", {k: f"{rel(mine[k], g64[k]):.1e}" for k in g64 if k in mine})
