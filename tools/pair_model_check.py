"""GPU box: one training-harness step (GATModel + MSE, scale_up = 100) in p_format 0 and 1: per-parameter gradient differences."""
import sys, torch
sys.path.insert(0, ".")
import spotv2net_b200 as sv
from spotv2net_b200 import gat_conv
from spotv2net_b200.train import _scaled
from oracle import synth
dev = torch.device("cuda", 0)
N, L, T = 30, 3, 46
vol, vv = synth.synthetic_matrices(T, N, seed=77)
ds = sv.WindowDataset(vol, vv, seq_length=L, device=dev, drop_first=2)
rel = lambda a, b: ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()
for scale in (None, 100.0):
    res = {}
    for pf in (0, 1):
        gat_conv.P_FORMAT = pf
        gat_conv._KEEP = {}
        torch.manual_seed(5)
        model = sv.GATModel(N * L, 3 * L, 3, 1, [16], concat_heads=True).to(dev)
        bt = _scaled(ds.collate(list(range(8))), scale)
        out = model(bt)
        loss = torch.nn.functional.mse_loss(out, bt.y_x)
        loss.backward()
        res[pf] = dict(loss=loss.detach(), out=out.detach(), **{k: p.grad.clone() for k, p in model.named_parameters()})
    print("scale", scale, {k: f"{rel(res[1][k], res[0][k]):.1e}" for k in res[0]})
# element-level: the ds | dd block of the scaled step
res = {}
for pf in (0, 1):
    gat_conv.P_FORMAT = pf
    gat_conv._KEEP = {}
    torch.manual_seed(5)
    model = sv.GATModel(N * L, 3 * L, 3, 1, [16], concat_heads=True).to(dev)
    bt = _scaled(ds.collate(list(range(8))), 100.0)
    torch.nn.functional.mse_loss(model(bt), bt.y_x).backward()
    k = gat_conv._KEEP
    hp, blk, H = k["head_pitch"], k["dp_blk"], 3
    full = k["dP16"][0].float() + k["dP16"][1].float()
    res[pf] = (full[:, H * hp:H * hp + 2 * H] * blk[3]).clone(), torch.stack([full[:, h * hp:h * hp + 16] for h in range(H)], 1) * blk[2]
sd0, dP0 = res[0]
sd1, dP1 = res[1]
print("dP rel diff", ((dP1 - dP0).abs().max() / dP0.abs().max()).item())
print("ds rel diff", ((sd1[:, :3] - sd0[:, :3]).abs().max() / sd0[:, :3].abs().max()).item(), "dd rel diff", ((sd1[:, 3:] - sd0[:, 3:]).abs().max() / sd0[:, 3:].abs().max()).item())
print("max|ds|", sd0[:, :3].abs().max().item(), "max|dd|", sd0[:, 3:].abs().max().item())
bad = ((sd1[:, 3:] - sd0[:, 3:]).abs() / sd0[:, 3:].abs().max() > 1e-3).nonzero().tolist()
print("dd entries off by > 1e-3 of max|dd|:", len(bad), [(i, h, sd0[i, 3 + h].item(), sd1[i, 3 + h].item()) for i, h in bad[:10]])
