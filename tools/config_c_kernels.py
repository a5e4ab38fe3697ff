"""GPU box: per-kernel device time of one config-C training step (torch.profiler), half and fp32 modes."""
import sys, torch
sys.path.insert(0, ".")
import spotv2net_b200 as sv
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda", 0)
B, N, L, H, Cc = 4096, 30, 42, 8, 256
g = torch.Generator(device=dev).manual_seed(1234)
mats = []
for _ in range(2):
    a = torch.randn(B + L + 1, N, N, device=dev, generator=g)
    mats.append((a + a.transpose(1, 2)) / 2 ** 0.5)
ds = sv.WindowDataset(mats[0], mats[1], seq_length=L, device=dev, drop_first=0)
bt = ds.collate(torch.arange(B))
torch.manual_seed(0)
model = sv.GATModel(N * L, 3 * L, H, 1, dim_hidden_layers=[Cc, Cc], concat_heads=True).to(dev)
def step():
    model.zero_grad(set_to_none=True)
    loss = torch.nn.functional.mse_loss(model(bt), bt.y_x)
    loss.backward()
for prec in ("half", "fp32"):
    model.set_precision(prec)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
    tot = sum(e.device_time_total for e in rows)
    print(f"== {prec}: {tot * 1e-3:.2f} ms of kernels")
    for e in rows[:16]:
        print(f"  {e.device_time_total * 1e-3:7.3f} ms x{e.count:<2d} {e.key[:110]}")
