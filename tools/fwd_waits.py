"""(Counters exist only in a bring-up build: `SPOTV2_BRINGUP=1 python -m spotv2net_b200.build -f` before shipping to the box.)
GPU box: run the forward attention kernel at B=4096 and print where each warp role spends its cycles
(spotv2_diag_counters), plus the single failing odd-shape case under CUDA_LAUNCH_BLOCKING."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import spotv2net_b200 as sv

dev = torch.device("cuda", 0)
hp = bench.HotPath(4096, dev, 1234)
lib = hp.lib
buf = (C.c_ulonglong * 32)()
for _ in range(3):
    hp.step()
lib.spotv2_diag_counters(buf, 1)
n = 5
ev = []
for _ in range(n):
    e = []
    hp.step(timed_events=e)
    ev.append(e)
torch.cuda.synchronize()
lib.spotv2_diag_counters(buf, 1)
names = ["ring_full(A)", "tile_empty(A)", "tile_full(B)", "ptile_full(B)", "ptile_empty(P)", "roleA_total", "roleB_total", "roleP_total", "A: mma loop", "A: logits call", "A: bar.sync"]
ctas = 148
ms = sum(a.elapsed_time(b) for e in ev for (n0, a), (n1, b) in zip(e[:-1], e[1:]) if n1 == "attn_fwd") / n
print(f"attn_fwd {ms:.3f} ms/launch")
for k, nm in enumerate(names):
    print(f"{nm:16s} {buf[k] / n / ctas / 1e3:10.1f} kcycles per CTA per launch")

msb = sum(a.elapsed_time(b) for e in ev for (n0, a), (n1, b) in zip(e[:-1], e[1:]) if n1 == "attn_bwd") / n
print(f"attn_bwd {msb:.3f} ms/launch (pipelined kernel, 1 CTA per SM: per-CTA phase totals; /27.7 = per graph)")
for k, nm in enumerate(["L logits", "S softmax", "A dalpha+smx bwd", "ds out", "D dP", "V dv"]):
    print(f"{nm:16s} {buf[16 + k] / n / 148 / 1e3:10.1f} kcycles per CTA per launch, of which waiting for data {buf[22 + k] / n / 148 / 1e3:10.1f}")
