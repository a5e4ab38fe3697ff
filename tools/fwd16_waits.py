"""(Counters exist only in a bring-up build: `SPOTV2_BRINGUP=1 python -m spotv2net_b200.build -f` before shipping to the box.)
GPU box: where each warp role of the p_format 1 forward spends its cycles (spotv2_diag_counters), materialised and structured."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
dev = torch.device("cuda", 0)
names = ["A: edge ring wait", "A: tile buffer free", "B: edge terms ready", "B: P tiles", "P: slot free", "A total", "B total", "P total",
         "A: logits arithmetic", "B: softmax", "B: conversions", "B: s|d tile wait", "A: chunk barriers", "A: edge-term copy-out"]
for structured in (False, True):
    hp = bench.HotPath(4096, dev, 1234, structured=structured)
    buf = (C.c_ulonglong * 32)()
    for _ in range(3):
        hp.step()
    hp.lib.spotv2_diag_counters(buf, 1)
    n, ev = 5, []
    for _ in range(n):
        e = []
        hp.step(timed_events=e)
        ev.append(e)
    torch.cuda.synchronize()
    hp.lib.spotv2_diag_counters(buf, 1)
    ms = sum(a.elapsed_time(b) for e in ev for (n0, a), (n1, b) in zip(e[:-1], e[1:]) if n1 == "attn_fwd") / n
    print(f"attn_fwd {ms:.3f} ms/launch ({'structured' if structured else 'edge rows'}); kcycles per CTA per graph (27.7 graphs per CTA):")
    for k, nm in enumerate(names):
        print(f"  {nm:24s} {buf[k] / n / 148 / 27.68 / 1e3:8.2f}")
    del hp
    torch.cuda.empty_cache()
