"""GPU box: run the hot path at B=4096 and print where each warp role of the tcgen05 attention backward (attn_bwd3.cu)
spends its cycles (spotv2_diag_counters entries 16..31), per CTA per graph."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

structured = "--structured" in sys.argv
dev = torch.device("cuda", 0)
hp = bench.HotPath(4096, dev, 1234, structured=structured)
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
ms = sum(a.elapsed_time(b) for e in ev for (n0, a), (n1, b) in zip(e[:-1], e[1:]) if n1 == "attn_bwd") / n
gpc = 4096 / 148
print(f"attn_bwd phase {ms:.3f} ms/launch  ({'structured' if structured else 'edge rows'}); kcycles per CTA per graph ({gpc:.1f} graphs per CTA):")
names = ["producer: slot empty", "MMA: lo ready (A)", "epilogue: dP^T ready", "MMA: alpha ready", "MMA: phase D operands", "stream: A/G slot full",
         "stream: lo buffer free", "stream: lo pass", "stream: dz' ready", "stream: V slot full", "stream: dv arithmetic",
         "softmax: T slot/tile/alpha free", "epilogue: work", "softmax: softmax", "softmax: dalpha ready", "softmax: dump + backward"]
for k, nm in enumerate(names):
    print(f"  {nm:34s} {buf[16 + k] / n / 148 / gpc / 1e3:8.2f}")
