"""Probe for DESIGN.md: is the forward projection GEMM bound by operand fill or by the tensor pipe?
gemm_algo 3 issues one product instead of three and loads only the hi tiles (half the fill bytes): a fill-bound kernel
should take ~1/2 of the 3-product time, a tensor-bound one ~1/3."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench

dev = torch.device("cuda", 0)
hp = bench.HotPath(4096, dev, 1234)
for algo in (0, 3):
    hp.desc.gemm_algo = algo
    for _ in range(3):
        hp.step()
    torch.cuda.synchronize()
    tot = {}
    n = 5
    for _ in range(n):
        ev = []
        hp.step(timed_events=ev)
        torch.cuda.synchronize()
        for (n0, a), (n1, b) in zip(ev[:-1], ev[1:]):
            tot[n1] = tot.get(n1, 0.0) + a.elapsed_time(b) / n
    print("gemm_algo", algo, {k: round(v, 3) for k, v in tot.items()})
