"""GPU box: does tcgen05 kind::tf32 truncate an fp32 bit pattern (then the lo pass need not round hi in place)?
Runs the hot path with attn_bwd_algo 3 (hi rounded onto the tf32 grid in place) and 4 (hi = raw TMA tile, lo = x - trunc(x))
and compares every gradient; agreement at ~1e-6 means the raw form is safe, ~1e-4 means the tensor core rounds."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
dev = torch.device("cuda", 0)
hp = bench.HotPath(592, dev, 1234)
res = {}
for algo in (2, 3, 4):
    hp.desc.attn_bwd_algo = algo
    hp.step(); torch.cuda.synchronize()
    res[algo] = [t.clone() for t in (hp.g_W, hp.g_as, hp.g_ad, hp.g_We, hp.g_ae, hp.g_b)]
rel = lambda a, b: ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()
for algo in (3, 4):
    print("algo", algo, "vs mma.sync kernel:", ["%.2e" % rel(a, b) for a, b in zip(res[algo], res[2])])
