"""GPU box: hammer the p_format 1 forward path (fold -> proj_fwd_pair -> attn_fwd_pair) and time every call."""
import sys, time, torch
sys.path.insert(0, ".")
import bench
from bench import HotPath
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 300
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
HotPath.P_FORMAT = 1
hp = HotPath(B, dev, 1234)
import ctypes as C
lib, d, p, L = hp.lib, C.byref(hp.desc), hp._lib.ptr, hp.layer
st = torch.cuda.current_stream(dev).cuda_stream
chk = hp._lib.check
chk(lib.spotv2_gat_fold(d, p(L.lin_src.weight), p(L.att_src), p(L.att_dst), p(L.lin_edge.weight), p(L.att_edge), p(hp.W_aug), p(hp.v), st), "fold")
chk(lib.spotv2_proj_fwd_pair(d, p(hp.x16[0]), p(hp.x16[1]), p(hp.x_blk), p(hp.W_aug), p(hp.P_aug[0]), p(hp.P_aug[1]), p(hp.p_amax), p(hp.sd32), p(hp.ws), hp.ws.numel(), st), "proj")
torch.cuda.synchronize()
ref = None
worst = 0.0
for i in range(iters):
    t0 = time.time()
    try:
        chk(lib.spotv2_gat_attn_fwd_pair(d, p(hp.P_aug[0]), p(hp.P_aug[1]), p(hp.p_amax), p(hp.sd32), p(hp.batch.edge_attr), p(hp.batch.spot_topology.table),
                                         p(hp.v), p(L.bias), p(hp.out), None, p(hp.edge_terms), st), "fwd")
        torch.cuda.synchronize()
    except Exception as ex:
        print("call", i, "FAILED after", round(time.time() - t0, 3), "s:", str(ex)[:200])
        sys.exit(1)
    dt = time.time() - t0
    worst = max(worst, dt)
    if ref is None:
        ref = hp.out.clone()
    elif not torch.equal(ref, hp.out):
        print("call", i, "differs from call 0: max abs", (ref - hp.out).abs().max().item())
print("ok", iters, "calls, worst", round(worst * 1e3, 3), "ms")
