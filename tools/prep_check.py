"""GPU box: the dout operand preparation against torch, several unit counts."""
import sys, torch
sys.path.insert(0, ".")
import spotv2net_b200 as sv
from spotv2net_b200 import _lib
lib = sv.load_library()
dev = torch.device("cuda", 0)
p = _lib.ptr
st = torch.cuda.current_stream(dev).cuda_stream
for (B, N, C, upg) in [(8, 30, 500, 1), (592, 30, 500, 1), (593, 30, 500, 1), (1184, 30, 500, 1), (4096, 30, 500, 1), (700, 30, 256, 8), (300, 7, 12, 3)]:
    ldo = upg * C
    g = torch.Generator(device=dev).manual_seed(B)
    dout = torch.randn(B * N, ldo, device=dev, generator=g) * torch.rand(B * N, 1, device=dev, generator=g)
    ld16 = lib.spotv2_gat_ld16(ldo)
    planes = torch.zeros(2, B * N, ld16, device=dev, dtype=torch.float16)
    scales = torch.zeros(B * upg, device=dev)
    blk = torch.zeros(8, device=dev)
    dbias = torch.zeros(ldo, device=dev)
    ws = torch.empty(lib.spotv2_diag_dout_pair_ws_bytes(B, C, upg), device=dev, dtype=torch.uint8)
    _lib.check(lib.spotv2_diag_dout_pair(p(dout), B, N, C, upg, p(planes[0]), p(planes[1]), ld16, p(scales), p(blk), p(dbias), p(ws), st), "prep")
    torch.cuda.synchronize()
    units = dout.view(B, N, upg, C)
    umax = units.abs().amax(dim=(1, 3))                       # [B, upg]
    want_s = torch.exp2(15 - torch.frexp(umax)[1].float())
    rec = (planes[0, :, :ldo].float() + planes[1, :, :ldo].float()).view(B, N, upg, C) / scales.view(B, 1, upg, 1)
    print(B, N, C, upg, "scale mismatches", int((scales.view(B, upg) != want_s).sum()), "max rel err of hi+lo", ((rec - units).abs().max() / units.abs().max()).item(),
          "per-unit worst", ((rec - units).abs().amax(dim=(1, 3)) / umax).max().item(), "amax", blk.view(torch.int32)[0].item() == units.abs().max().view(torch.int32).item(),
          "dbias", ((dbias - dout.sum(0)).abs().max() / dout.sum(0).abs().max()).item())
