"""Times the projection GEMM variants at the production shapes (GPU box): tile widths, k-block depth,
accumulation-chunk length.  Reports ms and the max-norm error against a float64 reference on a row slice."""
import ctypes as C
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spotv2net_b200 as sv
from spotv2net_b200._lib import check, ptr

lib = sv.load_library()
dev = "cuda:0"
st = lambda: torch.cuda.current_stream().cuda_stream


def time_ms(fn, iters=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def run(name, a_kc, b_kc, M, N, K, splits, variants):
    torch.manual_seed(0)
    A = torch.randn((M, K) if a_kc else (K, M), device=dev)
    B = torch.randn((N, K) if b_kc else (K, N), device=dev)
    Cm = torch.empty(M, N, device=dev)
    ws = torch.empty(8 * (A.numel() + B.numel()) + 4 * splits * M * N + (1 << 20), dtype=torch.uint8, device=dev)
    ld_a, ld_b = lib.spotv2_gat_ld16(A.shape[1]), lib.spotv2_gat_ld16(B.shape[1])
    pa = torch.empty(2, A.shape[0], ld_a, device=dev, dtype=torch.float16)
    pb = torch.empty(2, B.shape[0], ld_b, device=dev, dtype=torch.float16)
    blk = torch.empty(16, device=dev)
    t_split = time_ms(lambda: (check(lib.spotv2_split_f16(ptr(A), A.shape[0], A.shape[1], A.shape[1], 0, 0, ptr(pa[0]), ptr(pa[1]),
                                                          ld_a, ptr(blk[:8]), st()), "s"),
                               check(lib.spotv2_split_f16(ptr(B), B.shape[0], B.shape[1], B.shape[1], 0, 0, ptr(pb[0]), ptr(pb[1]),
                                                          ld_b, ptr(blk[8:]), st()), "s")))
    rows = slice(0, 256)
    if a_kc:
        ref = A[rows].double() @ (B.double().t() if b_kc else B.double())
    else:
        ref = A[:, rows].double().t() @ (B.double().t() if b_kc else B.double())
    flops = 2.0 * M * N * K
    print(f"== {name}: M={M} N={N} K={K} splits={splits}; operand split alone {t_split:.3f} ms")
    for algo, bn, kbc in variants:
        def call():
            check(lib.spotv2_diag_gemm(a_kc, b_kc, M, N, K, ptr(A), A.shape[1], ptr(B), B.shape[1], ptr(Cm), N, algo,
                                       splits, bn, kbc, ptr(ws), ws.numel(), st()), "gemm")
        try:
            t = time_ms(call)
        except Exception as ex:
            print(f"  algo={algo} bn={bn} chunk={kbc}: FAILED {ex}")
            continue
        err = ((Cm[rows].double() - ref).abs().max() / ref.abs().max()).item()
        t_g = t - (t_split if algo == 3 else 0.0)
        print(f"  algo={algo} bn={bn:3d} chunk={kbc}: total {t:7.3f} ms, gemm-only {t_g:7.3f} ms = "
              f"{flops / t_g / 1e9:7.1f} TFLOP/s fp32-equivalent ({3 * flops / t_g / 1e9:7.1f} issued), err {err:.2e}")


# algo 3 = fp16 pairs (production), algo 2 = 3xTF32 (its split time is NOT subtracted: ~0.3-0.9 ms), 1 = CUDA cores
V = [(3, 256, 0), (3, 272, 0), (3, 128, 0), (3, 256, 1), (3, 256, 4), (3, 272, 8), (2, 256, 0), (2, 272, 0)]
run("proj_fwd   P = x W^T", 1, 1, 122880, 3012, 1260, 1, V)
run("proj_bwd_w dW = dP^T x", 0, 0, 3012, 1260, 122880, 15, V)
run("proj_bwd_x dX = dP W", 1, 0, 122880, 1260, 3012, 1, V[:3])
