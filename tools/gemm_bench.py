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
    hi, lo = torch.empty_like(A), torch.empty_like(A)
    hb, lb = torch.empty_like(B), torch.empty_like(B)
    t_split = time_ms(lambda: (check(lib.spotv2_split_tf32(ptr(A), ptr(hi), ptr(lo), A.numel(), st()), "s"),
                               check(lib.spotv2_split_tf32(ptr(B), ptr(hb), ptr(lb), B.numel(), st()), "s")))
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
        t_g = t - (t_split if algo == 2 else 0.0)
        print(f"  algo={algo} bn={bn:3d} chunk={kbc}: total {t:7.3f} ms, gemm-only {t_g:7.3f} ms = "
              f"{flops / t_g / 1e9:7.1f} TFLOP/s fp32-equivalent ({3 * flops / t_g / 1e9:7.1f} tf32 issued), err {err:.2e}")


V = [(2, 256, 0), (2, 272, 0), (2, 128, 0), (2, 256, 2), (2, 256, 8), (2, 272, 16), (2, 256, 40), (1, 0, 0)]
run("proj_fwd   P = x W^T", 1, 1, 122880, 3012, 1260, 1, V)
run("proj_bwd_w dW = dP^T x", 0, 0, 3012, 1260, 122880, 15, V[:7])
run("proj_bwd_x dX = dP W", 1, 0, 122880, 1260, 3012, 1, V[:3])
