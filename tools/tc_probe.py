"""Probe (GPU box): does the tensor-core fp32 accumulator round to nearest or truncate?

Sums K exactly-representable positive products per output with a cuBLAS TF32 / BF16 GEMM and
compares with float64.  Truncation (RZ) shows as a negative mean signed error growing ~linearly
with K; round-to-nearest shows a zero-mean error growing ~sqrt(K).  Decides how the 3xTF32
projection GEMM must chain its accumulations."""
import torch

dev = "cuda:0"
torch.manual_seed(0)
print("K, mode, mean_signed_rel_err, rms_rel_err")
for K in (1024, 4096, 16384, 65536, 131072):
    M, N = 256, 256
    # values with <= 8 significant bits so every product is exact in tf32/bf16 and in fp32
    a = (torch.randint(128, 256, (M, K), device=dev).float() / 128.0)
    b = (torch.randint(128, 256, (K, N), device=dev).float() / 128.0)
    ref = (a.double() @ b.double())
    for mode in ("fp32", "tf32", "bf16"):
        if mode == "fp32":
            torch.backends.cuda.matmul.allow_tf32 = False
            c = a @ b
        elif mode == "tf32":
            torch.backends.cuda.matmul.allow_tf32 = True
            c = a @ b
        else:
            c = (a.bfloat16() @ b.bfloat16()).float()      # output rounded to bf16: only the sign of the bias is meaningful
            c = torch.matmul(a.bfloat16(), b.bfloat16(), out_dtype=torch.float32) if hasattr(torch, "_scaled_mm") and False else c
        err = (c.double() - ref) / ref
        print(K, mode, f"{err.mean().item():+.3e}", f"{err.pow(2).mean().sqrt().item():.3e}")
torch.backends.cuda.matmul.allow_tf32 = False
