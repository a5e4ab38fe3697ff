// Projection entry points: lin_src forward and its two backward products
// ([PyG] nn/dense/linear.py F.linear inside gat_conv.py; /root/reference/utils/models.py:146).
//
// gemm_algo: 0 = auto (tensor cores whenever the operands satisfy TMA's alignment rules, which the
// reference's shapes do: F = 30*seq_length, ldp % 4 == 0), 1 = exact-fp32 CUDA-core GEMM,
// 2 = tensor cores or error.  Both back ends are hand-written kernels of this library; the CUDA-core
// one exists for leading dimensions TMA cannot address (in_channels % 4 != 0).
#include "gemm.cuh"

using namespace spotv2;

namespace {

struct ProjShape {
  int rows, n_aug, F, ldp;
};

ProjShape shape_of(const spotv2_gat_desc* d) {
  return {d->B * d->N, d->H * d->C + 2 * d->H, d->F, d->ldp};
}

bool use_tc(const spotv2_gat_desc* d) {
  if (d->gemm_algo == 1) return false;
  return d->F % 4 == 0 && d->ldp % 4 == 0;
}

// Carve [hi | lo] pairs out of the workspace.
struct Carver {
  unsigned char* p;
  size_t left;
  float* take(size_t elems) {
    const size_t bytes = round_up(elems * sizeof(float), 256);
    if (bytes > left) return nullptr;
    float* r = reinterpret_cast<float*>(p);
    p += bytes;
    left -= bytes;
    return r;
  }
};

size_t pair_bytes(size_t elems) { return 2 * round_up(elems * sizeof(float), 256); }

}  // namespace

extern "C" int spotv2_gat_workspace_bytes(const spotv2_gat_desc* d, size_t* proj_fwd,
                                          size_t* attn_bwd, size_t* proj_bwd) {
  if (int rc = check_desc(d)) return rc;
  const ProjShape s = shape_of(d);
  const bool tc = use_tc(d);
  if (proj_fwd) *proj_fwd = 256 + (tc ? pair_bytes((size_t)s.rows * s.F) + pair_bytes((size_t)s.n_aug * s.F) : 0);
  if (attn_bwd) {
    // per-CTA partials of dv [H, Fe] and dbias [C or HC]; at most 2 CTAs per SM
    const size_t ctas = 2 * (size_t)sm_count();
    const size_t ldo = d->concat ? (size_t)d->H * d->C : (size_t)d->C;
    *attn_bwd = round_up(ctas * ((size_t)d->H * d->Fe + ldo) * sizeof(float), 256) + 256;
  }
  if (proj_bwd) {
    const int splits = weight_grad_splits(s.rows);
    size_t w = round_up((size_t)splits * s.n_aug * s.F * sizeof(float), 256) + 256;
    if (tc) w += pair_bytes((size_t)s.rows * s.ldp) + pair_bytes((size_t)s.rows * s.F) + pair_bytes((size_t)s.n_aug * s.F);
    *proj_bwd = w;
  }
  return SPOTV2_OK;
}

extern "C" int spotv2_gat_uses_tensor_cores(const spotv2_gat_desc* d) {
  if (check_desc(d)) return 0;
  return use_tc(d) ? 1 : 0;
}

extern "C" int spotv2_split_tf32(const float* src, float* hi, float* lo, size_t n, void* stream) {
  SPOTV2_REQUIRE(src && hi && lo && n > 0, "split_tf32: null pointer or empty");
  return split_tf32(src, hi, lo, n, as_stream(stream));
}

extern "C" int spotv2_proj_fwd(const spotv2_gat_desc* d, const float* x, const float* x_hi, const float* x_lo,
                               const float* W_aug, float* P_aug, void* ws, size_t ws_bytes, void* stream) {
  if (int rc = check_desc(d)) return rc;
  SPOTV2_REQUIRE(x && W_aug && P_aug, "proj_fwd: null pointer");
  const ProjShape s = shape_of(d);
  cudaStream_t st = as_stream(stream);
  if (!use_tc(d) || !aligned16(x) || !aligned16(W_aug) || !aligned16(P_aug)) {
    if (d->gemm_algo == 2) return fail(SPOTV2_ERR_UNSUPPORTED, "proj_fwd: operands do not meet the TMA alignment rules");
    return sgemm_simt(true, true, s.rows, s.n_aug, s.F, x, s.F, W_aug, s.F, P_aug, s.ldp, 1, ws, ws_bytes, st);
  }
  Carver c{static_cast<unsigned char*>(ws), ws ? ws_bytes : 0};
  const bool have_x = x_hi && x_lo;
  SPOTV2_REQUIRE(!have_x || (aligned16(x_hi) && aligned16(x_lo)), "proj_fwd: x_hi/x_lo must be 16-byte aligned");
  float* xh = have_x ? const_cast<float*>(x_hi) : c.take((size_t)s.rows * s.F);
  float* xl = have_x ? const_cast<float*>(x_lo) : c.take((size_t)s.rows * s.F);
  float* wh = c.take((size_t)s.n_aug * s.F);
  float* wl = c.take((size_t)s.n_aug * s.F);
  if (!xh || !xl || !wl) return fail(SPOTV2_ERR_WORKSPACE, "proj_fwd: workspace too small (%zu B)", ws_bytes);
  if (!have_x)
    if (int rc = split_tf32(x, xh, xl, (size_t)s.rows * s.F, st)) return rc;
  if (int rc = split_tf32(W_aug, wh, wl, (size_t)s.n_aug * s.F, st)) return rc;
  return gemm3x_tf32(true, true, s.rows, s.n_aug, s.F, xh, xl, s.F, wh, wl, s.F, P_aug, s.ldp, 1, 256, 4, nullptr, 0, st);
}

extern "C" int spotv2_proj_bwd_weight(const spotv2_gat_desc* d, const float* x, const float* x_hi, const float* x_lo,
                                      const float* dP_aug, const float* dP_lo, float* dW_aug, void* ws,
                                      size_t ws_bytes, void* stream) {
  if (int rc = check_desc(d)) return rc;
  SPOTV2_REQUIRE(x && dP_aug && dW_aug, "proj_bwd_weight: null pointer");
  const ProjShape s = shape_of(d);
  cudaStream_t st = as_stream(stream);
  const int splits = weight_grad_splits(s.rows);
  if (!use_tc(d) || !aligned16(x) || !aligned16(dP_aug) || !aligned16(dW_aug)) {
    if (d->gemm_algo == 2) return fail(SPOTV2_ERR_UNSUPPORTED, "proj_bwd_weight: operands do not meet the TMA alignment rules");
    SPOTV2_REQUIRE(!dP_lo, "proj_bwd_weight: a pre-split dP needs the tensor-core path");
    return sgemm_simt(false, false, s.n_aug, s.F, s.rows, dP_aug, s.ldp, x, s.F, dW_aug, s.F, splits, ws, ws_bytes, st);
  }
  Carver c{static_cast<unsigned char*>(ws), ws ? ws_bytes : 0};
  const bool have_x = x_hi && x_lo, have_p = dP_lo != nullptr;      // dP_aug is the hi part when dP_lo is given
  float* ph = have_p ? const_cast<float*>(dP_aug) : c.take((size_t)s.rows * s.ldp);
  float* pl = have_p ? const_cast<float*>(dP_lo) : c.take((size_t)s.rows * s.ldp);
  float* xh = have_x ? const_cast<float*>(x_hi) : c.take((size_t)s.rows * s.F);
  float* xl = have_x ? const_cast<float*>(x_lo) : c.take((size_t)s.rows * s.F);
  if (!ph || !pl || !xh || !xl) return fail(SPOTV2_ERR_WORKSPACE, "proj_bwd_weight: workspace too small (%zu B)", ws_bytes);
  if (!have_p)
    if (int rc = split_tf32(dP_aug, ph, pl, (size_t)s.rows * s.ldp, st)) return rc;
  if (!have_x)
    if (int rc = split_tf32(x, xh, xl, (size_t)s.rows * s.F, st)) return rc;
  // contraction over the B*N node rows: both operands are MN-major ([K, rows]) for this product
  return gemm3x_tf32(false, false, s.n_aug, s.F, s.rows, ph, pl, s.ldp, xh, xl, s.F, dW_aug, s.F, splits, 256 + 16, 0,
                     c.p, c.left, st);
}

extern "C" int spotv2_proj_bwd_input(const spotv2_gat_desc* d, const float* dP_aug, const float* dP_lo,
                                     const float* W_aug, float* dX, void* ws, size_t ws_bytes,
                                     void* stream) {
  if (int rc = check_desc(d)) return rc;
  SPOTV2_REQUIRE(dP_aug && W_aug && dX, "proj_bwd_input: null pointer");
  const ProjShape s = shape_of(d);
  cudaStream_t st = as_stream(stream);
  if (!use_tc(d) || !aligned16(dP_aug) || !aligned16(W_aug) || !aligned16(dX)) {
    if (d->gemm_algo == 2) return fail(SPOTV2_ERR_UNSUPPORTED, "proj_bwd_input: operands do not meet the TMA alignment rules");
    SPOTV2_REQUIRE(!dP_lo, "proj_bwd_input: a pre-split dP needs the tensor-core path");
    return sgemm_simt(true, false, s.rows, s.F, s.n_aug, dP_aug, s.ldp, W_aug, s.F, dX, s.F, 1, ws, ws_bytes, st);
  }
  Carver c{static_cast<unsigned char*>(ws), ws ? ws_bytes : 0};
  const bool have_p = dP_lo != nullptr;
  float* ph = have_p ? const_cast<float*>(dP_aug) : c.take((size_t)s.rows * s.ldp);
  float* pl = have_p ? const_cast<float*>(dP_lo) : c.take((size_t)s.rows * s.ldp);
  float* wh = c.take((size_t)s.n_aug * s.F);
  float* wl = c.take((size_t)s.n_aug * s.F);
  if (!ph || !pl || !wl) return fail(SPOTV2_ERR_WORKSPACE, "proj_bwd_input: workspace too small (%zu B)", ws_bytes);
  if (!have_p)
    if (int rc = split_tf32(dP_aug, ph, pl, (size_t)s.rows * s.ldp, st)) return rc;
  if (int rc = split_tf32(W_aug, wh, wl, (size_t)s.n_aug * s.F, st)) return rc;
  // dX[rows, F] = dP_aug[rows, n_aug] . W_aug[n_aug, F]: A K-major, B MN-major
  return gemm3x_tf32(true, false, s.rows, s.F, s.n_aug, ph, pl, s.ldp, wh, wl, s.F, dX, s.F, 1, 256 + 16, 0, nullptr, 0, st);
}
