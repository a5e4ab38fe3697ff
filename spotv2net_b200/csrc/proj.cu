// Projection entry points: lin_src forward and its two backward products
// ([PyG] nn/dense/linear.py F.linear inside gat_conv.py; /root/reference/utils/models.py:146).
#include "gemm.cuh"

using namespace spotv2;

extern "C" int spotv2_gat_workspace_bytes(const spotv2_gat_desc* d, size_t* proj_fwd,
                                          size_t* attn_bwd, size_t* proj_bwd) {
  if (int rc = check_desc(d)) return rc;
  const size_t n_aug = (size_t)d->H * d->C + 2 * d->H;
  if (proj_fwd) *proj_fwd = 256;
  if (attn_bwd) {
    // per-CTA partials of dv [H, Fe] and dbias [C or HC]; at most 2 CTAs per SM
    const size_t ctas = 2 * (size_t)sm_count();
    const size_t ldo = d->concat ? (size_t)d->H * d->C : (size_t)d->C;
    *attn_bwd = round_up(ctas * ((size_t)d->H * d->Fe + ldo) * sizeof(float), 256) + 256;
  }
  if (proj_bwd) {
    const int splits = weight_grad_splits(d->B * d->N);
    *proj_bwd = round_up((size_t)splits * n_aug * d->F * sizeof(float), 256) + 256;
  }
  return SPOTV2_OK;
}

extern "C" int spotv2_proj_fwd(const spotv2_gat_desc* d, const float* x, const float* W_aug,
                               float* P_aug, void* ws, size_t ws_bytes, void* stream) {
  if (int rc = check_desc(d)) return rc;
  SPOTV2_REQUIRE(x && W_aug && P_aug, "proj_fwd: null pointer");
  const int rows = d->B * d->N, n_aug = d->H * d->C + 2 * d->H;
  return sgemm_simt(true, true, rows, n_aug, d->F, x, d->F, W_aug, d->F, P_aug, d->ldp, 1, ws,
                    ws_bytes, as_stream(stream));
}

extern "C" int spotv2_proj_bwd_weight(const spotv2_gat_desc* d, const float* x, const float* dP_aug,
                                      float* dW_aug, void* ws, size_t ws_bytes, void* stream) {
  if (int rc = check_desc(d)) return rc;
  SPOTV2_REQUIRE(x && dP_aug && dW_aug, "proj_bwd_weight: null pointer");
  const int rows = d->B * d->N, n_aug = d->H * d->C + 2 * d->H;
  return sgemm_simt(false, false, n_aug, d->F, rows, dP_aug, d->ldp, x, d->F, dW_aug, d->F,
                    weight_grad_splits(rows), ws, ws_bytes, as_stream(stream));
}

extern "C" int spotv2_proj_bwd_input(const spotv2_gat_desc* d, const float* dP_aug,
                                     const float* W_aug, float* dX, void* ws, size_t ws_bytes,
                                     void* stream) {
  if (int rc = check_desc(d)) return rc;
  SPOTV2_REQUIRE(dP_aug && W_aug && dX, "proj_bwd_input: null pointer");
  const int rows = d->B * d->N, n_aug = d->H * d->C + 2 * d->H;
  return sgemm_simt(true, false, rows, d->F, n_aug, dP_aug, d->ldp, W_aug, d->F, dX, d->F, 1, ws,
                    ws_bytes, as_stream(stream));
}
