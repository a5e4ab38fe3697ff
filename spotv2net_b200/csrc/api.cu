// Error plumbing, descriptor validation and the small ABI queries of libspotv2_gat.so.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace spotv2 {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fail(spotv2_status st, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return (int)st;
}

int check_desc(const spotv2_gat_desc* d) {
  if (!d) return fail(SPOTV2_ERR_INVALID_ARG, "descriptor is null");
  if (d->B <= 0 || d->N <= 0 || d->F <= 0 || d->H <= 0 || d->C <= 0 || d->Fe < 0 || d->R < 0)
    return fail(SPOTV2_ERR_INVALID_ARG, "non-positive size in descriptor (B=%d N=%d F=%d Fe=%d H=%d C=%d R=%d)",
                d->B, d->N, d->F, d->Fe, d->H, d->C, d->R);
  if (d->N > 0xffff) return fail(SPOTV2_ERR_UNSUPPORTED, "N=%d does not fit the 16-bit row table", d->N);
  if (d->ldp < d->H * d->C + 2 * d->H || d->ldp % 4 != 0)
    return fail(SPOTV2_ERR_INVALID_ARG, "ldp=%d must be >= H*C+2H=%d and a multiple of 4", d->ldp,
                d->H * d->C + 2 * d->H);
  if (!(d->dropout_p >= 0.f && d->dropout_p < 1.f))
    return fail(SPOTV2_ERR_INVALID_ARG, "dropout_p=%g must be in [0, 1)", (double)d->dropout_p);
  if (d->Fe > 0 && d->R <= 0) return fail(SPOTV2_ERR_INVALID_ARG, "R must be positive when Fe > 0");
  if (d->p_format != 0 && d->p_format != 1) return fail(SPOTV2_ERR_INVALID_ARG, "p_format=%d must be 0 (fp32) or 1 (fp16 pair)", d->p_format);
  if (d->p_format == 1 && (d->N > 32 || d->gemm_algo == 1))
    return fail(SPOTV2_ERR_UNSUPPORTED, "p_format 1 covers N <= 32 on the tensor-core GEMM (gemm_algo 0, 2 or 3)");
  return SPOTV2_OK;
}

int sm_count() {
  static int cached[64] = {0};          // per device: ranks of one process may drive different GPUs
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev >= 0 && dev < 64 && cached[dev]) return cached[dev];
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
  if (dev >= 0 && dev < 64) cached[dev] = n;
  return n;
}

}  // namespace spotv2

using namespace spotv2;

extern "C" const char* spotv2_last_error(void) { return g_err; }
extern "C" int32_t spotv2_abi_version(void) { return SPOTV2_ABI_VERSION; }
// >= 32 floats so one P_aug row always spans a full 128-byte TMA tile row
extern "C" int32_t spotv2_gat_n_aug(const spotv2_gat_desc* d) { return d ? n_aug_of(d) : 0; }
extern "C" int32_t spotv2_gat_head_pitch(const spotv2_gat_desc* d) { return d ? head_pitch_of(d) : 0; }
extern "C" int32_t spotv2_gat_ldp(int32_t H, int32_t C) {
  const int32_t w = (H * C + 2 * H + 3) / 4 * 4;
  return w < 32 ? 32 : w;
}
