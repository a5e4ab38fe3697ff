// Operand preparation of the attention backward in p_format 1: the upstream gradient dout [B*N, ldo] (fp32, produced by
// autograd: 5_train_SpotV2Net.py:157 loss.backward()) becomes an fp16 operand pair in ONE pass - a unit (a graph, or a
// (graph, head) block of a concat layer) is read into registers, its largest magnitude fixes the unit's power-of-two scale,
// and the hi | lo planes are written.  The same pass yields the bias gradient (column sums: per-CTA partials, reduced in a
// fixed order) and max|dout| (sizes the scale of the dP pair), so the attention kernel computes neither.
#include <cuda_fp16.h>

#include <algorithm>

#include "attn_bwd.cuh"

namespace spotv2 {

namespace {

__device__ __forceinline__ float unit_scale(float amax) {   // same rule as gemm_f16.cu: amax * scale in [2^14, 2^15)
  if (!(amax > 0.f) || !(amax < INFINITY)) return 1.f;
  int ex;
  frexpf(amax, &ex);
  return exp2f((float)(15 - ex));
}

// KP: column pairs per thread (C <= 512 * KP).  grid: a multiple of units_per_graph, so a CTA only ever sees one head.
template <int KP>
__global__ void __launch_bounds__(256, KP == 1 ? 3 : 1)      // three CTAs per SM: one loading, one reducing, one storing
dout_pair_kernel(const float* __restrict__ dout, int n_units, int N, int C, int upg, int ldo, __half* __restrict__ hi,
                 __half* __restrict__ lo, int ld16, float* __restrict__ scales, unsigned* __restrict__ amax_bits,
                 float* __restrict__ dbias_part) {
  __shared__ float red[2][8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int h = upg > 1 ? (int)(blockIdx.x % (unsigned)upg) : 0;
  float2 bsum[KP];
#pragma unroll
  for (int kp = 0; kp < KP; ++kp) bsum[kp] = make_float2(0.f, 0.f);
  int flip = 0;
  for (int u = blockIdx.x; u < n_units; u += gridDim.x, flip ^= 1) {
    const int b = u / upg;
    const float* src = dout + (size_t)b * N * ldo + (size_t)h * C;
    float2 v[KP][32];
    float m = 0.f;
#pragma unroll
    for (int kp = 0; kp < KP; ++kp) {
      const int c = 2 * (tid + 256 * kp);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        v[kp][i] = (i < N && c < C) ? ldg_stream2(src + (size_t)i * ldo + c) : make_float2(0.f, 0.f);
        m = fmaxf(m, fmaxf(fabsf(v[kp][i].x), fabsf(v[kp][i].y)));
      }
    }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) red[flip][warp] = m;
    __syncthreads();                     // (two buffers: the next unit's writes cannot overtake this unit's reads)
    float mx = red[flip][0];
#pragma unroll
    for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[flip][w]);
    const float s = unit_scale(mx);
    if (tid == 0) {
      scales[u] = s;
      if (mx > 0.f) atomicMax(amax_bits, __float_as_uint(mx));
    }
#pragma unroll
    for (int kp = 0; kp < KP; ++kp) {
      const int c = 2 * (tid + 256 * kp);
      if (c < C) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i < N) {
            const float y0 = v[kp][i].x * s, y1 = v[kp][i].y * s;
            const __half2 hh = __floats2half2_rn(y0, y1);
            const size_t o = ((size_t)b * N + i) * ld16 + (size_t)h * C + c;
            *reinterpret_cast<__half2*>(hi + o) = hh;
            if (lo) {
              const float2 bk = __half22float2(hh);
              *reinterpret_cast<__half2*>(lo + o) = __floats2half2_rn(y0 - bk.x, y1 - bk.y);
            }
            bsum[kp].x += v[kp][i].x;
            bsum[kp].y += v[kp][i].y;
          }
        }
      }
    }
  }
  if (dbias_part) {
    float* mine = dbias_part + (size_t)blockIdx.x * ldo;
    for (int c = tid; c < ldo; c += 256) mine[c] = 0.f;
    __syncthreads();
#pragma unroll
    for (int kp = 0; kp < KP; ++kp) {
      const int c = 2 * (tid + 256 * kp);
      if (c < C) {
        mine[h * C + c] = bsum[kp].x;
        mine[h * C + c + 1] = bsum[kp].y;
      }
    }
  }
}

}  // namespace

int dout_pair_grid(int n_units, int upg) {
  int grid = 6 * sm_count();
  if (grid > n_units) grid = n_units;
  grid = grid / upg * upg;
  return grid < upg ? upg : grid;
}

// dout [B*N, ldo] -> hi | lo [B*N, ld16] (lo may be null), scales [B * upg], blk[0] = bits of max|dout| (blk zeroed here),
// dbias [ldo] (null: skipped) through dbias_part [grid][ldo].  upg = 1 (head mean: ldo == C) or H (concat: ldo == H * C).
int dout_pair_prepass(const float* dout, int B, int N, int C, int upg, __half* hi, __half* lo, int ld16, float* scales, float* blk,
                      float* dbias, float* dbias_part, cudaStream_t st, int* n_parts) {
  if (C % 2 != 0 || C > 1024) return fail(SPOTV2_ERR_UNSUPPORTED, "attn_bwd (p_format 1): even C <= 1024 required, got %d", C);
  const int ldo = upg * C, n_units = B * upg;
  const int grid = dout_pair_grid(n_units, upg);
  SPOTV2_CUDA_OK(cudaMemsetAsync(blk, 0, kScaleBlockFloats * sizeof(float), st));
  float* part = dbias_part;
  if (n_parts) *n_parts = grid;
  if (C <= 512)
    dout_pair_kernel<1><<<grid, 256, 0, st>>>(dout, n_units, N, C, upg, ldo, hi, lo, ld16, scales, reinterpret_cast<unsigned*>(blk), part);
  else
    dout_pair_kernel<2><<<grid, 256, 0, st>>>(dout, n_units, N, C, upg, ldo, hi, lo, ld16, scales, reinterpret_cast<unsigned*>(blk), part);
  SPOTV2_CUDA_OK(cudaGetLastError());
  if (dbias) return reduce_partials(part, grid, ldo, dbias, st);
  return SPOTV2_OK;
}

}  // namespace spotv2

using namespace spotv2;

extern "C" size_t spotv2_diag_dout_pair_ws_bytes(int32_t B, int32_t C, int32_t upg) {
  if (B <= 0 || C <= 0 || upg <= 0) return 0;
  return round_up((size_t)dout_pair_grid(B * upg, upg) * upg * C * sizeof(float), 256);
}

extern "C" int spotv2_diag_dout_pair(const float* dout, int32_t B, int32_t N, int32_t C, int32_t upg, void* hi, void* lo_or_null,
                                     int32_t ld16, float* scales, float* scale_block, float* dbias_or_null, void* ws, void* stream) {
  SPOTV2_REQUIRE(dout && hi && scales && scale_block && B > 0 && N > 0 && N <= 32 && C > 0 && upg > 0, "diag_dout_pair: bad argument");
  SPOTV2_REQUIRE(ld16 >= upg * C && ld16 % 8 == 0, "diag_dout_pair: ld16 >= upg * C and ld16 %% 8 == 0");
  SPOTV2_REQUIRE(!dbias_or_null || ws, "diag_dout_pair: dbias needs the partial-sum workspace");
  return dout_pair_prepass(dout, B, N, C, upg, static_cast<__half*>(hi), static_cast<__half*>(lo_or_null), ld16, scales, scale_block,
                           dbias_or_null, static_cast<float*>(ws), as_stream(stream));
}
