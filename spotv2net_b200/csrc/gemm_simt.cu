// Exact-fp32 CUDA-core GEMM (packed FFMA2), the projection path for shapes the tensor-core
// kernel does not cover (unaligned leading dimensions, tiny problems) and the parity yardstick
// for it.  C[M,N] = sum_k A(m,k) * B(n,k); each operand is either K-contiguous (row = m|n) or
// M|N-contiguous (row = k).  Optional split-K through a workspace, reduced in a fixed order.
#include "gemm.cuh"

namespace spotv2 {

constexpr int BM = 128, BN = 128, BK = 16, PADM = 4;

template <bool KC>
__device__ __forceinline__ void load_tile(float4 (&reg)[2], const float* __restrict__ P, int ld,
                                          int rows_total, int row0, int k0, int k_end, bool vec,
                                          int tid) {
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const int idx = tid + p * 256;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (KC) {            // P[(row0+r)*ld + k], 4 consecutive k per thread
      const int r = idx >> 2, k = k0 + (idx & 3) * 4;
      if (row0 + r < rows_total) {
        const float* src = P + (size_t)(row0 + r) * ld + k;
        if (vec && k + 3 < k_end) {
          v = *reinterpret_cast<const float4*>(src);
        } else {
          if (k + 0 < k_end) v.x = src[0];
          if (k + 1 < k_end) v.y = src[1];
          if (k + 2 < k_end) v.z = src[2];
          if (k + 3 < k_end) v.w = src[3];
        }
      }
    } else {             // P[k*ld + row0 + r], 4 consecutive rows per thread
      const int k = k0 + (idx >> 5), r = (idx & 31) * 4;
      if (k < k_end) {
        const float* src = P + (size_t)k * ld + row0 + r;
        if (vec && row0 + r + 3 < rows_total) {
          v = *reinterpret_cast<const float4*>(src);
        } else {
          if (row0 + r + 0 < rows_total) v.x = src[0];
          if (row0 + r + 1 < rows_total) v.y = src[1];
          if (row0 + r + 2 < rows_total) v.z = src[2];
          if (row0 + r + 3 < rows_total) v.w = src[3];
        }
      }
    }
    reg[p] = v;
  }
}

template <bool KC>
__device__ __forceinline__ void store_tile(float (*S)[BM + PADM], const float4 (&reg)[2], int tid) {
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const int idx = tid + p * 256;
    if (KC) {
      const int r = idx >> 2, k = (idx & 3) * 4;
      S[k + 0][r] = reg[p].x;
      S[k + 1][r] = reg[p].y;
      S[k + 2][r] = reg[p].z;
      S[k + 3][r] = reg[p].w;
    } else {
      const int k = idx >> 5, r = (idx & 31) * 4;
      *reinterpret_cast<float4*>(&S[k][r]) = reg[p];
    }
  }
}

template <bool A_KC, bool B_KC>
__global__ void __launch_bounds__(256, 2)
sgemm_kernel(int M, int N, int K, const float* __restrict__ A, int lda, const float* __restrict__ B,
             int ldb, float* __restrict__ C, int ldc, int k_per_split, size_t split_stride,
             int vecA, int vecB) {
  __shared__ __align__(16) float As[2][BK][BM + PADM];
  __shared__ __align__(16) float Bs[2][BK][BN + PADM];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int k_begin = blockIdx.z * k_per_split;
  const int k_end = min(K, k_begin + k_per_split);
  float* Cout = C + (size_t)blockIdx.z * split_stride;

  float2 acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);

  float4 ra[2], rb[2];
  load_tile<A_KC>(ra, A, lda, M, m0, k_begin, k_end, vecA, tid);
  load_tile<B_KC>(rb, B, ldb, N, n0, k_begin, k_end, vecB, tid);
  store_tile<A_KC>(As[0], ra, tid);
  store_tile<B_KC>(Bs[0], rb, tid);
  __syncthreads();

  int buf = 0;
  for (int k0 = k_begin; k0 < k_end; k0 += BK) {
    const bool more = k0 + BK < k_end;
    if (more) {
      load_tile<A_KC>(ra, A, lda, M, m0, k0 + BK, k_end, vecA, tid);
      load_tile<B_KC>(rb, B, ldb, N, n0, k0 + BK, k_end, vecB, tid);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float2 bp[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w),
                            make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 ad = make_float2(a[i], a[i]);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = ffma2(ad, bp[j], acc[i][j]);
      }
    }
    if (more) {
      store_tile<A_KC>(As[buf ^ 1], ra, tid);
      store_tile<B_KC>(Bs[buf ^ 1], rb, tid);
      __syncthreads();
      buf ^= 1;
    }
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + (j < 2 ? tx * 4 + 2 * j : 64 + tx * 4 + 2 * (j - 2));
      float* dst = Cout + (size_t)m * ldc + n;
      if (n < N) dst[0] = acc[i][j].x;
      if (n + 1 < N) dst[1] = acc[i][j].y;
    }
  }
}

// Batched form for the large-universe attention path (attn_large.cu): blockIdx.z = batch entry
// z -> (zo, zi) = (z / inner, z % inner); every operand has an element offset per zo and per zi.  The
// contraction runs over `segs` segments (heads) whose operands sit a_s / b_s elements apart and whose
// products are summed into one C tile; the epilogue applies C = scale * acc + bias[zi * bias_i + n].
template <bool A_KC, bool B_KC>
__global__ void __launch_bounds__(256, 2) bgemm_kernel(const BGemm g) {
  __shared__ __align__(16) float As[2][BK][BM + PADM];
  __shared__ __align__(16) float Bs[2][BK][BN + PADM];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int zo = blockIdx.z / g.inner, zi = blockIdx.z - zo * g.inner;
  const float* A0 = g.A + (size_t)zo * g.a_o + (size_t)zi * g.a_i;
  const float* B0 = g.B + (size_t)zo * g.b_o + (size_t)zi * g.b_i;
  float* C = g.C + (size_t)zo * g.c_o + (size_t)zi * g.c_i;

  float2 acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);

  for (int seg = 0; seg < g.segs; ++seg) {
    const float* A = A0 + (size_t)seg * g.a_s;
    const float* B = B0 + (size_t)seg * g.b_s;
    float4 ra[2], rb[2];
    __syncthreads();                    // the previous segment's last tiles are still being read
    load_tile<A_KC>(ra, A, g.lda, g.M, m0, 0, g.K, g.vecA, tid);
    load_tile<B_KC>(rb, B, g.ldb, g.N, n0, 0, g.K, g.vecB, tid);
    store_tile<A_KC>(As[0], ra, tid);
    store_tile<B_KC>(Bs[0], rb, tid);
    __syncthreads();
    int buf = 0;
    for (int k0 = 0; k0 < g.K; k0 += BK) {
      const bool more = k0 + BK < g.K;
      if (more) {
        load_tile<A_KC>(ra, A, g.lda, g.M, m0, k0 + BK, g.K, g.vecA, tid);
        load_tile<B_KC>(rb, B, g.ldb, g.N, n0, k0 + BK, g.K, g.vecB, tid);
      }
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
        const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
        const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float2 bp[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w),
                              make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float2 ad = make_float2(a[i], a[i]);
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = ffma2(ad, bp[j], acc[i][j]);
        }
      }
      if (more) {
        store_tile<A_KC>(As[buf ^ 1], ra, tid);
        store_tile<B_KC>(Bs[buf ^ 1], rb, tid);
        __syncthreads();
        buf ^= 1;
      }
    }
  }

  const float* bias = g.bias ? g.bias + (size_t)zi * g.bias_i : nullptr;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + (j < 2 ? tx * 4 + 2 * j : 64 + tx * 4 + 2 * (j - 2));
      float* dst = C + (size_t)m * g.ldc + n;
      if (n < g.N) dst[0] = fmaf(acc[i][j].x, g.scale, bias ? bias[n] : 0.f);
      if (n + 1 < g.N) dst[1] = fmaf(acc[i][j].y, g.scale, bias ? bias[n + 1] : 0.f);
    }
  }
}

// Tensor-core form of bgemm_kernel: same tiles, loaders and epilogue, the 128x128x16 block product on
// mma.sync.m16n8k8 TF32 with the 3-product split (lo*hi + hi*lo + hi*hi: fp32-accurate).  8 warps in a 2 x 4 grid,
// warp tile 64 x 32 = 4 x 4 MMA tiles.  The tensor core's fp32 accumulate truncates, so the MMA accumulators are
// folded into fp32 registers (round-to-nearest adds) every 64 elements of K.  Shared tiles are [k][m] with a pitch
// of 136 floats: a fragment load touches rows t, t+4 and columns g, g+8 -> banks 8t + g, all distinct.
constexpr int PADT = 8;

template <bool KC>
__device__ __forceinline__ void store_tile_t(float (*S)[BM + PADT], const float4 (&reg)[2], int tid) {
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const int idx = tid + p * 256;
    if (KC) {
      const int r = idx >> 2, k = (idx & 3) * 4;
      S[k + 0][r] = reg[p].x;
      S[k + 1][r] = reg[p].y;
      S[k + 2][r] = reg[p].z;
      S[k + 3][r] = reg[p].w;
    } else {
      const int k = idx >> 5, r = (idx & 31) * 4;
      *reinterpret_cast<float4*>(&S[k][r]) = reg[p];
    }
  }
}

template <bool A_KC, bool B_KC>
__global__ void __launch_bounds__(256, 1) bgemm_tc_kernel(const BGemm g) {
  __shared__ __align__(16) float As[2][BK][BM + PADT];
  __shared__ __align__(16) float Bs[2][BK][BN + PADT];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gq = lane >> 2, t = lane & 3;
  const int wm = (warp >> 2) * 64, wn = (warp & 3) * 32;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int zo = blockIdx.z / g.inner, zi = blockIdx.z - zo * g.inner;
  const float* A0 = g.A + (size_t)zo * g.a_o + (size_t)zi * g.a_i;
  const float* B0 = g.B + (size_t)zo * g.b_o + (size_t)zi * g.b_i;
  float* C = g.C + (size_t)zo * g.c_o + (size_t)zi * g.c_i;

  float run[4][4][4], acc[4][4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) run[i][j][q] = acc[i][j][q] = 0.f;
  int since_fold = 0;

  for (int seg = 0; seg < g.segs; ++seg) {
    const float* A = A0 + (size_t)seg * g.a_s;
    const float* B = B0 + (size_t)seg * g.b_s;
    float4 ra[2], rb[2];
    __syncthreads();
    load_tile<A_KC>(ra, A, g.lda, g.M, m0, 0, g.K, g.vecA, tid);
    load_tile<B_KC>(rb, B, g.ldb, g.N, n0, 0, g.K, g.vecB, tid);
    store_tile_t<A_KC>(As[0], ra, tid);
    store_tile_t<B_KC>(Bs[0], rb, tid);
    __syncthreads();
    int buf = 0;
    for (int k0 = 0; k0 < g.K; k0 += BK) {
      const bool more = k0 + BK < g.K;
      if (more) {
        load_tile<A_KC>(ra, A, g.lda, g.M, m0, k0 + BK, g.K, g.vecA, tid);
        load_tile<B_KC>(rb, B, g.ldb, g.N, n0, k0 + BK, g.K, g.vecB, tid);
      }
#pragma unroll
      for (int ks = 0; ks < BK / 8; ++ks) {
        uint32_t bh[4][2], bl[4][2];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          split_tf32_trunc(Bs[buf][ks * 8 + t][wn + j * 8 + gq], bh[j][0], bl[j][0]);
          split_tf32_trunc(Bs[buf][ks * 8 + t + 4][wn + j * 8 + gq], bh[j][1], bl[j][1]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint32_t ah[4], al[4];
          split_tf32_trunc(As[buf][ks * 8 + t][wm + i * 16 + gq], ah[0], al[0]);
          split_tf32_trunc(As[buf][ks * 8 + t][wm + i * 16 + gq + 8], ah[1], al[1]);
          split_tf32_trunc(As[buf][ks * 8 + t + 4][wm + i * 16 + gq], ah[2], al[2]);
          split_tf32_trunc(As[buf][ks * 8 + t + 4][wm + i * 16 + gq + 8], ah[3], al[3]);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            mma_tf32_16x8x8(acc[i][j], al, bh[j]);
            mma_tf32_16x8x8(acc[i][j], ah, bl[j]);
            mma_tf32_16x8x8(acc[i][j], ah, bh[j]);
          }
        }
      }
      since_fold += BK;
      if (since_fold >= 64) {
        since_fold = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int q = 0; q < 4; ++q) { run[i][j][q] += acc[i][j][q]; acc[i][j][q] = 0.f; }
      }
      if (more) {
        store_tile_t<A_KC>(As[buf ^ 1], ra, tid);
        store_tile_t<B_KC>(Bs[buf ^ 1], rb, tid);
        __syncthreads();
        buf ^= 1;
      }
    }
  }

  const float* bias = g.bias ? g.bias + (size_t)zi * g.bias_i : nullptr;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        // C fragment: (row gq | gq+8, col 2t | 2t+1)
        const int m = m0 + wm + i * 16 + gq + ((q & 2) ? 8 : 0);
        const int n = n0 + wn + j * 8 + 2 * t + (q & 1);
        if (m < g.M && n < g.N)
          C[(size_t)m * g.ldc + n] = fmaf(run[i][j][q] + acc[i][j][q], g.scale, bias ? bias[n] : 0.f);
      }
}

int bgemm_simt(bool a_kc, bool b_kc, BGemm g, int batches, cudaStream_t st) {
  if (g.inner < 1) g.inner = 1;
  if (g.segs < 1) g.segs = 1;
  auto mult4 = [](long long x) { return x % 4 == 0; };
  g.vecA = aligned16(g.A) && g.lda % 4 == 0 && mult4(g.a_o) && mult4(g.a_i) && mult4(g.a_s);
  g.vecB = aligned16(g.B) && g.ldb % 4 == 0 && mult4(g.b_o) && mult4(g.b_i) && mult4(g.b_s);
  dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, batches);
  if (g.tensor_cores) {
    if (a_kc && b_kc) bgemm_tc_kernel<true, true><<<grid, 256, 0, st>>>(g);
    else if (a_kc && !b_kc) bgemm_tc_kernel<true, false><<<grid, 256, 0, st>>>(g);
    else if (!a_kc && b_kc) bgemm_tc_kernel<false, true><<<grid, 256, 0, st>>>(g);
    else bgemm_tc_kernel<false, false><<<grid, 256, 0, st>>>(g);
  } else if (a_kc && b_kc) bgemm_kernel<true, true><<<grid, 256, 0, st>>>(g);
  else if (a_kc && !b_kc) bgemm_kernel<true, false><<<grid, 256, 0, st>>>(g);
  else if (!a_kc && b_kc) bgemm_kernel<false, true><<<grid, 256, 0, st>>>(g);
  else bgemm_kernel<false, false><<<grid, 256, 0, st>>>(g);
  SPOTV2_CUDA_OK(cudaGetLastError());
  return SPOTV2_OK;
}

__global__ void splitk_reduce_kernel(const float* __restrict__ ws, int splits, size_t split_stride,
                                     int M, int N, float* __restrict__ C, int ldc) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)M * N) return;
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += ws[(size_t)z * split_stride + idx];
  const size_t m = idx / N, n = idx - m * N;
  C[m * ldc + n] = s;
}

int sgemm_simt(bool a_kc, bool b_kc, int M, int N, int K, const float* A, int lda, const float* B,
               int ldb, float* C, int ldc, int splits, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (splits < 1) splits = 1;
  int k_per_split = ((K + splits - 1) / splits + BK - 1) / BK * BK;
  splits = (K + k_per_split - 1) / k_per_split;
  float* out = C;
  int ldo = ldc;
  size_t stride = 0;
  if (splits > 1) {
    stride = (size_t)M * N;
    if (!ws || ws_bytes < stride * splits * sizeof(float))
      return fail(SPOTV2_ERR_WORKSPACE, "split-K GEMM needs %zu B of workspace, got %zu",
                  stride * splits * sizeof(float), ws_bytes);
    out = static_cast<float*>(ws);
    ldo = N;
  }
  const int vecA = aligned16(A) && lda % 4 == 0, vecB = aligned16(B) && ldb % 4 == 0;
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, splits);
#define SPOTV2_LAUNCH(AK, BK_)                                                                 \
  sgemm_kernel<AK, BK_><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, out, ldo, k_per_split,  \
                                              stride, vecA, vecB)
  if (a_kc && b_kc) SPOTV2_LAUNCH(true, true);
  else if (a_kc && !b_kc) SPOTV2_LAUNCH(true, false);
  else if (!a_kc && b_kc) SPOTV2_LAUNCH(false, true);
  else SPOTV2_LAUNCH(false, false);
#undef SPOTV2_LAUNCH
  SPOTV2_CUDA_OK(cudaGetLastError());
  if (splits > 1) {
    const size_t total = (size_t)M * N;
    splitk_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(out, splits, stride, M, N, C, ldc);
    SPOTV2_CUDA_OK(cudaGetLastError());
  }
  return SPOTV2_OK;
}

}  // namespace spotv2
