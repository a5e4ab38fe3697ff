// Structured edge source.  In the reference's dataset the edge features are a pure function of the
// window of co-volatility matrices (/root/reference/utils/dataset.py:228-242, restated in SURVEY.md Appendix B):
//   edge (j -> i), feature k*L + t:   k = 0: vv[t0+t][min(i,j)][max(i,j)]   k = 1: vv[t0+t][j][j]   k = 2: vv[t0+t][i][i]
// so the folded edge term of the attention logit splits into one [N*N x L] . [L x H] product and two per-node sums:
//   g_ij,h = <e_ij, v_h> = sum_t vv[t0+t][r][c] v_h[t]  +  sum_t vv[t0+t][j][j] v_h[L+t]  +  sum_t vv[t0+t][i][i] v_h[2L+t]
// Reading the [L, N, N] window (151 KB per graph at the default geometry) replaces reading the [N(N-1), 3L] edge rows
// (438 KB) in the forward and in the backward, and the 3L-wide products shrink to L-wide ones.
//   spotv2_edge_terms_from_windows   windows, v -> edge-term tile (the buffer spotv2_gat_attn_fwd / _bwd consume with
//                                    edge_mode = 1)
//   spotv2_windows_dv                windows, d(edge terms) (written by spotv2_gat_attn_bwd) -> dv [H, 3L]
// Model specific by construction (it knows how CovarianceLaggedDataset builds edge_attr); the generic edge_attr path
// stays the default and the reference for parity.
#include "attn_bwd.cuh"

namespace spotv2 {

namespace {

constexpr int kWinThreads = 256;

// pair index m = j*N + i (source j, target i) -> offset of vv[.][min][max] inside one N x N matrix
__device__ __forceinline__ int pair_offset(int m, int N) {
  const int j = m / N, i = m - j * N;
  return j < i ? j * N + i : i * N + j;
}

__global__ void __launch_bounds__(kWinThreads)
win_edge_terms_kernel(const float* __restrict__ vv, const int32_t* __restrict__ t0, const float* __restrict__ v,
                      float* __restrict__ terms, int B, int N, int L, int H, int NS) {
  extern __shared__ __align__(16) float smem[];
  float* vs = smem;                       // [3L][8]: v_h[k] for k < 3L, heads padded to 8
  float* as = vs + 3 * L * 8;             // [8][N] source-node term
  float* bs = as + 8 * N;                 // [8][N] target-node term
  const int tid = threadIdx.x, NN = N * N;
  for (int idx = tid; idx < 3 * L * 8; idx += kWinThreads) {
    const int k = idx >> 3, h = idx & 7;
    vs[idx] = h < H ? v[(size_t)h * 3 * L + k] : 0.f;
  }
  __syncthreads();
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    const float* W = vv + (size_t)t0[b] * NN;
    for (int idx = tid; idx < 8 * N; idx += kWinThreads) {       // per-node sums over the matrices' diagonals
      const int h = idx / N, j = idx - h * N;
      float a = 0.f, c = 0.f;
      if (h < H)
        for (int t = 0; t < L; ++t) {
          const float dg = W[(size_t)t * NN + j * N + j];
          a = fmaf(dg, vs[(L + t) * 8 + h], a);
          c = fmaf(dg, vs[(2 * L + t) * 8 + h], c);
        }
      as[idx] = a;
      bs[idx] = c;
    }
    __syncthreads();
    float* out = terms + (size_t)b * H * N * NS;
    for (int m = tid; m < NN; m += kWinThreads) {
      const int j = m / N, i = m - j * N;
      float g[8];
#pragma unroll
      for (int h = 0; h < 8; ++h) g[h] = 0.f;
      if (i != j) {
        const float* src = W + pair_offset(m, N);
#pragma unroll 6
        for (int t = 0; t < L; ++t) {
          const float x = src[(size_t)t * NN];
          const float4 v0 = *reinterpret_cast<const float4*>(vs + t * 8);
          const float4 v1 = *reinterpret_cast<const float4*>(vs + t * 8 + 4);
          g[0] = fmaf(x, v0.x, g[0]); g[1] = fmaf(x, v0.y, g[1]); g[2] = fmaf(x, v0.z, g[2]); g[3] = fmaf(x, v0.w, g[3]);
          g[4] = fmaf(x, v1.x, g[4]); g[5] = fmaf(x, v1.y, g[5]); g[6] = fmaf(x, v1.z, g[6]); g[7] = fmaf(x, v1.w, g[7]);
        }
      }
#pragma unroll
      for (int h = 0; h < 8; ++h)
        if (h < H) out[((size_t)h * N + j) * NS + i] = (i != j) ? g[h] + as[h * N + j] + bs[h * N + i] : 0.f;
    }
    // the padding columns i in [N, NS) must be zero: the attention kernels copy the whole tile into shared memory and
    // their MMA fragments contract over all 32 target slots (alpha = 0 there is what masks the next graph's rows)
    for (int idx = tid; idx < H * N * (NS - N); idx += kWinThreads) {
      const int hj = idx / (NS - N), i = N + idx - hj * (NS - N);
      out[(size_t)hj * NS + i] = 0.f;
    }
    __syncthreads();
  }
}

// dv[h][k] partials of this CTA's graphs.  dterms = dz' in the tile layout ([H][N][NS], diagonal 0).
__global__ void __launch_bounds__(kWinThreads)
win_dv_kernel(const float* __restrict__ vv, const int32_t* __restrict__ t0, const float* __restrict__ dterms,
              float* __restrict__ part, int B, int N, int L, int H, int NS) {
  extern __shared__ __align__(16) float smem[];
  const int NN = N * N;
  float* dzs = smem;                       // [H][NN] compact, diagonal 0
  float* rs = dzs + H * NN;                // [H][N] sum over targets  (feeds the source-variance features)
  float* cs = rs + H * N;                  // [H][N] sum over sources  (feeds the target-variance features)
  float* acc = cs + H * N;                 // [H][3L] this CTA's running dv
  int* poff = reinterpret_cast<int*>(acc + H * 3 * L);          // [NN]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int idx = tid; idx < H * 3 * L; idx += kWinThreads) acc[idx] = 0.f;
  for (int m = tid; m < NN; m += kWinThreads) poff[m] = pair_offset(m, N);
  __syncthreads();
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    const float* W = vv + (size_t)t0[b] * NN;
    const float* src = dterms + (size_t)b * H * N * NS;
    for (int idx = tid; idx < H * NN; idx += kWinThreads) {
      const int h = idx / NN, m = idx - h * NN, j = m / N, i = m - j * N;
      dzs[idx] = (i != j) ? src[((size_t)h * N + j) * NS + i] : 0.f;
    }
    __syncthreads();
    for (int idx = tid; idx < 2 * H * N; idx += kWinThreads) {
      const int which = idx / (H * N), r = idx - which * H * N, h = r / N, n = r - h * N;
      float s = 0.f;
      if (which == 0) for (int i = 0; i < N; ++i) s += dzs[h * NN + n * N + i];
      else            for (int j = 0; j < N; ++j) s += dzs[h * NN + j * N + n];
      (which == 0 ? rs : cs)[r] = s;
    }
    __syncthreads();
    // k = 0 block: dv[h][t] += sum_m dz'[h][m] vv[t0+t][pair(m)]; one warp per lag, lanes over the pairs
    for (int t = warp; t < L; t += kWinThreads / 32) {
      const float* Wt = W + (size_t)t * NN;
      float a[8];
#pragma unroll
      for (int h = 0; h < 8; ++h) a[h] = 0.f;
      for (int m = lane; m < NN; m += 32) {
        const float x = Wt[poff[m]];
#pragma unroll
        for (int h = 0; h < 8; ++h)
          if (h < H) a[h] = fmaf(x, dzs[h * NN + m], a[h]);
      }
#pragma unroll
      for (int h = 0; h < 8; ++h) {
        if (h < H) {
          float s = a[h];
          for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
          if (lane == 0) acc[h * 3 * L + t] += s;
        }
      }
    }
    // k = 1, 2 blocks: the diagonals against the row / column sums
    for (int idx = tid; idx < 2 * H * L; idx += kWinThreads) {
      const int which = idx / (H * L), r = idx - which * H * L, h = r / L, t = r - h * L;
      const float* sums = (which == 0 ? rs : cs) + h * N;
      float s = 0.f;
      for (int n = 0; n < N; ++n) s = fmaf(W[(size_t)t * NN + n * N + n], sums[n], s);
      acc[h * 3 * L + (1 + which) * L + t] += s;
    }
    __syncthreads();
  }
  for (int idx = tid; idx < H * 3 * L; idx += kWinThreads) part[(size_t)blockIdx.x * H * 3 * L + idx] = acc[idx];
}

int check_windows(const spotv2_gat_desc* d, int32_t T, int32_t L) {
  if (int rc = check_desc(d)) return rc;
  SPOTV2_REQUIRE(L > 0 && T > L && d->Fe == 3 * L, "windows: edge_dim must equal 3 * seq_length (got Fe=%d, L=%d)", d->Fe, L);
  if (d->N > 32) return fail(SPOTV2_ERR_UNSUPPORTED, "windows: the structured edge source covers N <= 32 in this version");
  if (d->H > kMaxHeads) return fail(SPOTV2_ERR_UNSUPPORTED, "H=%d > %d", d->H, kMaxHeads);
  return SPOTV2_OK;
}

int win_grid(int B) {
  const int g = 4 * sm_count();
  return g < B ? g : B;
}

}  // namespace

}  // namespace spotv2

using namespace spotv2;

extern "C" int spotv2_edge_terms_from_windows(const spotv2_gat_desc* d, const float* M_vv, int32_t T, int32_t L,
                                              const int32_t* t0, const float* v, float* edge_terms, void* stream) {
  if (int rc = check_windows(d, T, L)) return rc;
  SPOTV2_REQUIRE(M_vv && t0 && v && edge_terms, "edge_terms_from_windows: null pointer");
  const size_t smem = ((size_t)3 * L * 8 + 16 * d->N) * sizeof(float);
  win_edge_terms_kernel<<<win_grid(d->B), kWinThreads, smem, as_stream(stream)>>>(M_vv, t0, v, edge_terms, d->B, d->N, L,
                                                                                 d->H, kEdgeTermNS);
  SPOTV2_CUDA_OK(cudaGetLastError());
  return SPOTV2_OK;
}

extern "C" int spotv2_windows_dv_workspace_bytes(const spotv2_gat_desc* d, size_t* bytes) {
  if (int rc = check_desc(d)) return rc;
  SPOTV2_REQUIRE(bytes, "windows_dv_workspace_bytes: null pointer");
  *bytes = round_up((size_t)4 * sm_count() * d->H * (d->Fe > 0 ? d->Fe : 1) * sizeof(float), 256);
  return SPOTV2_OK;
}

extern "C" int spotv2_windows_dv(const spotv2_gat_desc* d, const float* M_vv, int32_t T, int32_t L, const int32_t* t0,
                                 const float* d_edge_terms, float* dv, void* ws, size_t ws_bytes, void* stream) {
  if (int rc = check_windows(d, T, L)) return rc;
  SPOTV2_REQUIRE(M_vv && t0 && d_edge_terms && dv, "windows_dv: null pointer");
  const int grid = win_grid(d->B);
  const size_t need = (size_t)grid * d->H * d->Fe * sizeof(float);
  if (!ws || ws_bytes < need) return fail(SPOTV2_ERR_WORKSPACE, "windows_dv needs %zu B of workspace, got %zu", need, ws_bytes);
  const int NN = d->N * d->N;
  const size_t smem = ((size_t)d->H * NN + 2 * (size_t)d->H * d->N + (size_t)d->H * d->Fe + NN) * sizeof(float);
  SPOTV2_CUDA_OK(cudaFuncSetAttribute(win_dv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  win_dv_kernel<<<grid, kWinThreads, smem, as_stream(stream)>>>(M_vv, t0, d_edge_terms, static_cast<float*>(ws), d->B, d->N, L,
                                                               d->H, kEdgeTermNS);
  SPOTV2_CUDA_OK(cudaGetLastError());
  return reduce_partials(static_cast<float*>(ws), grid, d->H * d->Fe, dv, as_stream(stream));
}
