// fp32-accurate projection GEMM on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
//   C[M,N] = sum_k A(m,k) * B(n,k),   A = A_hi + A_lo, B = B_hi + B_lo  (tf32 pairs, see split_tf32)
//   C     += A_lo*B_hi + A_hi*B_lo + A_hi*B_hi        (3xTF32: the dropped lo*lo term is ~2^-22)
//
// Operands are pre-split into tf32 hi/lo arrays in HBM (split_tf32_kernel) so the main loop is pure
// TMA -> shared memory -> tcgen05.mma with no per-element work on the SM.  Each operand is either
// K-major ([rows, K], K contiguous) or MN-major ([K, rows]); both are legal tf32 UMMA layouts, which
// is what lets dW = dP^T x contract over the node dimension without materialising transposes.
//
// Accumulation.  Measured on B200 (tools/tc_probe.py, profiles/): the tensor core adds each k-step into
// its fp32 accumulator with TRUNCATION, a bias that grows linearly with the chain length (-9e-5 relative
// at K=16384).  To stay inside the 1e-5 parity budget the TMEM accumulator only ever holds a short
// chunk of K (kb_per_chunk k-blocks, 128 elements by default = 48 MMAs); the consumer warps drain each
// chunk with tcgen05.ld and add it round-to-nearest into fp32 registers, while the tensor core is
// already filling the other TMEM stage with the next chunk.
//
// Kernel shape: persistent, one CTA per SM, 384 threads:
//   warp 0   TMA producer (one elected lane)        warp 1   MMA issuer (one elected lane)
//   warp 2   TMEM allocator                          warps 4-11 accumulate + epilogue: warp w owns TMEM
//            lane quarter w%4 and columns [(w-4)/4 * BN/2, +BN/2) of the tile, BN/2 running sums per thread
// Tile 128 x BN x BK (BK = 32: 128-byte swizzle rows, or 16: 64-byte rows and a twice deeper ring),
// kStages-deep shared-memory ring, two TMEM stages.  Optional split-K writes
// partials to a workspace that a fixed-order reduce sums (deterministic).
#include "gemm.cuh"
#include "tc.cuh"
#include "tma.cuh"

namespace spotv2 {

constexpr int TBM = 128;       // UMMA M (cta_group::1)
constexpr int UMMA_K = 8;      // tf32
constexpr int kTcThreads = 384;
constexpr int kEpiWarps = 8;

struct TcParams {
  int M, N, K;
  int ldc;
  float* C;
  int m_tiles, n_tiles, splits, kb_per_split, kb_total, kb_per_chunk;
  size_t split_stride;   // elements between split partials (0 when splits == 1)
};

template <int BN, int TBK, bool A_KM, bool B_KM>
struct TcSmem {
  static constexpr int kAOp = TBM * TBK * 4;               // one A operand tile (hi or lo)
  static constexpr int kBOp = BN * TBK * 4;                // one B operand tile
  static constexpr int kStage = 2 * kAOp + 2 * kBOp;
  static constexpr int kStages = (200 * 1024) / kStage > 6 ? 6 : (200 * 1024) / kStage;   // 256x32: 2, 256x16: 4, 128x32: 3
  static constexpr int kBarOff = kStages * kStage;
  static constexpr int kTotal = kBarOff + 256 + 1024;      // barriers + tmem ptr + alignment slack
  static constexpr uint32_t kTxBytes = kStage;
};

template <int BN, int TBK, bool A_KM, bool B_KM>
__global__ void __launch_bounds__(kTcThreads, 1)
gemm3x_tf32_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                   const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                   const TcParams p) {
  using S = TcSmem<BN, TBK, A_KM, B_KM>;
  constexpr int kStages = S::kStages;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::kBarOff);
  uint64_t* empty = full + kStages;
  uint64_t* tfull = empty + kStages;     // [2] accumulator ready
  uint64_t* tempty = tfull + 2;          // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.m_tiles * p.n_tiles * p.splits;
  constexpr uint32_t kTmemCols = 2 * BN;   // 512 (BN=256) or 256 (BN=128): power of two >= 32

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmAh); prefetch_tmap(&tmAl); prefetch_tmap(&tmBh); prefetch_tmap(&tmBl);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], kEpiWarps); }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto tile_coords = [&](int tile, int& mt, int& nt, int& sp) {
    nt = tile % p.n_tiles;
    const int r = tile / p.n_tiles;
    mt = r % p.m_tiles;
    sp = r / p.m_tiles;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int mt, nt, sp; tile_coords(tile, mt, nt, sp);
        const int kb0 = sp * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          unsigned char* st = smem + stage * S::kStage;
          mbar_expect_tx(&full[stage], S::kTxBytes);
          const int k = kb * TBK;
          if (A_KM) {
            tma_load_2d(st, &tmAh, k, mt * TBM, &full[stage]);
            tma_load_2d(st + S::kAOp, &tmAl, k, mt * TBM, &full[stage]);
          } else {
#pragma unroll
            for (int blk = 0; blk < TBM / 32; ++blk) {
              tma_load_2d(st + blk * (TBK * 128), &tmAh, mt * TBM + blk * 32, k, &full[stage]);
              tma_load_2d(st + S::kAOp + blk * (TBK * 128), &tmAl, mt * TBM + blk * 32, k, &full[stage]);
            }
          }
          unsigned char* sb = st + 2 * S::kAOp;
          if (B_KM) {
            tma_load_2d(sb, &tmBh, k, nt * BN, &full[stage]);
            tma_load_2d(sb + S::kBOp, &tmBl, k, nt * BN, &full[stage]);
          } else {
#pragma unroll
            for (int blk = 0; blk < BN / 32; ++blk) {
              tma_load_2d(sb + blk * (TBK * 128), &tmBh, nt * BN + blk * 32, k, &full[stage]);
              tma_load_2d(sb + S::kBOp + blk * (TBK * 128), &tmBl, nt * BN + blk * 32, k, &full[stage]);
            }
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_KM ? 0u : 1u) << 15) |
                                 ((B_KM ? 0u : 1u) << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);
      // K-major: rows of TBK*4 bytes (128 B -> SWIZZLE_128B, 64 B -> SWIZZLE_64B), 8-row atoms, LBO unused.
      // MN-major: LBO = one 32-wide block = TBK rows * 128 B, SBO 512 (4-row atoms), SWIZZLE_128B_BASE32B.
      constexpr uint32_t k_sbo = 8 * TBK * 4, k_lt = (TBK == 32) ? 2 : 4;
      constexpr uint32_t a_lbo = A_KM ? 16 : TBK * 128, b_lbo = B_KM ? 16 : TBK * 128;
      constexpr uint32_t a_sbo = A_KM ? k_sbo : 512, b_sbo = B_KM ? k_sbo : 512;
      constexpr uint32_t a_lt = A_KM ? k_lt : 1, b_lt = B_KM ? k_lt : 1;
      constexpr uint32_t a_kstep = A_KM ? UMMA_K * 4 : UMMA_K * 128;   // bytes per k-step of 8
      constexpr uint32_t b_kstep = B_KM ? UMMA_K * 4 : UMMA_K * 128;
      int stage = 0; uint32_t phase = 0;
      uint32_t chunk_ctr = 0;                         // TMEM stage = chunk_ctr & 1, across tiles
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int mt, nt, sp; tile_coords(tile, mt, nt, sp);
        const int kb0 = sp * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        for (int kc = kb0; kc < kb1; kc += p.kb_per_chunk, ++chunk_ctr) {
          const int as = chunk_ctr & 1;
          mbar_wait(&tempty[as], ((chunk_ctr >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
          uint32_t accum = 0;                         // every chunk starts a fresh accumulation chain
          const int kce = min(kb1, kc + p.kb_per_chunk);
          for (int kb = kc; kb < kce; ++kb) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + stage * S::kStage);
            const uint32_t sb = sa + 2 * S::kAOp;
#pragma unroll
            for (int ks = 0; ks < TBK / UMMA_K; ++ks) {
              const uint64_t ah = make_desc(sa + ks * a_kstep, a_lbo, a_sbo, a_lt);
              const uint64_t al = make_desc(sa + S::kAOp + ks * a_kstep, a_lbo, a_sbo, a_lt);
              const uint64_t bh = make_desc(sb + ks * b_kstep, b_lbo, b_sbo, b_lt);
              const uint64_t bl = make_desc(sb + S::kBOp + ks * b_kstep, b_lbo, b_sbo, b_lt);
              umma_tf32(d_tmem, al, bh, idesc, accum);
              accum = 1;
              umma_tf32(d_tmem, ah, bl, idesc, 1);
              umma_tf32(d_tmem, ah, bh, idesc, 1);
            }
            umma_commit(&empty[stage]);               // frees the smem stage when these MMAs retire
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
          umma_commit(&tfull[as]);                     // chunk accumulator complete
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== accumulate + epilogue =====================
    constexpr int HALF = BN / 2;                       // columns owned by this warp
    const int q = warp & 3;                            // TMEM lane quarter this warp may read
    const int ch = (warp - 4) >> 2;                    // column half
    uint32_t chunk_ctr = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int mt, nt, sp; tile_coords(tile, mt, nt, sp);
      const int kb0 = sp * p.kb_per_split;
      const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
      float acc[HALF];
#pragma unroll
      for (int e = 0; e < HALF; ++e) acc[e] = 0.f;
      for (int kc = kb0; kc < kb1; kc += p.kb_per_chunk, ++chunk_ctr) {
        const int as = chunk_ctr & 1;
        mbar_wait(&tfull[as], (chunk_ctr >> 1) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + ch * HALF);
#pragma unroll
        for (int c0 = 0; c0 < HALF; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(taddr + c0, r);
#pragma unroll
          for (int e = 0; e < 32; ++e) acc[c0 + e] += __uint_as_float(r[e]);   // round-to-nearest adds
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[as]);
      }
      const int row = mt * TBM + q * 32 + lane;
      const int col0 = nt * BN + ch * HALF;
      if (row < p.M && col0 < p.N) {
        float* crow = p.C + (size_t)sp * p.split_stride + (size_t)row * p.ldc + col0;
        const bool vec_ok = (p.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) &&
                            (p.split_stride % 4 == 0) && (col0 + HALF <= p.N);
        if (vec_ok) {
#pragma unroll
          for (int v4 = 0; v4 < HALF / 4; ++v4)
            *reinterpret_cast<float4*>(crow + 4 * v4) =
                make_float4(acc[4 * v4], acc[4 * v4 + 1], acc[4 * v4 + 2], acc[4 * v4 + 3]);
        } else {
#pragma unroll
          for (int e = 0; e < HALF; ++e)
            if (col0 + e < p.N) crow[e] = acc[e];
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---- operand split ---------------------------------------------------------------------------------
__global__ void split_tf32_kernel(const float* __restrict__ src, float* __restrict__ hi, float* __restrict__ lo,
                                  size_t n4, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n4) {
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    float4 h, l;
    uint32_t a, b;
    split_tf32(v.x, a, b); h.x = __uint_as_float(a); l.x = __uint_as_float(b);
    split_tf32(v.y, a, b); h.y = __uint_as_float(a); l.y = __uint_as_float(b);
    split_tf32(v.z, a, b); h.z = __uint_as_float(a); l.z = __uint_as_float(b);
    split_tf32(v.w, a, b); h.w = __uint_as_float(a); l.w = __uint_as_float(b);
    reinterpret_cast<float4*>(hi)[i] = h;
    reinterpret_cast<float4*>(lo)[i] = l;
  }
  if (i == 0) {
    for (size_t k = n4 * 4; k < n; ++k) {
      uint32_t a, b;
      split_tf32(src[k], a, b);
      hi[k] = __uint_as_float(a);
      lo[k] = __uint_as_float(b);
    }
  }
}

int split_tf32(const float* src, float* hi, float* lo, size_t n, cudaStream_t st) {
  if (!aligned16(src) || !aligned16(hi) || !aligned16(lo))
    return fail(SPOTV2_ERR_INVALID_ARG, "split_tf32: pointers must be 16-byte aligned");
  const size_t n4 = n / 4;
  const size_t blocks = (std::max<size_t>(n4, 1) + 255) / 256;
  split_tf32_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, hi, lo, n4, n);
  SPOTV2_CUDA_OK(cudaGetLastError());
  return SPOTV2_OK;
}

// ---- host side (tensor-map helpers: tma.cuh) --------------------------------------------------------
bool tc_gemm_supported(bool a_kc, bool b_kc, int M, int N, int K, const float* A_hi, int lda, const float* B_hi,
                       int ldb) {
  if (!aligned16(A_hi) || !aligned16(B_hi)) return false;
  if (lda % 4 != 0 || ldb % 4 != 0) return false;          // TMA: 16-byte row pitch
  if (M < 1 || N < 1 || K < 1) return false;
  (void)a_kc; (void)b_kc;
  return tma_available();
}

template <int BN, int TBK, bool A_KM, bool B_KM>
static int launch_tc(const CUtensorMap& tAh, const CUtensorMap& tAl, const CUtensorMap& tBh, const CUtensorMap& tBl,
                     const TcParams& p, cudaStream_t st) {
  using S = TcSmem<BN, TBK, A_KM, B_KM>;
  auto kern = gemm3x_tf32_kernel<BN, TBK, A_KM, B_KM>;
  SPOTV2_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
  int grid = sm_count();
  const int total = p.m_tiles * p.n_tiles * p.splits;
  if (grid > total) grid = total;
  kern<<<grid, kTcThreads, S::kTotal, st>>>(tAh, tAl, tBh, tBl, p);
  SPOTV2_CUDA_OK(cudaGetLastError());
  return SPOTV2_OK;
}

__global__ void tc_splitk_reduce_kernel(const float* __restrict__ ws, int splits, size_t split_stride, int M, int N,
                                        float* __restrict__ C, int ldc) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)M * N) return;
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += ws[(size_t)z * split_stride + idx];
  const size_t m = idx / N, n = idx - m * N;
  C[m * ldc + n] = s;
}

// C[M,N] = A . B^T with pre-split operands.  a_kc / b_kc as in sgemm_simt.  bn = 128 or 256.
int gemm3x_tf32(bool a_kc, bool b_kc, int M, int N, int K, const float* A_hi, const float* A_lo, int lda,
                const float* B_hi, const float* B_lo, int ldb, float* C, int ldc, int splits, int bn,
                int kb_per_chunk, void* ws, size_t ws_bytes, cudaStream_t st) {
  // bn encodes the tile: 256 | 128 select BK = 32; 256 + 16 | 128 + 16 select BK = 16 (deeper ring).
  const int TBK = (bn & 16) ? 16 : 32;
  bn &= ~16;
  if (!tc_gemm_supported(a_kc, b_kc, M, N, K, A_hi, lda, B_hi, ldb))
    return fail(SPOTV2_ERR_UNSUPPORTED, "tensor-core GEMM needs 16-byte aligned operands and leading dimensions % 4 == 0");
  TcParams p;
  p.M = M; p.N = N; p.K = K;
  p.m_tiles = (M + TBM - 1) / TBM;
  p.n_tiles = (N + bn - 1) / bn;
  p.kb_total = (K + TBK - 1) / TBK;
  if (splits < 1) splits = 1;
  if (splits > p.kb_total) splits = p.kb_total;
  p.kb_per_split = (p.kb_total + splits - 1) / splits;
  p.splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  p.kb_per_chunk = kb_per_chunk < 1 ? 128 / TBK : kb_per_chunk;     // default: 128-element accumulation chains
  p.C = C; p.ldc = ldc; p.split_stride = 0;
  if (p.splits > 1) {
    const size_t need = (size_t)p.splits * M * N * sizeof(float);
    if (!ws || ws_bytes < need)
      return fail(SPOTV2_ERR_WORKSPACE, "split-K tensor-core GEMM needs %zu B of workspace, got %zu", need, ws_bytes);
    p.C = static_cast<float*>(ws); p.ldc = N; p.split_stride = (size_t)M * N;
  }
  CUtensorMap tAh, tAl, tBh, tBl;
  int rc;
  // K-major operand: tensor [rows = M|N, cols = K], box 32(k) x tile rows.
  // MN-major operand: tensor [rows = K, cols = M|N], box 32(m|n) x 32(k); one box per 32-wide block.
  const CUtensorMapSwizzle kSw = TBK == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  const CUtensorMapSwizzle mnSw = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  if (a_kc) {
    if ((rc = make_tmap(&tAh, A_hi, M, K, lda, TBK, TBM, kSw))) return rc;
    if ((rc = make_tmap(&tAl, A_lo, M, K, lda, TBK, TBM, kSw))) return rc;
  } else {
    if ((rc = make_tmap(&tAh, A_hi, K, M, lda, 32, TBK, mnSw))) return rc;
    if ((rc = make_tmap(&tAl, A_lo, K, M, lda, 32, TBK, mnSw))) return rc;
  }
  if (b_kc) {
    if ((rc = make_tmap(&tBh, B_hi, N, K, ldb, TBK, bn, kSw))) return rc;
    if ((rc = make_tmap(&tBl, B_lo, N, K, ldb, TBK, bn, kSw))) return rc;
  } else {
    if ((rc = make_tmap(&tBh, B_hi, K, N, ldb, 32, TBK, mnSw))) return rc;
    if ((rc = make_tmap(&tBl, B_lo, K, N, ldb, 32, TBK, mnSw))) return rc;
  }
#define SPOTV2_TC(BN_, TBK_, AK, BK_) rc = launch_tc<BN_, TBK_, AK, BK_>(tAh, tAl, tBh, tBl, p, st)
#define SPOTV2_TC_MAJ(BN_, TBK_)                            \
  do {                                                      \
    if (a_kc && b_kc) SPOTV2_TC(BN_, TBK_, true, true);     \
    else if (a_kc) SPOTV2_TC(BN_, TBK_, true, false);       \
    else if (b_kc) SPOTV2_TC(BN_, TBK_, false, true);       \
    else SPOTV2_TC(BN_, TBK_, false, false);                \
  } while (0)
  if (bn == 256 && TBK == 32) SPOTV2_TC_MAJ(256, 32);
  else if (bn == 256 && TBK == 16) SPOTV2_TC_MAJ(256, 16);
  else if (bn == 128 && TBK == 32) SPOTV2_TC_MAJ(128, 32);
  else return fail(SPOTV2_ERR_INVALID_ARG, "bn must be 128, 256 or 256+16");
#undef SPOTV2_TC_MAJ
#undef SPOTV2_TC
  if (rc) return rc;
  if (p.splits > 1) {
    const size_t total = (size_t)M * N;
    tc_splitk_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(p.C, p.splits, p.split_stride, M, N, C, ldc);
    SPOTV2_CUDA_OK(cudaGetLastError());
  }
  return SPOTV2_OK;
}

}  // namespace spotv2

// Test / bring-up hook (not part of the reference-facing ABI; declared in include/spotv2_gat.h under
// "diagnostics"): a generic 3xTF32 GEMM on caller-provided fp32 operands.
extern "C" int spotv2_diag_gemm(int a_kc, int b_kc, int M, int N, int K, const float* A, int lda, const float* B,
                                int ldb, float* C, int ldc, int algo, int splits, int bn, int kb_per_chunk, void* ws,
                                size_t ws_bytes, void* stream) {
  using namespace spotv2;
  cudaStream_t st = as_stream(stream);
  if (algo == 1) return sgemm_simt(a_kc, b_kc, M, N, K, A, lda, B, ldb, C, ldc, splits, ws, ws_bytes, st);
  if (algo == 3) {
    // fp16-pair path: [A_hi | A_lo | B_hi | B_lo | scale blocks | split-K partials], per-tensor scales
    const int a_rows = a_kc ? M : K, a_cols = a_kc ? K : M, b_rows = b_kc ? N : K, b_cols = b_kc ? K : N;
    const int lda16 = ld16_of(a_cols), ldb16 = ld16_of(b_cols);
    const size_t a_bytes = round_up((size_t)a_rows * lda16 * 2, 256), b_bytes = round_up((size_t)b_rows * ldb16 * 2, 256);
    const size_t need = 2 * a_bytes + 2 * b_bytes + 256;
    if (!ws || ws_bytes < need) return fail(SPOTV2_ERR_WORKSPACE, "diag_gemm needs at least %zu B of workspace", need);
    unsigned char* w = static_cast<unsigned char*>(ws);
    float* blk = reinterpret_cast<float*>(w + 2 * a_bytes + 2 * b_bytes);
    const int none = 0x7fffffff;
    if (int rc = split_f16(A, a_rows, a_cols, lda, 0, none, nullptr, 0, w, w + a_bytes, lda16, blk, st)) return rc;
    if (int rc = split_f16(B, b_rows, b_cols, ldb, 0, none, nullptr, 0, w + 2 * a_bytes, w + 2 * a_bytes + b_bytes, ldb16,
                           blk + kScaleBlockFloats, st))
      return rc;
    F16Operand opA{w, w + a_bytes, lda16, blk + 2, none};
    F16Operand opB{w + 2 * a_bytes, w + 2 * a_bytes + b_bytes, ldb16, blk + kScaleBlockFloats + 2, none};
    return gemm3x_f16(a_kc, b_kc, M, N, K, opA, opB, C, ldc, splits, bn, kb_per_chunk, w + need, ws_bytes - need, st);
  }
  // workspace layout: [A_hi | A_lo | B_hi | B_lo | split-K partials]
  const size_t a_elems = (size_t)(a_kc ? M : K) * lda, b_elems = (size_t)(b_kc ? N : K) * ldb;
  const size_t a_bytes = round_up(a_elems * 4, 256), b_bytes = round_up(b_elems * 4, 256);
  if (!ws || ws_bytes < 2 * a_bytes + 2 * b_bytes)
    return fail(SPOTV2_ERR_WORKSPACE, "diag_gemm needs at least %zu B of workspace", 2 * a_bytes + 2 * b_bytes);
  unsigned char* w = static_cast<unsigned char*>(ws);
  float* Ah = reinterpret_cast<float*>(w);
  float* Al = reinterpret_cast<float*>(w + a_bytes);
  float* Bh = reinterpret_cast<float*>(w + 2 * a_bytes);
  float* Bl = reinterpret_cast<float*>(w + 2 * a_bytes + b_bytes);
  if (int rc = split_tf32(A, Ah, Al, a_elems, st)) return rc;
  if (int rc = split_tf32(B, Bh, Bl, b_elems, st)) return rc;
  return gemm3x_tf32(a_kc, b_kc, M, N, K, Ah, Al, lda, Bh, Bl, ldb, C, ldc, splits, bn, kb_per_chunk,
                     w + 2 * a_bytes + 2 * b_bytes,
                     ws_bytes - (2 * a_bytes + 2 * b_bytes), st);
}
