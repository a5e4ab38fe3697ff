// Fused GAT attention backward, pipelined version (the default; attn_bwd.cu is the any-shape fallback).
// Same mathematics as attn_bwd.cu (attention recomputed from P_aug and the edge
// rows), organised like the forward kernel: one persistent CTA per SM, 12 compute warps + 1 producer warp,
// every operand arrives through asynchronous copies that run ahead of the arithmetic, across phases and
// across graphs, and every product runs on the tensor cores (mma.sync m16n8k8 TF32 with the 3x split).
//
//   producer warp   two independent streams, polled by one lane:
//                   * edge rows: 1-D bulk copies into a 2-stage ring (twice per graph: logits, then dv)
//                   * tile groups: 32x32 fp32 TMA tiles (128B-swizzled) of dout and P, 2 groups of <= 7 tiles
//   compute warps   per graph
//     L  g[e,h] = <edge row, v_h>        12 warps = 6 m16 row tiles x 2 k halves, red.shared into the tile
//     S  self-loop mean fill, LeakyReLU, softmax -> alpha[h][j][i], z>0 masks       (thread per (h, i))
//     A  dalpha_h = g dO_h P_h^T         warp = (head, 16-target tile); K = channels streamed as tile groups;
//        softmax/LeakyReLU backward directly on the accumulator fragments (row sums by 4-lane shuffles),
//        dd -> global, ds partials, dz' (mean-fill redistributed) -> shared D tile
//     D  dP_h = g alpha_h^T dO_h         warp = (head, 16-source tile), alpha^T fragments in registers,
//        results stored as scaled fp16 hi/lo pairs (tensor-core GEMM operand) or fp32; dbias column sums
//     V  dv^T[f,h] += T^T dz'            warp = (16-feature tile, row group) over the second edge pass
// Measured mma.sync facts this layout relies on (profiles/r1_mma_sync_latency_throughput.txt): 21-cycle
// dependent latency, one MMA per 8 cycles per SM sub-partition, so 3 chains per warp keep a tensor unit busy.
#include <string.h>

#include "attn_bwd.cuh"
#include "tma.cuh"

namespace spotv2 {

namespace {

constexpr int kW = 12;                  // compute warps
constexpr int kCT = kW * 32;            // compute threads
constexpr int kB2Threads = kCT + 32;    // + producer warp
constexpr int kTile = 4096;             // 32 rows x 32 fp32, 128B-swizzled
constexpr int kGrpTiles = 7;            // tiles per group slot
constexpr int kNS2 = 36;                // alpha / D tile row stride: (g*4 + t) fragment reads hit 32 banks
constexpr int kMaxDvUnits = 4;          // (feature tile, row group) units per warp in phase V

__device__ unsigned long long g_bwd2_counters[kNumCounters];

struct Bwd2Plan {
  int KS, NT, chunk_rows, nchunks, n_cb, hpr, n_rounds, n_mt_chunk, ksplit;
  int n_mtiles, dv_rg, dv_rpu, dv_units, cbs_per_grp_d, tma_ok;
  uint32_t off_bar, off_table, off_vfrag, off_sd, off_mask, off_dspart, off_dbias, off_tile, off_D, off_ring,
      ring_stage, off_grp, total;
};

Bwd2Plan make_plan(const AttnParams& p) {
  Bwd2Plan s{};
  const int N = p.N, H = p.H, C = p.C, Fe = p.Fe;
  s.KS = ((Fe + 7) / 8 + 7) / 8 * 8;
  s.NT = 1;
  s.n_cb = (C + 31) / 32;
  s.hpr = p.concat ? 3 : 6;
  s.n_rounds = (H + s.hpr - 1) / s.hpr;
  const int nh_max = H < s.hpr ? H : s.hpr;
  s.cbs_per_grp_d = p.concat ? (kGrpTiles / nh_max > 0 ? kGrpTiles / nh_max : 1) : kGrpTiles;
  s.tma_ok = (C % 4 == 0) ? 1 : 0;
  uint32_t o = 0;
  s.off_bar = o;    o += 128;
  s.off_table = o;  o += (uint32_t)round_up((size_t)(p.R > 0 ? p.R : 1) * 4, 16);
  s.off_vfrag = o;  o += (uint32_t)((Fe > 0 ? s.KS : 0) * 32 * 16);
  s.off_sd = o;     o += (uint32_t)round_up((size_t)N * 2 * H * 4, 16);
  s.off_mask = o;   o += (uint32_t)round_up((size_t)H * N * 4, 16);
  s.off_dspart = o; o += (uint32_t)(2 * H * 32 * 4);
  s.off_dbias = o;  o += (uint32_t)round_up((size_t)p.ldo * 4, 16);
  s.off_tile = o;   o += (uint32_t)round_up((size_t)H * N * kNS2 * 4, 16);
  s.off_D = o;      o += (uint32_t)round_up((size_t)H * N * kNS2 * 4, 16);
  s.off_grp = (uint32_t)round_up(o, 1024);
  o = s.off_grp + 2 * kGrpTiles * kTile;
  s.off_ring = o;                                        // 1024-aligned
  // largest edge ring stage (multiple of 16 rows, <= 96) that fits
  s.chunk_rows = 0;
  s.ring_stage = 0;
  if (Fe > 0) {
    for (int rows = 96; rows >= 16; rows -= 16) {
      const uint32_t st = (uint32_t)round_up((size_t)rows * Fe * 4, 128);
      if (o + 2 * st <= 227 * 1024) { s.chunk_rows = rows; s.ring_stage = st; break; }
    }
    if (s.chunk_rows == 0) { s.total = 0xffffffffu; return s; }
    if (s.chunk_rows > p.R) s.chunk_rows = (p.R + 15) / 16 * 16, s.ring_stage = (uint32_t)round_up((size_t)s.chunk_rows * Fe * 4, 128);
    s.nchunks = (p.R + s.chunk_rows - 1) / s.chunk_rows;
  }
  s.total = o + 2 * s.ring_stage;
  s.n_mt_chunk = s.chunk_rows / 16;
  s.ksplit = (s.n_mt_chunk > 0 && 2 * s.n_mt_chunk <= kW && s.KS >= 16) ? 2 : 1;
  // phase V units: (16-feature tile, row group)
  s.n_mtiles = (Fe + 15) / 16;
  if (Fe > 0) {
    int rg = (kMaxDvUnits * kW) / s.n_mtiles;
    const int rg_max = (s.chunk_rows + 7) / 8;
    if (rg > rg_max) rg = rg_max;
    if (rg > 3) rg = 3;
    if (rg < 1) { s.total = 0xffffffffu; return s; }
    s.dv_rg = rg;
    s.dv_rpu = ((s.chunk_rows + rg - 1) / rg + 7) / 8 * 8;
    s.dv_units = s.n_mtiles * rg;
  }
  return s;
}

__device__ __forceinline__ void bar_sync_compute() { asm volatile("bar.sync 1, 384;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive2(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float ld_tile(const unsigned char* t, int r, int c) {
  float v;
  const uint32_t a = smem_u32(t) + r * 128 + ((((c >> 2) ^ (r & 7)) << 4) | ((c & 3) << 2));
  asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void red_add_shared(float* p, float v) {
  asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(smem_u32(p)), "f"(v) : "memory");
}

// Edge terms of the 16 rows [m0, m0+16) of a staged chunk for the k-blocks kb0, kb0+kb_stride, ... (8 k-steps
// each): the same branch-free 3xTF32 loop as the forward (attn_common.cuh), results added into the tile.
template <class Sink>
__device__ __forceinline__ void edge_logits_part(const float* Ts, const float4* vfrag, int Fe, int KS, int m0, int lane,
                                                 int kb0, int kb_stride, Sink&& sink) {
  constexpr int kSlots = 8;
  const int g = lane >> 2, t = lane & 3;
  float acc[3][4];
#pragma unroll
  for (int pr = 0; pr < 3; ++pr)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[pr][q] = 0.f;
  const float* r0 = Ts + (size_t)(m0 + g) * Fe;
  const float* r1 = r0 + (size_t)8 * Fe;
  const int kmax = Fe - 1;
  for (int ks0 = kb0 * kSlots; ks0 < KS; ks0 += kb_stride * kSlots) {
    float a[kSlots][4];
    float4 bf[kSlots];
#pragma unroll
    for (int sl = 0; sl < kSlots; ++sl) {
      const int k0 = min((ks0 + sl) * 8 + t, kmax), k1 = min((ks0 + sl) * 8 + t + 4, kmax);
      a[sl][0] = lds_f32(r0 + k0);
      a[sl][1] = lds_f32(r1 + k0);
      a[sl][2] = lds_f32(r0 + k1);
      a[sl][3] = lds_f32(r1 + k1);
      bf[sl] = vfrag[(ks0 + sl) * 32 + lane];
    }
#pragma unroll
    for (int sl = 0; sl < kSlots; ++sl) {
      uint32_t ah[4], al[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) split_tf32_trunc(a[sl][q], ah[q], al[q]);
      const uint32_t bh[2] = {__float_as_uint(bf[sl].x), __float_as_uint(bf[sl].y)};
      const uint32_t bl[2] = {__float_as_uint(bf[sl].z), __float_as_uint(bf[sl].w)};
      mma_tf32_16x8x8(acc[0], al, bh);
      mma_tf32_16x8x8(acc[1], ah, bl);
      mma_tf32_16x8x8(acc[2], ah, bh);
    }
  }
  const int n = 2 * t;
#pragma unroll
  for (int q = 0; q < 4; ++q) sink(m0 + g + ((q & 2) ? 8 : 0), n + (q & 1), (acc[0][q] + acc[1][q]) + acc[2][q]);
}

__global__ void __launch_bounds__(kB2Threads, 1)
gat_attn_bwd2_kernel(const AttnBwdArgs args, const Bwd2Plan pl, const __grid_constant__ CUtensorMap tmP,
                     const __grid_constant__ CUtensorMap tmG) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const AttnParams& p = args.p;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = p.N, H = p.H, C = p.C, Fe = p.Fe, HC = H * C;
  constexpr int NS = kNS2;
  const int tile_floats = H * N * NS;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + pl.off_bar);
  uint64_t* edge_full = bars;          // [2]
  uint64_t* edge_empty = bars + 2;     // [2]
  uint64_t* grp_full = bars + 4;       // [2]
  uint64_t* grp_empty = bars + 6;      // [2]
  int32_t* table_s = reinterpret_cast<int32_t*>(smem_raw + pl.off_table);
  float4* vfrag = reinterpret_cast<float4*>(smem_raw + pl.off_vfrag);
  float* sd = reinterpret_cast<float*>(smem_raw + pl.off_sd);
  uint32_t* pos_mask = reinterpret_cast<uint32_t*>(smem_raw + pl.off_mask);
  float* ds_part = reinterpret_cast<float*>(smem_raw + pl.off_dspart);     // [2][H][32]
  float* dbias_s = reinterpret_cast<float*>(smem_raw + pl.off_dbias);
  float* tile = reinterpret_cast<float*>(smem_raw + pl.off_tile);          // alpha[h][j][i]
  float* D = reinterpret_cast<float*>(smem_raw + pl.off_D);                // dz'[h][j][i]
  unsigned char* grp0 = smem_raw + pl.off_grp;
  float* stage0 = reinterpret_cast<float*>(smem_raw + pl.off_ring);
  float* stage1 = reinterpret_cast<float*>(smem_raw + pl.off_ring + pl.ring_stage);
  __builtin_assume(__isShared(table_s)); __builtin_assume(__isShared(vfrag)); __builtin_assume(__isShared(sd));
  __builtin_assume(__isShared(pos_mask)); __builtin_assume(__isShared(ds_part)); __builtin_assume(__isShared(dbias_s));
  __builtin_assume(__isShared(tile)); __builtin_assume(__isShared(D)); __builtin_assume(__isShared(stage0));
  __builtin_assume(__isShared(stage1));

  const int nchunks = pl.nchunks;
  const int my_graphs = (p.B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int n_cb = pl.n_cb;
  auto rows_in = [&](int c) { const int r = p.R - c * pl.chunk_rows; return r < pl.chunk_rows ? r : pl.chunk_rows; };

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&edge_full[s], 1); mbar_init(&edge_empty[s], kW);
      mbar_init(&grp_full[s], 1);  mbar_init(&grp_empty[s], kW);
    }
    fence_mbar_init();
  }
  for (int r = tid; r < p.R; r += kB2Threads) table_s[r] = Fe > 0 ? p.table[r] : -1;
  if (Fe > 0) build_vfrag(vfrag, p.v, H, Fe, pl.KS, 1, tid, kB2Threads);
  for (int idx = tid; idx < tile_floats; idx += kB2Threads) { tile[idx] = 0.f; D[idx] = 0.f; }
  for (int idx = tid; idx < p.ldo; idx += kB2Threads) dbias_s[idx] = 0.f;
  __syncthreads();

  if (warp == kW) {
    // =========================================== producer ===========================================
    if (lane == 0) { prefetch_tmap(&tmP); prefetch_tmap(&tmG); }
    const long long n_rows = (long long)p.B * N;
    // one tile: TMA, or (C % 4 != 0: tile starts are not 16-byte aligned) a cooperative gather into the
    // same swizzled layout
    auto put_tile = [&](unsigned char* dst, const CUtensorMap* tm, const float* base, int ld, int col0, int row0, uint64_t* full) {
      if (pl.tma_ok) {
        if (lane == 0) tma_load_2d(dst, tm, col0, row0, full);
      } else {
        for (int idx = lane; idx < 32 * 32; idx += 32) {
          const int r = idx >> 5, c = idx & 31;
          float v = 0.f;
          if (row0 + r < n_rows && col0 + c < ld) v = base[((size_t)row0 + r) * ld + col0 + c];
          *reinterpret_cast<float*>(dst + r * 128 + ((((c >> 2) ^ (r & 7)) << 4) | ((c & 3) << 2))) = v;
        }
      }
    };
    // edge stream state
    int ek = 0, e_it = 0, e_pass = 0, e_c = 0;
    bool e_done = (nchunks == 0 || my_graphs == 0);
    // group stream state: phase 0 = A (per round, per channel block), phase 1 = D (per round, per dO group)
    int gk = 0, g_it = 0, g_phase = 0, g_r = 0, g_i = 0;
    bool g_done = (my_graphs == 0);
    const int n_grp_d = (n_cb + pl.cbs_per_grp_d - 1) / pl.cbs_per_grp_d;
    long long t_idle0 = clock64();
    while (!e_done || !g_done) {
      bool progressed = false;
      if (!e_done) {
        const int s = ek & 1;
        int ok = (lane == 0) ? (int)mbar_try_wait(&edge_empty[s], ((ek >> 1) & 1) ^ 1) : 0;
        ok = __shfl_sync(0xffffffffu, ok, 0);
        if (ok) {
          const int b = blockIdx.x + e_it * gridDim.x;
          const int rows = rows_in(e_c);
          const float* src = p.edge_rows + ((size_t)b * p.R + (size_t)e_c * pl.chunk_rows) * Fe;
          float* dst = s ? stage1 : stage0;
          if (p.bulk_ok) {
            if (lane == 0) {
              const uint32_t bytes = (uint32_t)rows * Fe * 4u;
              mbar_expect_tx(&edge_full[s], bytes);
              bulk_g2s(dst, src, bytes, &edge_full[s]);
            }
          } else {
            for (int idx = lane; idx < rows * Fe; idx += 32) dst[idx] = src[idx];
            __syncwarp();
            if (lane == 0) mbar_arrive2(&edge_full[s]);
          }
          ++ek;
          if (++e_c == nchunks) { e_c = 0; if (++e_pass == 2) { e_pass = 0; if (++e_it == my_graphs) e_done = true; } }
          progressed = true;
        }
      }
      if (!g_done) {
        const int s = gk & 1;
        int ok = (lane == 0) ? (int)mbar_try_wait(&grp_empty[s], ((gk >> 1) & 1) ^ 1) : 0;
        ok = __shfl_sync(0xffffffffu, ok, 0);
        if (ok) {
          const int b = blockIdx.x + g_it * gridDim.x;
          unsigned char* gb = grp0 + s * (kGrpTiles * kTile);
          const int h0 = g_r * pl.hpr;
          const int nh = min(pl.hpr, H - h0);
          int ntiles;
          if (g_phase == 0) ntiles = p.concat ? 2 * nh : 1 + nh;
          else {
            const int cb0 = g_i * pl.cbs_per_grp_d;
            const int ncb = min(pl.cbs_per_grp_d, n_cb - cb0);
            ntiles = p.concat ? ncb * nh : ncb;
          }
          if (pl.tma_ok && lane == 0) mbar_expect_tx(&grp_full[s], (uint32_t)ntiles * kTile);
          if (g_phase == 0) {
            const int cb = g_i;
            if (p.concat) {
              for (int hl = 0; hl < nh; ++hl) {
                put_tile(gb + (2 * hl) * kTile, &tmG, args.dout, p.ldo, (h0 + hl) * C + cb * 32, b * N, &grp_full[s]);
                put_tile(gb + (2 * hl + 1) * kTile, &tmP, p.P_aug, p.ldp, (h0 + hl) * C + cb * 32, b * N, &grp_full[s]);
              }
            } else {
              put_tile(gb, &tmG, args.dout, p.ldo, cb * 32, b * N, &grp_full[s]);
              for (int hl = 0; hl < nh; ++hl)
                put_tile(gb + (1 + hl) * kTile, &tmP, p.P_aug, p.ldp, (h0 + hl) * C + cb * 32, b * N, &grp_full[s]);
            }
          } else {
            const int cb0 = g_i * pl.cbs_per_grp_d;
            const int ncb = min(pl.cbs_per_grp_d, n_cb - cb0);
            for (int k = 0; k < ncb; ++k) {
              if (p.concat) {
                for (int hl = 0; hl < nh; ++hl)
                  put_tile(gb + (k * nh + hl) * kTile, &tmG, args.dout, p.ldo, (h0 + hl) * C + (cb0 + k) * 32, b * N, &grp_full[s]);
              } else {
                put_tile(gb + k * kTile, &tmG, args.dout, p.ldo, (cb0 + k) * 32, b * N, &grp_full[s]);
              }
            }
          }
          if (!pl.tma_ok) {
            __syncwarp();
            if (lane == 0) mbar_arrive2(&grp_full[s]);
          }
          ++gk;
          const int lim = (g_phase == 0) ? n_cb : n_grp_d;
          if (++g_i == lim) {
            g_i = 0;
            if (++g_r == pl.n_rounds) { g_r = 0; if (++g_phase == 2) { g_phase = 0; if (++g_it == my_graphs) g_done = true; } }
          }
          progressed = true;
        }
      }
      if (progressed) {
        t_idle0 = clock64();
      } else {
        __nanosleep(40);
        if (clock64() - t_idle0 > 4000000000LL) __trap();
      }
    }
    return;
  }

  // ============================================== compute ==============================================
  const int g = lane >> 2, t = lane & 3;
  const float g_scale = p.concat ? 1.f : 1.f / (float)H;
  const float inv_nm1 = 1.f / (float)(N > 1 ? N - 1 : 1);
  float dp_scale = 1.f;
  if (args.dP_hi16) {
    dp_scale = dp_scale_from_amax(__uint_as_float(*reinterpret_cast<const unsigned*>(args.dout_blk)) * args.bound);
    if (blockIdx.x == 0 && tid == 0) { args.dp_blk[2] = 1.f / dp_scale; args.dp_blk[4] = dp_scale; }
  }
  AttnSmem asm_{};                        // what softmax_phase reads
  asm_.NS = NS; asm_.KS = pl.KS; asm_.NT = 1;

  float dv_run[kMaxDvUnits][4];
#pragma unroll
  for (int u = 0; u < kMaxDvUnits; ++u)
#pragma unroll
    for (int q = 0; q < 4; ++q) dv_run[u][q] = 0.f;

  int ek = 0, gk = 0;                      // consumed edge chunks / tile groups (all compute warps in step)
  long long ph[6] = {0, 0, 0, 0, 0, 0};
  long long t_ph = clock64();
  auto lap = [&](int k) { const long long now = clock64(); ph[k] += now - t_ph; t_ph = now; };

  for (int it = 0; it < my_graphs; ++it) {
    const int b = blockIdx.x + it * gridDim.x;
    // ------------------------------------------------ L: edge logits ------------------------------------------------
    for (int idx = tid; idx < N * 2 * H; idx += kCT) {
      const int j = idx / (2 * H), k = idx - j * 2 * H;
      sd[idx] = p.P_aug[((size_t)b * N + j) * p.ldp + HC + k];
    }
    for (int idx = tid; idx < 2 * H * 32; idx += kCT) ds_part[idx] = 0.f;
    for (int c = 0; c < nchunks; ++c, ++ek) {
      const int s = ek & 1;
      mbar_wait(&edge_full[s], (ek >> 1) & 1);
      const int rows = rows_in(c);
      const int mt = warp % pl.n_mt_chunk, kh = warp / pl.n_mt_chunk;
      if (kh < pl.ksplit && mt * 16 < rows) {
        const int row_base = c * pl.chunk_rows;
        edge_logits_part(s ? stage1 : stage0, vfrag, Fe, pl.KS, mt * 16, lane, kh, pl.ksplit, [&](int r, int h, float val) {
          if (r < rows && h < H) {
            const int code = table_s[row_base + r];
            if (code >= 0) red_add_shared(&tile[(h * N + (code & 0xffff)) * NS + (code >> 16)], val);
          }
        });
      }
      __syncwarp();
      if (lane == 0) mbar_arrive2(&edge_empty[s]);
    }
    bar_sync_compute();
    lap(0);
    // ------------------------------------------------ S: softmax ------------------------------------------------
    softmax_phase(p, asm_, tile, sd, 1.f, nullptr, pos_mask, tid, kCT);
    bar_sync_compute();
    lap(1);
    // ------------------------------------------------ A: dalpha + softmax backward ------------------------------------------------
    for (int r = 0; r < pl.n_rounds; ++r) {
      const int h0 = r * pl.hpr;
      const int nh = min(pl.hpr, H - h0);
      const int hl = warp >> 1, m = warp & 1;
      const bool active = (warp < 2 * nh) && (16 * m < N);
      const int h = h0 + hl;
      const int dslot = p.concat ? 2 * hl : 0, pslot = p.concat ? 2 * hl + 1 : 1 + hl;
      const int i0 = 16 * m + g, i1 = i0 + 8;
      float cmain[4][4], ccorr[4][4], dacc[4][4];
#pragma unroll
      for (int n = 0; n < 4; ++n)
#pragma unroll
        for (int q = 0; q < 4; ++q) cmain[n][q] = ccorr[n][q] = dacc[n][q] = 0.f;
      for (int cb = 0; cb < n_cb; ++cb, ++gk) {
        const int s = gk & 1;
        mbar_wait(&grp_full[s], (gk >> 1) & 1);
        if (active) {
          const unsigned char* gb = grp0 + s * (kGrpTiles * kTile);
          const unsigned char* dOt = gb + dslot * kTile;
          const unsigned char* Pt = gb + pslot * kTile;
          const bool tail = (cb * 32 + 32 > C);          // channels beyond C: next head's columns (concat) or zero fill
          uint32_t ah[4][4], al[4][4];
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const int c0 = 8 * ks + t, c1 = c0 + 4;
            float a0 = ld_tile(dOt, i0, c0), a1 = ld_tile(dOt, i1, c0), a2 = ld_tile(dOt, i0, c1), a3 = ld_tile(dOt, i1, c1);
            if (tail) {
              if (cb * 32 + c0 >= C) a0 = a1 = 0.f;
              if (cb * 32 + c1 >= C) a2 = a3 = 0.f;
            }
            split_tf32_trunc(a0, ah[ks][0], al[ks][0]);
            split_tf32_trunc(a1, ah[ks][1], al[ks][1]);
            split_tf32_trunc(a2, ah[ks][2], al[ks][2]);
            split_tf32_trunc(a3, ah[ks][3], al[ks][3]);
          }
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
            for (int n = 0; n < 4; ++n) {
              const int j = 8 * n + g;
              float b0 = ld_tile(Pt, j, 8 * ks + t), b1 = ld_tile(Pt, j, 8 * ks + t + 4);
              if (tail) {                                  // never let padding bits (possibly NaN) meet the zeros above
                if (cb * 32 + 8 * ks + t >= C) b0 = 0.f;
                if (cb * 32 + 8 * ks + t + 4 >= C) b1 = 0.f;
              }
              uint32_t bh[2], bl[2];
              split_tf32_trunc(b0, bh[0], bl[0]);
              split_tf32_trunc(b1, bh[1], bl[1]);
              mma_tf32_16x8x8(ccorr[n], al[ks], bh);
              mma_tf32_16x8x8(cmain[n], ah[ks], bh);
              mma_tf32_16x8x8(ccorr[n], ah[ks], bl);
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive2(&grp_empty[s]);
        if (active && ((cb & 1) || cb == n_cb - 1)) {       // keep tensor-core accumulation chains short: fold in fp32 RN
#pragma unroll
          for (int n = 0; n < 4; ++n)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              dacc[n][q] += cmain[n][q] + ccorr[n][q];
              cmain[n][q] = ccorr[n][q] = 0.f;
            }
        }
      }
      if (active) {
        // softmax / LeakyReLU backward on the fragment: lane holds rows i0, i1 and columns j = 8n + 2t + {0,1}
        float al_[4][4];
        float dot0 = 0.f, dot1 = 0.f;
#pragma unroll
        for (int n = 0; n < 4; ++n)
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int i = (q & 2) ? i1 : i0, j = 8 * n + 2 * t + (q & 1);
            const bool valid = i < N && j < N;
            const float a = valid ? tile[(h * N + j) * NS + i] : 0.f;
            const float da = valid ? dacc[n][q] * g_scale : 0.f;
            al_[n][q] = a;
            dacc[n][q] = da;
            if (q & 2) dot1 = fmaf(a, da, dot1); else dot0 = fmaf(a, da, dot0);
          }
        dot0 += __shfl_xor_sync(0xffffffffu, dot0, 1); dot0 += __shfl_xor_sync(0xffffffffu, dot0, 2);
        dot1 += __shfl_xor_sync(0xffffffffu, dot1, 1); dot1 += __shfl_xor_sync(0xffffffffu, dot1, 2);
        const uint32_t mask0 = i0 < N ? pos_mask[h * N + i0] : 0u, mask1 = i1 < N ? pos_mask[h * N + i1] : 0u;
        float dd0 = 0.f, dd1 = 0.f, dii0 = 0.f, dii1 = 0.f;
        float dsc[4][2];
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          dsc[n][0] = dsc[n][1] = 0.f;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int i = (q & 2) ? i1 : i0, j = 8 * n + 2 * t + (q & 1);
            const float dl = al_[n][q] * (dacc[n][q] - ((q & 2) ? dot1 : dot0));
            const uint32_t mk = (q & 2) ? mask1 : mask0;
            const float dz = ((mk >> j) & 1u) ? dl : dl * p.slope;
            dacc[n][q] = dz;
            if (q & 2) { dd1 += dz; if (j == i) dii1 = dz; } else { dd0 += dz; if (j == i) dii0 = dz; }
            dsc[n][q & 1] += dz;
          }
        }
        dd0 += __shfl_xor_sync(0xffffffffu, dd0, 1); dd0 += __shfl_xor_sync(0xffffffffu, dd0, 2);
        dd1 += __shfl_xor_sync(0xffffffffu, dd1, 1); dd1 += __shfl_xor_sync(0xffffffffu, dd1, 2);
        dii0 += __shfl_xor_sync(0xffffffffu, dii0, 1); dii0 += __shfl_xor_sync(0xffffffffu, dii0, 2);
        dii1 += __shfl_xor_sync(0xffffffffu, dii1, 1); dii1 += __shfl_xor_sync(0xffffffffu, dii1, 2);
        if (t == 0) {
          if (i0 < N) {
            if (args.dsd) args.dsd[((size_t)b * N + i0) * 2 * H + H + h] = dd0;
            else args.dP_aug[((size_t)b * N + i0) * p.ldp + HC + H + h] = dd0;
          }
          if (i1 < N) {
            if (args.dsd) args.dsd[((size_t)b * N + i1) * 2 * H + H + h] = dd1;
            else args.dP_aug[((size_t)b * N + i1) * p.ldp + HC + H + h] = dd1;
          }
        }
        // ds partial over this warp's 16 targets: sum over the 8 row groups (lane bits 2..4)
#pragma unroll
        for (int n = 0; n < 4; ++n)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            float v = dsc[n][e];
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            if (g == 0) ds_part[(m * H + h) * 32 + 8 * n + 2 * t + e] = v;
          }
        // dz' = dz + dz_ii / (N - 1) off the diagonal, 0 on it (gradient through the mean fill)
        const float sh0 = dii0 * inv_nm1, sh1 = dii1 * inv_nm1;
#pragma unroll
        for (int n = 0; n < 4; ++n)
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int i = (q & 2) ? i1 : i0, j = 8 * n + 2 * t + (q & 1);
            if (i < N && j < N) D[(h * N + j) * NS + i] = (j == i) ? 0.f : dacc[n][q] + ((q & 2) ? sh1 : sh0);
          }
      }
    }
    bar_sync_compute();
    lap(2);
    for (int idx = tid; idx < H * N; idx += kCT) {       // ds_j = sum over both target tiles
      const int h = idx / N, j = idx - h * N;
      const float ds = ds_part[h * 32 + j] + ds_part[(H + h) * 32 + j];
      if (args.dsd) args.dsd[((size_t)b * N + j) * 2 * H + h] = ds;
      else args.dP_aug[((size_t)b * N + j) * p.ldp + HC + h] = ds;
    }
    lap(3);
    // ------------------------------------------------ D: dP = g alpha^T dO ------------------------------------------------
    for (int r = 0; r < pl.n_rounds; ++r) {
      const int h0 = r * pl.hpr;
      const int nh = min(pl.hpr, H - h0);
      const int hl = warp >> 1, m = warp & 1;
      const bool active = (warp < 2 * nh) && (16 * m < N);
      const int h = h0 + hl;
      const int j0 = 16 * m + g, j1 = j0 + 8;
      const bool bias_owner = active && m == 0 && (p.concat || (r == 0 && hl == 0));
      uint32_t ah[4][4], al[4][4];
      if (active) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const int ia = 8 * ks + t, ib = ia + 4;
          const float a0 = j0 < N ? tile[(h * N + j0) * NS + ia] * g_scale : 0.f;
          const float a1 = j1 < N ? tile[(h * N + j1) * NS + ia] * g_scale : 0.f;
          const float a2 = j0 < N ? tile[(h * N + j0) * NS + ib] * g_scale : 0.f;
          const float a3 = j1 < N ? tile[(h * N + j1) * NS + ib] * g_scale : 0.f;
          split_tf32_trunc(a0, ah[ks][0], al[ks][0]);
          split_tf32_trunc(a1, ah[ks][1], al[ks][1]);
          split_tf32_trunc(a2, ah[ks][2], al[ks][2]);
          split_tf32_trunc(a3, ah[ks][3], al[ks][3]);
        }
      }
      const int n_grp = (n_cb + pl.cbs_per_grp_d - 1) / pl.cbs_per_grp_d;
      for (int gi = 0; gi < n_grp; ++gi, ++gk) {
        const int s = gk & 1;
        mbar_wait(&grp_full[s], (gk >> 1) & 1);
        if (active) {
          const unsigned char* gb = grp0 + s * (kGrpTiles * kTile);
          const int cb0 = gi * pl.cbs_per_grp_d;
          const int ncb = min(pl.cbs_per_grp_d, n_cb - cb0);
          for (int k = 0; k < ncb; ++k) {
            const int cb = cb0 + k;
            const unsigned char* Bt = gb + (p.concat ? k * nh + hl : k) * kTile;
            float cmain[4][4], ccorr[4][4];
            float bsum[4];
#pragma unroll
            for (int n = 0; n < 4; ++n) {
              bsum[n] = 0.f;
#pragma unroll
              for (int q = 0; q < 4; ++q) cmain[n][q] = ccorr[n][q] = 0.f;
            }
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const int r0 = 8 * ks + t, r1 = r0 + 4;
#pragma unroll
              for (int n = 0; n < 4; ++n) {
                const int c = 8 * n + g;
                const float b0 = ld_tile(Bt, r0, c), b1 = ld_tile(Bt, r1, c);
                if (bias_owner) bsum[n] += (r0 < N ? b0 : 0.f) + (r1 < N ? b1 : 0.f);
                uint32_t bh[2], bl[2];
                split_tf32_trunc(b0, bh[0], bl[0]);
                split_tf32_trunc(b1, bh[1], bl[1]);
                mma_tf32_16x8x8(ccorr[n], al[ks], bh);
                mma_tf32_16x8x8(cmain[n], ah[ks], bh);
                mma_tf32_16x8x8(ccorr[n], ah[ks], bl);
              }
            }
            if (bias_owner) {
#pragma unroll
              for (int n = 0; n < 4; ++n) {
                float v = bsum[n];
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                const int c = cb * 32 + 8 * n + g;
                if (t == 0 && c < C) dbias_s[(p.concat ? h * C : 0) + c] += v;     // one owner lane per column
              }
            }
#pragma unroll
            for (int n = 0; n < 4; ++n)
#pragma unroll
              for (int hf = 0; hf < 2; ++hf) {
                const int j = hf ? j1 : j0;
                const int c = cb * 32 + 8 * n + 2 * t;
                if (j < N && c < C) {
                  const float v0 = cmain[n][2 * hf] + ccorr[n][2 * hf], v1 = cmain[n][2 * hf + 1] + ccorr[n][2 * hf + 1];
                  const bool has1 = c + 1 < C;
                  if (args.dP_hi16) {
                    const size_t off = ((size_t)b * N + j) * args.ldp16 + (size_t)h * C + c;
                    const float w0 = v0 * dp_scale, w1 = v1 * dp_scale;
                    const __half h0_ = __float2half_rn(w0), h1_ = __float2half_rn(w1);
                    const __half l0_ = __float2half_rn(w0 - __half2float(h0_)), l1_ = __float2half_rn(w1 - __half2float(h1_));
                    if (p.vec2_ok) {
                      *reinterpret_cast<__half2*>(args.dP_hi16 + off) = __halves2half2(h0_, h1_);
                      *reinterpret_cast<__half2*>(args.dP_lo16 + off) = __halves2half2(l0_, l1_);
                    } else {
                      args.dP_hi16[off] = h0_; args.dP_lo16[off] = l0_;
                      if (has1) { args.dP_hi16[off + 1] = h1_; args.dP_lo16[off + 1] = l1_; }
                    }
                  } else {
                    const size_t off = ((size_t)b * N + j) * p.ldp + (size_t)h * C + c;
                    if (p.vec2_ok) {
                      *reinterpret_cast<float2*>(args.dP_aug + off) = make_float2(v0, v1);
                    } else {
                      args.dP_aug[off] = v0;
                      if (has1) args.dP_aug[off + 1] = v1;
                    }
                  }
                }
              }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive2(&grp_empty[s]);
      }
    }
    bar_sync_compute();                                   // every warp is done with alpha; D is complete
    lap(4);
    // ------------------------------------------------ V: dv += dz'^T . edge rows ------------------------------------------------
    for (int idx = tid; idx < tile_floats; idx += kCT) tile[idx] = 0.f;      // next graph's logits accumulate into zeros
    for (int c = 0; c < nchunks; ++c, ++ek) {
      const int s = ek & 1;
      mbar_wait(&edge_full[s], (ek >> 1) & 1);
      const int rows = rows_in(c);
      const float* Ts = s ? stage1 : stage0;
      const int row_base = c * pl.chunk_rows;
#pragma unroll
      for (int uu = 0; uu < kMaxDvUnits; ++uu) {
        const int u = warp + uu * kW;
        if (u >= pl.dv_units) continue;                   // warp-uniform
        const int mt = u % pl.n_mtiles, rg = u / pl.n_mtiles;
        const int rbeg = rg * pl.dv_rpu, rend = min(rows, rbeg + pl.dv_rpu);
        if (rbeg >= rend) continue;
        const int f0 = mt * 16 + g, f1 = f0 + 8;
        float acc[3][4];
#pragma unroll
        for (int pr = 0; pr < 3; ++pr)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[pr][q] = 0.f;
        for (int r0 = rbeg; r0 < rend; r0 += 8) {
          // B fragment (k = edge row, n = head): b0 = dz'[r0+t][g], b1 = dz'[r0+t+4][g]
          float bv[2];
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int r = r0 + t + 4 * half;
            float val = 0.f;
            if (r < rend && g < H) {
              const int code = table_s[row_base + r];
              if (code >= 0) val = D[(g * N + (code & 0xffff)) * NS + (code >> 16)];
            }
            bv[half] = val;
          }
          uint32_t bh[2], bl[2];
          split_tf32_trunc(bv[0], bh[0], bl[0]);
          split_tf32_trunc(bv[1], bh[1], bl[1]);
          const bool k0_ok = r0 + t < rend, k1_ok = r0 + t + 4 < rend;
          const float* t0p = Ts + (size_t)(r0 + t) * Fe;
          const float* t1p = t0p + (size_t)4 * Fe;
          float a[4];
          a[0] = (k0_ok && f0 < Fe) ? lds_f32(t0p + f0) : 0.f;
          a[1] = (k0_ok && f1 < Fe) ? lds_f32(t0p + f1) : 0.f;
          a[2] = (k1_ok && f0 < Fe) ? lds_f32(t1p + f0) : 0.f;
          a[3] = (k1_ok && f1 < Fe) ? lds_f32(t1p + f1) : 0.f;
          uint32_t ah[4], al[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) split_tf32_trunc(a[q], ah[q], al[q]);
          mma_tf32_16x8x8(acc[0], al, bh);
          mma_tf32_16x8x8(acc[1], ah, bl);
          mma_tf32_16x8x8(acc[2], ah, bh);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) dv_run[uu][q] += (acc[0][q] + acc[1][q]) + acc[2][q];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive2(&edge_empty[s]);
    }
    bar_sync_compute();                                   // tile zeroed, D free for the next graph
    lap(5);
  }
  if (tid == 0)
    for (int k = 0; k < 6; ++k) atomicAdd(&g_bwd2_counters[k], (unsigned long long)ph[k]);

  // ---- per-CTA partials: dv_part[(cta * dv_rg + rg)][h][f], dbias_part[cta][col] ----
  if (Fe > 0) {
#pragma unroll
    for (int uu = 0; uu < kMaxDvUnits; ++uu) {
      const int u = warp + uu * kW;
      if (u < pl.dv_units) {
        const int mt = u % pl.n_mtiles, rg = u / pl.n_mtiles;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int f = mt * 16 + g + ((q & 2) ? 8 : 0), h = 2 * t + (q & 1);
          if (f < Fe && h < H)
            args.dv_part[((size_t)blockIdx.x * pl.dv_rg + rg) * H * Fe + (size_t)h * Fe + f] = dv_run[uu][q];
        }
      }
    }
  }
  for (int idx = tid; idx < p.ldo; idx += kCT) args.dbias_part[(size_t)blockIdx.x * p.ldo + idx] = dbias_s[idx];
}

}  // namespace

bool attn_bwd2_fits(const AttnParams& p) {
  if (p.N > 32 || p.H > kMaxHeads || p.Fe > kMaxFe) return false;
  const Bwd2Plan pl = make_plan(p);
  return pl.total <= 227 * 1024;
}

size_t attn_bwd2_partials_bytes(const spotv2_gat_desc* d) {
  const size_t ctas = (size_t)sm_count();
  const size_t ldo = d->concat ? (size_t)d->H * d->C : (size_t)d->C;
  return round_up(ctas * (3 * (size_t)d->H * d->Fe + ldo) * sizeof(float), 256);
}

int launch_attn_bwd2(AttnBwdArgs& a, float* dv, float* dbias, void* ws, size_t ws_bytes, cudaStream_t st) {
  const AttnParams& p = a.p;
  const Bwd2Plan pl = make_plan(p);
  if (pl.total > 227 * 1024) return fail(SPOTV2_ERR_UNSUPPORTED, "attn_bwd2: shared-memory plan does not fit");
  int grid = sm_count();
  if (grid > p.B) grid = p.B;
  const int rg = p.Fe > 0 ? pl.dv_rg : 0;
  const size_t need = ((size_t)grid * ((size_t)rg * p.H * p.Fe + p.ldo)) * sizeof(float);
  if (!ws || ws_bytes < need) return fail(SPOTV2_ERR_WORKSPACE, "attn_bwd needs %zu B of workspace, got %zu", need, ws_bytes);
  a.dv_part = static_cast<float*>(ws);
  a.dbias_part = a.dv_part + (size_t)grid * rg * p.H * p.Fe;
  CUtensorMap tmP, tmG;
  memset(&tmP, 0, sizeof(tmP));
  memset(&tmG, 0, sizeof(tmG));
  if (pl.tma_ok) {
    if (int rc = make_tmap(&tmP, p.P_aug, (uint64_t)p.B * p.N, (uint64_t)p.ldp, (uint64_t)p.ldp, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B))
      return rc;
    if (int rc = make_tmap(&tmG, a.dout, (uint64_t)p.B * p.N, (uint64_t)p.ldo, (uint64_t)p.ldo, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B))
      return rc;
  }
  auto kern = gat_attn_bwd2_kernel;
  SPOTV2_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.total));
  kern<<<grid, kB2Threads, pl.total, st>>>(a, pl, tmP, tmG);
  SPOTV2_CUDA_OK(cudaGetLastError());
  if (dv && p.Fe > 0) {
    if (int rc = reduce_partials(a.dv_part, grid * rg, p.H * p.Fe, dv, st)) return rc;
  }
  if (dbias)
    if (int rc = reduce_partials(a.dbias_part, grid, p.ldo, dbias, st)) return rc;
  return SPOTV2_OK;
}

int bwd2_diag_add(unsigned long long* host_out, int reset) {
  unsigned long long tmp[kNumCounters];
  SPOTV2_CUDA_OK(cudaMemcpyFromSymbol(tmp, g_bwd2_counters, sizeof(tmp)));
  for (int k = 0; k < kNumCounters; ++k) host_out[k] += tmp[k];
  if (reset) {
    unsigned long long zeros[kNumCounters] = {0};
    SPOTV2_CUDA_OK(cudaMemcpyToSymbol(g_bwd2_counters, zeros, sizeof(zeros)));
  }
  return SPOTV2_OK;
}

}  // namespace spotv2
