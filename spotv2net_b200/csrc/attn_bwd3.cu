// Fused GAT attention backward on the 5th-generation tensor cores (tcgen05 + TMEM + TMA): the default for head-mean
// layers on small graphs (attn_bwd2.cu / attn_bwd.cu cover the rest).  Same mathematics as the other two kernels
// (attention recomputed from P_aug and the forward's edge terms, nothing of size E x H stored).
//
// Why a third kernel: the mma.sync version (attn_bwd2.cu) sat at 40 % of the HBM roofline, bound by instruction issue -
// every warp converted its own operand fragments and turned its own accumulator fragments into output pairs.  Here no
// thread touches an MMA operand fragment: operands are TMA tiles in shared memory, one thread issues tcgen05.mma, the
// accumulators live in TMEM and the per-element work that remains (the "lo" halves of the split operands, the dP
// epilogue, dv) is spread once over the CTA.
//
// fp32 accuracy on tf32 tensor cores: every operand is hi + lo with hi = RN_tf32(x) (written back in place over the TMA
// tile) and lo = RN_tf32(x - hi) (a second buffer of the same layout), and a product is lo*hi + hi*lo + hi*hi.
//
// One persistent CTA per SM, 22 warps, five roles connected by mbarriers; every role walks the graphs of this CTA in
// the same order:
//   producer (1 warp)   ONE in-order stream of shared-memory slots, per graph:
//                         T  the forward's edge-term tile                         (1-D bulk copies)
//                         A  per 16 channels: P tiles of all heads + the dout tile (K-major, 64-byte swizzled rows)
//                         G  per 128 channels: dout tiles in MN-major form         (128B/32B-atom swizzle)
//                         V  edge-row chunks                                       (1-D bulk copies; not in edge_mode 1)
//   stream group (8 w)  A/G slots: the lo pass (hi rounded in place, lo written to one of two lo buffers) -> MMA warp;
//                       V slots: dv += dz'^T . edge rows on the CUDA cores (exact fp32, FFMA2), two-level accumulation
//   MMA warp (1 thread) phase A  dalpha[(h,j), i] = sum_c P[j,h,c] dout[i,c]      M = (head, source) rows, N = targets
//                       phase D  dP^T[c, (h,j)]   = sum_i dout[i,c] alpha_h[i,j]  M = channels, N = (head, source)
//                       (one spare source row of head 0 holds ones, so the same product yields the bias gradient)
//   softmax group (8 w) thread = (head, target): self-loop mean fill, LeakyReLU, softmax -> alpha as a tf32 hi/lo pair
//                       in UMMA layout; dalpha from TMEM -> shared; softmax / LeakyReLU backward; ds, dd, dz'
//   epilogue group (4w) dP^T from TMEM: scale, fp16 hi/lo pair (or fp32), 64-byte row pieces to global; dbias
#include <string.h>

#include "attn_bwd.cuh"
#include "tc.cuh"
#include "tma.cuh"

namespace spotv2 {

namespace {

constexpr int kSmWarps = 8, kDeWarps = 4, kStWarps = 8;
constexpr int kSmT = kSmWarps * 32, kDeT = kDeWarps * 32, kStT = kStWarps * 32;
constexpr int kWarpDe0 = kSmWarps, kWarpSt0 = kSmWarps + kDeWarps, kWarpProd = kWarpSt0 + kStWarps, kWarpMma = kWarpProd + 1;
constexpr int kB3Threads = (kWarpMma + 1) * 32;      // 704
constexpr int kNS3 = kEdgeTermNS;                    // work-tile row stride (the forward's edge-term layout)
constexpr int kKB = 16;                              // channels per A slot (64-byte rows)
constexpr int kPTile = 32 * kKB * 4;                 // one (head, k-block) tile: 32 source rows x 64 B
constexpr int kCB = 128;                             // channels per G slot / phase-D block
constexpr int kGTile = 32 * 32 * 4;                  // one MN-major dout box: 32 target rows x 32 channels
constexpr int kMaxSlots3 = 8;
constexpr int kMaxChunkRows = 64;
constexpr int kRowGroups = 4;                        // phase V: row groups per chunk (x 2 feature halves = 8 warps)

// barrier block layout (uint64 each)
enum { kBarFull = 0, kBarEmpty = kMaxSlots3, kBarLoFull = 2 * kMaxSlots3, kBarLoEmpty = kBarLoFull + 2, kBarDAFull = kBarLoEmpty + 2,
       kBarDAEmpty, kBarAlphaFull, kBarAlphaEmpty, kBarDTFull, kBarDTEmpty = kBarDTFull + 2, kBarDzFull = kBarDTEmpty + 2, kBarDzEmpty,
       kNumBars };

struct Bwd3Plan {
  int n_kb, n_cb, chunk_rows, nchunks, t_pieces, n_slots, slots_per_graph, n_blk, d_cols, n_dbuf;
  uint32_t slot_bytes, lo_bytes, a_bytes, tile_bytes, alpha_bytes;
  uint32_t off_bar, off_table, off_sd, off_dzr, off_dbias, off_work, off_ahi, off_alo, off_lo, off_slots, total;
};

Bwd3Plan make_plan3(const AttnParams& p) {
  Bwd3Plan s{};
  const int N = p.N, H = p.H, C = p.C, Fe = p.Fe;
  s.n_kb = (C + kKB - 1) / kKB;
  s.n_cb = (C + kCB - 1) / kCB;
  s.n_blk = (32 * H + 127) / 128;
  s.d_cols = 32 * H;
  s.n_dbuf = (128 + 2 * s.d_cols <= 512) ? 2 : 1;
  s.a_bytes = (uint32_t)(H + 1) * kPTile;
  s.lo_bytes = (uint32_t)round_up(s.a_bytes + kPTile > 4u * kGTile ? s.a_bytes + kPTile : 4u * kGTile, 1024);
  // block 1 of phase A reads 128 rows from row 128 on whatever H is: keep that inside the buffers
  if (s.lo_bytes < (uint32_t)s.n_blk * 128 * 64) s.lo_bytes = (uint32_t)s.n_blk * 128 * 64;
  s.slot_bytes = s.lo_bytes;
  s.tile_bytes = (uint32_t)(H * N * kNS3 * 4);
  s.alpha_bytes = (uint32_t)(32 * H * 128);
  s.t_pieces = (int)((s.tile_bytes + s.slot_bytes - 1) / s.slot_bytes);
  if (!p.terms_in) {
    int rows = (int)(s.slot_bytes / ((uint32_t)Fe * 4u)) / 8 * 8;
    if (rows > kMaxChunkRows) rows = kMaxChunkRows;
    if (rows < 8) { s.total = 0xffffffffu; return s; }
    s.chunk_rows = rows;
    s.nchunks = (p.R + rows - 1) / rows;
  }
  s.slots_per_graph = s.t_pieces + s.n_kb + s.n_cb + s.nchunks;
  uint32_t o = 0;
  s.off_bar = o;    o += 512;
  s.off_table = o;  o += (uint32_t)round_up((size_t)(p.R > 0 ? p.R : 1) * 4, 16);
  s.off_sd = o;     o += (uint32_t)round_up((size_t)N * 2 * H * 4, 16);
  s.off_dzr = o;    o += (uint32_t)(kMaxChunkRows * 8 * 8);
  s.off_dbias = o;  o += (uint32_t)round_up((size_t)s.n_cb * kCB * 4, 16);
  s.off_work = o;   o += (uint32_t)round_up(s.tile_bytes, 16);
  o = (uint32_t)round_up(o, 1024);
  s.off_ahi = o;    o += s.alpha_bytes;
  s.off_alo = o;    o += s.alpha_bytes;
  s.off_lo = o;     o += 2 * s.lo_bytes;
  s.off_slots = o;
  const uint32_t cap = 227 * 1024;
  const uint32_t avail = cap > o ? cap - o : 0;
  s.n_slots = (int)(avail / s.slot_bytes);
  if (s.n_slots > kMaxSlots3) s.n_slots = kMaxSlots3;
  if (s.n_slots < 3) { s.total = 0xffffffffu; return s; }
  s.total = o + (uint32_t)s.n_slots * s.slot_bytes;
  return s;
}

// ---- small PTX helpers -----------------------------------------------------------------------------------------------
__device__ __forceinline__ void bar_group(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ float4 lds128a(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts128a(uint32_t a, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float2 lds64a(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ float ldsa(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void stsa(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
// round-to-nearest (ties away) onto the tf32 grid: what the tensor core would see of x, made explicit
__device__ __forceinline__ float rn_tf32(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ void split4(const float4 x, float4& hi, float4& lo) {
  hi = make_float4(rn_tf32(x.x), rn_tf32(x.y), rn_tf32(x.z), rn_tf32(x.w));
  lo = make_float4(rn_tf32(x.x - hi.x), rn_tf32(x.y - hi.y), rn_tf32(x.z - hi.z), rn_tf32(x.w - hi.w));
}
__device__ __forceinline__ void tma_load_3d_hint(uint32_t smem_dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar,
                                                 uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], "
      "[%2], %6;" ::"r"(smem_dst),
      "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "l"(hint)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// mbarrier by 32-bit shared address
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (int spin = 0; spin < 64 && !ok; ++spin)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  if (ok) return;
  const long long t0 = clock64();
  while (true) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return;
    __nanosleep(32);
    if (clock64() - t0 > 4000000000LL) __trap();        // a lost arrival fails loudly instead of hanging the box
  }
}
__device__ __forceinline__ void bar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void bar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_commit(uint32_t bar) {      // arrives when every tcgen05.mma issued so far has retired
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_g2s_a(uint32_t smem_dst, const void* gmem_src, uint32_t bytes, uint32_t bar, uint64_t hint) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_dst),
      "l"(gmem_src), "r"(bytes), "r"(bar), "l"(hint)
      : "memory");
}

// a cursor over the slot ring: every role steps through the producer's slot order
struct SlotCursor {
  int slot;
  uint32_t ph;
  int n;
  __device__ __forceinline__ void advance(int k = 1) {
    slot += k;
    while (slot >= n) { slot -= n; ph ^= 1u; }
  }
};

// DROP: attention dropout in training mode (mask regenerated from the descriptor's Philox key); a separate instantiation.
template <bool DROP>
__global__ void __launch_bounds__(kB3Threads, 1)
gat_attn_bwd3_kernel(const AttnBwdArgs args, const Bwd3Plan pl, const __grid_constant__ CUtensorMap tmP,
                     const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmGt) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const AttnParams& p = args.p;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = p.N, H = p.H, C = p.C, Fe = p.Fe, HC = H * C;
  const uint32_t sbase = smem_u32(smem_raw);
  const uint32_t a_bar = sbase + pl.off_bar, a_table = sbase + pl.off_table, a_sd = sbase + pl.off_sd, a_dzr = sbase + pl.off_dzr;
  const uint32_t a_dbias = sbase + pl.off_dbias, a_work = sbase + pl.off_work, a_ahi = sbase + pl.off_ahi, a_alo = sbase + pl.off_alo;
  const uint32_t a_lo = sbase + pl.off_lo, a_slots = sbase + pl.off_slots;
  auto BAR = [&](int k) { return a_bar + (uint32_t)k * 8u; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + pl.off_bar + kNumBars * 8);
  const int my_graphs = (p.B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const bool has_v = !p.terms_in;
  const int tile_floats = H * N * kNS3;

  // ---------------------------------------------------------------- setup
  if (tid == 0) {
    for (int k = 0; k < kNumBars; ++k) mbar_init(reinterpret_cast<uint64_t*>(smem_raw + pl.off_bar) + k, 1);
    fence_mbar_init();
  }
  // zero everything the tensor core or a padded loop may read before it is written
  for (uint32_t o = pl.off_table + (uint32_t)tid * 16u; o < pl.total; o += kB3Threads * 16u)
    *reinterpret_cast<float4*>(smem_raw + o) = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  // row -> float offset of (source j, target i) inside one head of the work tile (-1: row skipped)
  for (int r = tid; r < p.R && has_v; r += kB3Threads) {
    const int code = p.table[r];
    reinterpret_cast<int32_t*>(smem_raw + pl.off_table)[r] = code >= 0 ? (code & 0xffff) * kNS3 + (code >> 16) : -1;
  }
  // the ones row (head 0, spare source row 31): phase D then also produces sum_i dout[i, c] = the bias gradient
  if (tid < N) {
    const int r = 31, i = tid;
    *reinterpret_cast<float*>(smem_raw + pl.off_ahi + r * 128 + ((((i >> 2) ^ (r & 7)) << 4) | ((i & 3) << 2))) = 1.f;
  }
  fence_proxy_async();
  if (warp == kWarpMma) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tA = tmem_base, tD = tmem_base + 128u;

  if (warp == kWarpProd) {
    // =========================================== producer ===========================================
    if (lane == 0) {
      prefetch_tmap(&tmP); prefetch_tmap(&tmG); prefetch_tmap(&tmGt);
      SlotCursor cur{0, 0u, pl.n_slots};
      auto acquire = [&]() -> uint32_t {
        bar_wait(BAR(kBarEmpty + cur.slot), cur.ph ^ 1u);
        return a_slots + (uint32_t)cur.slot * pl.slot_bytes;
      };
      for (int it = 0; it < my_graphs; ++it) {
        const int b = blockIdx.x + it * gridDim.x;
        for (int pc = 0; pc < pl.t_pieces; ++pc) {                               // T: the forward's edge terms
          const uint32_t dst = acquire(), full = BAR(kBarFull + cur.slot);
          const uint32_t off = (uint32_t)pc * pl.slot_bytes;
          const uint32_t bytes = min(pl.slot_bytes, pl.tile_bytes - off);
          bar_expect_tx(full, bytes);
          bulk_g2s_a(dst, reinterpret_cast<const unsigned char*>(p.edge_terms + (size_t)b * tile_floats) + off, bytes, full, kEvictFirst);
          cur.advance();
        }
        for (int kb = 0; kb < pl.n_kb; ++kb) {                                   // A: P tiles of all heads + dout, 16 channels
          const uint32_t dst = acquire(), full = BAR(kBarFull + cur.slot);
          bar_expect_tx(full, pl.a_bytes);
          for (int h = 0; h < H; ++h) tma_load_3d_hint(dst + (uint32_t)h * kPTile, &tmP, h * C + kb * kKB, 0, b, full, kEvictFirst);
          tma_load_3d_hint(dst + (uint32_t)H * kPTile, &tmG, kb * kKB, 0, b, full, kEvictNormal);
          cur.advance();
        }
        for (int cb = 0; cb < pl.n_cb; ++cb) {                                   // G: dout again, MN-major boxes (L2 hits)
          const uint32_t dst = acquire(), full = BAR(kBarFull + cur.slot);
          bar_expect_tx(full, 4u * kGTile);
          for (int k = 0; k < 4; ++k) tma_load_3d_hint(dst + (uint32_t)k * kGTile, &tmGt, cb * kCB + k * 32, 0, b, full, kEvictFirst);
          cur.advance();
        }
        for (int c = 0; c < pl.nchunks; ++c) {                                   // V: edge rows, once
          const uint32_t dst = acquire(), full = BAR(kBarFull + cur.slot);
          int rows = p.R - c * pl.chunk_rows;
          if (rows > pl.chunk_rows) rows = pl.chunk_rows;
          const uint32_t bytes = (uint32_t)rows * (uint32_t)Fe * 4u;
          bar_expect_tx(full, bytes);
          bulk_g2s_a(dst, p.edge_rows + ((size_t)b * p.R + (size_t)c * pl.chunk_rows) * Fe, bytes, full, kEvictFirst);
          cur.advance();
        }
      }
    }
  } else if (warp == kWarpMma) {
    // =========================================== MMA issuer ===========================================
    if (lane == 0) {
      // instruction descriptors: D fp32, A/B tf32; bit 15 = A is MN-major; N >> 3 at bit 17, M >> 4 at bit 24
      const uint32_t id_base = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t id_a64 = id_base | ((uint32_t)(64 >> 3) << 17), id_a32 = id_base | ((uint32_t)(32 >> 3) << 17);
      const uint32_t id_d = id_base | (1u << 15) | ((uint32_t)(pl.d_cols >> 3) << 17);
      SlotCursor cur{0, 0u, pl.n_slots};
      uint32_t lo_n = 0, d_n = 0;                       // lo-buffer uses and phase-D blocks so far
      for (int it = 0; it < my_graphs; ++it) {
        const uint32_t gph = (uint32_t)it & 1u;
        cur.advance(pl.t_pieces);
        // ---- phase A: dalpha.  TMEM columns per 128-row block: [0,32) small terms (hi*lo + lo*hi), [32,64) hi*hi
        bar_wait(BAR(kBarDAEmpty), gph ^ 1u);
        tc_fence_after();
        for (int kb = 0; kb < pl.n_kb; ++kb, ++lo_n) {
          const uint32_t li = lo_n & 1u;
          bar_wait(BAR(kBarLoFull + li), (lo_n >> 1) & 1u);
          tc_fence_after();
          const uint32_t sa = a_slots + (uint32_t)cur.slot * pl.slot_bytes, lo = a_lo + li * pl.lo_bytes;
#pragma unroll
          for (int ks = 0; ks < kKB / 8; ++ks) {
            const uint64_t b_cat = make_desc(lo + (uint32_t)H * kPTile + ks * 32, 16, 512, 4);    // dout lo rows | hi rows
            const uint64_t b_hi = make_desc(sa + (uint32_t)H * kPTile + ks * 32, 16, 512, 4);
            for (int blk = 0; blk < pl.n_blk; ++blk) {
              const uint64_t a_hi = make_desc(sa + (uint32_t)blk * 8192u + ks * 32, 16, 512, 4);
              const uint64_t a_lo_ = make_desc(lo + (uint32_t)blk * 8192u + ks * 32, 16, 512, 4);
              umma_tf32(tA + (uint32_t)blk * 64u, a_hi, b_cat, id_a64, (kb | ks) ? 1u : 0u);
              umma_tf32(tA + (uint32_t)blk * 64u, a_lo_, b_hi, id_a32, 1u);
            }
          }
          bar_commit(BAR(kBarEmpty + cur.slot));
          bar_commit(BAR(kBarLoEmpty + li));
          cur.advance();
        }
        bar_commit(BAR(kBarDAFull));
        // ---- phase D: dP^T per block of 128 channels
        bar_wait(BAR(kBarAlphaFull), gph);
        for (int cb = 0; cb < pl.n_cb; ++cb, ++lo_n, ++d_n) {
          const uint32_t li = lo_n & 1u;
          const uint32_t buf = pl.n_dbuf == 2 ? (d_n & 1u) : 0u;
          const uint32_t dph = pl.n_dbuf == 2 ? ((d_n >> 1) & 1u) : (d_n & 1u);
          bar_wait(BAR(kBarLoFull + li), (lo_n >> 1) & 1u);
          bar_wait(BAR(kBarDTEmpty + buf), dph ^ 1u);
          tc_fence_after();
          const uint32_t sa = a_slots + (uint32_t)cur.slot * pl.slot_bytes, lo = a_lo + li * pl.lo_bytes;
          const uint32_t td = tD + buf * (uint32_t)pl.d_cols;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t a_hi = make_desc(sa + ks * 1024, kGTile, 512, 1);       // MN-major: 8 target rows per k-step
            const uint64_t a_lo_ = make_desc(lo + ks * 1024, kGTile, 512, 1);
            const uint64_t b_hi = make_desc(a_ahi + ks * 32, 16, 1024, 2);
            const uint64_t b_lo = make_desc(a_alo + ks * 32, 16, 1024, 2);
            umma_tf32(td, a_lo_, b_hi, id_d, ks ? 1u : 0u);                        // small terms first
            umma_tf32(td, a_hi, b_lo, id_d, 1u);
            umma_tf32(td, a_hi, b_hi, id_d, 1u);
          }
          bar_commit(BAR(kBarEmpty + cur.slot));
          bar_commit(BAR(kBarLoEmpty + li));
          bar_commit(BAR(kBarDTFull + buf));
          cur.advance();
        }
        bar_commit(BAR(kBarAlphaEmpty));
        cur.advance(pl.nchunks);
      }
    }
  } else if (warp >= kWarpSt0) {
    // =========================================== stream group ===========================================
    const int st = tid - kWarpSt0 * 32, sw = st >> 5;
    SlotCursor cur{0, 0u, pl.n_slots};
    uint32_t lo_n = 0;
    const int fhalf = sw & 1, rq = sw >> 1;
    const int f0 = 64 * fhalf + 2 * lane;
    float2 run[kMaxHeads], acc[kMaxHeads];
#pragma unroll
    for (int h = 0; h < kMaxHeads; ++h) run[h] = acc[h] = make_float2(0.f, 0.f);
    auto lo_pass = [&](uint32_t src, uint32_t dst, int n16, int dup_from) {
      // hi = RN_tf32(x) back in place, lo = RN_tf32(x - hi) to the lo buffer; 16-byte pieces at or past dup_from (the
      // dout tile of an A slot) also copy hi one tile further into the lo buffer (B operand "lo rows | hi rows")
      for (int idx = st; idx < n16; idx += kStT) {
        const float4 x = lds128a(src + (uint32_t)idx * 16u);
        float4 hi, lo;
        split4(x, hi, lo);
        sts128a(src + (uint32_t)idx * 16u, hi);
        sts128a(dst + (uint32_t)idx * 16u, lo);
        if (idx >= dup_from) sts128a(dst + (uint32_t)idx * 16u + kPTile, hi);
      }
    };
    for (int it = 0; it < my_graphs; ++it) {
      const uint32_t gph = (uint32_t)it & 1u;
      cur.advance(pl.t_pieces);
      for (int kb = 0; kb < pl.n_kb + pl.n_cb; ++kb, ++lo_n) {                    // A slots, then G slots
        const uint32_t li = lo_n & 1u;
        bar_wait(BAR(kBarFull + cur.slot), cur.ph);
        bar_wait(BAR(kBarLoEmpty + li), ((lo_n >> 1) & 1u) ^ 1u);
        const uint32_t sa = a_slots + (uint32_t)cur.slot * pl.slot_bytes, lo = a_lo + li * pl.lo_bytes;
        if (kb < pl.n_kb) lo_pass(sa, lo, (int)(pl.a_bytes / 16u), H * (kPTile / 16));
        else lo_pass(sa, lo, 4 * kGTile / 16, 0x7fffffff);
        fence_proxy_async();
        bar_group(3, kStT);
        if (st == 0) bar_arrive(BAR(kBarLoFull + li));
        cur.advance();
      }
      if (has_v) {
        bar_wait(BAR(kBarDzFull), gph);                                           // dz' of this graph is in the work tile
#pragma unroll
        for (int h = 0; h < kMaxHeads; ++h) acc[h] = make_float2(0.f, 0.f);
        const int rpg = pl.chunk_rows / kRowGroups;
        for (int c = 0; c < pl.nchunks; ++c) {
          bar_wait(BAR(kBarFull + cur.slot), cur.ph);
          const uint32_t sa = a_slots + (uint32_t)cur.slot * pl.slot_bytes;
          int rows = p.R - c * pl.chunk_rows;
          if (rows > pl.chunk_rows) rows = pl.chunk_rows;
          // dz' of the chunk's rows, gathered through the row table into [row][head] pairs (v, v)
          for (int idx = st; idx < rows * H; idx += kStT) {
            const int r = idx / H, h = idx - r * H;
            int off;
            asm volatile("ld.shared.s32 %0, [%1];" : "=r"(off) : "r"(a_table + (uint32_t)(c * pl.chunk_rows + r) * 4u));
            const float v = off >= 0 ? ldsa(a_work + (uint32_t)(h * N * kNS3 + off) * 4u) : 0.f;
            asm volatile("st.shared.v2.f32 [%0], {%1,%1};" ::"r"(a_dzr + (uint32_t)(r * 8 + h) * 8u), "f"(v) : "memory");
          }
          bar_group(3, kStT);
          const int r0 = rq * rpg, r1 = min(rows, r0 + rpg);
          if (f0 < Fe) {
            for (int r = r0; r < r1; ++r) {
              const float2 e = lds64a(sa + (uint32_t)(r * Fe + f0) * 4u);
#pragma unroll
              for (int hp = 0; hp < kMaxHeads / 2; ++hp) {
                if (2 * hp < H) {
                  const float4 dz = lds128a(a_dzr + (uint32_t)(r * 8 + 2 * hp) * 8u);
                  acc[2 * hp] = ffma2(e, make_float2(dz.x, dz.y), acc[2 * hp]);
                  acc[2 * hp + 1] = ffma2(e, make_float2(dz.z, dz.w), acc[2 * hp + 1]);
                }
              }
            }
          }
          bar_group(3, kStT);
          if (st == 0) bar_arrive(BAR(kBarEmpty + cur.slot));
          cur.advance();
        }
#pragma unroll
        for (int h = 0; h < kMaxHeads; ++h) { run[h].x += acc[h].x; run[h].y += acc[h].y; }   // two-level sum: per graph, then total
        if (st == 0) bar_arrive(BAR(kBarDzEmpty));
      }
    }
    // per-CTA partials: dv_part[cta * kRowGroups + rq][h][f]
    if (has_v && f0 < Fe) {
      float* dst = args.dv_part + ((size_t)blockIdx.x * kRowGroups + rq) * H * Fe;
#pragma unroll
      for (int h = 0; h < kMaxHeads; ++h)
        if (h < H) {
          dst[(size_t)h * Fe + f0] = run[h].x;
          if (f0 + 1 < Fe) dst[(size_t)h * Fe + f0 + 1] = run[h].y;
        }
    }
  } else if (warp >= kWarpDe0) {
    // =========================================== epilogue group: dP ===========================================
    const int de = tid - kWarpDe0 * 32, q = de >> 5;
    float dp_scale = 1.f;
    if (args.dP_hi16) {
      dp_scale = dp_scale_from_amax(__uint_as_float(*reinterpret_cast<const unsigned*>(args.dout_blk)) * args.bound);
      if (blockIdx.x == 0 && de == 0) { args.dp_blk[2] = 1.f / dp_scale; args.dp_blk[4] = dp_scale; }
    }
    const float k_dp = dp_scale / (float)H * (DROP ? p.drop.scale : 1.f);
    uint32_t d_n = 0;
    for (int it = 0; it < my_graphs; ++it) {
      const int b = blockIdx.x + it * gridDim.x;
      for (int cb = 0; cb < pl.n_cb; ++cb, ++d_n) {
        const uint32_t buf = pl.n_dbuf == 2 ? (d_n & 1u) : 0u;
        const uint32_t dph = pl.n_dbuf == 2 ? ((d_n >> 1) & 1u) : (d_n & 1u);
        bar_wait(BAR(kBarDTFull + buf), dph);
        tc_fence_after();
        const int c = cb * kCB + q * 32 + lane;
        const bool cok = c < C;
        const uint32_t tbase = tD + buf * (uint32_t)pl.d_cols + ((uint32_t)(q * 32) << 16);
        for (int h = 0; h < H; ++h) {
          uint32_t r[32];
          tmem_ld32(tbase + (uint32_t)h * 32u, r);
          if (h == 0 && cok) stsa(a_dbias + (uint32_t)c * 4u, ldsa(a_dbias + (uint32_t)c * 4u) + __uint_as_float(r[31]));
          if (args.dP_hi16) {
            __half* ph = args.dP_hi16 + ((size_t)b * N) * args.ldp16 + (size_t)h * C + c;
            __half* pq = args.dP_lo16 + ((size_t)b * N) * args.ldp16 + (size_t)h * C + c;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (j < N && cok) {
                const float w = __uint_as_float(r[j]) * k_dp;
                const __half hh = __float2half_rn(w);
                ph[(size_t)j * args.ldp16] = hh;
                pq[(size_t)j * args.ldp16] = __float2half_rn(w - __half2float(hh));
              }
            }
          } else {
            float* pf = args.dP_aug + ((size_t)b * N) * p.ldp + (size_t)h * C + c;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < N && cok) pf[(size_t)j * p.ldp] = __uint_as_float(r[j]) * k_dp;
          }
        }
        tc_fence_before();
        bar_group(2, kDeT);
        if (de == 0) bar_arrive(BAR(kBarDTEmpty + buf));
      }
    }
    for (int cb = 0; cb < pl.n_cb; ++cb) {
      const int c = cb * kCB + q * 32 + lane;
      if (c < C) args.dbias_part[(size_t)blockIdx.x * p.ldo + c] = ldsa(a_dbias + (uint32_t)c * 4u);
    }
  } else {
    // =========================================== softmax group ===========================================
    const int h = warp, i = lane;                     // thread = (head, target) for the column passes, (head, source) for rows
    const bool on = h < H && i < N;
    const float g_scale = 1.f / (float)H;
    const float inv_nm1 = 1.f / (float)(N > 1 ? N - 1 : 1);
    SlotCursor cur{0, 0u, pl.n_slots};
    const uint32_t col = a_work + (uint32_t)(h * N * kNS3 + i) * 4u;             // + j * kNS3 * 4
    // element (row r = 32 h + j, k = i) of the alpha operand tiles (128-byte rows, 16-byte chunks XOR (r & 7))
    auto aoff = [&](int j, int k) -> uint32_t {
      const int r = 32 * h + j;
      return (uint32_t)(r * 128 + ((((k >> 2) ^ (r & 7)) << 4) | ((k & 3) << 2)));
    };
    for (int it = 0; it < my_graphs; ++it) {
      const int b = blockIdx.x + it * gridDim.x;
      const uint32_t gph = (uint32_t)it & 1u;
      // s | d columns of this graph's nodes (2H floats per node behind the H*C projection columns)
      float sd_reg[2] = {0.f, 0.f};
      for (int k = 0; k < 2; ++k) {
        const int idx = tid + k * kSmT;
        if (idx < N * 2 * H) sd_reg[k] = p.P_aug[((size_t)b * N + idx / (2 * H)) * p.ldp + HC + idx % (2 * H)];
      }
      if (has_v) bar_wait(BAR(kBarDzEmpty), gph ^ 1u);                           // phase V of the previous graph is done with the tile
      for (int pc = 0; pc < pl.t_pieces; ++pc) {
        bar_wait(BAR(kBarFull + cur.slot), cur.ph);
        const uint32_t sa = a_slots + (uint32_t)cur.slot * pl.slot_bytes, off = (uint32_t)pc * pl.slot_bytes;
        const int n16 = (int)(min(pl.slot_bytes, pl.tile_bytes - off) / 16u);
        for (int idx = tid; idx < n16; idx += kSmT) sts128a(a_work + off + (uint32_t)idx * 16u, lds128a(sa + (uint32_t)idx * 16u));
        bar_group(1, kSmT);
        if (tid == 0) bar_arrive(BAR(kBarEmpty + cur.slot));
        cur.advance();
      }
      cur.advance(pl.slots_per_graph - pl.t_pieces);
      for (int k = 0; k < 2; ++k) {
        const int idx = tid + k * kSmT;
        if (idx < N * 2 * H) stsa(a_sd + (uint32_t)idx * 4u, sd_reg[k]);
      }
      bar_group(1, kSmT);
      // ------------------------------------------------ softmax (thread = head h, target i) ------------------------------------------------
      bar_wait(BAR(kBarAlphaEmpty), gph ^ 1u);                                   // phase D of the previous graph has read alpha
      uint32_t mask = 0u, keep = 0xffffffffu;
      if (on) {
        float gsum = 0.f;
        for (int j = 0; j < N; ++j) gsum += (j != i) ? ldsa(col + (uint32_t)j * (kNS3 * 4)) : 0.f;
        const float gii = gsum / (float)(N > 1 ? N - 1 : 1);
        const float di = ldsa(a_sd + (uint32_t)(i * 2 * H + H + h) * 4u);
        float mx = -INFINITY;
        for (int j = 0; j < N; ++j) {
          const float z = (j == i ? gii : ldsa(col + (uint32_t)j * (kNS3 * 4))) + ldsa(a_sd + (uint32_t)(j * 2 * H + h) * 4u) + di;
          if (z > 0.f) mask |= 1u << j;
          const float l = z > 0.f ? z : z * p.slope;
          mx = fmaxf(mx, l);
          stsa(col + (uint32_t)j * (kNS3 * 4), l);
        }
        float sum = 0.f;
        for (int j = 0; j < N; ++j) {
          const float e = expf(ldsa(col + (uint32_t)j * (kNS3 * 4)) - mx);
          sum += e;
          stsa(col + (uint32_t)j * (kNS3 * 4), e);
        }
        const float inv = 1.f / (sum + 1e-16f);
        if (DROP) keep = dropout_keep_bits(p.drop, (((unsigned long long)b * H + h) * N + i) * N, N);
        for (int j = 0; j < N; ++j) {
          const float a = ldsa(col + (uint32_t)j * (kNS3 * 4)) * inv;
          stsa(col + (uint32_t)j * (kNS3 * 4), a);                              // un-dropped alpha stays in the work tile for now
          const float am = ((keep >> j) & 1u) ? a : 0.f;                         // the MMA operand carries the mask, not 1/(1-p)
          const float hi = rn_tf32(am);
          const uint32_t o = aoff(j, i);
          stsa(a_ahi + o, hi);
          stsa(a_alo + o, rn_tf32(am - hi));
        }
      }
      fence_proxy_async();
      bar_group(1, kSmT);
      if (tid == 0) bar_arrive(BAR(kBarAlphaFull));
      // alpha column of this thread: kept in registers across the TMEM dump (which overwrites the work tile)
      float al[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) al[j] = (on && j < N) ? ldsa(col + (uint32_t)j * (kNS3 * 4)) : 0.f;
      // ------------------------------------------------ dalpha: TMEM -> work tile (thread = head h, source j) ------------------------------------------------
      bar_wait(BAR(kBarDAFull), gph);
      tc_fence_after();
      bar_group(1, kSmT);                                                        // every alpha column is in registers
      if (h < H) {
        const uint32_t taddr = tA + (uint32_t)(h >> 2) * 64u + ((uint32_t)((h & 3) * 32) << 16);
        const uint32_t row = a_work + (uint32_t)((h * N + lane) * kNS3) * 4u;    // lane = source j
#pragma unroll
        for (int c0 = 0; c0 < 32; c0 += 16) {
          uint32_t sm_[16], mn_[16];
          tmem_ld16(taddr + (uint32_t)c0, sm_);
          tmem_ld16(taddr + 32u + (uint32_t)c0, mn_);
          if (lane < N) {
#pragma unroll
            for (int e = 0; e < 16; e += 4)
              sts128a(row + (uint32_t)(c0 + e) * 4u,
                      make_float4((__uint_as_float(mn_[e]) + __uint_as_float(sm_[e])) * g_scale,
                                  (__uint_as_float(mn_[e + 1]) + __uint_as_float(sm_[e + 1])) * g_scale,
                                  (__uint_as_float(mn_[e + 2]) + __uint_as_float(sm_[e + 2])) * g_scale,
                                  (__uint_as_float(mn_[e + 3]) + __uint_as_float(sm_[e + 3])) * g_scale));
          }
        }
      }
      tc_fence_before();
      bar_group(1, kSmT);
      if (tid == 0) bar_arrive(BAR(kBarDAEmpty));
      // ------------------------------------------------ softmax / LeakyReLU backward (thread = head h, target i) ------------------------------------------------
      float share = 0.f;
      if (on) {
        float dot = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < N) {
            float da = ldsa(col + (uint32_t)j * (kNS3 * 4));
            if (DROP) da = ((keep >> j) & 1u) ? da * p.drop.scale : 0.f;         // dalpha = m * d(alpha m)
            dot = fmaf(al[j], da, dot);
          }
        float dd = 0.f, dii = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < N) {
            float da = ldsa(col + (uint32_t)j * (kNS3 * 4));
            if (DROP) da = ((keep >> j) & 1u) ? da * p.drop.scale : 0.f;
            const float dl = al[j] * (da - dot);
            const float dz = ((mask >> j) & 1u) ? dl : dl * p.slope;
            stsa(col + (uint32_t)j * (kNS3 * 4), dz);
            dd += dz;
            if (j == i) dii = dz;
          }
        share = dii * inv_nm1;
        if (args.dsd) args.dsd[((size_t)b * N + i) * 2 * H + H + h] = dd;
        else args.dP_aug[((size_t)b * N + i) * p.ldp + HC + H + h] = dd;
      }
      bar_group(1, kSmT);
      if (on) {                                                                  // ds_j = row sum of dz (thread = head h, source j = lane)
        const uint32_t row = a_work + (uint32_t)((h * N + lane) * kNS3) * 4u;
        float ds = 0.f;
        for (int k = 0; k < N; ++k) ds += ldsa(row + (uint32_t)k * 4u);
        if (args.dsd) args.dsd[((size_t)b * N + lane) * 2 * H + h] = ds;
        else args.dP_aug[((size_t)b * N + lane) * p.ldp + HC + h] = ds;
      }
      bar_group(1, kSmT);
      if (on) {                                                                  // dz' = dz + dz_ii / (N - 1) off the diagonal, 0 on it
        for (int j = 0; j < N; ++j) {
          const uint32_t a = col + (uint32_t)j * (kNS3 * 4);
          stsa(a, j == i ? 0.f : ldsa(a) + share);
        }
      }
      bar_group(1, kSmT);
      if (p.dterms_out) {                                                        // edge_mode 1: d(edge terms) leaves in the tile layout
        float4* dst = reinterpret_cast<float4*>(p.dterms_out + (size_t)b * tile_floats);
        for (int idx = tid; idx < tile_floats / 4; idx += kSmT) dst[idx] = lds128a(a_work + (uint32_t)idx * 16u);
        bar_group(1, kSmT);
      }
      if (has_v && tid == 0) bar_arrive(BAR(kBarDzFull));
    }
  }
  // ---------------------------------------------------------------- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == kWarpMma) tmem_dealloc(tmem_base, 512);
}

int make_tmap3_f32(CUtensorMap* tm, const float* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_elems,
                   uint64_t stride2_elems, uint32_t b0, uint32_t b1, CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(SPOTV2_ERR_NO_DEVICE, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_elems * sizeof(float), stride2_elems * sizeof(float)};
  cuuint32_t box[3] = {b0, b1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SPOTV2_ERR_CUDA, "cuTensorMapEncodeTiled (attention backward, 3-D) failed with CUresult %d", (int)r);
  return SPOTV2_OK;
}

}  // namespace

// Head-mean layers on graphs of up to 31 nodes with the forward's edge terms at hand; everything else stays on the
// mma.sync kernels.
bool attn_bwd3_applies(const AttnParams& p) {
  if (p.N > 31 || p.N < 1 || p.H > kMaxHeads || p.concat || p.C % 4 != 0 || p.ldp % 4 != 0) return false;
  if (p.Fe <= 0 || !p.edge_terms) return false;
  if (!p.terms_in && (p.Fe > 128 || p.Fe % 2 != 0 || !p.bulk_ok || !p.edge_rows || !p.table)) return false;
  if (!tma_available()) return false;
  return make_plan3(p).total <= 227 * 1024;
}

size_t attn_bwd3_partials_bytes(const spotv2_gat_desc* d) {
  const size_t ctas = (size_t)sm_count();
  const size_t ldo = d->concat ? (size_t)d->H * d->C : (size_t)d->C;
  return round_up(ctas * (kRowGroups * (size_t)d->H * d->Fe + ldo) * sizeof(float), 256);
}

int launch_attn_bwd3(AttnBwdArgs& a, float* dv, float* dbias, void* ws, size_t ws_bytes, cudaStream_t st) {
  const AttnParams& p = a.p;
  const Bwd3Plan pl = make_plan3(p);
  if (pl.total > 227 * 1024) return fail(SPOTV2_ERR_UNSUPPORTED, "attn_bwd3: shared-memory plan does not fit");
  int grid = sm_count();
  if (grid > p.B) grid = p.B;
  const int rg = p.terms_in ? 0 : kRowGroups;
  const size_t need = ((size_t)grid * ((size_t)rg * p.H * p.Fe + p.ldo)) * sizeof(float);
  if (!ws || ws_bytes < need) return fail(SPOTV2_ERR_WORKSPACE, "attn_bwd needs %zu B of workspace, got %zu", need, ws_bytes);
  a.dv_part = static_cast<float*>(ws);
  a.dbias_part = a.dv_part + (size_t)grid * rg * p.H * p.Fe;
  CUtensorMap tmP, tmG, tmGt;
  // P_aug and dout as [graph][node][column] so that the two node rows a 32-row box reaches past a graph read as zeros
  // (columns past H*C + 2H are row padding that may hold anything: outside the map, they read as zeros too)
  if (int rc = make_tmap3_f32(&tmP, p.P_aug, (uint64_t)(p.H * p.C + 2 * p.H), (uint64_t)p.N, (uint64_t)p.B, (uint64_t)p.ldp, (uint64_t)p.N * p.ldp, kKB, 32,
                              CU_TENSOR_MAP_SWIZZLE_64B))
    return rc;
  if (int rc = make_tmap3_f32(&tmG, a.dout, (uint64_t)p.ldo, (uint64_t)p.N, (uint64_t)p.B, (uint64_t)p.ldo, (uint64_t)p.N * p.ldo, kKB, 32,
                              CU_TENSOR_MAP_SWIZZLE_64B))
    return rc;
  if (int rc = make_tmap3_f32(&tmGt, a.dout, (uint64_t)p.ldo, (uint64_t)p.N, (uint64_t)p.B, (uint64_t)p.ldo, (uint64_t)p.N * p.ldo, 32, 32,
                              CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
    return rc;
  auto kern = p.drop.p > 0.f ? gat_attn_bwd3_kernel<true> : gat_attn_bwd3_kernel<false>;
  SPOTV2_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.total));
  kern<<<grid, kB3Threads, pl.total, st>>>(a, pl, tmP, tmG, tmGt);
  SPOTV2_CUDA_OK(cudaGetLastError());
  if (dv && rg > 0)
    if (int rc = reduce_partials(a.dv_part, grid * rg, p.H * p.Fe, dv, st)) return rc;
  if (dbias)
    if (int rc = reduce_partials(a.dbias_part, grid, p.ldo, dbias, st)) return rc;
  return SPOTV2_OK;
}

}  // namespace spotv2
