// Fused GAT attention backward on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).  OPT-IN (attn_bwd_algo = 3):
// parity-green on every case the pipelined mma.sync kernel (attn_bwd2.cu) passes, but measured SLOWER on B200 at the
// default geometry (2.0 - 2.4 ms against 1.95 ms per 4096-graph batch), so attn_bwd2.cu stays the default.  Same
// mathematics as the other two kernels (attention recomputed from P_aug and the forward's edge
// terms, nothing of size E x H stored).
//
// What it does differently: no thread touches an MMA operand fragment.  Operands are TMA tiles in shared memory, one
// thread issues tcgen05.mma, accumulators live in TMEM, and the per-element work that remains (the "lo" halves of the
// split operands, the dP epilogue, dv) is spread once over the CTA instead of being repeated per warp.
//
// What bounds it (profiles/r2_bwd3_*): the shared-memory port.  With fp32 inputs in HBM, an fp32-accurate product on
// tf32 tensor cores needs each operand as hi + lo; per element of P that is 4 B written by TMA, 4 B read and 4 B
// written by the lo pass, and 8 B x 4/3 read by the MMAs (K = 8 per instruction, 128-row blocks of which a quarter is
// padding): ~23 B of shared-memory traffic per 4 B that arrive from HBM, i.e. ~2.8 MB per graph for phase A alone at
// 128 B/clk, on top of the edge-row chunks, the alpha tiles and the epilogue staging - at or above the time the same
// graph's bytes take to arrive from HBM.  The fix is upstream, not here: the projection GEMM has to emit P already as
// an operand pair (then phase A costs 8 B per element and no lo pass); see DESIGN.md section 9.
//
// fp32 accuracy on tf32 tensor cores: every operand is hi + lo; hi is the TMA tile as it is (the tensor core reads the
// top 19 bits of an fp32 pattern - measured, tools/bwd3_rawhi.py), lo = RN_tf32(x - trunc_tf32(x)) is written to a
// second buffer of the same layout, and a product is lo*hi + hi*lo + hi*hi.
//
// One persistent CTA per SM, 16 warps (128 registers per thread), five roles connected by mbarriers; every role walks
// the graphs of this CTA in the same order:
//   producer (1 warp)   ONE in-order stream of shared-memory slots, per graph:
//                         T  the forward's edge-term tile                         (1-D bulk copies)
//                         A  per 16 channels: P tiles of all heads + the dout tile (K-major, 64-byte swizzled rows)
//                         G  per 128 channels: dout tiles in MN-major form         (128B/32B-atom swizzle)
//                         V  edge-row chunks                                       (1-D bulk copies; not in edge_mode 1)
//   stream group (2 w)  A/G slots: the lo pass; warp w owns lo buffer w and every second slot
//   MMA warp (1 thread) phase A  dalpha[(h,j), i] = sum_c P[j,h,c] dout[i,c]      M = (head, source) rows, N = targets;
//                                B = [dout_lo ; dout_hi] so that one MMA yields hi*lo and hi*hi, a second adds lo*hi
//                       phase D  dP[(h,j), c] = sum_i alpha_h[i,j] dout[i,c]      M = (head, source) rows, N = 128 channels
//                       (one spare source row of head 0 holds ones, so the same product yields the bias gradient)
//   softmax group (8 w) thread = (head, target): self-loop mean fill, LeakyReLU, softmax -> alpha as a tf32 hi/lo pair
//                       in UMMA layout; dalpha from TMEM -> shared; softmax / LeakyReLU backward; ds, dd, dz';
//                       then V: dv += dz'^T . edge rows on the CUDA cores (exact fp32, FFMA2, two-level sums)
//   epilogue group (4w) dP from TMEM (thread = one (head, source) row, 32 channels per tcgen05.ld): scale, fp16 hi/lo
//                       pair by packed converts, a 4 KB staging tile per warp, 64-byte row pieces to global
#include <string.h>

#include "attn_bwd.cuh"
#include "tc.cuh"
#include "tma.cuh"

namespace spotv2 {

namespace {

constexpr int kSmWarps = 8, kDeWarps = 4, kStWarps = 2;
constexpr int kSmT = kSmWarps * 32, kDeT = kDeWarps * 32, kStT = kStWarps * 32;
constexpr int kWarpDe0 = kSmWarps, kWarpSt0 = kSmWarps + kDeWarps, kWarpProd = kWarpSt0 + kStWarps, kWarpMma = kWarpProd + 1;
constexpr int kB3Threads = (kWarpMma + 1) * 32;      // 512: 16 warps, 128 registers per thread
constexpr int kNS3 = kEdgeTermNS;                    // work-tile row stride (the forward's edge-term layout)
constexpr int kKB = 16;                              // channels per A slot (64-byte rows)
constexpr int kPTile = 32 * kKB * 4;                 // one (head, k-block) tile: 32 source rows x 64 B
constexpr int kCB = 128;                             // channels per G slot / phase-D block
constexpr int kGTile = 32 * 32 * 4;                  // one MN-major dout box: 32 target rows x 32 channels
constexpr int kMaxSlots3 = 8;
constexpr int kMaxChunkRows = 64;

// barrier block layout (uint64 each)
enum { kBarFull = 0, kBarEmpty = kMaxSlots3, kBarLoFull = 2 * kMaxSlots3, kBarLoEmpty = kBarLoFull + 2, kBarDAFull = kBarLoEmpty + 2,
       kBarDAEmpty, kBarAlphaFull, kBarAlphaEmpty, kBarDTFull, kBarDTEmpty, kBarVDone, kBarTDone, kNumBars };

struct Bwd3Plan {
  int n_kb, n_cb, chunk_rows, nchunks, t_pieces, n_slots, slots_per_graph, n_blk;
  uint32_t slot_bytes, lo_bytes, a_bytes, tile_bytes, alpha_bytes;
  uint32_t off_bar, off_table, off_sd, off_stage, off_dbias, off_work, off_ahi, off_alo, off_lo, off_slots, total;
};

Bwd3Plan make_plan3(const AttnParams& p) {
  Bwd3Plan s{};
  const int N = p.N, H = p.H, C = p.C, Fe = p.Fe;
  s.n_kb = (C + kKB - 1) / kKB;
  s.n_cb = (C + kCB - 1) / kCB;
  s.n_blk = (32 * H + 127) / 128;
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
  s.off_stage = o;  o += (uint32_t)(kDeWarps * 4096);          // per epilogue warp: [32 sources][64 B] fp16, hi | lo
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
// what the tensor core reads of an fp32 bit pattern handed to it as tf32: the top 19 bits
__device__ __forceinline__ float tr_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
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

// Diagnostics (spotv2_diag_counters, entries 16..31 when this kernel ran): cycles one sampling thread per role spent
// waiting on each class of barrier / inside each block of work, summed over CTAs.
//  0 producer: slot empty     1 MMA: lo buffer ready (phase A; incl. TMEM free)   2 epilogue: dP^T ready   3 MMA: alpha ready
//  4 MMA: phase D operands    5 stream: A/G slot full    6 stream: lo buffer free   7 stream: lo pass      8 stream: dz' ready
//  9 stream: V slot full     10 stream: dv arithmetic   11 softmax: T slot / tile / alpha free              12 epilogue: work
// 13 softmax: softmax        14 softmax: dalpha ready   15 softmax: TMEM dump + softmax backward
__device__ unsigned long long g_bwd3_counters[kNumCounters];
__device__ __forceinline__ void bar_wait_t(uint32_t bar, uint32_t parity, long long& acc) {
  const long long t0 = clock64();
  bar_wait(bar, parity);
  acc += clock64() - t0;
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
// HT: the head count as a compile-time constant (0 = run-time H; the default geometry's 6 gets its own instantiation).
template <bool DROP, int HT>
__global__ void __launch_bounds__(kB3Threads, 1)
gat_attn_bwd3_kernel(const AttnBwdArgs args, const Bwd3Plan pl, const __grid_constant__ CUtensorMap tmP,
                     const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmGt) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const AttnParams& p = args.p;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = p.N, H = HT ? HT : p.H, C = p.C, Fe = p.Fe, HC = H * C;
  const uint32_t sbase = smem_u32(smem_raw);
  const uint32_t a_bar = sbase + pl.off_bar, a_table = sbase + pl.off_table, a_sd = sbase + pl.off_sd;
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
  const uint32_t tA = tmem_base, tD = tmem_base + 128u;        // phase A: 64 columns per 128-row block; phase D: 128 per block

  // Ring discipline.  A wait on a slot's "full" barrier tells phases apart by parity alone, so a role may start such a
  // wait only when the previous use of that slot has been filled already (else the wait returns at once) and before the
  // next one can be (else it never returns).  Every role therefore waits only on slots it consumes itself, and each
  // jump over slots consumed by others is gated by an event that implies the fills in between:
  //   softmax group  T(g): after alpha_empty(g-1) (all MMAs of g-1 done, so all its A/G fills) and its own V(g-1);
  //                  V(g): after dalpha_full(g) (all A(g) fills); the G(g) fills it jumps are issued before V(g)'s
  //   stream warps   A/G(g): after v_done(g-1) and t_done(g); inside the run, lo_empty of the warp's previous slot
  //                  (a commit, so every earlier MMA and with it every earlier fill) comes BEFORE the full wait
  if (warp == kWarpProd) {
    // =========================================== producer ===========================================
    if (lane == 0) {
      prefetch_tmap(&tmP); prefetch_tmap(&tmG); prefetch_tmap(&tmGt);
      SlotCursor cur{0, 0u, pl.n_slots};
      long long w_empty = 0;
      auto acquire = [&]() -> uint32_t {
        bar_wait_t(BAR(kBarEmpty + cur.slot), cur.ph ^ 1u, w_empty);
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
      atomicAdd(&g_bwd3_counters[0], (unsigned long long)w_empty);
    }
  } else if (warp == kWarpMma) {
    // =========================================== MMA issuer ===========================================
    if (lane == 0) {
      // instruction descriptors: D fp32, A/B tf32; bit 16 = B is MN-major; N >> 3 at bit 17, M >> 4 at bit 24
      const uint32_t id_base = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t id_a64 = id_base | ((uint32_t)(64 >> 3) << 17), id_a32 = id_base | ((uint32_t)(32 >> 3) << 17);
      const uint32_t id_d = id_base | (1u << 16) | ((uint32_t)(kCB >> 3) << 17);
      SlotCursor cur{0, 0u, pl.n_slots};
      uint32_t lo_n = 0, d_n = 0;                       // lo-buffer uses and phase-D blocks so far
      long long w_a = 0, w_al = 0, w_d = 0;
      for (int it = 0; it < my_graphs; ++it) {
        const uint32_t gph = (uint32_t)it & 1u;
        cur.advance(pl.t_pieces);
        // ---- phase A: dalpha.  TMEM columns per 128-row block: [0,32) small terms (hi*lo + lo*hi), [32,64) hi*hi
        bar_wait_t(BAR(kBarDAEmpty), gph ^ 1u, w_a);
        tc_fence_after();
        for (int kb = 0; kb < pl.n_kb; ++kb, ++lo_n) {
          const uint32_t li = lo_n & 1u;
          bar_wait_t(BAR(kBarLoFull + li), (lo_n >> 1) & 1u, w_a);
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
        // ---- phase D: dP[(h,j), c] per block of 128 channels; A = alpha^T (K-major), B = dout (channel-major)
        bar_wait_t(BAR(kBarAlphaFull), gph, w_al);
        for (int cb = 0; cb < pl.n_cb; ++cb, ++lo_n, ++d_n) {
          const uint32_t li = lo_n & 1u;
          bar_wait_t(BAR(kBarLoFull + li), (lo_n >> 1) & 1u, w_d);
          bar_wait_t(BAR(kBarDTEmpty), (d_n & 1u) ^ 1u, w_d);
          tc_fence_after();
          const uint32_t sa = a_slots + (uint32_t)cur.slot * pl.slot_bytes, lo = a_lo + li * pl.lo_bytes;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t b_hi = make_desc(sa + ks * 1024, kGTile, 512, 1);       // 8 target rows per k-step
            const uint64_t b_lo = make_desc(lo + ks * 1024, kGTile, 512, 1);
            for (int blk = 0; blk < pl.n_blk; ++blk) {
              const uint64_t a_hi = make_desc(a_ahi + (uint32_t)blk * 16384u + ks * 32, 16, 1024, 2);
              const uint64_t a_lo_ = make_desc(a_alo + (uint32_t)blk * 16384u + ks * 32, 16, 1024, 2);
              const uint32_t td = tD + (uint32_t)blk * kCB;
              umma_tf32(td, a_lo_, b_hi, id_d, ks ? 1u : 0u);                      // small terms first
              umma_tf32(td, a_hi, b_lo, id_d, 1u);
              umma_tf32(td, a_hi, b_hi, id_d, 1u);
            }
          }
          bar_commit(BAR(kBarEmpty + cur.slot));
          bar_commit(BAR(kBarLoEmpty + li));
          bar_commit(BAR(kBarDTFull));
          cur.advance();
        }
        bar_commit(BAR(kBarAlphaEmpty));
        cur.advance(pl.nchunks);
      }
      atomicAdd(&g_bwd3_counters[1], (unsigned long long)w_a);
      atomicAdd(&g_bwd3_counters[3], (unsigned long long)w_al);
      atomicAdd(&g_bwd3_counters[4], (unsigned long long)w_d);
    }
  } else if (warp >= kWarpSt0) {
    // =========================================== stream group: the lo pass ===========================================
    // Warp w owns lo buffer w and takes every second A / G slot: two slots are in flight and nothing but the warp itself
    // has to be synchronised before the MMA warp is told.  hi stays where TMA put it (the tensor core reads the top 19
    // bits of an fp32 pattern: measured, tools/bwd3_rawhi.py); lo = RN_tf32(x - trunc_tf32(x)) goes to the lo buffer.
    // 16-byte pieces at or past dup_from (the dout tile of an A slot) also copy hi one tile further into the lo buffer
    // (B operand "lo rows | hi rows").  Four pieces per lane per round, loads first.
    const int sw = warp - kWarpSt0;
    SlotCursor cur{0, 0u, pl.n_slots};
    uint32_t lo_n = 0;
    long long w_full = 0, w_loe = 0, t_lo = 0;
    auto lo_pass = [&](uint32_t src, uint32_t dst, int n16, int dup_from) {
      for (int base = 0; base < n16; base += 4 * 32) {
        float4 x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int idx = base + u * 32 + lane;
          x[u] = idx < n16 ? lds128a(src + (uint32_t)idx * 16u) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int idx = base + u * 32 + lane;
          if (idx < n16) {
            const float4 lo = make_float4(rn_tf32(x[u].x - tr_tf32(x[u].x)), rn_tf32(x[u].y - tr_tf32(x[u].y)),
                                          rn_tf32(x[u].z - tr_tf32(x[u].z)), rn_tf32(x[u].w - tr_tf32(x[u].w)));
            sts128a(dst + (uint32_t)idx * 16u, lo);
            if (idx >= dup_from) sts128a(dst + (uint32_t)idx * 16u + kPTile, x[u]);
          }
        }
      }
    };
    for (int it = 0; it < my_graphs; ++it) {
      bar_group(3, kStT);                                                         // both warps know every fill either has seen
      if (has_v && it > 0) bar_wait_t(BAR(kBarVDone), (uint32_t)(it - 1) & 1u, w_full);
      bar_wait_t(BAR(kBarTDone), (uint32_t)it & 1u, w_full);
      cur.advance(pl.t_pieces);
      for (int kb = 0; kb < pl.n_kb + pl.n_cb; ++kb, ++lo_n) {                    // A slots, then G slots
        if ((int)(lo_n & 1u) == sw) {
          bar_wait_t(BAR(kBarLoEmpty + sw), ((lo_n >> 1) & 1u) ^ 1u, w_loe);
          bar_wait_t(BAR(kBarFull + cur.slot), cur.ph, w_full);
          const long long t0 = clock64();
          const uint32_t sa = a_slots + (uint32_t)cur.slot * pl.slot_bytes, lo = a_lo + (uint32_t)sw * pl.lo_bytes;
          if (kb < pl.n_kb) lo_pass(sa, lo, (int)(pl.a_bytes / 16u), H * (kPTile / 16));
          else lo_pass(sa, lo, 4 * kGTile / 16, 0x7fffffff);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) bar_arrive(BAR(kBarLoFull + sw));
          t_lo += clock64() - t0;
        }
        cur.advance();
      }
      cur.advance(pl.nchunks);
    }
    if (lane == 0) {
      atomicAdd(&g_bwd3_counters[5], (unsigned long long)w_full);
      atomicAdd(&g_bwd3_counters[6], (unsigned long long)w_loe);
      atomicAdd(&g_bwd3_counters[7], (unsigned long long)t_lo);
    }
  } else if (warp >= kWarpDe0) {
    // =========================================== epilogue group: dP ===========================================
    // Warp q owns TMEM lanes [32q, 32q + 32) = the source rows of head q (block 0) and head 4 + q (block 1): a thread
    // holds one (head, source) row and takes 32 consecutive channels per tcgen05.ld, so the fp16 hi/lo pair of those
    // channels leaves as 64 contiguous bytes per array (8-byte stores: a head's columns start on 8-byte boundaries).
    const int de = tid - kWarpDe0 * 32, q = de >> 5;
    float dp_scale = 1.f;
    if (args.dP_hi16) {
      dp_scale = dp_scale_from_amax(__uint_as_float(*reinterpret_cast<const unsigned*>(args.dout_blk)) * args.bound);
      if (blockIdx.x == 0 && de == 0) { args.dp_blk[2] = 1.f / dp_scale; args.dp_blk[4] = dp_scale; }
    }
    const float k_dp = dp_scale / (float)H * (DROP ? p.drop.scale : 1.f);
    const uint32_t stg = sbase + pl.off_stage + (uint32_t)q * 4096u;
    uint32_t d_n = 0;
    long long w_dt = 0, t_de = 0;
    for (int it = 0; it < my_graphs; ++it) {
      const int b = blockIdx.x + it * gridDim.x;
      for (int cb = 0; cb < pl.n_cb; ++cb, ++d_n) {
        bar_wait_t(BAR(kBarDTFull), d_n & 1u, w_dt);
        const long long t0 = clock64();
        tc_fence_after();
        for (int blk = 0; blk < pl.n_blk; ++blk) {
          const int h = blk * 4 + q;
          if (h >= H) break;                                                      // warp-uniform
          const uint32_t tbase = tD + (uint32_t)blk * kCB + ((uint32_t)(q * 32) << 16);
          const size_t row = ((size_t)b * N + lane);                              // lane = source j
#pragma unroll 1
          for (int c0 = 0; c0 < kCB; c0 += 32) {
            const int c = cb * kCB + c0;
            if (c >= C) break;
            uint32_t r[32];
            tmem_ld32(tbase + (uint32_t)c0, r);
            if (h == 0 && lane == 31) {                                           // the ones row: column sums of dout
#pragma unroll
              for (int k = 0; k < 32; ++k)
                if (c + k < C) stsa(a_dbias + (uint32_t)(c + k) * 4u, ldsa(a_dbias + (uint32_t)(c + k) * 4u) + __uint_as_float(r[k]));
            }
            if (args.dP_hi16) {
              // the row's 32 channels as fp16 hi | lo (packed converts), parked in the warp's staging tile ([32 sources][64 B]
              // per array, 16-byte chunks XOR (j >> 1) & 3: conflict-free both ways) and read back so that 8 lanes cover one
              // row: a store instruction writes four 64-byte row pieces instead of thirty-two 8-byte ones
              uint32_t hi[16], lo[16];
#pragma unroll
              for (int k = 0; k < 16; ++k) {
                const float w0 = __uint_as_float(r[2 * k]) * k_dp, w1 = __uint_as_float(r[2 * k + 1]) * k_dp;
                const __half2 hh = __floats2half2_rn(w0, w1);
                const float2 bk = __half22float2(hh);
                const __half2 ll = __floats2half2_rn(w0 - bk.x, w1 - bk.y);
                hi[k] = *reinterpret_cast<const uint32_t*>(&hh);
                lo[k] = *reinterpret_cast<const uint32_t*>(&ll);
              }
              const uint32_t sw_ = (uint32_t)((lane >> 1) & 3);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint32_t o = (uint32_t)lane * 64u + (((uint32_t)k ^ sw_) << 4);
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(stg + o), "r"(hi[4 * k]), "r"(hi[4 * k + 1]), "r"(hi[4 * k + 2]), "r"(hi[4 * k + 3]) : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(stg + 2048u + o), "r"(lo[4 * k]), "r"(lo[4 * k + 1]), "r"(lo[4 * k + 2]), "r"(lo[4 * k + 3]) : "memory");
              }
              __syncwarp();
              const int pc = (lane & 7) * 4;                                        // first of this lane's 4 channels in the piece
              if (c + pc < C) {                                                    // C % 4 == 0: four channels are in or out together
                uint2 vh[8], vl[8];
#pragma unroll
                for (int it8 = 0; it8 < 8; ++it8) {
                  const int j = 4 * it8 + (lane >> 3);
                  const uint32_t o = (uint32_t)j * 64u + ((((uint32_t)(lane & 7) >> 1) ^ (uint32_t)((j >> 1) & 3)) << 4) + (uint32_t)(lane & 1) * 8u;
                  asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(vh[it8].x), "=r"(vh[it8].y) : "r"(stg + o));
                  asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(vl[it8].x), "=r"(vl[it8].y) : "r"(stg + 2048u + o));
                }
#pragma unroll
                for (int it8 = 0; it8 < 8; ++it8) {
                  const int j = 4 * it8 + (lane >> 3);
                  if (j < N) {
                    const size_t off = ((size_t)b * N + j) * args.ldp16 + (size_t)h * C + c + pc;
                    *reinterpret_cast<uint2*>(args.dP_hi16 + off) = vh[it8];
                    *reinterpret_cast<uint2*>(args.dP_lo16 + off) = vl[it8];
                  }
                }
              }
              __syncwarp();
            } else if (lane < N) {
              float* pf = args.dP_aug + row * p.ldp + (size_t)h * C + c;
#pragma unroll
              for (int k = 0; k < 32; k += 4)
                if (c + k < C)
                  *reinterpret_cast<float4*>(pf + k) = make_float4(__uint_as_float(r[k]) * k_dp, __uint_as_float(r[k + 1]) * k_dp,
                                                                   __uint_as_float(r[k + 2]) * k_dp, __uint_as_float(r[k + 3]) * k_dp);
            }
          }
        }
        tc_fence_before();
        bar_group(2, kDeT);
        if (de == 0) bar_arrive(BAR(kBarDTEmpty));
        t_de += clock64() - t0;
      }
    }
    bar_group(2, kDeT);
    for (int c = de; c < C; c += kDeT) args.dbias_part[(size_t)blockIdx.x * p.ldo + c] = ldsa(a_dbias + (uint32_t)c * 4u);
    if (de == 0) {
      atomicAdd(&g_bwd3_counters[2], (unsigned long long)w_dt);
      atomicAdd(&g_bwd3_counters[12], (unsigned long long)t_de);
    }
  } else {
    // =========================================== softmax group ===========================================
    const int h = warp, i = lane;                     // thread = (head, target) for the column passes, (head, source) for rows
    const bool on = h < H && i < N;
    const float g_scale = 1.f / (float)H;
    const float inv_nm1 = 1.f / (float)(N > 1 ? N - 1 : 1);
    SlotCursor cur{0, 0u, pl.n_slots};
    long long w_t = 0, t_sm = 0, w_da = 0, t_sb = 0, w_vfull = 0, t_v = 0;
    float dsd_max = 0.f;                 // max |ds|, |dd| written by this thread
    // phase V: a warp takes every 8th row of a chunk; a lane owns features 2l, 2l+1, 64+2l, 64+2l+1 and all heads
    constexpr int kHP = HT ? (HT + 1) / 2 : kMaxHeads / 2;      // head pairs
    const int f0 = 2 * lane, f1 = 64 + 2 * lane;
    float2 run[4][kHP];
#pragma unroll
    for (int f = 0; f < 4; ++f)
#pragma unroll
      for (int k = 0; k < kHP; ++k) run[f][k] = make_float2(0.f, 0.f);
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
      bar_wait_t(BAR(kBarAlphaEmpty), gph ^ 1u, w_t);     // every MMA of the previous graph is done: its slots were all filled
      for (int pc = 0; pc < pl.t_pieces; ++pc) {
        bar_wait_t(BAR(kBarFull + cur.slot), cur.ph, w_t);
        const uint32_t sa = a_slots + (uint32_t)cur.slot * pl.slot_bytes, off = (uint32_t)pc * pl.slot_bytes;
        const int n16 = (int)(min(pl.slot_bytes, pl.tile_bytes - off) / 16u);
        for (int idx = tid; idx < n16; idx += kSmT) sts128a(a_work + off + (uint32_t)idx * 16u, lds128a(sa + (uint32_t)idx * 16u));
        bar_group(1, kSmT);
        if (tid == 0) bar_arrive(BAR(kBarEmpty + cur.slot));
        cur.advance();
      }
      if (tid == 0) bar_arrive(BAR(kBarTDone));
      cur.advance(pl.n_kb + pl.n_cb);                                            // on to this graph's V slots
      for (int k = 0; k < 2; ++k) {
        const int idx = tid + k * kSmT;
        if (idx < N * 2 * H) stsa(a_sd + (uint32_t)idx * 4u, sd_reg[k]);
      }
      bar_group(1, kSmT);
      // ------------------------------------------------ softmax (thread = head h, target i) ------------------------------------------------
      const long long t_s0 = clock64();
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
          const float hi = tr_tf32(am);
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
      t_sm += clock64() - t_s0;
      bar_wait_t(BAR(kBarDAFull), gph, w_da);
      const long long t_b0 = clock64();
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
        dsd_max = fmaxf(dsd_max, fabsf(dd));
        if (args.dsd) args.dsd[((size_t)b * N + i) * 2 * H + H + h] = dd;
        else args.dP_aug[((size_t)b * N + i) * p.ldp + HC + H + h] = dd;
      }
      bar_group(1, kSmT);
      if (on) {                                                                  // ds_j = row sum of dz (thread = head h, source j = lane)
        const uint32_t row = a_work + (uint32_t)((h * N + lane) * kNS3) * 4u;
        float ds = 0.f;
        for (int k = 0; k < N; ++k) ds += ldsa(row + (uint32_t)k * 4u);
        dsd_max = fmaxf(dsd_max, fabsf(ds));
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
      t_sb += clock64() - t_b0;
      // ------------------------------------------------ V: dv += dz'^T . edge rows (exact fp32 on the CUDA cores) ------------------------------------------------
      if (has_v) {
        // (a ring of fewer slots than G blocks + 2 could still hold an unfilled G block where the first V chunk goes)
        if (pl.n_slots <= pl.n_cb + 1) bar_wait_t(BAR(kBarAlphaEmpty), gph, w_vfull);
        float2 acc[4][kHP];
#pragma unroll
        for (int f = 0; f < 4; ++f)
#pragma unroll
          for (int k = 0; k < kHP; ++k) acc[f][k] = make_float2(0.f, 0.f);
        const uint32_t hstride = (uint32_t)(N * kNS3) * 4u;
        const bool f0_on = f0 < Fe, f1_on = f1 < Fe;
        for (int c = 0; c < pl.nchunks; ++c) {
          bar_wait_t(BAR(kBarFull + cur.slot), cur.ph, w_vfull);
          const long long t0 = clock64();
          const uint32_t sa = a_slots + (uint32_t)cur.slot * pl.slot_bytes;
          int rows = p.R - c * pl.chunk_rows;
          if (rows > pl.chunk_rows) rows = pl.chunk_rows;
          for (int rb = warp; rb < rows; rb += 2 * kSmWarps) {
            // two rows per round (rb, rb + 8), branch-free, loads first: table entries and the lane's four edge features, then
            // the H dz' values each entry points at.  A row past the chunk's end or skipped by the table contributes e = 0.
            int off[2];
            float2 e0[2], e1[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int r = min(rb + u * kSmWarps, rows - 1);
              asm volatile("ld.shared.s32 %0, [%1];" : "=r"(off[u]) : "r"(a_table + (uint32_t)(c * pl.chunk_rows + r) * 4u));
              e0[u] = f0_on ? lds64a(sa + (uint32_t)(r * Fe + f0) * 4u) : make_float2(0.f, 0.f);
              e1[u] = f1_on ? lds64a(sa + (uint32_t)(r * Fe + f1) * 4u) : make_float2(0.f, 0.f);
            }
            float z[2][2 * kHP];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const bool ok = rb + u * kSmWarps < rows && off[u] >= 0;
              if (!ok) e0[u] = e1[u] = make_float2(0.f, 0.f);
              const uint32_t zb = a_work + (uint32_t)max(off[u], 0) * 4u;
#pragma unroll
              for (int hh = 0; hh < 2 * kHP; ++hh) z[u][hh] = hh < H ? ldsa(zb + (uint32_t)hh * hstride) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
#pragma unroll
              for (int k = 0; k < kHP; ++k) {
                const float2 zz = make_float2(z[u][2 * k], z[u][2 * k + 1]);
                acc[0][k] = ffma2(make_float2(e0[u].x, e0[u].x), zz, acc[0][k]);
                acc[1][k] = ffma2(make_float2(e0[u].y, e0[u].y), zz, acc[1][k]);
                acc[2][k] = ffma2(make_float2(e1[u].x, e1[u].x), zz, acc[2][k]);
                acc[3][k] = ffma2(make_float2(e1[u].y, e1[u].y), zz, acc[3][k]);
              }
            }
          }
          bar_group(1, kSmT);
          if (tid == 0) bar_arrive(BAR(kBarEmpty + cur.slot));
          cur.advance();
          t_v += clock64() - t0;
        }
        if (tid == 0) bar_arrive(BAR(kBarVDone));
#pragma unroll
        for (int f = 0; f < 4; ++f)
#pragma unroll
          for (int k = 0; k < kHP; ++k) { run[f][k].x += acc[f][k].x; run[f][k].y += acc[f][k].y; }   // two-level sum
      }
    }
    if (args.dsd_amax) {
      for (int o = 16; o > 0; o >>= 1) dsd_max = fmaxf(dsd_max, __shfl_xor_sync(0xffffffffu, dsd_max, o));
      if (lane == 0 && dsd_max > 0.f) atomicMax(args.dsd_amax, __float_as_uint(dsd_max));
    }
    // per-CTA partials: dv_part[cta * kSmWarps + warp][h][f]
    if (has_v) {
      float* dst = args.dv_part + ((size_t)blockIdx.x * kSmWarps + warp) * H * Fe;
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        const int feat = (f < 2 ? f0 : f1) + (f & 1);
        if (feat < Fe) {
#pragma unroll
          for (int k = 0; k < kHP; ++k) {
            if (2 * k < H) dst[(size_t)(2 * k) * Fe + feat] = run[f][k].x;
            if (2 * k + 1 < H) dst[(size_t)(2 * k + 1) * Fe + feat] = run[f][k].y;
          }
        }
      }
    }
    if (tid == 0) {
      atomicAdd(&g_bwd3_counters[9], (unsigned long long)w_vfull);
      atomicAdd(&g_bwd3_counters[10], (unsigned long long)t_v);
      atomicAdd(&g_bwd3_counters[11], (unsigned long long)w_t);
      atomicAdd(&g_bwd3_counters[13], (unsigned long long)t_sm);
      atomicAdd(&g_bwd3_counters[14], (unsigned long long)w_da);
      atomicAdd(&g_bwd3_counters[15], (unsigned long long)t_sb);
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
  return round_up(ctas * (kSmWarps * (size_t)d->H * d->Fe + ldo) * sizeof(float), 256);
}

int launch_attn_bwd3(AttnBwdArgs& a, float* dv, float* dbias, void* ws, size_t ws_bytes, cudaStream_t st) {
  const AttnParams& p = a.p;
  const Bwd3Plan pl = make_plan3(p);
  if (pl.total > 227 * 1024) return fail(SPOTV2_ERR_UNSUPPORTED, "attn_bwd3: shared-memory plan does not fit");
  int grid = sm_count();
  if (grid > p.B) grid = p.B;
  const int rg = p.terms_in ? 0 : kSmWarps;
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
  const bool drop = p.drop.p > 0.f;
  auto kern = p.H == 6 ? (drop ? gat_attn_bwd3_kernel<true, 6> : gat_attn_bwd3_kernel<false, 6>)
                       : (drop ? gat_attn_bwd3_kernel<true, 0> : gat_attn_bwd3_kernel<false, 0>);
  SPOTV2_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.total));
  kern<<<grid, kB3Threads, pl.total, st>>>(a, pl, tmP, tmG, tmGt);
  SPOTV2_CUDA_OK(cudaGetLastError());
  return reduce_partials2(a.dv_part, grid * rg, (dv && rg > 0) ? p.H * p.Fe : 0, dv, a.dbias_part, grid, dbias ? p.ldo : 0, dbias, st);
}

int bwd3_diag_add(unsigned long long* host_out, int reset) {
  unsigned long long tmp[kNumCounters];
  SPOTV2_CUDA_OK(cudaMemcpyFromSymbol(tmp, g_bwd3_counters, sizeof(tmp)));
  for (int k = 0; k < kNumCounters; ++k) host_out[k] += tmp[k];
  if (reset) {
    unsigned long long zeros[kNumCounters] = {0};
    SPOTV2_CUDA_OK(cudaMemcpyToSymbol(g_bwd3_counters, zeros, sizeof(zeros)));
  }
  return SPOTV2_OK;
}

}  // namespace spotv2
