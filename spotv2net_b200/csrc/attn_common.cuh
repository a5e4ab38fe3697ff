// Pieces shared by the fused attention forward and backward kernels:
// shared-memory carve-up, the edge-row ring (1-D bulk async copies), the 3xTF32
// mma.sync edge-logit phase and the per-(head, target) softmax.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace spotv2 {

constexpr int kAttnThreads = 256;
constexpr int kFwdChunkRows = 64; // edge rows per ring stage, forward  (4 m16 tiles)
constexpr int kBwdChunkRows = 48; // edge rows per ring stage, backward (3 m16 tiles; smem budget)
constexpr int kMaxHeads = 8;     // one n8 MMA tile of heads (reference HPO range is 2..7; config C uses 8)
constexpr int kMaxFe = 512;

// Attention dropout (F.dropout(alpha, p, training) in [PyG] gat_conv.py message): a counter-based mask that the
// forward and the recomputing backward regenerate independently.  Element e = ((b*H + h)*N + i)*N + j takes lane
// e & 3 of Philox4x32-10(counter = e >> 2, key = seed); kept iff the 32-bit draw >= thresh = p * 2^32.
struct DropoutParams {
  float p;          // 0 = off
  float scale;      // 1 / (1 - p)
  uint32_t thresh, k0, k1;
};

inline DropoutParams dropout_params(const spotv2_gat_desc* d) {
  DropoutParams r{0.f, 1.f, 0u, d->dropout_seed_lo, d->dropout_seed_hi};
  if (d->dropout_p > 0.f) {
    r.p = d->dropout_p;
    r.scale = 1.f / (1.f - d->dropout_p);
    const double t = (double)d->dropout_p * 4294967296.0;
    r.thresh = t >= 4294967295.0 ? 0xffffffffu : (uint32_t)t;
  }
  return r;
}

// What edge_terms carries between spotv2_gat_attn_fwd_pair and spotv2_gat_attn_bwd_pair (AttnParams::alpha_rec): 1 = the
// signed attention coefficients.  A function of the descriptor only; attention dropout keeps the edge terms (the backward
// needs the un-dropped coefficients of dropped edges, which the forward's tile no longer has).
inline int attn_record_of(const spotv2_gat_desc* d) {
  return (d->p_format == 1 && d->Fe > 0 && d->N <= 32 && !(d->dropout_p > 0.f)) ? 1 : 0;
}

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t k0, uint32_t k1) {
  uint32_t c2 = 0u, c3 = 0u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
    const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
    c0 = h1 ^ c1 ^ k0;
    c1 = l1;
    c2 = h0 ^ c3 ^ k1;
    c3 = l0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ bool dropout_keep(const DropoutParams& d, unsigned long long e) {
  const uint4 r = philox4x32_10((uint32_t)(e >> 2), (uint32_t)(e >> 34), d.k0, d.k1);
  const uint32_t lane = (uint32_t)e & 3u;
  const uint32_t v = lane == 0 ? r.x : lane == 1 ? r.y : lane == 2 ? r.z : r.w;
  return v >= d.thresh;
}

// Bit j = 1 iff element base + j is kept, j < n <= 32 (one Philox call per 4 consecutive elements).
__device__ __forceinline__ uint32_t dropout_keep_bits(const DropoutParams& d, unsigned long long base, int n) {
  uint32_t bits = 0u;
  unsigned long long cur = ~0ull;
  uint4 r = make_uint4(0u, 0u, 0u, 0u);
  for (int j = 0; j < n; ++j) {
    const unsigned long long e = base + (unsigned long long)j;
    if ((e >> 2) != cur) {
      cur = e >> 2;
      r = philox4x32_10((uint32_t)cur, (uint32_t)(cur >> 32), d.k0, d.k1);
    }
    const uint32_t lane = (uint32_t)e & 3u;
    const uint32_t v = lane == 0 ? r.x : lane == 1 ? r.y : lane == 2 ? r.z : r.w;
    bits |= (v >= d.thresh ? 1u : 0u) << j;
  }
  return bits;
}

struct AttnParams {
  int B, N, F, Fe, H, C, R, concat, ldp, ldo;
  float slope;
  const float* P_aug;
  const float* edge_rows;
  const int32_t* table;
  const float* v;
  int bulk_ok;     // edge block 16-byte aligned and R*Fe % 4 == 0
  int vec2_ok;     // C even (8-byte aligned channel pairs)
  DropoutParams drop;
  // Edge-term tile g[b][h][j][i] = <e_ij, v_h> in the kernels' shared-memory layout ([H][N][kEdgeTermNS] floats per
  // graph for N <= 32, [H][N][N] for larger graphs).  The forward writes it when given; the backward reads it instead
  // of streaming the edge rows a first time and recomputing the logits (6 floats per edge instead of Fe).  Null = off.
  float* edge_terms;
  int terms_in;          // edge_mode 1: edge_terms is an INPUT (spotv2_edge_terms_from_windows); there are no edge rows
  float* dterms_out;     // backward, edge_mode 1: receives dz' (gradient w.r.t. the edge terms) in the same tile layout
  // p_format 1 without attention dropout: what the forward leaves in edge_terms for the backward is not the edge terms but
  // its RESULT - the attention coefficients, in the same tile layout, with the LeakyReLU side in the sign bit (+alpha: z > 0,
  // -alpha: z <= 0; -0.0 counts).  The backward then has no logits, no s | d and no softmax to redo.  Both pair entry points
  // derive the flag from the descriptor alone (attn_record_of), so the two calls of one step always agree.
  int alpha_rec = 0;
  // p_format 1: the projection arrives as an fp16 operand pair (planes [B*N, ldp16], head pitch hp, scale block p_blk:
  // [2],[3] inverse scales of the projection / s|d column groups, [4],[5] the scales); P_aug is null then.  P_lo null =
  // half-precision class (hi plane only).
  const __half* P_hi = nullptr;
  const __half* P_lo = nullptr;
  const float* p_blk = nullptr;
  const float* sd32 = nullptr;     // [B*N, 2H] fp32: the logit terms s | d (the planes' own s|d columns are not read)
  int ldp16 = 0, hp = 0;
  int lg_tensor_cores;   // large-universe path: batched GEMMs on mma.sync (3xTF32) unless gemm_algo == 1 (exact-fp32 FFMA2)
};

constexpr int kEdgeTermNS = 36;     // == the alpha-tile row stride of attn_fwd.cu / attn_bwd2.cu

// Shared-memory plan common to both directions.  All offsets in bytes, 16-aligned.
struct AttnSmem {
  int NS;          // tile row stride over targets i (multiple of 4)
  int KS;          // k-steps of 8 over Fe
  int NT;          // n-tiles of 8 over H
  int chunk_rows;  // edge rows per ring stage (multiple of 16)
  size_t off_bar, off_table, off_vfrag, off_sd, off_tile, off_ring, ring_stage_bytes, base_total;
};

inline AttnSmem attn_smem_plan(int N, int Fe, int H, int R, int npairs, int chunk_rows) {
  AttnSmem s;
  s.chunk_rows = chunk_rows;
  s.NS = (2 * npairs + 3) / 4 * 4;
  s.KS = ((Fe + 7) / 8 + 7) / 8 * 8;     // k-steps of 8 over Fe, padded to a multiple of 8 (zero B fragments)
  s.NT = (H + 7) / 8;
  size_t o = 0;
  s.off_bar = o;   o += 512;
  s.off_table = o; o += round_up((size_t)R * 4, 16);
  s.off_vfrag = o; o += (size_t)s.NT * s.KS * 32 * 16;
  s.off_sd = o;    o += round_up((size_t)N * 2 * H * 4, 16);
  s.off_tile = o;  o += round_up((size_t)H * N * s.NS * 4, 16);
  s.off_ring = o;
  s.ring_stage_bytes = round_up((size_t)chunk_rows * Fe * 4, 128);
  s.base_total = o;
  return s;
}

// ---------------------------------------------------------------------------------------
// Edge-row ring: two stages of chunk_rows rows, filled by cp.async.bulk (thread 0 issues,
// an mbarrier counts the bytes) or, for unaligned blocks, by a cooperative copy.
struct EdgeRing {
  float* stage[2];
  uint64_t* full;       // [2]
  uint32_t uses[2];     // completed waits per stage (uniform across the CTA)
  int nchunks;
  int chunk_rows;
  const AttnParams* p;

  __device__ __forceinline__ int rows_in(int c) const {
    int r = p->R - c * chunk_rows;
    return r < chunk_rows ? r : chunk_rows;
  }
  __device__ __forceinline__ void issue(int b, int c) {   // call from ONE thread
    const int s = c & 1;
    const uint32_t bytes = (uint32_t)rows_in(c) * p->Fe * 4u;
    const float* src = p->edge_rows + ((size_t)b * p->R + (size_t)c * chunk_rows) * p->Fe;
    mbar_expect_tx(&full[s], bytes);
    bulk_g2s(stage[s], src, bytes, &full[s]);
  }
  __device__ __forceinline__ void prefetch_first(int b) {  // call from ONE thread
    issue(b, 0);
    if (nchunks > 1) issue(b, 1);
  }
};

// v [H, Fe] -> B fragments of mma.m16n8k8 (col-major 8x8: b0 = (k=t, n=g), b1 = (k=t+4, n=g)),
// pre-split into tf32 hi/lo:  vfrag[(nt*KS + ks)*32 + lane] = {b0_hi, b1_hi, b0_lo, b1_lo}.
__device__ __forceinline__ void build_vfrag(float4* vfrag, const float* v, int H, int Fe, int KS,
                                            int NT, int tid, int nthreads) {
  for (int idx = tid; idx < NT * KS * 32; idx += nthreads) {
    const int lane = idx & 31, ks = (idx >> 5) % KS, nt = (idx >> 5) / KS;
    const int g = lane >> 2, t = lane & 3;
    const int n = nt * 8 + g, k0 = ks * 8 + t, k1 = k0 + 4;
    const float b0 = (n < H && k0 < Fe) ? v[n * Fe + k0] : 0.f;
    const float b1 = (n < H && k1 < Fe) ? v[n * Fe + k1] : 0.f;
    uint32_t h0, l0, h1, l1;
    split_tf32(b0, h0, l0);
    split_tf32(b1, h1, l1);
    vfrag[idx] = make_float4(__uint_as_float(h0), __uint_as_float(h1), __uint_as_float(l0),
                             __uint_as_float(l1));
  }
}

// One warp: edge terms g[row, h] = <edge_row, v_h> for the 16 rows [m0, m0+16) of a staged chunk,
// fp32-accurate through the 3xTF32 split (lo*hi + hi*lo + hi*hi, small terms first).
// Results go to sink(row_in_chunk, head, value) for the rows/heads this lane owns.
template <int NT_MAX, int kSlots, class Sink>
__device__ __forceinline__ void warp_edge_logits(const float* Ts, const float4* vfrag, int Fe, int KS,
                                                 int NT, int m0, int lane, Sink&& sink, long long* t_mma = nullptr) {
  static_assert(NT_MAX == 1, "one n8 tile of heads (H <= 8)");
  const int g = lane >> 2, t = lane & 3;
  const long long t_begin = t_mma ? clock64() : 0;
  // One warp per scheduler runs this loop, so it has to carry its own instruction-level parallelism:
  // the body is branch-free (KS is padded to a multiple of kSlots with zero B fragments; out-of-range
  // feature indices are clamped onto real data and multiplied by those zeros), all fragment loads of a
  // block of kSlots k-steps are issued up front, and every k-step slot and product has its own
  // accumulator so no mma.sync waits on another (their latency is long on sm_100).
  float acc[kSlots][3][4];
#pragma unroll
  for (int sl = 0; sl < kSlots; ++sl)
#pragma unroll
    for (int pr = 0; pr < 3; ++pr)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[sl][pr][q] = 0.f;
  const float* r0 = Ts + (size_t)(m0 + g) * Fe;
  const float* r1 = r0 + (size_t)8 * Fe;
  const int kmax = Fe - 1;
  (void)NT;
  for (int ks0 = 0; ks0 < KS; ks0 += kSlots) {
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
      mma_tf32_16x8x8(acc[sl][0], al, bh);
      mma_tf32_16x8x8(acc[sl][1], ah, bl);
      mma_tf32_16x8x8(acc[sl][2], ah, bh);
    }
  }
  float c[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float corr = 0.f, mainp = 0.f;
#pragma unroll
    for (int sl = 0; sl < kSlots; ++sl) {
      corr += acc[sl][0][q] + acc[sl][1][q];
      mainp += acc[sl][2][q];
    }
    c[q] = corr + mainp;
  }
  if (t_mma) *t_mma += clock64() - t_begin;          // includes waiting for the accumulators
  const int n = 2 * t;
  sink(m0 + g, n, c[0]);
  sink(m0 + g, n + 1, c[1]);
  sink(m0 + g + 8, n, c[2]);
  sink(m0 + g + 8, n + 1, c[3]);
}

// Phase 1 for one graph: stream the R edge rows through the ring and scatter the edge terms
// into tile[h][j][i] (i contiguous, stride NS).  Ends with a __syncthreads().
__device__ __forceinline__ void edge_logit_phase(EdgeRing& ring, const AttnParams& p,
                                                 const AttnSmem& sm, float* tile,
                                                 const int32_t* table_s, const float4* vfrag,
                                                 int b, int tid) {
  const int warp = tid >> 5, lane = tid & 31;
  const int N = p.N, H = p.H, NS = sm.NS;
  for (int c = 0; c < ring.nchunks; ++c) {
    const int s = c & 1;
    const int rows = ring.rows_in(c);
    if (p.bulk_ok) {
      mbar_wait(&ring.full[s], ring.uses[s] & 1);
      ring.uses[s]++;
    } else {
      const float* src = p.edge_rows + ((size_t)b * p.R + (size_t)c * ring.chunk_rows) * p.Fe;
      for (int idx = tid; idx < rows * p.Fe; idx += kAttnThreads) ring.stage[s][idx] = src[idx];
      __syncthreads();
    }
    if (warp * 16 < rows) {
      const int row_base = c * ring.chunk_rows;
      warp_edge_logits<1, 4>(ring.stage[s], vfrag, p.Fe, sm.KS, sm.NT, warp * 16, lane,
                          [&](int r, int h, float val) {
                            if (r < rows && h < H) {
                              const int code = table_s[row_base + r];
                              if (code >= 0) tile[(h * N + (code & 0xffff)) * NS + (code >> 16)] = val;
                            }
                          });
    }
    __syncthreads();
    if (p.bulk_ok && tid == 0 && c + 2 < ring.nchunks) ring.issue(b, c + 2);
  }
}

// Phase 2: thread (h, i) turns the edge terms of target i into attention coefficients in place.
//   g_ii = mean_{j != i} g_ij (PyG's fill_value='mean' self loop),  z = s_j + d_i + g_ij,
//   l = leaky_relu(z),  alpha = softmax_j(l)  (max-subtracted, true division as in PyG).
// tile[h][j][i] <- alpha * out_scale ; optional raw alpha to global ; optional z>0 mask bits.
__device__ __forceinline__ void softmax_phase(const AttnParams& p, const AttnSmem& sm, float* tile,
                                              const float* sd, float out_scale, float* alpha_out_b,
                                              uint32_t* pos_mask, int tid, int nthreads = kAttnThreads,
                                              int sd_j_stride = -1, int sd_swizzled = 0,
                                              const float* tile_add = nullptr, int drop_graph = -1) {
  // drop_graph >= 0: apply this call's attention dropout to graph `drop_graph` (the forward); the backward
  // passes -1, recomputes the un-dropped coefficients and applies the mask itself.
  const int N = p.N, H = p.H, NS = sm.NS;
  // s_j = sd(j, h), d_i = sd(i, H + h).  sd is either a packed [N][2H] array or (sd_swizzled) a
  // 128B-swizzled TMA tile of 32 rows x 32 floats holding the 2H augmented columns of P_aug.
  auto sd_at = [&](int j, int k) -> float {
    if (sd_swizzled) return sd[j * 32 + ((((k >> 2) ^ (j & 7)) << 2) | (k & 3))];
    return sd[j * (sd_j_stride < 0 ? 2 * H : sd_j_stride) + k];
  };
  for (int idx = tid; idx < H * N; idx += nthreads) {
    const int h = idx / N, i = idx - h * N;
    float* col = tile + (size_t)h * N * NS + i;
    if (tile_add) {            // edge terms arrive as two partial sums (k halves): fold the second one in first
      const float* col2 = tile_add + (size_t)h * N * NS + i;
#pragma unroll 6
      for (int j = 0; j < N; ++j)
        if (j != i) col[j * NS] += col2[j * NS];
    }
    float gsum = 0.f;
#pragma unroll 6
    for (int j = 0; j < N; ++j) gsum += (j != i) ? col[j * NS] : 0.f;
    const float gii = gsum / (float)(N > 1 ? N - 1 : 1);
    const float di = sd_at(i, H + h);
    float mx = -INFINITY;
    uint32_t mask = 0;
#pragma unroll 6
    for (int j = 0; j < N; ++j) {
      const float z = (j == i ? gii : col[j * NS]) + sd_at(j, h) + di;
      if (z > 0.f) mask |= 1u << j;
      const float l = z > 0.f ? z : z * p.slope;
      mx = fmaxf(mx, l);
      col[j * NS] = l;
    }
    float sum = 0.f;
#pragma unroll 6
    for (int j = 0; j < N; ++j) {
      const float e = expf(col[j * NS] - mx);
      sum += e;
      col[j * NS] = e;
    }
    // PyG divides by (sum + 1e-16); one reciprocal + multiplies differ from that by <= 1 ulp
    const float inv = 1.f / (sum + 1e-16f);
    uint32_t keep = 0xffffffffu;
    float kscale = 1.f;
    if (drop_graph >= 0 && p.drop.p > 0.f) {
      keep = dropout_keep_bits(p.drop, (((unsigned long long)drop_graph * H + h) * N + i) * N, N);
      kscale = p.drop.scale;
    }
#pragma unroll 6
    for (int j = 0; j < N; ++j) {
      const float a = ((keep >> j) & 1u) ? col[j * NS] * inv * kscale : 0.f;
      if (alpha_out_b) alpha_out_b[((size_t)h * N + j) * N + i] = a;
      col[j * NS] = a * out_scale;
    }
    if (pos_mask) pos_mask[idx] = mask;
  }
}


// The same phase with the row held in registers between the passes (N <= 32): one pass of shared-memory loads and one of
// stores per row instead of four and three.  Arithmetic and its order are those of softmax_phase - the results are
// bit-identical (the LeakyReLU kinks and the attention coefficients of the forward and of the recomputing backward must
// not depend on which of the two a kernel calls).  sd is the packed [N][2H] array.
// signed_rec: the tile receives the record of AttnParams::alpha_rec (callers pass out_scale 1 with it).
__device__ __forceinline__ void softmax_phase_regs(const AttnParams& p, int NS, float* tile, const float* sd, float out_scale,
                                                   float* alpha_out_b, uint32_t* pos_mask, int tid, int nthreads,
                                                   const float* tile_add = nullptr, int drop_graph = -1, bool signed_rec = false) {
  const int N = p.N, H = p.H;
  for (int idx = tid; idx < H * N; idx += nthreads) {
    const int h = idx / N, i = idx - h * N;
    float* col = tile + (size_t)h * N * NS + i;
    float r[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) r[j] = j < N ? col[j * NS] : 0.f;
    if (tile_add) {
      const float* col2 = tile_add + (size_t)h * N * NS + i;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < N && j != i) r[j] += col2[j * NS];
    }
    float gsum = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < N) gsum += (j != i) ? r[j] : 0.f;
    const float gii = gsum / (float)(N > 1 ? N - 1 : 1);
    const float di = sd[i * 2 * H + H + h];
    float mx = -INFINITY;
    uint32_t mask = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < N) {
        const float z = (j == i ? gii : r[j]) + sd[j * 2 * H + h] + di;
        if (z > 0.f) mask |= 1u << j;
        const float l = z > 0.f ? z : z * p.slope;
        mx = fmaxf(mx, l);
        r[j] = l;
      }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < N) {
        const float e = expf(r[j] - mx);
        sum += e;
        r[j] = e;
      }
    const float inv = 1.f / (sum + 1e-16f);
    uint32_t keep = 0xffffffffu;
    float kscale = 1.f;
    if (drop_graph >= 0 && p.drop.p > 0.f) {
      keep = dropout_keep_bits(p.drop, (((unsigned long long)drop_graph * H + h) * N + i) * N, N);
      kscale = p.drop.scale;
    }
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < N) {
        const float a = ((keep >> j) & 1u) ? r[j] * inv * kscale : 0.f;
        if (alpha_out_b) alpha_out_b[((size_t)h * N + j) * N + i] = a;
        col[j * NS] = (signed_rec && !((mask >> j) & 1u)) ? -(a * out_scale) : a * out_scale;
      }
    if (pos_mask) pos_mask[idx] = mask;
  }
}

}  // namespace spotv2
