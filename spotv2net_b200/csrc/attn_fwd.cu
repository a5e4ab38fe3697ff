// Fused GAT attention forward for batches of small complete graphs (N <= 32).
// Replaces, per graph, PyG 2.3.0 GATConv's edge_update (gathers, lin_edge, leaky_relu),
// utils.softmax, message + 'add' aggregation, head mean/concat and bias
// ([PyG] nn/conv/gat_conv.py; reached from /root/reference/utils/models.py:146).
//
// One persistent CTA per SM, 672 threads, warp-specialised and pipelined ACROSS graphs:
//
//   group A (warps 0-3)   for graph b+1: edge rows stream through a 2-stage shared-memory ring
//                         (cp.async.bulk + mbarrier, issued two chunks ahead, across graph boundaries);
//                         g[e,h] = <edge_attr[e], v_h> on mma.sync m16n8k8 with a 3xTF32 split;
//                         self-loop mean fill, s_j + d_i + g_ij, LeakyReLU, softmax over sources
//                         -> attention tile[buf] in shared memory.
//   group B (warps 4-19)  for graph b: out[i, c] = sum_{h,j} alpha_h[i,j] P[j,h,c].  A thread owns a
//                         channel pair and half of the targets (packed FFMA2, alpha broadcast from shared
//                         memory, next alpha row prefetched into registers): 4 warps per scheduler hide
//                         the shared-memory latency.
//   warp 20               P-row producer: one cp.async.bulk per source row (all heads, 12 KB) into a
//                         kPRows-deep shared-memory ring, running ahead across graph boundaries.
//
// The two groups hand tiles over through mbarriers (tile_full / tile_empty), so the edge stream and the
// P stream keep HBM busy at the same time and the softmax never sits on the aggregation's critical path.
#include "attn_common.cuh"

namespace spotv2 {

constexpr int kGroupA = 128;
constexpr int kGroupB = 512;          // 16 warps: 256 channel pairs x 2 target halves
constexpr int kItemsPerPass = kGroupB / 2;
constexpr int kFwdThreads = kGroupA + kGroupB + 32;
constexpr int kPRows = 4;          // P-row ring depth (rows in flight)

struct AttnFwdArgs {
  AttnParams p;
  const float* bias;
  float* out;
  float* alpha_out;
};

__device__ __forceinline__ void bar_sync_group_a() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int NPAIRS, bool VEC2>
__global__ void __launch_bounds__(kFwdThreads, 1)
gat_attn_fwd_kernel(const AttnFwdArgs args, const AttnSmem sm, const uint32_t off_prow, const uint32_t prow_bytes) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const AttnParams& p = args.p;
  const int tid = threadIdx.x;
  const int N = p.N, H = p.H, C = p.C, NS = sm.NS;
  const int HC = H * C;
  const int tile_floats = H * N * NS;
  const int sd_floats = N * 2 * H;

  // [0,1] edge ring, [2,3] tile_full, [4,5] tile_empty, [6..6+kPRows) prow_full, then prow_empty
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + sm.off_bar);
  uint64_t* tile_full = bars + 2;
  uint64_t* tile_empty = bars + 4;
  uint64_t* prow_full = bars + 6;
  uint64_t* prow_empty = bars + 6 + kPRows;
  const int CP = (C + 1) / 2;
  const int n_items = p.concat ? H * CP : CP;
  const int n_pass = (n_items + kItemsPerPass - 1) / kItemsPerPass;
  int32_t* table_s = reinterpret_cast<int32_t*>(smem_raw + sm.off_table);
  float4* vfrag = reinterpret_cast<float4*>(smem_raw + sm.off_vfrag);
  float* sd0 = reinterpret_cast<float*>(smem_raw + sm.off_sd);
  float* tile0 = reinterpret_cast<float*>(smem_raw + sm.off_tile);

  const int nchunks = p.Fe > 0 ? (p.R + sm.chunk_rows - 1) / sm.chunk_rows : 0;
  const int my_graphs = (p.B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&tile_full[0], kGroupA);
    mbar_init(&tile_full[1], kGroupA);
    mbar_init(&tile_empty[0], kGroupB);
    mbar_init(&tile_empty[1], kGroupB);
    for (int r = 0; r < kPRows; ++r) { mbar_init(&prow_full[r], 1); mbar_init(&prow_empty[r], kGroupB / 32); }
    fence_mbar_init();
  }
  for (int r = tid; r < p.R; r += kFwdThreads) table_s[r] = p.Fe > 0 ? p.table[r] : -1;
  if (p.Fe > 0) build_vfrag(vfrag, p.v, H, p.Fe, sm.KS, sm.NT, tid, kFwdThreads);
  for (int idx = tid; idx < 2 * tile_floats; idx += kFwdThreads) tile0[idx] = 0.f;
  __syncthreads();

  if (tid < kGroupA) {
    // ================================ group A: logits + softmax ================================
    const int warp = tid >> 5, lane = tid & 31;
    float* stage[2] = {reinterpret_cast<float*>(smem_raw + sm.off_ring),
                       reinterpret_cast<float*>(smem_raw + sm.off_ring + sm.ring_stage_bytes)};
    const int total_chunks = my_graphs * nchunks;      // this CTA's global chunk stream
    auto rows_in = [&](int c) { const int r = p.R - c * sm.chunk_rows; return r < sm.chunk_rows ? r : sm.chunk_rows; };
    auto issue = [&](int k) {                           // one thread: start the k-th chunk of the stream
      const int it = k / nchunks, c = k - it * nchunks;
      const int b = blockIdx.x + it * gridDim.x;
      const uint32_t bytes = (uint32_t)rows_in(c) * p.Fe * 4u;
      mbar_expect_tx(&bars[k & 1], bytes);
      bulk_g2s(stage[k & 1], p.edge_rows + ((size_t)b * p.R + (size_t)c * sm.chunk_rows) * p.Fe, bytes, &bars[k & 1]);
    };
    if (p.bulk_ok && tid == 0) {
      if (total_chunks > 0) issue(0);
      if (total_chunks > 1) issue(1);
    }
    const float out_scale = p.concat ? 1.f : 1.f / (float)H;
    int k = 0;                                          // global chunk counter (stage = k & 1, parity = (k >> 1) & 1)
    for (int it = 0; it < my_graphs; ++it) {
      const int b = blockIdx.x + it * gridDim.x;
      const int buf = it & 1;
      float* tile = tile0 + buf * tile_floats;
      float* sd = sd0 + buf * sd_floats;
      mbar_wait(&tile_empty[buf], ((it >> 1) & 1) ^ 1);   // group B has finished reading this buffer
      for (int idx = tid; idx < sd_floats; idx += kGroupA) {
        const int j = idx / (2 * H), kk = idx - j * 2 * H;
        sd[idx] = p.P_aug[((size_t)b * N + j) * p.ldp + HC + kk];
      }
      if (nchunks == 0) {
        for (int idx = tid; idx < tile_floats; idx += kGroupA) tile[idx] = 0.f;
      }
      for (int c = 0; c < nchunks; ++c, ++k) {
        const int s = k & 1;
        const int rows = rows_in(c);
        if (p.bulk_ok) {
          mbar_wait(&bars[s], (k >> 1) & 1);
        } else {
          const float* src = p.edge_rows + ((size_t)b * p.R + (size_t)c * sm.chunk_rows) * p.Fe;
          for (int idx = tid; idx < rows * p.Fe; idx += kGroupA) stage[s][idx] = src[idx];
          bar_sync_group_a();
        }
        if (warp * 16 < rows) {
          const int row_base = c * sm.chunk_rows;
          warp_edge_logits<1>(stage[s], vfrag, p.Fe, sm.KS, sm.NT, warp * 16, lane, [&](int r, int h, float val) {
            if (r < rows && h < H) {
              const int code = table_s[row_base + r];
              if (code >= 0) tile[(h * N + (code & 0xffff)) * NS + (code >> 16)] = val;
            }
          });
        }
        bar_sync_group_a();                              // stage s consumed by all four warps
        if (p.bulk_ok && tid == 0 && k + 2 < total_chunks) issue(k + 2);
      }
      if (nchunks == 0) bar_sync_group_a();
      softmax_phase(p, sm, tile, sd, out_scale, args.alpha_out ? args.alpha_out + (size_t)b * H * N * N : nullptr,
                    nullptr, tid, kGroupA);
      mbar_arrive_cta(&tile_full[buf]);                  // release: alpha tile visible to group B
    }
  } else if (tid < kGroupA + kGroupB) {
    // ================================ group B: aggregation ================================
    constexpr int HP = (NPAIRS + 1) / 2;                   // target pairs per half
    constexpr int HQ = (HP + 1) / 2;                       // float4 loads per alpha half-row
    const int t = tid - kGroupA;
    const int half = t / kItemsPerPass;                    // which half of the targets
    const int tt = t - half * kItemsPerPass;
    const int h_loop = p.concat ? 1 : H;
    const int a_off = half * 2 * HP;                       // first target of this half (multiple of 4 floats)
    uint32_t rowctr = 0;                                   // position in the P-row stream (graph, pass, j)
    for (int it = 0; it < my_graphs; ++it) {
      const int b = blockIdx.x + it * gridDim.x;
      const int buf = it & 1;
      const float* tile = tile0 + buf * tile_floats + a_off;
      mbar_wait(&tile_full[buf], (it >> 1) & 1);
      for (int pass = 0; pass < n_pass; ++pass) {
        const int item = pass * kItemsPerPass + tt;
        const bool valid = item < n_items;
        const int h0 = (p.concat && valid) ? item / CP : 0;
        const int cp = valid ? (p.concat ? item - h0 * CP : item) : 0;
        const int c0 = 2 * cp;
        const bool has1 = c0 + 1 < C;
        float2 acc[HP][2];
#pragma unroll
        for (int ip = 0; ip < HP; ++ip) acc[ip][0] = acc[ip][1] = make_float2(0.f, 0.f);
        for (int j = 0; j < N; ++j, ++rowctr) {
          const int slot = rowctr % kPRows;
          mbar_wait(&prow_full[slot], (rowctr / kPRows) & 1);
          if (valid) {
            const float* prow = reinterpret_cast<const float*>(smem_raw + off_prow + (size_t)slot * prow_bytes) + h0 * C + c0;
            const float* ar = tile + (size_t)(h0 * N + j) * NS;
            float4 an[HQ];                                 // alpha half-row of the NEXT head, prefetched
#pragma unroll
            for (int q = 0; q < HQ; ++q) an[q] = *reinterpret_cast<const float4*>(ar + 4 * q);
            float2 pn;
            if (VEC2) pn = *reinterpret_cast<const float2*>(prow);
            else { pn.x = prow[0]; pn.y = has1 ? prow[1] : 0.f; }
            for (int hh = 0; hh < h_loop; ++hh) {
              float4 ac[HQ];
#pragma unroll
              for (int q = 0; q < HQ; ++q) ac[q] = an[q];
              const float2 px = make_float2(pn.x, pn.x), py = make_float2(pn.y, pn.y);
              if (hh + 1 < h_loop) {
                ar += (size_t)N * NS;
                prow += C;
#pragma unroll
                for (int q = 0; q < HQ; ++q) an[q] = *reinterpret_cast<const float4*>(ar + 4 * q);
                if (VEC2) pn = *reinterpret_cast<const float2*>(prow);
                else { pn.x = prow[0]; pn.y = has1 ? prow[1] : 0.f; }
              }
#pragma unroll
              for (int q = 0; q < HQ; ++q) {
                const float2 a0 = make_float2(ac[q].x, ac[q].y), a1 = make_float2(ac[q].z, ac[q].w);
                acc[2 * q][0] = ffma2(a0, px, acc[2 * q][0]);
                acc[2 * q][1] = ffma2(a0, py, acc[2 * q][1]);
                if (2 * q + 1 < HP) {
                  acc[2 * q + 1][0] = ffma2(a1, px, acc[2 * q + 1][0]);
                  acc[2 * q + 1][1] = ffma2(a1, py, acc[2 * q + 1][1]);
                }
              }
            }
          }
          __syncwarp();
          if ((tid & 31) == 0) mbar_arrive_cta(&prow_empty[slot]);   // this warp is done with the row
        }
        if (valid) {
          // epilogue: + bias, targets a_off + 2ip and a_off + 2ip + 1
          const int col = h0 * C + c0;
          const float b0 = args.bias ? args.bias[col] : 0.f;
          const float b1 = (args.bias && has1) ? args.bias[col + 1] : 0.f;
          float* orow = args.out + (size_t)b * N * p.ldo + col;
#pragma unroll
          for (int ip = 0; ip < HP; ++ip) {
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              const int i = a_off + 2 * ip + hf;
              if (i < N) {
                const float o0 = (hf ? acc[ip][0].y : acc[ip][0].x) + b0;
                const float o1 = (hf ? acc[ip][1].y : acc[ip][1].x) + b1;
                float* dst = orow + (size_t)i * p.ldo;
                if (VEC2) {
                  *reinterpret_cast<float2*>(dst) = make_float2(o0, o1);
                } else {
                  dst[0] = o0;
                  if (has1) dst[1] = o1;
                }
              }
            }
          }
        }
      }
      mbar_arrive_cta(&tile_empty[buf]);                 // this thread is done reading the tile
    }
  } else if (tid == kGroupA + kGroupB) {
    // ================================ last warp, lane 0: P-row producer ================================
    uint32_t rowctr = 0;
    for (int it = 0; it < my_graphs; ++it) {
      const int b = blockIdx.x + it * gridDim.x;
      for (int pass = 0; pass < n_pass; ++pass) {
        for (int j = 0; j < N; ++j, ++rowctr) {
          const int slot = rowctr % kPRows;
          mbar_wait(&prow_empty[slot], ((rowctr / kPRows) & 1) ^ 1);
          mbar_expect_tx(&prow_full[slot], prow_bytes);
          bulk_g2s(smem_raw + off_prow + (size_t)slot * prow_bytes, p.P_aug + ((size_t)b * N + j) * p.ldp, prow_bytes,
                   &prow_full[slot]);
        }
      }
    }
  }
}

template <int NPAIRS, bool VEC2>
static int launch_fwd(const AttnFwdArgs& a, cudaStream_t st) {
  const AttnParams& p = a.p;
  const size_t tile_bytes = round_up((size_t)p.H * p.N * ((2 * NPAIRS + 3) / 4 * 4) * 4, 16);
  const size_t sd_bytes = round_up((size_t)p.N * 2 * p.H * 4, 16);
  // the common plan reserves one tile + one sd; this kernel double-buffers both
  auto finish = [&](AttnSmem s) {
    s.off_tile = s.off_sd + 2 * sd_bytes;
    s.off_ring = round_up(s.off_tile + 2 * tile_bytes, 128);
    s.base_total = s.off_ring;
    return s;
  };
  const size_t prow_bytes = (size_t)p.ldp * 4;           // one P_aug row (ldp % 4 == 0 -> multiple of 16)
  auto total = [&](const AttnSmem& s) { return round_up(s.off_ring + 2 * s.ring_stage_bytes, 128) + kPRows * prow_bytes; };
  AttnSmem sm = finish(attn_smem_plan(p.N, p.Fe, p.H, p.R, NPAIRS, kFwdChunkRows));
  for (int rows = kFwdChunkRows - 16; rows >= 16 && total(sm) > 227 * 1024; rows -= 16)
    sm = finish(attn_smem_plan(p.N, p.Fe, p.H, p.R, NPAIRS, rows));
  const size_t smem = total(sm);
  if (smem > 227 * 1024)
    return fail(SPOTV2_ERR_UNSUPPORTED, "attn_fwd needs %zu B shared memory (> 227 KB)", smem);
  const uint32_t off_prow = (uint32_t)round_up(sm.off_ring + 2 * sm.ring_stage_bytes, 128);
  auto kern = gat_attn_fwd_kernel<NPAIRS, VEC2>;
  SPOTV2_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = sm_count();
  if (grid > p.B) grid = p.B;
  kern<<<grid, kFwdThreads, smem, st>>>(a, sm, off_prow, (uint32_t)prow_bytes);
  SPOTV2_CUDA_OK(cudaGetLastError());
  return SPOTV2_OK;
}

template <int NPAIRS>
static int dispatch_vec(const AttnFwdArgs& a, cudaStream_t st) {
  return a.p.vec2_ok ? launch_fwd<NPAIRS, true>(a, st) : launch_fwd<NPAIRS, false>(a, st);
}

int attn_fwd_dispatch(const AttnFwdArgs& a, cudaStream_t st) {
  const int np = (a.p.N + 1) / 2;
  if (np <= 4) return dispatch_vec<4>(a, st);
  if (np <= 8) return dispatch_vec<8>(a, st);
  if (np <= 15) return dispatch_vec<15>(a, st);
  if (np <= 16) return dispatch_vec<16>(a, st);
  return fail(SPOTV2_ERR_UNSUPPORTED, "N=%d > 32: the one-CTA-per-graph kernel covers N <= 32", a.p.N);
}

}  // namespace spotv2

using namespace spotv2;

extern "C" int spotv2_gat_attn_fwd(const spotv2_gat_desc* d, const float* P_aug,
                                   const float* edge_rows, const int32_t* table, const float* v,
                                   const float* bias_or_null, float* out, float* alpha_or_null,
                                   void* stream) {
  if (int rc = check_desc(d)) return rc;
  SPOTV2_REQUIRE(P_aug && out, "attn_fwd: P_aug and out must be non-null");
  SPOTV2_REQUIRE(d->Fe == 0 || (edge_rows && table && v),
                 "attn_fwd: edge_rows, table and v are required when Fe > 0");
  SPOTV2_REQUIRE(aligned16(P_aug) && aligned16(out), "attn_fwd: P_aug/out must be 16-byte aligned");
  if (d->H > kMaxHeads) return fail(SPOTV2_ERR_UNSUPPORTED, "H=%d > %d", d->H, kMaxHeads);
  if (d->Fe > kMaxFe) return fail(SPOTV2_ERR_UNSUPPORTED, "Fe=%d > %d", d->Fe, kMaxFe);
  AttnFwdArgs a;
  a.p.B = d->B; a.p.N = d->N; a.p.F = d->F; a.p.Fe = d->Fe; a.p.H = d->H; a.p.C = d->C;
  a.p.R = d->R; a.p.concat = d->concat; a.p.ldp = d->ldp;
  a.p.ldo = d->concat ? d->H * d->C : d->C;
  a.p.slope = d->negative_slope;
  a.p.P_aug = P_aug; a.p.edge_rows = edge_rows; a.p.table = table; a.p.v = v;
  a.p.bulk_ok = d->Fe > 0 && aligned16(edge_rows) && ((size_t)d->R * d->Fe) % 4 == 0;
  a.p.vec2_ok = (d->C % 2 == 0);
  a.bias = bias_or_null; a.out = out; a.alpha_out = alpha_or_null;
  return attn_fwd_dispatch(a, as_stream(stream));
}
