// Fused GAT attention forward for batches of small complete graphs (N <= 32), one graph per
// CTA iteration.  Replaces, per graph, PyG 2.3.0 GATConv's edge_update (gathers, lin_edge,
// leaky_relu), utils.softmax, message + 'add' aggregation, head mean/concat and bias
// ([PyG] nn/conv/gat_conv.py; reached from /root/reference/utils/models.py:146).
//
//   phase 1  edge rows stream through a 2-stage shared-memory ring (1-D bulk async copies);
//            g[e,h] = <edge_attr[e], v_h> on mma.sync m16n8k8 with a 3xTF32 split (fp32-accurate);
//            scattered by the row table into tile[h][j][i].
//   phase 2  thread (h,i): self-loop mean fill, s_j + d_i + g_ij, LeakyReLU, softmax over j.
//   phase 3  out[i, c] = sum_{h,j} alpha_h[i,j] * P[j, h, c]: each thread owns a channel pair
//            and all targets (packed FFMA2, alpha broadcast from shared memory, P streamed
//            straight from HBM with a register double buffer).
#include "attn_common.cuh"

namespace spotv2 {

struct AttnFwdArgs {
  AttnParams p;
  const float* bias;
  float* out;
  float* alpha_out;
};

template <int NPAIRS>
__global__ void __launch_bounds__(kAttnThreads, 2)
gat_attn_fwd_kernel(const AttnFwdArgs args, const AttnSmem sm) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const AttnParams& p = args.p;
  const int tid = threadIdx.x;
  const int N = p.N, H = p.H, C = p.C, NS = sm.NS;
  const int HC = H * C;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + sm.off_bar);
  int32_t* table_s = reinterpret_cast<int32_t*>(smem_raw + sm.off_table);
  float4* vfrag = reinterpret_cast<float4*>(smem_raw + sm.off_vfrag);
  float* sd = reinterpret_cast<float*>(smem_raw + sm.off_sd);
  float* tile = reinterpret_cast<float*>(smem_raw + sm.off_tile);

  EdgeRing ring;
  ring.stage[0] = reinterpret_cast<float*>(smem_raw + sm.off_ring);
  ring.stage[1] = reinterpret_cast<float*>(smem_raw + sm.off_ring + sm.ring_stage_bytes);
  ring.full = bars;
  ring.uses[0] = ring.uses[1] = 0;
  ring.chunk_rows = sm.chunk_rows;
  ring.nchunks = p.Fe > 0 ? (p.R + sm.chunk_rows - 1) / sm.chunk_rows : 0;
  ring.p = &p;

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  for (int r = tid; r < p.R; r += kAttnThreads) table_s[r] = p.Fe > 0 ? p.table[r] : -1;
  if (p.Fe > 0) build_vfrag(vfrag, p.v, H, p.Fe, sm.KS, sm.NT, tid, kAttnThreads);
  for (int idx = tid; idx < H * N * NS; idx += kAttnThreads) tile[idx] = 0.f;
  __syncthreads();

  int b = blockIdx.x;
  if (p.bulk_ok && tid == 0 && b < p.B && ring.nchunks > 0) ring.prefetch_first(b);

  const float out_scale = p.concat ? 1.f : 1.f / (float)H;
  const int CP = (C + 1) / 2;
  const int n_items = p.concat ? H * CP : CP;
  const int h_loop = p.concat ? 1 : H;

  for (; b < p.B; b += gridDim.x) {
    // s_j, d_i of this graph: the 2H augmented columns of P_aug
    for (int idx = tid; idx < N * 2 * H; idx += kAttnThreads) {
      const int j = idx / (2 * H), k = idx - j * 2 * H;
      sd[idx] = p.P_aug[((size_t)b * N + j) * p.ldp + HC + k];
    }
    if (ring.nchunks > 0) {
      edge_logit_phase(ring, p, sm, tile, table_s, vfrag, b, tid);   // ends with a barrier
    } else {
      for (int idx = tid; idx < H * N * NS; idx += kAttnThreads) tile[idx] = 0.f;
      __syncthreads();
    }
    softmax_phase(p, sm, tile, sd, out_scale,
                  args.alpha_out ? args.alpha_out + (size_t)b * H * N * N : nullptr, nullptr, tid);
    __syncthreads();
    // the ring is idle from here on: start the next graph's first chunks under phase 3
    if (p.bulk_ok && tid == 0 && b + (int)gridDim.x < p.B && ring.nchunks > 0)
      ring.prefetch_first(b + gridDim.x);

    // ---- phase 3: aggregation -------------------------------------------------------------
    constexpr int JU = 5;
    for (int item = tid; item < n_items; item += kAttnThreads) {
      const int h0 = p.concat ? item / CP : 0;
      const int cp = p.concat ? item - h0 * CP : item;
      const int c0 = 2 * cp;
      const bool has1 = c0 + 1 < C;
      float2 acc[NPAIRS][2];
#pragma unroll
      for (int ip = 0; ip < NPAIRS; ++ip) acc[ip][0] = acc[ip][1] = make_float2(0.f, 0.f);

      const int total = h_loop * N;                     // flattened (h, j)
      const float* prow = p.P_aug + (size_t)b * N * p.ldp + (size_t)h0 * C + c0;
      const float* arow = tile + (size_t)h0 * N * NS;
      // running load cursor (one (h,j) row ahead of the math by JU)
      int lj = 0;
      const float* lptr = prow;
      auto load_next = [&](bool valid) -> float2 {
        float2 v = make_float2(0.f, 0.f);
        if (valid) {
          if (p.vec2_ok) {
            v = ldg_stream2(lptr);
          } else {
            v.x = __ldg(lptr);
            if (has1) v.y = __ldg(lptr + 1);
          }
          ++lj;
          lptr += p.ldp;
          if (lj == N) { lj = 0; lptr += (ptrdiff_t)C - (ptrdiff_t)N * p.ldp; }
        }
        return v;
      };
      float2 cur[JU], nxt[JU];
#pragma unroll
      for (int u = 0; u < JU; ++u) cur[u] = load_next(u < total);
      for (int base = 0; base < total; base += JU) {
#pragma unroll
        for (int u = 0; u < JU; ++u) nxt[u] = load_next(base + JU + u < total);
#pragma unroll
        for (int u = 0; u < JU; ++u) {
          if (base + u < total) {
            const float* ar = arow + (size_t)(base + u) * NS;
            const float2 px = make_float2(cur[u].x, cur[u].x);
            const float2 py = make_float2(cur[u].y, cur[u].y);
#pragma unroll
            for (int q = 0; q < NPAIRS / 2; ++q) {
              const float4 a4 = *reinterpret_cast<const float4*>(ar + 4 * q);
              const float2 a0 = make_float2(a4.x, a4.y), a1 = make_float2(a4.z, a4.w);
              acc[2 * q][0] = ffma2(a0, px, acc[2 * q][0]);
              acc[2 * q][1] = ffma2(a0, py, acc[2 * q][1]);
              acc[2 * q + 1][0] = ffma2(a1, px, acc[2 * q + 1][0]);
              acc[2 * q + 1][1] = ffma2(a1, py, acc[2 * q + 1][1]);
            }
            if (NPAIRS & 1) {
              const float2 a0 = *reinterpret_cast<const float2*>(ar + 2 * (NPAIRS - 1));
              acc[NPAIRS - 1][0] = ffma2(a0, px, acc[NPAIRS - 1][0]);
              acc[NPAIRS - 1][1] = ffma2(a0, py, acc[NPAIRS - 1][1]);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < JU; ++u) cur[u] = nxt[u];
      }
      // epilogue: + bias, store rows 2ip and 2ip+1
      const int col = h0 * C + c0;
      const float b0 = args.bias ? args.bias[col] : 0.f;
      const float b1 = (args.bias && has1) ? args.bias[col + 1] : 0.f;
      float* orow = args.out + (size_t)b * N * p.ldo + col;
#pragma unroll
      for (int ip = 0; ip < NPAIRS; ++ip) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int i = 2 * ip + half;
          if (i < N) {
            const float o0 = (half ? acc[ip][0].y : acc[ip][0].x) + b0;
            const float o1 = (half ? acc[ip][1].y : acc[ip][1].x) + b1;
            float* dst = orow + (size_t)i * p.ldo;
            if (p.vec2_ok) {
              *reinterpret_cast<float2*>(dst) = make_float2(o0, o1);
            } else {
              dst[0] = o0;
              if (has1) dst[1] = o1;
            }
          }
        }
      }
    }
    __syncthreads();   // tile and sd are rewritten by the next graph
  }
}

template <int NPAIRS>
static int launch_fwd(const AttnFwdArgs& a, cudaStream_t st) {
  const AttnParams& p = a.p;
  AttnSmem sm = attn_smem_plan(p.N, p.Fe, p.H, p.R, NPAIRS, kFwdChunkRows);
  for (int rows = kFwdChunkRows - 16; rows >= 16 && sm.base_total + 2 * sm.ring_stage_bytes > 113 * 1024; rows -= 16)
    sm = attn_smem_plan(p.N, p.Fe, p.H, p.R, NPAIRS, rows);
  const size_t smem = sm.base_total + 2 * sm.ring_stage_bytes;
  if (smem > 227 * 1024)
    return fail(SPOTV2_ERR_UNSUPPORTED, "attn_fwd needs %zu B shared memory (> 227 KB)", smem);
  auto kern = gat_attn_fwd_kernel<NPAIRS>;
  SPOTV2_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = 2 * sm_count();
  if (grid > p.B) grid = p.B;
  kern<<<grid, kAttnThreads, smem, st>>>(a, sm);
  SPOTV2_CUDA_OK(cudaGetLastError());
  return SPOTV2_OK;
}

int attn_fwd_dispatch(const AttnFwdArgs& a, cudaStream_t st) {
  const int np = (a.p.N + 1) / 2;
  if (np <= 4) return launch_fwd<4>(a, st);
  if (np <= 8) return launch_fwd<8>(a, st);
  if (np <= 15) return launch_fwd<15>(a, st);
  if (np <= 16) return launch_fwd<16>(a, st);
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
