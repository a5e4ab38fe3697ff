// Probe (GPU box): is the TMA engine's cost per box ROW (a 64-byte row costs as much as a 128-byte one) what bounds the
// dP stores / P loads of the attention backward?  One CTA per SM, 12 warps; each warp streams the boxes of its (head, 16-row
// tile) of every graph of the CTA, either as 32-column boxes (64-byte rows) or 64-column boxes (128-byte rows); the two
// fp16 planes in one instruction (4-D map: column-in-head, row, head, plane), as csrc/attn_bwd2.cu does.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I spotv2net_b200/csrc -o /tmp/tma_row_rate tools/ubench/tma_row_rate_probe.cu spotv2net_b200/csrc/api.cu -lcuda
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include "tma.cuh"
using namespace spotv2;

static int make_map(CUtensorMap* tm, void* base, uint64_t plane_stride, uint64_t rows, uint64_t Cp, uint64_t H, uint64_t ld, uint32_t box_cols,
                    uint32_t box_rows, uint32_t box_heads) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return 1;
  cuuint64_t dims[4] = {Cp, rows, H, 2};
  cuuint64_t strides[3] = {ld * 2, Cp * 2, plane_stride * 2};
  cuuint32_t box[4] = {box_cols, box_rows, box_heads, 2};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  box_cols == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 2;
}

// mode 0: stores, one box per (warp, column block);  mode 1: loads, one box of all heads per column block (one thread issues)
// shift_odd: the kernel's rule for heads that start 16 bytes into a sector - boxes after the first start 8 columns early
__global__ void __launch_bounds__(384, 1) k(const __grid_constant__ CUtensorMap tm, int B, int N, int n_cb, int box_cols, int mode, int depth, int shift_odd) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ uint64_t bar[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (mode == 0) {
    const int h = warp >> 1, m = warp & 1;
    const uint32_t stg = smem_u32(sm) + warp * 2 * (uint32_t)(box_cols * 2 * 16 * 2);      // two staging pieces per warp
    const uint32_t piece = (uint32_t)(box_cols * 2 * 16 * 2);
    int k_ = 0;
    for (int b = blockIdx.x; b < B; b += gridDim.x)
      for (int cb = 0; cb < n_cb; ++cb, ++k_) {
        if (lane == 0) {
          if (depth == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(&tm),
                       "r"(stg + (depth == 1 ? 0u : (uint32_t)(k_ & 1) * piece)), "r"(cb * box_cols - ((shift_odd && (h & 1) && cb > 0) ? 8 : 0)),
                       "r"(b * N + 16 * m), "r"(h), "r"(0) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        __syncwarp();
      }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  } else {
    // loads: ring of `depth` slots, each one box (box_cols x 32 rows x 6 heads x 2 planes); thread 0 issues, waits in order
    if (threadIdx.x == 0) { for (int s = 0; s < 8; ++s) mbar_init(&bar[s], 1); fence_mbar_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
      const uint32_t bytes = (uint32_t)box_cols * 2u * 32u * 6u * 2u;
      int issued = 0, done = 0;
      const int my = (B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
      const int total = my * n_cb;
      while (done < total) {
        while (issued < total && issued - done < depth) {
          const int it = issued / n_cb, cb = issued - it * n_cb, b = blockIdx.x + it * gridDim.x, s = issued % depth;
          mbar_expect_tx(&bar[s], bytes);
          tma_load_4d(sm + (size_t)s * bytes, &tm, cb * box_cols, b * N, 0, 0, &bar[s]);
          ++issued;
        }
        mbar_wait(&bar[done % depth], (done / depth) & 1);
        ++done;
      }
    }
  }
}

int main(int argc, char** argv) {
  const int B = 4096, N = 30, H = 6;
  const int Cp = argc > 1 ? atoi(argv[1]) : 504, ld = argc > 2 ? atoi(argv[2]) : 3040;
  const uint64_t rows = (uint64_t)B * N;
  __half* d;
  const size_t plane = rows * ld;
  cudaMalloc(&d, plane * 2 * sizeof(__half));
  cudaMemset(d, 0, plane * 2 * sizeof(__half));
  {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) { cudaEventRecord(e0); cudaMemsetAsync(d, 0, plane * 2 * sizeof(__half)); cudaEventRecord(e1); cudaDeviceSynchronize(); }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("Cp %d ld %d; cudaMemset of both planes (%.2f GB): %.3f ms, %.0f GB/s\n", Cp, ld, plane * 4 * 1e-9, ms, plane * 4 / ms * 1e-6);
  }
  for (int mode = 0; mode < 3; ++mode)
    for (int box_cols = 32; box_cols <= (mode == 2 ? 32 : 64); box_cols *= 2)
      for (int depth = 1; depth <= (mode == 1 ? 4 : 2); depth *= 2) {
        const int shift_odd = mode == 2;
        const int kmode = mode == 2 ? 0 : mode;
        CUtensorMap tm;
        if (make_map(&tm, d, plane, rows, Cp, H, ld, box_cols, kmode ? 32 : 16, kmode ? 6 : 1)) { printf("encode failed\n"); return 1; }
        const int n_cb = (Cp + box_cols - 1) / box_cols;
        const size_t smem = kmode ? (size_t)depth * box_cols * 2 * 32 * 6 * 2 + 1024 : (size_t)12 * 2 * box_cols * 2 * 16 * 2 + 1024;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int rep = 0; rep < 3; ++rep) {
          cudaEventRecord(e0);
          k<<<148, 384, smem>>>(tm, B, N, n_cb, box_cols, kmode, depth, shift_odd);
          cudaEventRecord(e1);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double bytes = (double)rows * H * Cp * 2 * 2;
        printf("%s box %2d cols (%3d-byte rows) depth %d: %.3f ms, %.0f GB/s, %.2f cycles per box row per SM at 1.9 GHz\n", mode == 1 ? "load " : (mode == 2 ? "store (odd heads shifted by 8 columns)" : "store"),
               box_cols, box_cols * 2, depth, ms, bytes / ms * 1e-6, ms * 1e-3 * 1.9e9 / ((double)rows * H * 2 * n_cb / 148.0));
      }
  return 0;
}
