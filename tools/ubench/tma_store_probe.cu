// Probe (GPU box): does a TMA tensor STORE accept an inner coordinate that is not 16-byte aligned?
// fp16 matrix [64 x 1024]; a 32-row x 64-column box (128-byte rows, SWIZZLE_128B) is stored at column c0, row 5.
// nvcc -gencode arch=compute_100a,code=sm_100a -I spotv2net_b200/csrc -o /tmp/tma_store_probe tools/ubench/tma_store_probe.cu ../../spotv2net_b200/csrc/api.cu
#include <cuda_fp16.h>
#include <vector>
#include "tma.cuh"
using namespace spotv2;
__global__ void k(const __grid_constant__ CUtensorMap tm, int c0, int r0) {
  extern __shared__ __align__(1024) unsigned char sm[];
  for (int idx = threadIdx.x; idx < 32 * 64; idx += blockDim.x) {
    const int r = idx / 64, c = idx % 64;
    const uint32_t off = r * 128 + ((((c >> 3) ^ (r & 7)) << 4) | ((c & 7) << 1));
    *reinterpret_cast<__half*>(sm + off) = __float2half((float)(r * 64 + c));
  }
  fence_proxy_async();
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tm), "r"(smem_u32(sm)), "r"(c0), "r"(r0) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}
int main() {
  const int R = 64, Cc = 1024;
  __half* d; cudaMalloc(&d, R * Cc * 2);
  int c0s[] = {512, 500, 1001, 996};
  for (int c0 : c0s) {
    cudaMemset(d, 0, R * Cc * 2);
    CUtensorMap tm;
    if (make_tmap_typed(&tm, d, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, R, Cc, Cc, 64, 32, CU_TENSOR_MAP_SWIZZLE_128B)) { printf("encode failed: %s\n", spotv2_last_error()); return 1; }
    k<<<1, 128, 4096 + 1024>>>(tm, c0, 5);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<__half> h(R * Cc);
    cudaMemcpy(h.data(), d, R * Cc * 2, cudaMemcpyDeviceToHost);
    int bad = 0, stray = 0;
    for (int r = 0; r < R; ++r) for (int c = 0; c < Cc; ++c) {
      const float v = __half2float(h[r * Cc + c]);
      const bool in = r >= 5 && r < 37 && c >= c0 && c < c0 + 64;
      const float want = in ? (float)((r - 5) * 64 + (c - c0)) : 0.f;
      if (v != want) { if (in) ++bad; else ++stray; }
    }
    printf("c0=%d (byte offset %% 16 = %d): %s, wrong inside %d, stray outside %d (columns past 1024 must be clipped)\n", c0, (c0 * 2) % 16,
           cudaGetErrorString(e), bad, stray);
  }
  return 0;
}
