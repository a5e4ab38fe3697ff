// Probe (GPU box): does a TMA tensor LOAD accept an inner coordinate that is not 16-byte aligned?
// fp16 matrix [64 x 1024]; a 32-row x 32-column box (64-byte rows, SWIZZLE_64B) is loaded from column c0, row 5.
// nvcc -gencode arch=compute_100a,code=sm_100a -I spotv2net_b200/csrc -o /tmp/tma_load_probe tools/ubench/tma_load_probe.cu spotv2net_b200/csrc/api.cu -lcuda
#include <cuda_fp16.h>
#include <vector>
#include "tma.cuh"
using namespace spotv2;
__global__ void k(const __grid_constant__ CUtensorMap tm, int c0, int r0, __half* out) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, 32 * 64);
    tma_load_2d(sm, &tm, c0, r0, &bar);
  }
  mbar_wait(&bar, 0);
  for (int idx = threadIdx.x; idx < 32 * 32; idx += blockDim.x) {
    const int r = idx / 32, c = idx % 32;
    // SWIZZLE_64B: 16-byte chunk index (2 bits) XOR bits [7:8) of the byte address >> 7
    const uint32_t off = r * 64 + ((((c >> 3) ^ ((r >> 1) & 3)) << 4) | ((c & 7) << 1));
    out[idx] = *reinterpret_cast<__half*>(sm + off);
  }
}
int main() {
  const int R = 64, Cc = 1024;
  std::vector<__half> h(R * Cc);
  for (int i = 0; i < R * Cc; ++i) h[i] = __float2half((float)(i % 2039));
  __half *d, *o; cudaMalloc(&d, R * Cc * 2); cudaMalloc(&o, 32 * 32 * 2);
  cudaMemcpy(d, h.data(), R * Cc * 2, cudaMemcpyHostToDevice);
  int c0s[] = {512, 500, 1001, 996, 4};
  for (int c0 : c0s) {
    CUtensorMap tm;
    if (make_tmap_typed(&tm, d, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, R, Cc, Cc, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B)) { printf("encode failed: %s\n", spotv2_last_error()); return 1; }
    cudaMemset(o, 0, 32 * 32 * 2);
    k<<<1, 128, 4096 + 1024>>>(tm, c0, 5, o);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<__half> got(32 * 32);
    cudaMemcpy(got.data(), o, 32 * 32 * 2, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r = 0; r < 32; ++r) for (int c = 0; c < 32; ++c) {
      const float want = (c0 + c < Cc) ? (float)(((r + 5) * Cc + c0 + c) % 2039) : 0.f;
      if (__half2float(got[r * 32 + c]) != want) ++bad;
    }
    printf("load c0=%d (byte offset %% 16 = %d): %s, wrong %d of 1024\n", c0, (c0 * 2) % 16, cudaGetErrorString(e), bad);
  }
  return 0;
}
