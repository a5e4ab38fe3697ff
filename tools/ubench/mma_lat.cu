// Microbenchmark (GPU box): latency of a dependent mma.sync chain and throughput with independent chains,
// for m16n8k8 tf32 and m16n8k16 f16 on sm_100a.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_lat mma_lat.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

template <int CH, bool F16>
__global__ void k(float* out, int iters, long long* cyc) {
  float c[CH][4];
  for (int i = 0; i < CH; ++i) for (int q = 0; q < 4; ++q) c[i][q] = 0.f;
  uint32_t a[4] = {0x3f800000u + threadIdx.x, 0x3f800000u, 0x3f800000u, 0x3f800000u};
  uint32_t b[2] = {0x3f800000u, 0x3f800000u + threadIdx.x};
  if (F16) { a[0] = 0x3c003c00u; a[1] = a[0]; a[2] = a[0]; a[3] = a[0]; b[0] = a[0]; b[1] = a[0]; }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      if (F16)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
      else
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
  }
  long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < CH; ++i) for (int q = 0; q < 4; ++q) s += c[i][q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int CH, bool F16>
void run(int warps) {
  float* out; long long* cyc; long long h;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  k<CH, F16><<<148, warps * 32>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  k<CH, F16><<<148, warps * 32>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  double per = (double)h / iters / CH;
  printf("%s chains=%2d warps/SM=%2d: %.1f cycles per MMA per warp -> %.2f cycles per MMA per SM (%s)\n", F16 ? "f16 m16n8k16" : "tf32 m16n8k8",
         CH, warps, per, per / warps, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<1, false>(1); run<2, false>(1); run<4, false>(1); run<8, false>(1); run<16, false>(1);
  run<8, false>(4); run<8, false>(8); run<8, false>(12); run<16, false>(12); run<8, false>(16);
  run<1, true>(1); run<4, true>(1); run<8, true>(1); run<16, true>(1);
  run<8, true>(4); run<8, true>(8); run<8, true>(12); run<16, true>(12);
  return 0;
}
