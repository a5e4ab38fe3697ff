// Shared helpers for libspotv2_gat.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/spotv2_gat.h"

namespace spotv2 {

// Thread-local error text behind spotv2_last_error().
void set_error(const char* fmt, ...);
int fail(spotv2_status st, const char* fmt, ...);

#define SPOTV2_CUDA_OK(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess)                                                                \
      return ::spotv2::fail(SPOTV2_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,              \
                            cudaGetErrorString(_e), __FILE__, __LINE__);                  \
  } while (0)

#define SPOTV2_REQUIRE(cond, ...)                                                         \
  do {                                                                                    \
    if (!(cond)) return ::spotv2::fail(SPOTV2_ERR_INVALID_ARG, __VA_ARGS__);              \
  } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline size_t round_up(size_t x, size_t m) { return (x + m - 1) / m * m; }

int check_desc(const spotv2_gat_desc* d);
// p_format 1 pads every head's channel block to a multiple of 8 columns (16-byte aligned TMA tile starts for fp16 planes)
inline int head_pitch_of(const spotv2_gat_desc* d) { return d->p_format == 1 ? (d->C + 7) / 8 * 8 : d->C; }
inline int n_aug_of(const spotv2_gat_desc* d) { return d->H * head_pitch_of(d) + 2 * d->H; }
int sm_count();
size_t attn_bwd_ws_bytes(const spotv2_gat_desc* d);      // attn_bwd.cu

// ---- device helpers ---------------------------------------------------------------------
// Packed fp32 pair FMA (Blackwell FFMA2): d = a * b + c on both halves.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra, rb, rc, rd;
  ra = *reinterpret_cast<unsigned long long*>(&a);
  rb = *reinterpret_cast<unsigned long long*>(&b);
  rc = *reinterpret_cast<unsigned long long*>(&c);
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}

__device__ __forceinline__ uint32_t tf32_rna(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return u;
}

// Split an fp32 value into a tf32 "hi" and a tf32 "lo" with hi + lo == x to ~2^-22.
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = tf32_rna(x);
  lo = tf32_rna(x - __uint_as_float(hi));
}

// Cheap split for hot loops (cvt.rna.tf32 is emulated with ~5 ALU ops on sm_100): hi = x truncated to
// tf32, lo = (x - hi) truncated to tf32; x - hi is exact, so hi + lo == x to 2^-21 |x|.
__device__ __forceinline__ void split_tf32_trunc(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi)) & 0xffffe000u;
}

// D(16x8, f32) += A(16x8, tf32, row) * B(8x8, tf32, col)
__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], const uint32_t (&a)[4],
                                                const uint32_t (&b)[2]) {
  asm(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
      "{%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// Explicit shared-memory load: pointers that reach a helper through arrays or structs lose their
// address space and compile to generic LD, which is several times slower than LDS (measured: the edge
// logit loop ran at ~290 cycles per k-step with generic loads).
__device__ __forceinline__ float lds_f32(const float* p) {
  float v;
  asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(smem_u32(p)));      // not volatile: free to be hoisted
  return v;
}

// ---- mbarrier + 1-D bulk async copy (TMA engine, no tensor map) ---------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait.  try_wait parks the thread for a short hardware-defined interval; after a few misses we
// back off with short sleeps so parked warps do not eat issue slots (a suspend-time HINT compiles to a
// 200 us NANOSLEEP and serialises the pipeline - measured).  A lost transaction traps after ~2 s
// instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (int spin = 0; spin < 16; ++spin)
    if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(64);
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
// Diagnostics: cycles spent parked on each class of barrier (one sampling thread per role adds its
// totals at kernel exit); read and reset through spotv2_diag_counters().
enum { kCntRingFull = 0, kCntTileEmpty, kCntTileFull, kCntPtileFull, kCntPtileEmpty, kCntRoleA, kCntRoleB, kCntRoleP,
       kCntAux0, kCntAux1, kCntAux2, kCntAux3, kNumCounters = 16 };
__device__ __forceinline__ void mbar_wait_timed(uint64_t* bar, uint32_t parity, long long& acc) {
  const long long t0 = clock64();
  mbar_wait(bar, parity);
  acc += clock64() - t0;
}

__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// L2 eviction-priority descriptors for the .L2::cache_hint forms of the bulk copies (the encodings CUTLASS
// ships as TMA::CacheHintSm90)
constexpr uint64_t kEvictNormal = 0x1000000000000000ull, kEvictFirst = 0x12F0000000000000ull, kEvictLast = 0x14F0000000000000ull;
__device__ __forceinline__ void bulk_g2s_hint(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar,
                                              uint64_t hint) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
          "r"(smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(hint)
      : "memory");
}

__device__ __forceinline__ float2 ldg_stream2(const float* p) {
  float2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

}  // namespace spotv2
