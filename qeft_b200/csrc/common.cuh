// Shared device/host helpers for the qeft_b200 kernels (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/qeft_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "qeft_b200 kernels are written for sm_100a (B200) only"
#endif

namespace qeft {

// one counter for every kernel launch made through the C ABI (bench.py: gpu_launches)
extern unsigned long long g_launch_count;
inline void count_launch(int n = 1) { g_launch_count += (unsigned long long)n; }

__host__ __device__ constexpr int cdiv(int a, int b) { return (a + b - 1) / b; }

// ---- global loads ----------------------------------------------------------------------
// streamed-once data (packed weights): bypass L1 (ptxas accepts the L2::evict_first hint only on
// 256-bit loads, so the 128-bit form carries none)
__device__ __forceinline__ uint4 ldg_stream_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint2 ldg_nc_v2(const void* p) {
  uint2 r;
  asm volatile("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ unsigned short ldg_nc_u16(const void* p) {
  unsigned short r;
  asm volatile("ld.global.nc.u16 %0, [%1];" : "=h"(r) : "l"(p));
  return r;
}

// ---- programmatic dependent launch -------------------------------------------------------
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// Column-sharded chains: a launch advances every rank's arrival counter by kArrivalsPerLaunch in total, CTA i of a
// grid of n contributing floor(A (i + 1) / n) - floor(A i / n) (the shares telescope to exactly A for any grid size,
// so a waiting launch needs to know neither the producer's grid nor its kernel: it waits for epoch x ranks x A).
constexpr unsigned kArrivalsPerLaunch = 1u << 16;
__device__ __forceinline__ unsigned arrival_share(unsigned i, unsigned n) {
  return (unsigned)(((unsigned long long)kArrivalsPerLaunch * (i + 1)) / n) - (unsigned)(((unsigned long long)kArrivalsPerLaunch * i) / n);
}

// Release side of the arrival-counter protocol.  __threadfence_system() is a SEQUENTIALLY CONSISTENT fence
// (MEMBAR.SC.SYS); ordering earlier stores before a later relaxed signal needs only acquire-release (MEMBAR.ALL.SYS),
// which does not have to be ordered against every other CTA's fence.
__device__ __forceinline__ void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- int4 -> fp16 ------------------------------------------------------------------------
// One 32-bit word of the packed layout holds 8 nibbles n0..n7.  Returns four half2:
//   h[0] = {n0, n4}  h[1] = {n1, n5}  h[2] = {n2, n6}  h[3] = {n3, n7}    (exact integers 0..15)
// which, for word g of a 16-byte chunk, are the k-pairs (2g,2g+1) (8+2g,9+2g) (16+2g,17+2g) (24+2g,25+2g).
// Same result as the reference's dequantize_s4_to_fp16x2 (kernel/quantization_new/dequantize.cuh:14-77):
// OR the nibble into the mantissa of 1024.0 (0x6400) and subtract the bias.
__device__ __forceinline__ void unpack_word_to_half2(uint32_t w, uint32_t (&h)[4]) {
  constexpr uint32_t kLo = 0x000f000fu, kHi = 0x00f000f0u, kMagic = 0x64006400u;
  const uint32_t t = w >> 8;
  uint32_t a, b, c, d;
  asm("lop3.b32 %0, %1, %2, %3, 0xea;" : "=r"(a) : "r"(w), "n"(kLo), "n"(kMagic));  // (w & lo) | magic
  asm("lop3.b32 %0, %1, %2, %3, 0xea;" : "=r"(b) : "r"(w), "n"(kHi), "n"(kMagic));
  asm("lop3.b32 %0, %1, %2, %3, 0xea;" : "=r"(c) : "r"(t), "n"(kLo), "n"(kMagic));
  asm("lop3.b32 %0, %1, %2, %3, 0xea;" : "=r"(d) : "r"(t), "n"(kHi), "n"(kMagic));
  constexpr uint32_t k1024 = 0x64006400u;   // {1024, 1024}
  constexpr uint32_t kSixteenth = 0x2c002c00u;  // {1/16, 1/16}
  constexpr uint32_t kNeg64 = 0xd400d400u;  // {-64, -64}
  asm("sub.f16x2 %0, %1, %2;" : "=r"(h[0]) : "r"(a), "r"(k1024));
  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(h[1]) : "r"(b), "r"(kSixteenth), "r"(kNeg64));
  asm("sub.f16x2 %0, %1, %2;" : "=r"(h[2]) : "r"(c), "r"(k1024));
  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(h[3]) : "r"(d), "r"(kSixteenth), "r"(kNeg64));
}

__device__ __forceinline__ uint32_t hfma2_u32(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// bf16 flavour of the dequant: {q0, q1} exact fp16 integers, s / sz the group's fp16 scale and scaled zero as fp32.
// w = fma(q, s, sz) in fp32, ONE rounding to bf16 (round to nearest even), packed {lo, hi}.
__device__ __forceinline__ uint32_t dequant_pair_bf16(uint32_t q2, float s, float sz) {
  const __half2 h = *reinterpret_cast<const __half2*>(&q2);
  const float2 q = __half22float2(h);
  const __nv_bfloat162 b = __floats2bfloat162_rn(fmaf(q.x, s, sz), fmaf(q.y, s, sz));
  return *reinterpret_cast<const uint32_t*>(&b);
}

__device__ __forceinline__ float2 half2_bits_to_float2(uint32_t v) {
  __half2 h = *reinterpret_cast<__half2*>(&v);
  return __half22float2(h);
}

// legacy warp-level MMA used as the dot-product engine of the decode GEMV (16 rows x 8 batch columns)
__device__ __forceinline__ void mma_m16n8k16_f16f32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2,
                                                    uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

inline int check_align16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace qeft
