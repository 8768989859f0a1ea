// Microbenchmark: how fast can one SM-resident CTA stream bytes from HBM into shared memory / registers on B200,
// as a function of the mechanism and the size of one request?  Guides the decode GEMV's data path.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_stream tools/microbench_stream.cu
//   ./tools/microbench_stream
//
// Modes:
//   bulk   : cp.async.bulk (1-D) of `sz` bytes per op, `lanes` lanes of one producer warp issue ops in parallel,
//            ring of `depth` slots per CTA, a second warp waits on the full barrier and frees the slot at once
//   ldg    : every thread keeps `unroll` 16-byte ld.global.nc in flight (registers), 256 threads per CTA
//   ldgsts : cp.async 16 B per thread into smem, commit/wait groups, `unroll` groups in flight
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// each CTA streams [cta * per_cta, (cta+1) * per_cta) of `src`
__global__ void __launch_bounds__(64) k_bulk(const uint8_t* src, size_t per_cta, int sz, int lanes, int depth, int slot_bytes, unsigned* sink) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);   // full[64], empty[64]
  uint8_t* ring = smem + 1024;
  const uint32_t full = smem_u32(bars), empty = smem_u32(bars + 64), ring_u = smem_u32(ring);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < depth; ++i) { mbar_init(full + 8 * i, 1); mbar_init(empty + 8 * i, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint8_t* base = src + (size_t)blockIdx.x * per_cta;
  const int per_stage = sz * lanes;
  const int nstages = (int)(per_cta / per_stage);
  if (warp == 0) {
    for (int i = 0; i < nstages; ++i) {
      const int slot = i % depth, use = i / depth;
      if (use > 0) mbar_wait(empty + 8 * slot, (use - 1) & 1);
      if (lane == 0) mbar_expect_tx(full + 8 * slot, per_stage);
      __syncwarp();
      if (lane < lanes) bulk_g2s(ring_u + slot * slot_bytes + lane * sz, base + (size_t)i * per_stage + (size_t)lane * sz, sz, full + 8 * slot);
    }
  } else {
    unsigned acc = 0;
    for (int i = 0; i < nstages; ++i) {
      const int slot = i % depth, use = i / depth;
      mbar_wait(full + 8 * slot, use & 1);
      acc += *reinterpret_cast<volatile unsigned*>(ring + slot * slot_bytes + lane * 4);
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + 8 * slot);
    }
    if (acc == 0x12345678u) sink[0] = acc;
  }
}

template <int UNROLL>
__global__ void __launch_bounds__(256) k_ldg(const uint8_t* src, size_t per_cta, unsigned* sink) {
  const uint4* base = reinterpret_cast<const uint4*>(src + (size_t)blockIdx.x * per_cta);
  const int n = (int)(per_cta / 16);
  unsigned acc = 0;
  for (int i = threadIdx.x; i + (UNROLL - 1) * 256 < n; i += UNROLL * 256) {
    uint4 v[UNROLL];
#pragma unroll
    for (int j = 0; j < UNROLL; ++j)
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[j].x), "=r"(v[j].y), "=r"(v[j].z), "=r"(v[j].w) : "l"(base + i + j * 256));
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) acc += v[j].x ^ v[j].y ^ v[j].z ^ v[j].w;
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

template <int GROUPS>
__global__ void __launch_bounds__(256) k_ldgsts(const uint8_t* src, size_t per_cta, unsigned* sink) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint8_t* base = src + (size_t)blockIdx.x * per_cta;
  const int n = (int)(per_cta / (16 * 256));     // groups of 4 KB per CTA
  const uint32_t dst0 = smem_u32(smem) + threadIdx.x * 16;
  unsigned acc = 0;
  for (int i = 0; i < GROUPS - 1 && i < n; ++i) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + (i % GROUPS) * 4096), "l"(base + (size_t)i * 4096 + threadIdx.x * 16) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int i = 0; i < n; ++i) {
    const int j = i + GROUPS - 1;
    if (j < n)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + (j % GROUPS) * 4096), "l"(base + (size_t)j * 4096 + threadIdx.x * 16) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group %0;" ::"n"(GROUPS - 1) : "memory");
    acc += *reinterpret_cast<volatile unsigned*>(smem + (i % GROUPS) * 4096 + threadIdx.x * 16);
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const size_t per_cta = 8u << 20;              // 8 MiB per CTA -> 1.16 GiB total, far beyond L2
  const size_t total = per_cta * sms;
  uint8_t* buf;
  unsigned* sink;
  CK(cudaMalloc(&buf, total));
  CK(cudaMalloc(&sink, 4));
  CK(cudaMemset(buf, 1, total));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  auto report = [&](const char* name, int a, int b, int c, float ms) {
    printf("{\"mode\": \"%s\", \"p0\": %d, \"p1\": %d, \"p2\": %d, \"ms\": %.4f, \"GBps\": %.1f}\n", name, a, b, c, ms, total / ms / 1e6);
    fflush(stdout);
  };
  CK(cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  // bulk: size per op x lanes x ring bytes
  const int sizes[] = {256, 512, 1024, 2048, 4096, 8192, 16384};
  for (int ctas_per_sm = 1; ctas_per_sm <= 2; ++ctas_per_sm)
    for (int sz : sizes)
      for (int lanes : {1, 4, 8}) {
        const int slot_bytes = sz * lanes;
        if (slot_bytes > 65536) continue;
        int depth = (96 * 1024 / ctas_per_sm) / slot_bytes;
        if (depth > 64) depth = 64;
        if (depth < 2) continue;
        const size_t smem = 1024 + (size_t)depth * slot_bytes;
        for (int rep = 0; rep < 2; ++rep) {
          CK(cudaEventRecord(e0));
          k_bulk<<<sms * ctas_per_sm, 64, smem>>>(buf, per_cta / ctas_per_sm, sz, lanes, depth, slot_bytes, sink);
          CK(cudaEventRecord(e1));
          CK(cudaEventSynchronize(e1));
          float ms;
          CK(cudaEventElapsedTime(&ms, e0, e1));
          if (rep == 1) report(ctas_per_sm == 1 ? "bulk_1cta" : "bulk_2cta", sz, lanes, depth, ms);
        }
      }
  // ldg
  for (int bps : {1, 2, 4}) {
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaEventRecord(e0));
      k_ldg<4><<<sms * bps, 256>>>(buf, per_cta / bps, sink);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      if (rep == 1) report("ldg_unroll4", bps, 256, 4, ms);
    }
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaEventRecord(e0));
      k_ldg<8><<<sms * bps, 256>>>(buf, per_cta / bps, sink);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      if (rep == 1) report("ldg_unroll8", bps, 256, 8, ms);
    }
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaEventRecord(e0));
      k_ldg<16><<<sms * bps, 256>>>(buf, per_cta / bps, sink);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      if (rep == 1) report("ldg_unroll16", bps, 256, 16, ms);
    }
  }
  CK(cudaFuncSetAttribute(k_ldgsts<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  CK(cudaFuncSetAttribute(k_ldgsts<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  for (int bps : {1, 2}) {
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaEventRecord(e0));
      k_ldgsts<8><<<sms * bps, 256, 8 * 4096>>>(buf, per_cta / bps, sink);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      if (rep == 1) report("ldgsts_g8", bps, 256, 8, ms);
    }
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaEventRecord(e0));
      k_ldgsts<16><<<sms * bps, 256, 16 * 4096>>>(buf, per_cta / bps, sink);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      if (rep == 1) report("ldgsts_g16", bps, 256, 16, ms);
    }
  }
  return 0;
}
