// Microbenchmark: throughput / latency of the legacy warp-level mma.sync.m16n8k16 (HMMA) and of lop3 on B200,
// per SM sub-partition, as a function of warps per SM and independent accumulator chains per warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/microbench_hmma tools/microbench_hmma.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void mma(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_f16acc(uint32_t (&d)[2], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f16.f16.f16.f16 {%0,%1}, {%2,%3,%4,%5}, {%6,%7}, {%0,%1};"
               : "+r"(d[0]), "+r"(d[1]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int CHAINS, bool F16ACC>
__global__ void k_hmma(int iters, float* out, long long* cycles) {
  float acc[CHAINS][4];
  uint32_t hacc[CHAINS][2];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) { acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0.f; hacc[c][0] = hacc[c][1] = 0; }
  uint32_t a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 0x3c003c00, a3 = 0x3c003c00, b0 = 0x3c003c00, b1 = 0x38003800;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
      if (F16ACC) mma_f16acc(hacc[c], a0, a1, a2, a3, b0, b1);
      else mma(acc[c], a0, a1, a2, a3, b0, b1);
    }
  }
  const long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += acc[c][0] + acc[c][1] + acc[c][2] + acc[c][3] + __uint_as_float(hacc[c][0]) + __uint_as_float(hacc[c][1]);
  if (s == 123.456f) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int CHAINS>
__global__ void k_imma(int iters, float* out, long long* cycles) {
  int acc[CHAINS][4];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0;
  uint32_t a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 0x01010101, a3 = 0x02020202, b0 = 0x01020304, b1 = 0x01010101;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c)
      asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+r"(acc[c][0]), "+r"(acc[c][1]), "+r"(acc[c][2]), "+r"(acc[c][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  const long long t1 = clock64();
  int s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += acc[c][0] + acc[c][1] + acc[c][2] + acc[c][3];
  if (s == 123456789) out[0] = (float)s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int CHAINS>
__global__ void k_lop3(int iters, uint32_t* out, long long* cycles) {
  uint32_t v[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) v[c] = threadIdx.x + c;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) asm volatile("lop3.b32 %0, %0, %1, %2, 0xea;" : "+r"(v[c]) : "r"(0x000f000f + i), "r"(0x64006400));
  }
  const long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s ^= v[c];
  if (s == 0x12345) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

int main() {
  float* out; long long* cyc;
  CK(cudaMalloc(&out, 16)); CK(cudaMalloc(&cyc, 8));
  const int iters = 4096;
  auto run = [&](const char* name, int chains, int warps, auto kern) {
    kern<<<148, warps * 32>>>(iters, out, cyc);
    CK(cudaDeviceSynchronize());
    kern<<<148, warps * 32>>>(iters, out, cyc);
    CK(cudaDeviceSynchronize());
    long long c; CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
    const double per_warp = (double)c / ((double)iters * chains);               // cycles per op seen by one warp
    const double per_smsp = per_warp / ((warps + 3) / 4);                       // cycles per op per sub-partition
    printf("{\"op\": \"%s\", \"chains\": %d, \"warps_per_sm\": %d, \"cycles_per_op_per_warp\": %.2f, \"cycles_per_op_per_smsp\": %.2f}\n",
           name, chains, warps, per_warp, per_smsp);
  };
  for (int warps : {1, 4, 8, 16, 32}) {
    run("hmma_f32acc", 1, warps, k_hmma<1, false>);
    run("hmma_f32acc", 2, warps, k_hmma<2, false>);
    run("hmma_f32acc", 4, warps, k_hmma<4, false>);
    run("hmma_f32acc", 8, warps, k_hmma<8, false>);
    run("hmma_f16acc", 1, warps, k_hmma<1, true>);
    run("hmma_f16acc", 4, warps, k_hmma<4, true>);
    run("hmma_f16acc", 8, warps, k_hmma<8, true>);
  }
  for (int warps : {1, 8, 16}) {
    run("imma_u8s8_k32", 1, warps, k_imma<1>);
    run("imma_u8s8_k32", 2, warps, k_imma<2>);
    run("imma_u8s8_k32", 4, warps, k_imma<4>);
  }
  uint32_t* o2 = reinterpret_cast<uint32_t*>(out);
  auto run2 = [&](int chains, int warps, auto kern) {
    kern<<<148, warps * 32>>>(iters, o2, cyc);
    CK(cudaDeviceSynchronize());
    long long c; CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
    const double per_warp = (double)c / ((double)iters * chains);
    printf("{\"op\": \"lop3\", \"chains\": %d, \"warps_per_sm\": %d, \"cycles_per_op_per_warp\": %.2f, \"cycles_per_op_per_smsp\": %.2f}\n",
           chains, warps, per_warp, per_warp / ((warps + 3) / 4));
  };
  for (int warps : {4, 8, 16}) { run2(1, warps, k_lop3<1>); run2(8, warps, k_lop3<8>); }
  return 0;
}
