// Device packers / unpackers for the QEFT packed layout (bit-exact with qeft/qlinear.py:70-121 of the
// reference; the layout is restated in SURVEY.md appendix A and include/qeft_b200.h).
//
// All kernels are pure byte/integer movers bound by HBM; one thread handles one 16-byte chunk
// (32 nibbles of one row) so that global accesses on the packed side are 128-bit and coalesced.
#include <type_traits>

#include "common.cuh"

namespace qeft {

// nibble slot n of word w inside a 16-byte chunk  <->  column offset inside the 32-column chunk
//   n = 4*odd + q8,  column = 8*q8 + 2*w + odd            (see oracle/qeft_oracle.py: tile_index_map)
__device__ __forceinline__ int chunk_col(int w, int n) { return 8 * (n & 3) + 2 * w + (n >> 2); }

__global__ void pack_w4_kernel(const int32_t* __restrict__ q, uint8_t* __restrict__ out, int N, int K) {
  // chunk id -> (qrow b, tile T, row-in-tile j, half h); 16 bytes at b*2K + T*128 + j*32 + h*16
  const size_t nchunks = (size_t)N * (size_t)(K >> 5);
  for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < nchunks; c += (size_t)gridDim.x * blockDim.x) {
    const int per_qrow = K >> 3;               // 16-byte chunks per qweight row (2K bytes)
    const int b = (int)(c / (size_t)per_qrow);
    const int rem = (int)(c - (size_t)b * per_qrow);
    const int T = rem >> 3, j = (rem >> 1) & 3, h = rem & 1;
    const int32_t* src = q + (size_t)(4 * b + j) * K + 64 * T + 32 * h;
    uint32_t words[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      // no clamp / no mask, exactly like the reference (qlinear.py:109-114): each int16 is the OR of four
      // shifted int32 values truncated to 16 bits, so an out-of-range value spills inside its int16 only
      uint32_t lo = 0, hi = 0;
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        lo |= (uint32_t)src[chunk_col(w, n)] << (4 * n);
        hi |= (uint32_t)src[chunk_col(w, n + 4)] << (4 * n);
      }
      words[w] = (lo & 0xFFFFu) | (hi << 16);
    }
    *reinterpret_cast<uint4*>(out + (size_t)b * (size_t)(2 * K) + (size_t)T * 128 + j * 32 + h * 16) =
        make_uint4(words[0], words[1], words[2], words[3]);
  }
}

__global__ void unpack_w4_kernel(const uint8_t* __restrict__ qw, int32_t* __restrict__ out, int N, int K) {
  const size_t nchunks = (size_t)N * (size_t)(K >> 5);
  for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < nchunks; c += (size_t)gridDim.x * blockDim.x) {
    const int per_qrow = K >> 3;
    const int b = (int)(c / (size_t)per_qrow);
    const int rem = (int)(c - (size_t)b * per_qrow);
    const int T = rem >> 3, j = (rem >> 1) & 3, h = rem & 1;
    const uint4 v = *reinterpret_cast<const uint4*>(qw + (size_t)b * (size_t)(2 * K) + (size_t)T * 128 + j * 32 + h * 16);
    const uint32_t words[4] = {v.x, v.y, v.z, v.w};
    int32_t* dst = out + (size_t)(4 * b + j) * K + 64 * T + 32 * h;
#pragma unroll
    for (int w = 0; w < 4; ++w)
#pragma unroll
      for (int n = 0; n < 8; ++n) dst[chunk_col(w, n)] = (int32_t)((words[w] >> (4 * n)) & 0xFu);
  }
}

// dense W[N, K] = fma(q, s, sz) (fp16, single rounding) or its bf16 rounding; outlier columns from oweight
template <typename OutT>
__global__ void dequant_w4_kernel(const uint8_t* __restrict__ qw, const __half* __restrict__ scales,
                                  const __half* __restrict__ szeros, const __half* __restrict__ ow,
                                  OutT* __restrict__ out, int N, int K, int r, int G) {
  const size_t nchunks = (size_t)N * (size_t)(K >> 5);
  for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < nchunks; c += (size_t)gridDim.x * blockDim.x) {
    const int per_qrow = K >> 3;
    const int b = (int)(c / (size_t)per_qrow);
    const int rem = (int)(c - (size_t)b * per_qrow);
    const int T = rem >> 3, j = (rem >> 1) & 3, h = rem & 1;
    const int row = 4 * b + j, k0 = 64 * T + 32 * h;
    OutT* dst = out + (size_t)row * K + k0;
    if (ow != nullptr && k0 >= K - r) {
      const __half* src = ow + (size_t)row * r + (k0 - (K - r));
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        if constexpr (sizeof(OutT) == 2 && std::is_same<OutT, __half>::value) dst[i] = src[i];
        else dst[i] = OutT(__half2float(src[i]));
      }
      continue;
    }
    const uint4 v = *reinterpret_cast<const uint4*>(qw + (size_t)b * (size_t)(2 * K) + (size_t)T * 128 + j * 32 + h * 16);
    const uint32_t words[4] = {v.x, v.y, v.z, v.w};
    const int gi = k0 / G;
    const __half s = scales[(size_t)gi * N + row], z = szeros[(size_t)gi * N + row];
    const __half2 s2 = __half2half2(s), z2 = __half2half2(z);
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      uint32_t hh[4];
      unpack_word_to_half2(words[w], hh);
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const __half2 wv = __hfma2(*reinterpret_cast<__half2*>(&hh[p]), s2, z2);
        const int col = 8 * p + 2 * w;
        if constexpr (std::is_same<OutT, __half>::value) {
          *reinterpret_cast<__half2*>(dst + col) = wv;
        } else {
          dst[col] = OutT(__low2float(wv));
          dst[col + 1] = OutT(__high2float(wv));
        }
      }
    }
  }
}

// pack_oweight: out[(n/8)*4 + n%4][64*(j/32) + 2*(j%32) + (n%8)/4] = ow[n][j]
template <typename InT>
__global__ void interleave_oweight_kernel(const InT* __restrict__ ow, __half* __restrict__ out, int N, int r) {
  const size_t total = (size_t)N * r;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    // iterate over OUTPUT elements so stores are coalesced
    const int R = (int)(i / (size_t)(2 * r));
    const int col = (int)(i - (size_t)R * (2 * r));
    const int c = col >> 6, tt = (col & 63) >> 1, s = col & 1;
    const int n = 8 * (R >> 2) + (R & 3) + 4 * s, j = 32 * c + tt;
    float v = (float)ow[(size_t)n * r + j];
    out[i] = __float2half_rn(v);
  }
}

static inline int grid_for(size_t work, int block) {
  size_t g = (work + block - 1) / block;
  const size_t cap = 148 * 16;
  return (int)(g < cap ? (g ? g : 1) : cap);
}

}  // namespace qeft

using namespace qeft;

#define QEFT_CHECK_LAUNCH()                          \
  do {                                               \
    cudaError_t e__ = cudaGetLastError();            \
    if (e__ != cudaSuccess) return (int)e__;         \
    count_launch();                                  \
  } while (0)

extern "C" int qeft_pack_w4(const int32_t* intweight, void* qweight, int N, int K, qeft_stream_t stream) {
  if (!intweight || !qweight) return QEFT_E_NULL;
  if (N <= 0 || K <= 0 || N % 4 != 0 || K % 64 != 0) return QEFT_E_SHAPE;
  if (!check_align16(qweight)) return QEFT_E_ALIGN;
  const size_t chunks = (size_t)N * (K >> 5);
  pack_w4_kernel<<<grid_for(chunks, 256), 256, 0, (cudaStream_t)stream>>>(intweight, (uint8_t*)qweight, N, K);
  QEFT_CHECK_LAUNCH();
  return QEFT_OK;
}

extern "C" int qeft_unpack_w4(const void* qweight, int32_t* intweight, int N, int K, qeft_stream_t stream) {
  if (!intweight || !qweight) return QEFT_E_NULL;
  if (N <= 0 || K <= 0 || N % 4 != 0 || K % 64 != 0) return QEFT_E_SHAPE;
  if (!check_align16(qweight)) return QEFT_E_ALIGN;
  const size_t chunks = (size_t)N * (K >> 5);
  unpack_w4_kernel<<<grid_for(chunks, 256), 256, 0, (cudaStream_t)stream>>>((const uint8_t*)qweight, intweight, N, K);
  QEFT_CHECK_LAUNCH();
  return QEFT_OK;
}

extern "C" int qeft_dequant_w4(const void* qweight, const void* scales, const void* scaled_zeros, const void* oweight,
                               void* w_dense, int N, int K, int r, int G, int dtype, qeft_stream_t stream) {
  if (!qweight || !scales || !scaled_zeros || !w_dense) return QEFT_E_NULL;
  if (G == -1) G = K;
  if (N <= 0 || K <= 0 || N % 4 != 0 || K % 64 != 0 || G <= 0 || K % G != 0 || G % 32 != 0) return QEFT_E_SHAPE;
  if (r < 0 || r % 32 != 0 || r >= K) return QEFT_E_SHAPE;
  if (!check_align16(qweight) || !check_align16(w_dense)) return QEFT_E_ALIGN;
  const size_t chunks = (size_t)N * (K >> 5);
  const __half* ow = r > 0 ? (const __half*)oweight : nullptr;
  if (dtype == QEFT_DT_F16) {
    dequant_w4_kernel<__half><<<grid_for(chunks, 256), 256, 0, (cudaStream_t)stream>>>(
        (const uint8_t*)qweight, (const __half*)scales, (const __half*)scaled_zeros, ow, (__half*)w_dense, N, K, r, G);
  } else if (dtype == QEFT_DT_BF16) {
    dequant_w4_kernel<__nv_bfloat16><<<grid_for(chunks, 256), 256, 0, (cudaStream_t)stream>>>(
        (const uint8_t*)qweight, (const __half*)scales, (const __half*)scaled_zeros, ow, (__nv_bfloat16*)w_dense, N, K, r, G);
  } else {
    return QEFT_E_DTYPE;
  }
  QEFT_CHECK_LAUNCH();
  return QEFT_OK;
}

extern "C" int qeft_interleave_oweight(const void* oweight, void* oweight_interleaved, int N, int r, int src_fp32,
                                       qeft_stream_t stream) {
  if (!oweight || !oweight_interleaved) return QEFT_E_NULL;
  if (N <= 0 || N % 8 != 0 || r <= 0 || r % 32 != 0) return QEFT_E_SHAPE;
  const size_t total = (size_t)N * r;
  if (src_fp32)
    interleave_oweight_kernel<float><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
        (const float*)oweight, (__half*)oweight_interleaved, N, r);
  else
    interleave_oweight_kernel<__half><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
        (const __half*)oweight, (__half*)oweight_interleaved, N, r);
  QEFT_CHECK_LAUNCH();
  return QEFT_OK;
}
