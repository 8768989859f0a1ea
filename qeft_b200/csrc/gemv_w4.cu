// Decode-path dequant + GEMV for the packed QEFT QuantLinear (sm_100a).
//
// Replaces gemv_kernel / gemv_kernel_qeft of the reference
// (qeft/kernel/quantization_new/gemv/gemv_cuda.cu:73-204, gemv_cuda_qeft.cu:75-222).
//
// Design (see DESIGN.md "GEMV"):
//   * one CTA = 16 output rows (4 consecutive qweight rows) x all of K, WARPS warps; warp w owns the
//     128-column k-steps w, w+WARPS, ... ;  grid = ceil(N/16) so even a 4096-row layer gives 256 CTAs that
//     are all resident at once on 148 SMs: the whole matrix is requested from HBM at t=0.
//   * per k-step a thread issues two 128-bit streaming loads (the 32 nibbles of row g and of row g+8 that
//     belong to its quarter of the step); a warp-level load instruction covers 2 x 256 contiguous bytes.
//     Loads for DEPTH steps are in flight per thread before the first one is consumed, and ALL of them are
//     issued before `griddepcontrol.wait`, so with programmatic dependent launch the weight stream of
//     layer i+1 overlaps the tail of layer i (weights do not depend on the previous kernel's output).
//   * nibbles are unpacked in registers (lop3 + one f16x2 op per pair, exact 0..15) and fed, without any
//     shuffle, as the A fragment of mma.m16n8k16 (the packed order IS that fragment order); x is the B
//     fragment (batch m <= 8 columns), accumulation is fp32.  Scale and zero point are applied once per
//     128-column group in fp32:  y += s * sum(q x) + sz * sum(x).
//   * the fp16 outlier columns are a CUDA-core dot product reduced with warp shuffles; the k-split
//     partial sums of the warps meet in shared memory; fp16 store.
#include "common.cuh"

namespace qeft {

unsigned long long g_launch_count = 0;

struct GemvPart {
  const uint8_t* qw;      // int16 [N/4, K] as bytes, row pitch 2K
  const __half* scales;   // [K/G, N]
  const __half* szeros;   // [K/G, N]
  const __half* ow;       // plain [N, r] or interleaved [N/2, 2r]
  const __half* bias;     // [N] or null
  __half* y;              // [m, N]
  int N;
  int cta_begin;          // first blockIdx.x of this part
};

struct GemvParams {
  GemvPart part[QEFT_GEMV_MAX_PARTS];
  int nparts;
  const __half* x;        // [m, K]
  const int32_t* gather;  // [K] or null
  int m, K, r, G;
  int ow_layout;
};

template <int WARPS>
struct GemvSmem {
  // floats
  static constexpr int kRed = WARPS * 16 * 8;   // k-split partial sums [warp][row][batch]
};

struct StepRegs {
  uint4 wa, wb;                 // 32 nibbles of row g / row g+8
  unsigned short sa, sb, za, zb;  // fp16 bits of scale / scaled zero of the two rows
};

template <int WARPS, int DEPTH, bool XS>
__global__ void __launch_bounds__(WARPS * 32)
gemv_w4_kernel(const GemvParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;

  // ---- which part / which 16 rows -------------------------------------------------------
  int pi = 0;
#pragma unroll
  for (int i = 1; i < QEFT_GEMV_MAX_PARTS; ++i)
    if (i < p.nparts && (int)blockIdx.x >= p.part[i].cta_begin) pi = i;
  const GemvPart& P = p.part[pi];
  const int n0 = ((int)blockIdx.x - P.cta_begin) * 16;
  const int N = P.N, K = p.K, r = p.r, m = p.m;
  const int KQ = K - r;                      // int4 columns that are live
  const int nsteps = cdiv(KQ, 128);          // 128-column k-steps
  const int nchunks = KQ >> 5;               // live 32-column chunks
  const bool rowB_ok = (n0 + 8) < N;         // N % 8 == 0: a CTA has 16 or 8 live rows

  // ---- shared memory carve-up -----------------------------------------------------------
  float* red = reinterpret_cast<float*>(smem_raw);              // [WARPS][16][8]
  float* xsum = red + WARPS * 128;                              // [nsteps][8]   sum of x per k-step
  float* opart = xsum + nsteps * 8;                             // [r/32][16][8] outlier partial sums
  __half* xs = reinterpret_cast<__half*>(opart + (r >> 5) * 128);  // XS only: gathered x [m][K]

  // ---- weight stream: issue before waiting on the previous kernel --------------------------
  const int my_cnt = (nsteps > warp) ? (nsteps - warp + WARPS - 1) / WARPS : 0;
  const uint8_t* rowA = P.qw + (size_t)((n0 >> 2) + (g >> 2)) * (size_t)(2 * K) + (g & 3) * 32 + (t >> 1) * 128 + (t & 1) * 16;
  const uint8_t* rowB = rowA + (size_t)2 * (size_t)(2 * K);
  const int gshift_n = N;  // scales row pitch
  const __half* scA = P.scales + n0 + g;
  const __half* szA = P.szeros + n0 + g;

  auto load_step = [&](StepRegs& R, int i) {
    const int s = warp + i * WARPS;
    const bool live = (i < my_cnt) && ((4 * s + t) < nchunks);
    R.wa = make_uint4(0, 0, 0, 0);
    R.wb = make_uint4(0, 0, 0, 0);
    R.sa = R.sb = R.za = R.zb = 0;
    if (live) {
      R.wa = ldg_stream_v4(rowA + (size_t)s * 256);
      if (rowB_ok) R.wb = ldg_stream_v4(rowB + (size_t)s * 256);
    }
    if (i < my_cnt) {
      const size_t go = (size_t)((s * 128) / p.G) * (size_t)gshift_n;
      R.sa = ldg_nc_u16(scA + go);
      R.za = ldg_nc_u16(szA + go);
      if (rowB_ok) {
        R.sb = ldg_nc_u16(scA + go + 8);
        R.zb = ldg_nc_u16(szA + go + 8);
      }
    }
  };

  StepRegs ring[DEPTH];
#pragma unroll
  for (int d = 0; d < DEPTH; ++d) load_step(ring[d], d);

  // outlier weights of this CTA: 16 rows x r fp16 = 2r pieces of 16 bytes (r = 128 -> one per thread)
  constexpr int kMaxOwIters = 2;  // r <= 256 with 256 threads; larger r loops without prefetch
  const int live_rows = rowB_ok ? 16 : 8;
  const int npieces = (r * live_rows) >> 3;
  uint4 owv[kMaxOwIters];
#pragma unroll
  for (int it = 0; it < kMaxOwIters; ++it) {
    const int piece = tid + it * WARPS * 32;
    owv[it] = make_uint4(0, 0, 0, 0);
    if (piece < npieces) {
      const uint8_t* base = (p.ow_layout == QEFT_OW_INTERLEAVED)
                                ? reinterpret_cast<const uint8_t*>(P.ow) + (size_t)(n0 >> 1) * (size_t)(4 * r)
                                : reinterpret_cast<const uint8_t*>(P.ow) + (size_t)n0 * (size_t)(2 * r);
      owv[it] = ldg_stream_v4(base + (size_t)piece * 16);
    }
  }

  pdl_launch_dependents();
  pdl_wait();   // x (and y as a reused buffer) belong to the previous kernel until here

  // ---- x: per-step sums (and, for the gathered variant, the staged copy) ---------------------
  const __half* xg = p.x;
  if (XS) {
    for (int i = tid; i < m * K; i += WARPS * 32) {
      const int b = i / K, k = i - b * K;
      xs[i] = xg[(size_t)b * K + p.gather[k]];
    }
    __syncthreads();
  }
  {
    // 8 threads per (batch row, k-step): 16 halves each, fp32 sum, 3 shuffles
    const int total = m * nsteps * 8;
    for (int i = tid; i < cdiv(total, 32) * 32; i += WARPS * 32) {
      float acc = 0.f;
      const int unit = i >> 3, sub = i & 7;
      const int b = unit / nsteps, s = unit - b * nsteps;
      const int k = s * 128 + sub * 16;
      if (i < total && k < KQ) {
        uint4 v0, v1;
        if (XS) {
          v0 = *reinterpret_cast<const uint4*>(xs + (size_t)b * K + k);
          v1 = *reinterpret_cast<const uint4*>(xs + (size_t)b * K + k + 8);
        } else {
          v0 = ldg_nc_v4(xg + (size_t)b * K + k);
          v1 = ldg_nc_v4(xg + (size_t)b * K + k + 8);
        }
        const uint32_t w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float2 f = half2_bits_to_float2(w[j]);
          acc += f.x + f.y;
        }
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      if (i < total && sub == 0) xsum[s * 8 + b] = acc;
    }
    // columns m..7 of xsum are never read with a non-zero multiplier, but keep them defined
    for (int i = tid; i < nsteps * 8; i += WARPS * 32)
      if ((i & 7) >= m) xsum[i] = 0.f;
  }
  __syncthreads();

  // ---- main loop ------------------------------------------------------------------------------
  float yacc[4] = {0.f, 0.f, 0.f, 0.f};   // rows g, g+8 x batch columns 2t, 2t+1
  const __half* xrow = (XS ? xs : xg) + (size_t)g * K + t * 32;

  auto consume = [&](const StepRegs& R, int i) {
    const int s = warp + i * WARPS;
    // B fragments: x[g][128 s + 32 t .. +32] as 16 half2 (zero for batch rows >= m and dead chunks)
    uint32_t xb[16];
    const bool xlive = (g < m) && ((4 * s + t) < nchunks);
    if (xlive) {
      const __half* xp = xrow + (size_t)s * 128;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 v = XS ? *reinterpret_cast<const uint4*>(xp + 8 * j) : ldg_nc_v4(xp + 8 * j);
        xb[4 * j + 0] = v.x; xb[4 * j + 1] = v.y; xb[4 * j + 2] = v.z; xb[4 * j + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) xb[j] = 0u;
    }
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const uint32_t wa[4] = {R.wa.x, R.wa.y, R.wa.z, R.wa.w};
    const uint32_t wb[4] = {R.wb.x, R.wb.y, R.wb.z, R.wb.w};
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      uint32_t ha[4], hb[4];
      unpack_word_to_half2(wa[w], ha);
      unpack_word_to_half2(wb[w], hb);
      // k-pairs of word w: offsets 2w (+0), 8+2w, 16+2w, 24+2w inside the chunk -> half2 index w, 4+w, 8+w, 12+w
      mma_m16n8k16_f16f32(acc, ha[0], hb[0], ha[1], hb[1], xb[w], xb[4 + w]);
      mma_m16n8k16_f16f32(acc, ha[2], hb[2], ha[3], hb[3], xb[8 + w], xb[12 + w]);
    }
    // group epilogue: y += s * sum(q x) + sz * sum(x)
    const float sa = __half2float(__ushort_as_half(R.sa)), sb = __half2float(__ushort_as_half(R.sb));
    const float za = __half2float(__ushort_as_half(R.za)), zb = __half2float(__ushort_as_half(R.zb));
    const float2 xs2 = *reinterpret_cast<const float2*>(xsum + s * 8 + 2 * t);
    yacc[0] = fmaf(sa, acc[0], fmaf(za, xs2.x, yacc[0]));
    yacc[1] = fmaf(sa, acc[1], fmaf(za, xs2.y, yacc[1]));
    yacc[2] = fmaf(sb, acc[2], fmaf(zb, xs2.x, yacc[2]));
    yacc[3] = fmaf(sb, acc[3], fmaf(zb, xs2.y, yacc[3]));
  };

  for (int i0 = 0; i0 < my_cnt; i0 += DEPTH) {
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) {
      const int i = i0 + d;
      if (i < my_cnt) {
        consume(ring[d], i);
        load_step(ring[d], i + DEPTH);
      }
    }
  }

  // ---- outlier columns (CUDA cores, fp32), reduced with warp shuffles ---------------------------
  if (r > 0) {
    const __half* xo = (XS ? xs : xg) + (K - r);
#pragma unroll
    for (int it = 0; it < kMaxOwIters; ++it) {
      const int piece = tid + it * WARPS * 32;
      const bool live = piece < npieces;
      const uint32_t w4[4] = {owv[it].x, owv[it].y, owv[it].z, owv[it].w};
      if (p.ow_layout == QEFT_OW_INTERLEAVED) {
        // interleaved row R (0..7 local) holds rows nl and nl+4; 16 bytes = columns j0..j0+3 of both rows
        const int per_row = r >> 2;                 // pieces per interleaved row
        const int R = live ? piece / per_row : 0, pp = live ? piece - R * per_row : 0;
        const int c = pp >> 3, j0 = 32 * c + 4 * (pp & 7);
        const int nl = 8 * (R >> 2) + (R & 3);
        for (int b = 0; b < m; ++b) {
          float s0 = 0.f, s1 = 0.f;
          if (live) {
            uint2 xv = XS ? *reinterpret_cast<const uint2*>(xo + (size_t)b * K + j0) : ldg_nc_v2(xo + (size_t)b * K + j0);
            const float2 x01 = half2_bits_to_float2(xv.x), x23 = half2_bits_to_float2(xv.y);
            const float xf[4] = {x01.x, x01.y, x23.x, x23.y};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 wv = half2_bits_to_float2(w4[j]);   // {row nl, row nl+4} at column j0+j
              s0 = fmaf(wv.x, xf[j], s0);
              s1 = fmaf(wv.y, xf[j], s1);
            }
          }
          s0 += __shfl_xor_sync(0xffffffffu, s0, 4); s1 += __shfl_xor_sync(0xffffffffu, s1, 4);
          s0 += __shfl_xor_sync(0xffffffffu, s0, 2); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
          s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
          if (live && (pp & 7) == 0) {
            opart[(c * 16 + nl) * 8 + b] = s0;
            opart[(c * 16 + nl + 4) * 8 + b] = s1;
          }
        }
      } else {
        // plain [N, r]: 16 bytes = 8 consecutive columns of one row
        const int per_row = r >> 3;
        const int nl = live ? piece / per_row : 0, pp = live ? piece - nl * per_row : 0;
        const int c = pp >> 2, j0 = 8 * pp;
        for (int b = 0; b < m; ++b) {
          float s0 = 0.f;
          if (live) {
            uint4 xv = XS ? *reinterpret_cast<const uint4*>(xo + (size_t)b * K + j0) : ldg_nc_v4(xo + (size_t)b * K + j0);
            const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 wv = half2_bits_to_float2(w4[j]);
              const float2 xf = half2_bits_to_float2(xw[j]);
              s0 = fmaf(wv.x, xf.x, s0);
              s0 = fmaf(wv.y, xf.y, s0);
            }
          }
          s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
          s0 += __shfl_xor_sync(0xffffffffu, s0, 1);
          if (live && (pp & 3) == 0) opart[(c * 16 + nl) * 8 + b] = s0;
        }
      }
    }
  }

  // ---- meet the k-split partial sums ----------------------------------------------------------------
  {
    float* my = red + warp * 128;
    *reinterpret_cast<float2*>(my + g * 8 + 2 * t) = make_float2(yacc[0], yacc[1]);
    *reinterpret_cast<float2*>(my + (g + 8) * 8 + 2 * t) = make_float2(yacc[2], yacc[3]);
  }
  __syncthreads();
  if (tid < 16 * m) {
    const int b = tid >> 4, nl = tid & 15;
    if (n0 + nl < N) {
      float acc = 0.f;
#pragma unroll
      for (int w = 0; w < WARPS; ++w) acc += red[w * 128 + nl * 8 + b];
      for (int c = 0; c < (r >> 5); ++c) acc += opart[(c * 16 + nl) * 8 + b];
      if (P.bias) acc += __half2float(P.bias[n0 + nl]);
      P.y[(size_t)b * N + n0 + nl] = __float2half_rn(acc);
    }
  }
}

// ----------------------------------------------------------------------------------------------------
constexpr int kGemvWarps = 8;
constexpr int kGemvDepth = 4;

static size_t gemv_smem_bytes(int m, int K, int r, bool xs) {
  const int nsteps = cdiv(K - r, 128);
  size_t b = sizeof(float) * ((size_t)kGemvWarps * 128 + (size_t)nsteps * 8 + (size_t)(r >> 5) * 128);
  if (xs) b += sizeof(__half) * (size_t)m * (size_t)K;
  return (b + 15) & ~(size_t)15;
}

template <bool XS>
static int launch_gemv(const GemvParams& prm, int total_ctas, unsigned flags, cudaStream_t stream) {
  auto kern = gemv_w4_kernel<kGemvWarps, kGemvDepth, XS>;
  const size_t smem = gemv_smem_bytes(prm.m, prm.K, prm.r, XS);
  if (smem > 227 * 1024) return QEFT_E_UNSUPPORTED;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)total_ctas);
  cfg.blockDim = dim3(kGemvWarps * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (flags & QEFT_F_PDL) ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, prm);
  if (e != cudaSuccess) return (int)e;
  count_launch();
  return QEFT_OK;
}

}  // namespace qeft

using namespace qeft;

extern "C" int qeft_gemv_w4_multi(const void* x, const qeft_gemv_part_t* parts, int nparts, int ow_layout,
                                  const int32_t* x_gather, int m, int K, int r, int G, unsigned flags,
                                  qeft_stream_t stream) {
  if (!x || !parts) return QEFT_E_NULL;
  if (nparts < 1 || nparts > QEFT_GEMV_MAX_PARTS) return QEFT_E_SHAPE;
  if (m < 1 || m > 8) return QEFT_E_BATCH;
  if (G == -1) G = K;
  if (K <= 0 || K % 64 != 0 || G <= 0 || K % G != 0 || (G % 128 != 0 && G != K)) return QEFT_E_SHAPE;
  if (r < 0 || r % 32 != 0 || r >= K) return QEFT_E_SHAPE;
  if (G != K && (K - r) > 0 && G % 128 != 0) return QEFT_E_SHAPE;
  if (G == K && K % 128 != 0 && cdiv(K - r, 128) > 1) {
    // per-channel scales: any K % 64 == 0 works (the group index is always 0)
  }
  if (r > 0 && ow_layout != QEFT_OW_PLAIN && ow_layout != QEFT_OW_INTERLEAVED) return QEFT_E_DTYPE;
  if (r == 0) ow_layout = QEFT_OW_NONE;
  if (r > 256) return QEFT_E_UNSUPPORTED;   // TODO(next): loop the outlier pieces
  if (!check_align16(x)) return QEFT_E_ALIGN;
  GemvParams prm = {};
  int ctas = 0;
  for (int i = 0; i < nparts; ++i) {
    const qeft_gemv_part_t& q = parts[i];
    if (!q.qweight || !q.scales || !q.scaled_zeros || !q.y) return QEFT_E_NULL;
    if (r > 0 && !q.oweight) return QEFT_E_NULL;
    if (q.N <= 0 || q.N % 8 != 0) return QEFT_E_SHAPE;
    if (!check_align16(q.qweight) || (r > 0 && !check_align16(q.oweight))) return QEFT_E_ALIGN;
    GemvPart& d = prm.part[i];
    d.qw = static_cast<const uint8_t*>(q.qweight);
    d.scales = static_cast<const __half*>(q.scales);
    d.szeros = static_cast<const __half*>(q.scaled_zeros);
    d.ow = static_cast<const __half*>(q.oweight);
    d.bias = static_cast<const __half*>(q.bias);
    d.y = static_cast<__half*>(q.y);
    d.N = q.N;
    d.cta_begin = ctas;
    ctas += cdiv(q.N, 16);
  }
  prm.nparts = nparts;
  prm.x = static_cast<const __half*>(x);
  prm.gather = x_gather;
  prm.m = m; prm.K = K; prm.r = r; prm.G = G;
  prm.ow_layout = ow_layout;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return x_gather ? launch_gemv<true>(prm, ctas, flags, st) : launch_gemv<false>(prm, ctas, flags, st);
}

extern "C" int qeft_gemv_w4(const void* x, const void* qweight, const void* scales, const void* scaled_zeros,
                            const void* oweight, int ow_layout, const void* bias, const int32_t* x_gather,
                            void* y, int m, int N, int K, int r, int G, unsigned flags, qeft_stream_t stream) {
  qeft_gemv_part_t part = {qweight, scales, scaled_zeros, oweight, bias, y, N};
  return qeft_gemv_w4_multi(x, &part, 1, ow_layout, x_gather, m, K, r, G, flags, stream);
}
