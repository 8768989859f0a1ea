// Decode-path dequant + GEMV for the packed QEFT QuantLinear (sm_100a).
//
// Replaces gemv_kernel / gemv_kernel_qeft of the reference
// (qeft/kernel/quantization_new/gemv/gemv_cuda.cu:73-204, gemv_cuda_qeft.cu:75-222).
//
// Design (see DESIGN.md "GEMV"):
//   * one CTA = 16 output rows (4 consecutive qweight rows) x all of K; 8 consumer warps + 1 producer warp.
//     Consumer warp w owns the 128-column k-steps w, w+8, ... ;  grid = ceil(N/16), two CTAs per SM.
//   * the producer warp streams the CTA's packed bytes into a shared-memory ring with 1-D bulk async copies
//     (cp.async.bulk, completion on an mbarrier): one ring stage = one "round" of 8 k-steps = 4 x 2 KB of
//     contiguous qweight bytes + the 8 steps' scales / scaled zeros.  Up to 16 stages (128 KB) are in flight
//     per CTA without costing a register, which is what it takes to cover HBM latency at 6.5 TB/s.  The
//     producer never waits for the previous kernel: with programmatic dependent launch the weight stream of
//     layer i+1 overlaps the tail of layer i (weights do not depend on the previous kernel's output).
//   * a consumer thread reads its two 16-byte chunks (32 nibbles of row g and of row g+8) from the ring,
//     unpacks them in registers (lop3 + one f16x2 op per pair, exact 0..15) and feeds them, without any
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
  int m, K, r;
  int g128;               // G / 128 (1 for the common G = 128), 0 for per-channel scales (G == K)
  int ow_layout;
  int nsteps;             // ceil((K - r) / 128)
  int nfull;              // (K - r) / 128: steps whose four 32-column chunks are all live
  int nchunks;            // (K - r) / 32 live 32-column chunks
  int xstride;            // halves between batch rows of the staged x (K + 8: rows start 4 banks apart)
  int ngroups;            // scale groups that cover the live int4 columns
  int stages;             // ring depth (rounds in flight)
  int rounds;             // ceil(nfull / 8)
};

// ---- mbarrier / bulk-copy primitives (shared::cta addresses as 32-bit) -----------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}"
      ::"r"(bar), "r"(parity) : "memory");
}
// global -> shared 1-D bulk copy, bytes multiple of 16, completion counted on `bar`
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// mma with a zero accumulator input (first k-slice of a group)
__device__ __forceinline__ void mma_m16n8k16_zero(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                                  uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
      : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "f"(0.f));
}

// XS: x is staged in shared memory (always when it fits; required for the fused o_proj gather).
// G128: one scale group per 128-column step (the common G = 128); otherwise the group index is s / g128.
constexpr int kStepBytes = 256;                       // one qweight row's bytes of a 128-column step
template <int WARPS> struct GemvStage {
  static constexpr int kRowBytes = WARPS * kStepBytes;            // one round of one qweight row
  static constexpr int kBytes = 4 * kRowBytes;                    // 4 qweight rows = 16 output rows
};
constexpr int kMaxStages = 16;

template <int WARPS, bool XS, bool G128, int MINB>
__global__ void __launch_bounds__((WARPS + 1) * 32, MINB)
gemv_w4_kernel(const GemvParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  using Stage = GemvStage<WARPS>;
  constexpr int kConsumers = WARPS * 32;
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
  const int nsteps = p.nsteps, nchunks = p.nchunks, nfull = p.nfull;
  const int stages = p.stages, rounds = p.rounds;
  const bool rowB_ok = (n0 + 8) < N;         // N % 8 == 0: a CTA has 16 or 8 live rows

  // ---- shared memory carve-up -----------------------------------------------------------
  uint8_t* ring = smem_raw;                                                   // [stages][Stage::kBytes]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)stages * Stage::kBytes);   // full[16], empty[16]
  float* red = reinterpret_cast<float*>(bars + 2 * kMaxStages);               // [WARPS][16][8]
  float* xsum = red + WARPS * 128;                                            // [nsteps][8]   sum of x per k-step
  float* opart = xsum + nsteps * 8;                                           // [r/32][16][8] outlier partial sums
  __half* sctab = reinterpret_cast<__half*>(opart + (r >> 5) * 128);          // [ngroups][16 scales | 16 scaled zeros]
  __half* xs = sctab + (size_t)p.ngroups * 32;                                // XS only: staged x [m][xstride]
  const uint32_t ring_u32 = smem_u32(ring);
  const uint32_t full_u32 = smem_u32(bars), empty_u32 = smem_u32(bars + kMaxStages);

  if (tid == 0) {
    for (int i = 0; i < stages; ++i) {
      mbar_init(full_u32 + 8 * i, 1);
      mbar_init(empty_u32 + 8 * i, WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  pdl_launch_dependents();

  // =====================================================================================================
  // producer warp: stream this CTA's packed bytes into the ring.  Nothing here depends on the previous
  // kernel, so it does not wait for it.
  // =====================================================================================================
  if (warp == WARPS) {
    const uint8_t* qrow0 = P.qw + (size_t)(n0 >> 2) * (size_t)(2 * K);
    const int nq = rowB_ok ? 4 : 2;
    for (int rd = 0; rd < rounds; ++rd) {
      const int st = rd % stages;
      if (rd >= stages) mbar_wait(empty_u32 + 8 * st, (uint32_t)((rd / stages - 1) & 1));
      const int steps = min(WARPS, nfull - rd * WARPS);
      const uint32_t fb = full_u32 + 8 * st;
      const uint32_t sbase = ring_u32 + (uint32_t)st * Stage::kBytes;
      if (lane == 0) mbar_expect_tx(fb, (uint32_t)(nq * steps * kStepBytes));
      __syncwarp();
      if (lane < nq)
        bulk_g2s(sbase + lane * Stage::kRowBytes, qrow0 + (size_t)lane * (size_t)(2 * K) + (size_t)rd * Stage::kRowBytes,
                 (uint32_t)(steps * kStepBytes), fb);
    }
    return;
  }

  // =====================================================================================================
  // consumer warps
  // =====================================================================================================
  // outlier weights of this CTA: 16 rows x r fp16 = 2r pieces of 16 bytes (r = 128 -> one per thread)
  constexpr int kMaxOwIters = 2;
  const int live_rows = rowB_ok ? 16 : 8;
  const int npieces = (r * live_rows) >> 3;
  uint4 owv[kMaxOwIters];
#pragma unroll
  for (int it = 0; it < kMaxOwIters; ++it) {
    const int piece = tid + it * kConsumers;
    owv[it] = make_uint4(0, 0, 0, 0);
    if (piece < npieces) {
      const uint8_t* base = (p.ow_layout == QEFT_OW_INTERLEAVED)
                                ? reinterpret_cast<const uint8_t*>(P.ow) + (size_t)(n0 >> 1) * (size_t)(4 * r)
                                : reinterpret_cast<const uint8_t*>(P.ow) + (size_t)n0 * (size_t)(2 * r);
      owv[it] = ldg_stream_v4(base + (size_t)piece * 16);
    }
  }

  // scale table: per group 16 scales | 16 scaled zeros (32 B + 32 B); one 16-byte load per thread and piece
  {
    const int pieces_per_group = rowB_ok ? 4 : 2;           // 16-byte pieces: s[0:8] s[8:16] z[0:8] z[8:16] (or s[0:8] z[0:8])
    const int npc = p.ngroups * pieces_per_group;
    for (int i = tid; i < npc; i += kConsumers) {
      const int gi = i / pieces_per_group, q = i - gi * pieces_per_group;
      const int which = rowB_ok ? (q >> 1) : q, half8 = rowB_ok ? (q & 1) : 0;
      const __half* src = (which ? P.szeros : P.scales) + (size_t)gi * N + n0 + 8 * half8;
      *reinterpret_cast<uint4*>(sctab + gi * 32 + which * 16 + 8 * half8) = ldg_nc_v4(src);
    }
  }

  pdl_wait();   // x (and y as a reused buffer) belong to the previous kernel until here

  // ---- x: staged copy (natural order) + per-step sums ------------------------------------------
  const __half* xg = p.x;
  const int xstride = XS ? p.xstride : K;
  {
    // units of 16 halves; 8 consecutive units = one 128-column step -> fp32 sum with 3 shuffles
    const int upr = cdiv(K, 128) * 8;                       // units per batch row, padded to whole steps
    const int live_k = nchunks * 32;
    for (int b = 0; b < m; ++b) {
      const __half* xrow = xg + (size_t)b * K;
      for (int u = tid; u < ((upr + 31) & ~31); u += kConsumers) {   // whole warps enter together (full-mask shuffles)
        const int k = u * 16;
        float acc = 0.f;
        if (k < K) {
          uint4 v0, v1;
          if (XS && p.gather) {
            __half tmp[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) tmp[j] = xrow[p.gather[k + j]];
            v0 = *reinterpret_cast<uint4*>(tmp);
            v1 = *reinterpret_cast<uint4*>(tmp + 8);
          } else {
            v0 = ldg_nc_v4(xrow + k);
            v1 = ldg_nc_v4(xrow + k + 8);
          }
          if (XS) {
            *reinterpret_cast<uint4*>(xs + (size_t)b * xstride + k) = v0;
            *reinterpret_cast<uint4*>(xs + (size_t)b * xstride + k + 8) = v1;
          }
          if (k < live_k) {
            const uint32_t w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float2 f = half2_bits_to_float2(w[j]);
              acc += f.x + f.y;
            }
          }
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        if ((u & 7) == 0 && (u >> 3) < nsteps) xsum[(u >> 3) * 8 + b] = acc;
      }
    }
    for (int i = tid; i < nsteps * 8; i += kConsumers)
      if ((i & 7) >= m) xsum[i] = 0.f;
  }
  named_bar_sync(1, kConsumers);

  // ---- main loop ------------------------------------------------------------------------------
  float yacc[4] = {0.f, 0.f, 0.f, 0.f};   // rows g, g+8 x batch columns 2t, 2t+1
  const bool xrow_ok = g < m;
  const int toff = (t >> 1) * 128 + (t & 1) * 16;               // this lane's 16-byte chunk inside the 256-byte step

  // one 128-column step: 8 words (4 of row g, 4 of row g+8) -> 8 mma, then the group epilogue.
  // xb[4j + c] is the natural-order half2 (k = 8j + 2c, +1) of this lane's 32-column chunk, i.e. the k-pair
  // that word c's j-th half2 multiplies.  mma (j, cc) takes the k-slots (2t, 2t+1) from word 2cc and
  // (2t+8, 2t+9) from word 2cc+1, so its B registers are the adjacent pair xb[4j + 2cc], xb[4j + 2cc + 1].
  auto step_math = [&](const uint4& va, const uint4& vb, const uint32_t (&xb)[16], float mine, int s) {
    float acc0[4], acc1[4];
    const uint32_t wa_[4] = {va.x, va.y, va.z, va.w};
    const uint32_t wb_[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      uint32_t a0[4], a1[4], b0[4], b1[4];
      unpack_word_to_half2(wa_[2 * cc], a0);
      unpack_word_to_half2(wa_[2 * cc + 1], a1);
      unpack_word_to_half2(wb_[2 * cc], b0);
      unpack_word_to_half2(wb_[2 * cc + 1], b1);
      float (&acc)[4] = cc ? acc1 : acc0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j == 0) mma_m16n8k16_zero(acc, a0[j], b0[j], a1[j], b1[j], xb[4 * j + 2 * cc], xb[4 * j + 2 * cc + 1]);
        else mma_m16n8k16_f16f32(acc, a0[j], b0[j], a1[j], b1[j], xb[4 * j + 2 * cc], xb[4 * j + 2 * cc + 1]);
      }
    }
    // group epilogue: y += s * sum(q x) + sz * sum(x).  Lane l holds scale (l < 16) / scaled zero (l >= 16) of row l % 16.
    const float sa = __shfl_sync(0xffffffffu, mine, g), sb = __shfl_sync(0xffffffffu, mine, g + 8);
    const float za = __shfl_sync(0xffffffffu, mine, g + 16), zb = __shfl_sync(0xffffffffu, mine, g + 24);
    const float2 xs2 = *reinterpret_cast<const float2*>(xsum + s * 8 + 2 * t);
    yacc[0] = fmaf(sa, acc0[0] + acc1[0], fmaf(za, xs2.x, yacc[0]));
    yacc[1] = fmaf(sa, acc0[1] + acc1[1], fmaf(za, xs2.y, yacc[1]));
    yacc[2] = fmaf(sb, acc0[2] + acc1[2], fmaf(zb, xs2.x, yacc[2]));
    yacc[3] = fmaf(sb, acc0[3] + acc1[3], fmaf(zb, xs2.y, yacc[3]));
  };

  // B fragments of step s: x[g][128 s + 32 t .. +32], natural order (zero for batch rows >= m and dead chunks)
  auto load_x = [&](uint32_t (&xb)[16], int s, bool live) {
    if (live) {
      const __half* xp = (XS ? xs : xg) + (size_t)g * xstride + s * 128 + t * 32;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 v = XS ? *reinterpret_cast<const uint4*>(xp + 8 * j) : ldg_nc_v4(xp + 8 * j);
        xb[4 * j + 0] = v.x; xb[4 * j + 1] = v.y; xb[4 * j + 2] = v.z; xb[4 * j + 3] = v.w;
      }
    }
  };

  {
    uint32_t xb[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) xb[j] = 0u;
    // this lane's bytes inside a stage: rows g (qweight row g/4) and g+8 (two qweight rows further)
    const uint32_t offA = (uint32_t)((g >> 2) * Stage::kRowBytes + warp * kStepBytes + (g & 3) * 32 + toff);
    const uint32_t offB = rowB_ok ? offA + 2 * Stage::kRowBytes : offA;
    // scale (lanes 0-15) / scaled zero (lanes 16-31) of row lane % 16 (lane % 8 when only 8 rows are live)
    const __half* my_sc = sctab + (lane >> 4) * 16 + (rowB_ok ? (lane & 15) : (lane & 7));
    int st = 0;
    uint32_t parity = 0;
    for (int s = warp; s < nfull; s += WARPS) {
      load_x(xb, s, xrow_ok);
      const int grp = G128 ? s : (p.g128 == 0 ? 0 : s / p.g128);
      const float mine = __half2float(my_sc[grp * 32]);
      mbar_wait(full_u32 + 8 * st, parity);
      const uint8_t* sb = ring + (size_t)st * Stage::kBytes;
      const uint4 va = *reinterpret_cast<const uint4*>(sb + offA);
      const uint4 vb = *reinterpret_cast<const uint4*>(sb + offB);
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_u32 + 8 * st);        // the stage's bytes of this warp are in registers
      step_math(va, vb, xb, mine, s);
      if (++st == stages) { st = 0; parity ^= 1u; }
    }
    // partial last step (K - r not a multiple of 128): chunks beyond K - r are dead
    if (nfull < nsteps && warp == (nfull % WARPS)) {
      const int sl = nfull;
      const bool live = (4 * sl + t) < nchunks;
      // a dead lane re-reads chunk 0 of the step (always mapped) and multiplies it by x = 0
      const uint8_t* a = P.qw + (size_t)((n0 >> 2) + (g >> 2)) * (size_t)(2 * K) + (g & 3) * 32 + (size_t)sl * 256 +
                         (live ? toff : 0);
      const uint4 va = ldg_stream_v4(a), vb = ldg_stream_v4(a + (rowB_ok ? (size_t)4 * K : 0));
      const int grp = G128 ? sl : (p.g128 == 0 ? 0 : sl / p.g128);
      const float mine = __half2float(my_sc[grp * 32]);
#pragma unroll
      for (int j = 0; j < 16; ++j) xb[j] = 0u;
      load_x(xb, sl, xrow_ok && live);
      step_math(va, vb, xb, mine, sl);
    }
  }

  // ---- outlier columns (CUDA cores, fp32), reduced with warp shuffles ---------------------------
  if (r > 0) {
    const __half* xo = (XS ? xs : xg) + (K - r);
#pragma unroll
    for (int it = 0; it < kMaxOwIters; ++it) {
      const int piece = tid + it * kConsumers;
      if (it > 0 && it * kConsumers >= npieces) break;   // uniform
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
            uint2 xv = XS ? *reinterpret_cast<const uint2*>(xo + (size_t)b * xstride + j0)
                          : ldg_nc_v2(xo + (size_t)b * xstride + j0);
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
            uint4 xv = XS ? *reinterpret_cast<const uint4*>(xo + (size_t)b * xstride + j0)
                          : ldg_nc_v4(xo + (size_t)b * xstride + j0);
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
  named_bar_sync(1, kConsumers);
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
constexpr int kGemvMinBlocks = 2;
constexpr size_t kStageXMaxBytes = 72 * 1024;   // stage x in shared memory when it is at most this big
constexpr size_t kSmemTwoPerSm = 113 * 1024;    // two CTAs per SM
constexpr size_t kSmemOnePerSm = 226 * 1024;

static size_t gemv_fixed_smem(int m, int K, int r, int ngroups, bool xs) {
  const int nsteps = cdiv(K - r, 128);
  size_t b = 2 * kMaxStages * sizeof(uint64_t) +
             sizeof(float) * ((size_t)kGemvWarps * 128 + (size_t)nsteps * 8 + (size_t)(r >> 5) * 128) +
             sizeof(__half) * (size_t)ngroups * 32;
  if (xs) b += sizeof(__half) * (size_t)m * (size_t)(K + 8);
  return (b + 127) & ~(size_t)127;
}

template <bool XS, bool G128>
static int launch_gemv(GemvParams& prm, int total_ctas, unsigned flags, cudaStream_t stream) {
  auto kern = gemv_w4_kernel<kGemvWarps, XS, G128, kGemvMinBlocks>;
  using Stage = GemvStage<kGemvWarps>;
  const size_t fixed = gemv_fixed_smem(prm.m, prm.K, prm.r, prm.ngroups, XS);
  prm.rounds = cdiv(prm.nfull, kGemvWarps);
  const size_t budget = (fixed + 2 * Stage::kBytes <= kSmemTwoPerSm) ? kSmemTwoPerSm : kSmemOnePerSm;
  if (fixed + Stage::kBytes > budget) return QEFT_E_UNSUPPORTED;
  int stages = (int)((budget - fixed) / Stage::kBytes);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages > prm.rounds) stages = prm.rounds;
  if (stages < 1) stages = 1;
  prm.stages = stages;
  const size_t smem = fixed + (size_t)stages * Stage::kBytes;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)total_ctas);
  cfg.blockDim = dim3((kGemvWarps + 1) * 32);
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
  prm.m = m; prm.K = K; prm.r = r;
  prm.g128 = (G == K) ? 0 : G / 128;
  prm.ow_layout = ow_layout;
  prm.nsteps = cdiv(K - r, 128);
  prm.nfull = (K - r) / 128;
  prm.nchunks = (K - r) / 32;
  prm.xstride = K + 8;
  prm.ngroups = (prm.g128 == 0) ? 1 : cdiv(prm.nsteps, prm.g128);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool stage = x_gather != nullptr || (size_t)m * (size_t)(K + 8) * 2 <= kStageXMaxBytes;
  if (prm.g128 == 1)
    return stage ? launch_gemv<true, true>(prm, ctas, flags, st) : launch_gemv<false, true>(prm, ctas, flags, st);
  return stage ? launch_gemv<true, false>(prm, ctas, flags, st) : launch_gemv<false, false>(prm, ctas, flags, st);
}

extern "C" int qeft_gemv_w4(const void* x, const void* qweight, const void* scales, const void* scaled_zeros,
                            const void* oweight, int ow_layout, const void* bias, const int32_t* x_gather,
                            void* y, int m, int N, int K, int r, int G, unsigned flags, qeft_stream_t stream) {
  qeft_gemv_part_t part = {qweight, scales, scaled_zeros, oweight, bias, y, N};
  return qeft_gemv_w4_multi(x, &part, 1, ow_layout, x_gather, m, K, r, G, flags, stream);
}
