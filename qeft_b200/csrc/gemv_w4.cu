// Decode-path dequant + GEMV for the packed QEFT QuantLinear (sm_100a).
//
// Replaces gemv_kernel / gemv_kernel_qeft of the reference
// (qeft/kernel/quantization_new/gemv/gemv_cuda.cu:73-204, gemv_cuda_qeft.cu:75-222).
//
// Design (see DESIGN.md "GEMV"):
//   * one CTA = RG x 16 output rows (RG x 4 consecutive qweight rows) x all of K; W = 4 consumer warps + 1
//     producer warp.  Consumer warp w owns the 128-column k-steps w, w+W, ... of every row group.
//   * the producer warp streams the CTA's packed bytes into a shared-memory ring with 1-D bulk async copies
//     (cp.async.bulk, completion on an mbarrier): one ring stage = one "round" of 8 k-steps = 2 KB of
//     contiguous bytes from each qweight row.  Many stages are in flight per CTA without costing a
//     register.  x itself arrives by one bulk copy per batch row.  The weight stream never waits for the
//     previous kernel: with programmatic dependent launch the weights of layer i+1 stream in while layer i
//     drains (weights do not depend on the previous kernel's output); only the x copy waits.
//   * a consumer thread reads its two 16-byte chunks (32 nibbles of row g and of row g+8) from the ring and
//     turns every nibble pair into an fp16 pair with ONE lop3 (the nibble is OR-ed into the mantissa of
//     1024.0, giving 1024+q or 1024+16q exactly).  These go, without any shuffle or conversion, into the A
//     fragment of mma.m16n8k16 (the packed order IS that fragment order); x is the B fragment (batch
//     m <= 8 columns); accumulation is fp32 in two chains, one per nibble position, and the 1024 bias is
//     removed per 128-column group with the group's x sums:
//         sum(q x) = acc_lo + acc_hi / 16 - (1024 sum_lo(x) + 64 sum_hi(x))
//         y += s * sum(q x) + sz * sum(x)                       (fp32, once per group)
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
  int rounds;             // ceil(nfull / consumer warps)
  int pdl;                // launched with programmatic dependent launch: x is not ready when the CTA starts
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

// mma with a zero accumulator input (first k-slice of a chain)
__device__ __forceinline__ void mma_m16n8k16_zero(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                                  uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
      : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "f"(0.f));
}

// One 32-bit word (8 nibbles n0..n7) -> four half2 WITHOUT removing the 1024 bias:
//   h[0] = {1024+n0, 1024+n4}  h[1] = {1024+16 n1, 1024+16 n5}  h[2] = {1024+n2, 1024+n6}  h[3] = {1024+16 n3, 1024+16 n7}
// For word c of a chunk these are the k-pairs (2c, 2c+1) + 8j, j = 0..3; even j carry q, odd j carry 16 q.
__device__ __forceinline__ void unpack_word_biased(uint32_t w, uint32_t (&h)[4]) {
  constexpr uint32_t kLo = 0x000f000fu, kHi = 0x00f000f0u, kMagic = 0x64006400u;
  const uint32_t t = w >> 8;
  asm("lop3.b32 %0, %1, %2, %3, 0xea;" : "=r"(h[0]) : "r"(w), "n"(kLo), "n"(kMagic));
  asm("lop3.b32 %0, %1, %2, %3, 0xea;" : "=r"(h[1]) : "r"(w), "n"(kHi), "n"(kMagic));
  asm("lop3.b32 %0, %1, %2, %3, 0xea;" : "=r"(h[2]) : "r"(t), "n"(kLo), "n"(kMagic));
  asm("lop3.b32 %0, %1, %2, %3, 0xea;" : "=r"(h[3]) : "r"(t), "n"(kHi), "n"(kMagic));
}

constexpr int kStepBytes = 256;                       // one qweight row's bytes of a 128-column step
constexpr int kMaxStages = 16;

// RG: 16-row groups per CTA (1 or 2).  XS: x staged in shared memory.  G128: one scale group per k-step.
template <int WARPS, int RG, bool XS, bool G128, int MINB>
__global__ void __launch_bounds__((WARPS + 1) * 32, MINB)
gemv_w4_kernel(const GemvParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  constexpr int kConsumers = WARPS * 32;
  constexpr int kRowBytes = WARPS * kStepBytes;          // one round of one qweight row
  constexpr int kStageBytes = RG * 4 * kRowBytes;        // RG x 4 qweight rows
  constexpr int kRows = RG * 16;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;

  // ---- which part / which rows ------------------------------------------------------------
  int pi = 0;
#pragma unroll
  for (int i = 1; i < QEFT_GEMV_MAX_PARTS; ++i)
    if (i < p.nparts && (int)blockIdx.x >= p.part[i].cta_begin) pi = i;
  const GemvPart& P = p.part[pi];
  const int n0 = ((int)blockIdx.x - P.cta_begin) * kRows;
  const int N = P.N, K = p.K, r = p.r, m = p.m;
  const int nsteps = p.nsteps, nchunks = p.nchunks, nfull = p.nfull;
  const int stages = p.stages, rounds = p.rounds;
  const int live_rows = min(kRows, N - n0);          // multiple of 8 (N % 8 == 0)
  const int live_q = live_rows >> 2;                 // live qweight rows (multiple of 2)

  // ---- shared memory carve-up -----------------------------------------------------------
  uint8_t* ring = smem_raw;                                                   // [stages][kStageBytes]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)stages * kStageBytes);   // full[16], empty[16], xbar
  float* red = reinterpret_cast<float*>(bars + 2 * kMaxStages + 2);           // [WARPS][kRows][8]
  float* xsum = red + WARPS * kRows * 8;                                      // [nsteps][8]  sum(x) per k-step
  float* csum = xsum + nsteps * 8;                                            // [nsteps][8]  1024 sum_lo(x) + 64 sum_hi(x)
  float* opart = csum + nsteps * 8;                                           // [r/32][kRows][8] outlier partial sums
  __half* sctab = reinterpret_cast<__half*>(opart + (r >> 5) * kRows * 8);    // [ngroups][RG][16 scales | 16 scaled zeros]
  __half* xs = sctab + (size_t)p.ngroups * RG * 32;                           // staged x [m][xstride]
  const uint32_t ring_u32 = smem_u32(ring);
  const uint32_t full_u32 = smem_u32(bars), empty_u32 = smem_u32(bars + kMaxStages);
  const uint32_t xbar_u32 = smem_u32(bars + 2 * kMaxStages);
  const bool x_by_bulk = XS && (p.gather == nullptr);

  if (tid == 0) {
    for (int i = 0; i < stages; ++i) {
      mbar_init(full_u32 + 8 * i, 1);
      mbar_init(empty_u32 + 8 * i, WARPS);
    }
    mbar_init(xbar_u32, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  pdl_launch_dependents();

  // =====================================================================================================
  // producer warp
  // =====================================================================================================
  if (warp == WARPS) {
    const uint8_t* qrow0 = P.qw + (size_t)(n0 >> 2) * (size_t)(2 * K);
    bool x_sent = !x_by_bulk;
    auto send_x = [&]() {
      x_sent = true;
      pdl_wait();            // x belongs to the previous kernel until here
      if (lane == 0) mbar_expect_tx(xbar_u32, (uint32_t)(m * K * 2));
      __syncwarp();
      if (lane < m)
        bulk_g2s(smem_u32(xs + (size_t)lane * p.xstride), p.x + (size_t)lane * K, (uint32_t)(K * 2), xbar_u32);
    };
    if (!x_sent && !p.pdl) send_x();   // x is ready: fetch it ahead of the weight stream
    for (int rd = 0; rd < rounds; ++rd) {
      const int st = rd % stages;
      if (rd >= stages) {
        if (!x_sent) send_x();     // the ring is full: fetch x before waiting for the consumers
        mbar_wait(empty_u32 + 8 * st, (uint32_t)((rd / stages - 1) & 1));
      }
      const int steps = min(WARPS, nfull - rd * WARPS);
      const uint32_t fb = full_u32 + 8 * st;
      const uint32_t sbase = ring_u32 + (uint32_t)st * kStageBytes;
      if (lane == 0) mbar_expect_tx(fb, (uint32_t)(live_q * steps * kStepBytes));
      __syncwarp();
      if (lane < live_q)
        bulk_g2s(sbase + lane * kRowBytes, qrow0 + (size_t)lane * (size_t)(2 * K) + (size_t)rd * kRowBytes,
                 (uint32_t)(steps * kStepBytes), fb);
    }
    if (!x_sent) send_x();
    return;
  }

  // =====================================================================================================
  // consumer warps
  // =====================================================================================================
  // outlier weights of this CTA: live_rows x r fp16 in 16-byte pieces (r = 128, 16 rows -> one per thread)
  constexpr int kMaxOwIters = (32 * kRows) / kConsumers;   // r <= 256: at most 32 pieces per row
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
  // scale table: per (group, row group) 16 scales | 16 scaled zeros; one 16-byte load per 8 rows
  {
    const int per_group = live_rows >> 2;                   // 16-byte pieces per group: (live_rows / 8) x {s, z}
    const int npc = p.ngroups * per_group;
    for (int i = tid; i < npc; i += kConsumers) {
      const int gi = i / per_group, q = i - gi * per_group;
      const int oct = q >> 1, which = q & 1;                // rows 8 oct .. 8 oct + 7; scales / scaled zeros
      const __half* src = (which ? P.szeros : P.scales) + (size_t)gi * N + n0 + 8 * oct;
      const uint32_t dst = smem_u32(sctab + (gi * RG + (oct >> 1)) * 32 + which * 16 + 8 * (oct & 1));
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }

  pdl_wait();   // x (and y as a reused buffer) belong to the previous kernel until here

  // ---- x: per-step sums (and, when not bulk-copied, the staged copy) ---------------------------------
  const __half* xg = p.x;
  const int xstride = XS ? p.xstride : K;
  {
    if (x_by_bulk) mbar_wait(xbar_u32, 0);
    // units of 16 halves; 8 consecutive units = one 128-column step.  Inside a unit the first 8 halves sit in
    // "low nibble" k-slots (k % 16 < 8) and the last 8 in "high nibble" slots.
    const int upr = cdiv(K, 128) * 8;                       // units per batch row, padded to whole steps
    const int live_k = nchunks * 32;
    for (int b = 0; b < m; ++b) {
      const __half* xrow = xg + (size_t)b * K;
      for (int u = tid; u < ((upr + 31) & ~31); u += kConsumers) {   // whole warps enter together (full-mask shuffles)
        const int k = u * 16;
        float lo = 0.f, hi = 0.f;
        if (k < K) {
          uint4 v0, v1;
          if (x_by_bulk) {
            v0 = *reinterpret_cast<const uint4*>(xs + (size_t)b * xstride + k);
            v1 = *reinterpret_cast<const uint4*>(xs + (size_t)b * xstride + k + 8);
          } else {
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
          }
          if (k < live_k) {
            const float2 a0 = half2_bits_to_float2(v0.x), a1 = half2_bits_to_float2(v0.y);
            const float2 a2 = half2_bits_to_float2(v0.z), a3 = half2_bits_to_float2(v0.w);
            const float2 c0 = half2_bits_to_float2(v1.x), c1 = half2_bits_to_float2(v1.y);
            const float2 c2 = half2_bits_to_float2(v1.z), c3 = half2_bits_to_float2(v1.w);
            lo = ((a0.x + a0.y) + (a1.x + a1.y)) + ((a2.x + a2.y) + (a3.x + a3.y));
            hi = ((c0.x + c0.y) + (c1.x + c1.y)) + ((c2.x + c2.y) + (c3.x + c3.y));
          }
        }
        lo += __shfl_xor_sync(0xffffffffu, lo, 4); hi += __shfl_xor_sync(0xffffffffu, hi, 4);
        lo += __shfl_xor_sync(0xffffffffu, lo, 2); hi += __shfl_xor_sync(0xffffffffu, hi, 2);
        lo += __shfl_xor_sync(0xffffffffu, lo, 1); hi += __shfl_xor_sync(0xffffffffu, hi, 1);
        if ((u & 7) == 0 && (u >> 3) < nsteps) {
          xsum[(u >> 3) * 8 + b] = lo + hi;
          csum[(u >> 3) * 8 + b] = fmaf(1024.f, lo, 64.f * hi);
        }
      }
    }
    for (int i = tid; i < nsteps * 8; i += kConsumers)
      if ((i & 7) >= m) { xsum[i] = 0.f; csum[i] = 0.f; }
  }
  asm volatile("cp.async.wait_all;" ::: "memory");     // scale table
  named_bar_sync(1, kConsumers);

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
        // interleaved row R (local) holds rows nl and nl+4; 16 bytes = columns j0..j0+3 of both rows
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
            opart[(c * kRows + nl) * 8 + b] = s0;
            opart[(c * kRows + nl + 4) * 8 + b] = s1;
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
          if (live && (pp & 3) == 0) opart[(c * kRows + nl) * 8 + b] = s0;
        }
      }
    }
  }

  // ---- main loop ------------------------------------------------------------------------------
  float yacc[RG][4];                       // per row group: rows g, g+8 x batch columns 2t, 2t+1
#pragma unroll
  for (int q = 0; q < RG; ++q) yacc[q][0] = yacc[q][1] = yacc[q][2] = yacc[q][3] = 0.f;
  const bool xrow_ok = g < m;
  const int toff = (t >> 1) * 128 + (t & 1) * 16;               // this lane's 16-byte chunk inside the 256-byte step

  // One 128-column step of one row group: 8 words (4 of row g, 4 of row g+8) -> 8 mma in two chains.
  // xb[4j + c] is the natural-order half2 (k = 8j + 2c, +1) of this lane's 32-column chunk, i.e. the k-pair
  // that word c's j-th half2 multiplies.  mma (j, cc) takes k-slots (2t, 2t+1) from word 2cc and (2t+8, 2t+9)
  // from word 2cc+1, so its B registers are the adjacent pair xb[4j + 2cc], xb[4j + 2cc + 1]; even j (low
  // nibbles, 1024+q) accumulate in `lo`, odd j (high nibbles, 1024+16q) in `hi`.
  auto step_math = [&](float (&ya)[4], const uint4& va, const uint4& vb, const uint32_t (&xb)[16], float mine,
                       const float2& xs2, const float2& cs2) {
    float lo[4], hi[4];
    const uint32_t wa_[4] = {va.x, va.y, va.z, va.w};
    const uint32_t wb_[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      uint32_t a0[4], a1[4], b0[4], b1[4];
      unpack_word_biased(wa_[2 * cc], a0);
      unpack_word_biased(wa_[2 * cc + 1], a1);
      unpack_word_biased(wb_[2 * cc], b0);
      unpack_word_biased(wb_[2 * cc + 1], b1);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float (&acc)[4] = (j & 1) ? hi : lo;
        if (cc == 0 && j < 2)
          mma_m16n8k16_zero(acc, a0[j], b0[j], a1[j], b1[j], xb[4 * j + 2 * cc], xb[4 * j + 2 * cc + 1]);
        else
          mma_m16n8k16_f16f32(acc, a0[j], b0[j], a1[j], b1[j], xb[4 * j + 2 * cc], xb[4 * j + 2 * cc + 1]);
      }
    }
    // group epilogue.  Lane l holds scale (l < 16) / scaled zero (l >= 16) of row l % 16.
    const float sa = __shfl_sync(0xffffffffu, mine, g), sb = __shfl_sync(0xffffffffu, mine, g + 8);
    const float za = __shfl_sync(0xffffffffu, mine, g + 16), zb = __shfl_sync(0xffffffffu, mine, g + 24);
    // y += s * (lo + hi/16 - c) + z * X
    ya[0] = fmaf(sa, fmaf(hi[0], 0.0625f, lo[0]) - cs2.x, fmaf(za, xs2.x, ya[0]));
    ya[1] = fmaf(sa, fmaf(hi[1], 0.0625f, lo[1]) - cs2.y, fmaf(za, xs2.y, ya[1]));
    ya[2] = fmaf(sb, fmaf(hi[2], 0.0625f, lo[2]) - cs2.x, fmaf(zb, xs2.x, ya[2]));
    ya[3] = fmaf(sb, fmaf(hi[3], 0.0625f, lo[3]) - cs2.y, fmaf(zb, xs2.y, ya[3]));
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
    // this lane's bytes inside a stage: rows g (qweight row g/4) and g+8 (two qweight rows further) of row group q
    const uint32_t offA = (uint32_t)((g >> 2) * kRowBytes + warp * kStepBytes + (g & 3) * 32 + toff);
    // scale (lanes 0-15) / scaled zero (lanes 16-31) of row lane % 16
    const __half* my_sc = sctab + (lane >> 4) * 16 + (lane & 15);
    int st = 0;
    uint32_t parity = 0;
    for (int s = warp; s < nfull; s += WARPS) {
      load_x(xb, s, xrow_ok);
      const int grp = G128 ? s : (p.g128 == 0 ? 0 : s / p.g128);
      float mine[RG];
#pragma unroll
      for (int q = 0; q < RG; ++q) mine[q] = __half2float(my_sc[(grp * RG + q) * 32]);
      const float2 xs2 = *reinterpret_cast<const float2*>(xsum + s * 8 + 2 * t);
      const float2 cs2 = *reinterpret_cast<const float2*>(csum + s * 8 + 2 * t);
      mbar_wait(full_u32 + 8 * st, parity);
      const uint8_t* sb = ring + (size_t)st * kStageBytes;
      uint4 va[RG], vb[RG];
#pragma unroll
      for (int q = 0; q < RG; ++q) {
        // dead rows (beyond N) read whatever the ring holds: finite garbage that is never stored
        va[q] = *reinterpret_cast<const uint4*>(sb + offA + q * 4 * kRowBytes);
        vb[q] = *reinterpret_cast<const uint4*>(sb + offA + q * 4 * kRowBytes + 2 * kRowBytes);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_u32 + 8 * st);        // the stage's bytes of this warp are in registers
#pragma unroll
      for (int q = 0; q < RG; ++q) step_math(yacc[q], va[q], vb[q], xb, mine[q], xs2, cs2);
      if (++st == stages) { st = 0; parity ^= 1u; }
    }
    // partial last step (K - r not a multiple of 128): chunks beyond K - r are dead
    if (nfull < nsteps && warp == (nfull % WARPS)) {
      const int sl = nfull;
      const bool live = (4 * sl + t) < nchunks;
      const int grp = G128 ? sl : (p.g128 == 0 ? 0 : sl / p.g128);
#pragma unroll
      for (int j = 0; j < 16; ++j) xb[j] = 0u;
      load_x(xb, sl, xrow_ok && live);
      const float2 xs2 = *reinterpret_cast<const float2*>(xsum + sl * 8 + 2 * t);
      const float2 cs2 = *reinterpret_cast<const float2*>(csum + sl * 8 + 2 * t);
#pragma unroll
      for (int q = 0; q < RG; ++q) {
        // a dead lane (or a dead row) re-reads chunk 0 of a live row: always mapped, multiplied by x = 0 / never stored
        const int qa = min((n0 >> 2) + 4 * q + (g >> 2), (N >> 2) - 1);
        const int qb = min((n0 >> 2) + 4 * q + 2 + (g >> 2), (N >> 2) - 1);
        const size_t inrow = (size_t)(g & 3) * 32 + (size_t)sl * 256 + (live ? toff : 0);
        const uint4 va = ldg_stream_v4(P.qw + (size_t)qa * (size_t)(2 * K) + inrow);
        const uint4 vb = ldg_stream_v4(P.qw + (size_t)qb * (size_t)(2 * K) + inrow);
        const float mine = __half2float(my_sc[(grp * RG + q) * 32]);
        step_math(yacc[q], va, vb, xb, mine, xs2, cs2);
      }
    }
  }

  // ---- meet the k-split partial sums ----------------------------------------------------------------
  {
    float* my = red + warp * kRows * 8;
#pragma unroll
    for (int q = 0; q < RG; ++q) {
      *reinterpret_cast<float2*>(my + (16 * q + g) * 8 + 2 * t) = make_float2(yacc[q][0], yacc[q][1]);
      *reinterpret_cast<float2*>(my + (16 * q + g + 8) * 8 + 2 * t) = make_float2(yacc[q][2], yacc[q][3]);
    }
  }
  named_bar_sync(1, kConsumers);
  for (int i = tid; i < kRows * m; i += kConsumers) {
    const int b = i / kRows, nl = i - b * kRows;
    if (n0 + nl < N) {
      float acc = 0.f;
#pragma unroll
      for (int w = 0; w < WARPS; ++w) acc += red[(w * kRows + nl) * 8 + b];
      for (int c = 0; c < (r >> 5); ++c) acc += opart[(c * kRows + nl) * 8 + b];
      if (P.bias) acc += __half2float(P.bias[n0 + nl]);
      P.y[(size_t)b * N + n0 + nl] = __float2half_rn(acc);
    }
  }
}

// ----------------------------------------------------------------------------------------------------
constexpr int kGemvWarps = 8;
constexpr int kGemvMinBlocks = 2;
constexpr int kGemvStageCap = 4;                // rounds in flight per CTA: enough to cover HBM latency with two CTAs per
                                                // SM, small enough that the dependent x fetch does not queue behind them
constexpr size_t kStageXMaxBytes = 72 * 1024;   // stage x in shared memory when it is at most this big
constexpr size_t kSmemPerSm = 226 * 1024;
constexpr int kNumSms = 148;

static size_t gemv_fixed_smem(int rg, int m, int K, int r, int ngroups, bool xs) {
  const int nsteps = cdiv(K - r, 128);
  const int rows = rg * 16;
  size_t b = (2 * kMaxStages + 2) * sizeof(uint64_t) +
             sizeof(float) * ((size_t)kGemvWarps * rows * 8 + 2 * (size_t)nsteps * 8 + (size_t)(r >> 5) * rows * 8) +
             sizeof(__half) * (size_t)ngroups * rg * 32;
  if (xs) b += sizeof(__half) * (size_t)m * (size_t)(K + 8);
  return (b + 127) & ~(size_t)127;
}

template <int RG, bool XS, bool G128>
static int launch_gemv(GemvParams& prm, int total_ctas, unsigned flags, cudaStream_t stream) {
  auto kern = gemv_w4_kernel<kGemvWarps, RG, XS, G128, kGemvMinBlocks>;
  constexpr size_t stage_bytes = (size_t)RG * 4 * kGemvWarps * kStepBytes;
  const size_t fixed = gemv_fixed_smem(RG, prm.m, prm.K, prm.r, prm.ngroups, XS);
  prm.rounds = cdiv(prm.nfull, kGemvWarps);
  // CTAs per SM: 4 (so that two consecutive launches are co-resident and the next layer's weights stream in
  // under programmatic dependent launch while this one computes), else 2, else 1 -- whatever leaves the ring
  // at least four stages
  size_t budget = kSmemPerSm / 2;
  if (fixed + 3 * stage_bytes > budget) budget = kSmemPerSm;
  if (fixed + stage_bytes > budget) return QEFT_E_UNSUPPORTED;
  int stages = (int)((budget - fixed) / stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages > kGemvStageCap) stages = kGemvStageCap;
  if (stages > prm.rounds) stages = prm.rounds;
  if (stages < 1) stages = 1;
  prm.stages = stages;
  const size_t smem = fixed + (size_t)stages * stage_bytes;
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
  prm.pdl = (flags & QEFT_F_PDL) ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, prm);
  if (e != cudaSuccess) return (int)e;
  count_launch();
  return QEFT_OK;
}

template <int RG>
static int dispatch_gemv(GemvParams& prm, int ctas, bool stage, unsigned flags, cudaStream_t st) {
  if (prm.g128 == 1)
    return stage ? launch_gemv<RG, true, true>(prm, ctas, flags, st) : launch_gemv<RG, false, true>(prm, ctas, flags, st);
  return stage ? launch_gemv<RG, true, false>(prm, ctas, flags, st) : launch_gemv<RG, false, false>(prm, ctas, flags, st);
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
  if (r > 0 && ow_layout != QEFT_OW_PLAIN && ow_layout != QEFT_OW_INTERLEAVED) return QEFT_E_DTYPE;
  if (r == 0) ow_layout = QEFT_OW_NONE;
  if (r > 256) return QEFT_E_UNSUPPORTED;   // TODO(next): loop the outlier pieces
  if (!check_align16(x)) return QEFT_E_ALIGN;
  GemvParams prm = {};
  long total_rows = 0;
  for (int i = 0; i < nparts; ++i) {
    const qeft_gemv_part_t& q = parts[i];
    if (!q.qweight || !q.scales || !q.scaled_zeros || !q.y) return QEFT_E_NULL;
    if (r > 0 && !q.oweight) return QEFT_E_NULL;
    if (q.N <= 0 || q.N % 8 != 0) return QEFT_E_SHAPE;
    if (!check_align16(q.qweight) || !check_align16(q.scales) || !check_align16(q.scaled_zeros) ||
        (r > 0 && !check_align16(q.oweight)))
      return QEFT_E_ALIGN;
    total_rows += q.N;
  }
  // two 16-row groups per CTA once that still gives every SM at least two CTAs (halves the per-CTA x work)
  const int rg = (total_rows >= 2L * kNumSms * 32) ? 2 : 1;
  int ctas = 0;
  for (int i = 0; i < nparts; ++i) {
    const qeft_gemv_part_t& q = parts[i];
    GemvPart& d = prm.part[i];
    d.qw = static_cast<const uint8_t*>(q.qweight);
    d.scales = static_cast<const __half*>(q.scales);
    d.szeros = static_cast<const __half*>(q.scaled_zeros);
    d.ow = static_cast<const __half*>(q.oweight);
    d.bias = static_cast<const __half*>(q.bias);
    d.y = static_cast<__half*>(q.y);
    d.N = q.N;
    d.cta_begin = ctas;
    ctas += cdiv(q.N, 16 * rg);
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
  return rg == 2 ? dispatch_gemv<2>(prm, ctas, stage, flags, st) : dispatch_gemv<1>(prm, ctas, stage, flags, st);
}

extern "C" int qeft_gemv_w4(const void* x, const void* qweight, const void* scales, const void* scaled_zeros,
                            const void* oweight, int ow_layout, const void* bias, const int32_t* x_gather,
                            void* y, int m, int N, int K, int r, int G, unsigned flags, qeft_stream_t stream) {
  qeft_gemv_part_t part = {qweight, scales, scaled_zeros, oweight, bias, y, N};
  return qeft_gemv_w4_multi(x, &part, 1, ow_layout, x_gather, m, K, r, G, flags, stream);
}
