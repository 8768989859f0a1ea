// Prefill / fine-tune GEMM for the packed QEFT QuantLinear on tcgen05 tensor cores (sm_100a).
//
//     y[M, N] = x[M, K] . Wdense^T (+ bias)          M tokens, N output features, K input features
//
// Replaces gemm_4bit (qeft/kernel/quantization_new/gemm/gemm_cuda.cu:929-1033: cp.async + ldmatrix +
// mma.sync.m16n8k16, dequant between ldmatrix and mma) PLUS the separate cuBLAS GEMM for the outlier columns and
// the bias add of qeft/qlinear.py:264-268 -- one kernel, one pass over y.
//
// Design (DESIGN.md "GEMM"):
//   * The WEIGHTS are the A operand of the UMMA (M_umma = 128 output features) and live in TENSOR MEMORY: dequant
//     warps read the packed int4 bytes straight from global/L2 (every lane owns one output feature: 32 bytes =
//     64 columns per k-block), turn them into fp16 with the lop3 magic-number trick + one HFMA2 per pair
//     (w = fma(q, s, sz), the reference's single rounding) and write them with tcgen05.st.  Dequantised weights
//     never touch shared memory, so shared-memory bandwidth is left to the activation tiles.
//   * The ACTIVATIONS are the B operand (N_umma = 128 tokens): 128 x 64 fp16 tiles, K-major, 128-byte swizzle,
//     brought in by TMA (cp.async.bulk.tensor) into a 6-stage ring.
//   * One CTA computes 256 features x 128 tokens: two fp32 accumulators of 128 TMEM columns each (y^T tiles),
//     so one activation tile feeds two MMAs (halves the L2 -> SM activation traffic).
//   * The dense fp16 outlier columns are simply the last r/64 k-blocks: same pipeline, the dequant warps copy
//     oweight[f, 64 b .. 64 b + 63] to TMEM unchanged; the dead int4 columns K-r..K-1 are never read.
//   * Warp roles (384 threads): warp 0 TMA producer, warp 1 MMA issuer (one elected lane, tcgen05.commit to
//     mbarriers), warp 2 TMEM allocator, warps 4-11 dequant (two warpgroups, one per 128-feature block) and then
//     epilogue (tcgen05.ld, + bias, fp16, transposed through shared memory, 16-byte stores).
#include "common.cuh"

#include <cuda.h>

namespace qeft {

constexpr int kBM = 256;          // output features per CTA (two UMMA M = 128 blocks)
constexpr int kBN = 128;          // tokens per CTA (UMMA N)
constexpr int kBK = 64;           // input columns per k-block (one 128-byte swizzle row of fp16)
constexpr int kXStages = 6;       // activation ring, 16 KB each
constexpr int kAStages = 4;       // weight ring in TMEM, 64 columns each (2 blocks x 32 columns = 64 fp16 per lane)
constexpr int kGemmThreads = 384;
constexpr int kXStageBytes = kBN * kBK * 2;
constexpr int kTmemCols = 512;
constexpr int kTmemA0 = 256;      // first TMEM column of the weight ring (accumulators: 0..127, 128..255)

struct GemmParams {
  const uint8_t* qw;       // int16 [N/4, K] as bytes
  const __half* scales;    // [K/G, N]
  const __half* szeros;    // [K/G, N]
  const __half* ow;        // [N, r] or null
  const __half* bias;      // [N] or null
  __half* y;               // [M, N]
  int M, N, K, r, G;
  int nkb_q;               // int4 k-blocks = (K - r) / 64
  int nkb;                 // + outlier k-blocks r / 64
};

// ---- PTX wrappers ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]   (A: 128 lanes x 16 fp16 = 8 columns; B: K-major 128-byte-swizzled tile)
__device__ __forceinline__ void umma_ts_f16(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
        "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
        "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 format): rows of 128 bytes, 8-row groups 1024
// bytes apart.  Advancing by one UMMA_K (16 fp16 = 32 bytes) inside the swizzle row adds 2 to the address field.
__device__ __forceinline__ uint64_t make_b_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);          // start address, bits [0, 14)
  d |= (uint64_t)1 << 16;                           // leading byte offset (ignored for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                 // stride byte offset, bits [32, 46)
  d |= (uint64_t)1 << 46;                           // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                           // SWIZZLE_128B
  return d;
}

// kind::f16 instruction descriptor: D = F32, A = B = F16, both K-major, M = 128, N = kBN
constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(kBN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_w4_kernel(const __grid_constant__ CUtensorMap xmap, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * kXStages + 2 * kAStages + 1];
  __shared__ uint32_t s_tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t xs0 = (smem_addr(smem_raw) + 1023u) & ~1023u;          // 1024-byte aligned stage ring
  uint8_t* xs_gen = smem_raw + (xs0 - smem_addr(smem_raw));
  const uint32_t bar0 = smem_addr(bars);
  auto x_full = [&](int s) { return bar0 + 8 * s; };
  auto x_empty = [&](int s) { return bar0 + 8 * (kXStages + s); };
  auto a_full = [&](int s) { return bar0 + 8 * (2 * kXStages + s); };
  auto a_empty = [&](int s) { return bar0 + 8 * (2 * kXStages + kAStages + s); };
  const uint32_t acc_full = bar0 + 8 * (2 * kXStages + 2 * kAStages);

  const int tok0 = blockIdx.x * kBN;
  const int n0 = blockIdx.y * kBM;
  const int nrb = (p.N - n0) >= kBM ? 2 : 1;          // 128-feature blocks of this tile (N % 128 == 0)
  const int nkb = p.nkb;

  if (tid == 0) {
    for (int s = 0; s < kXStages; ++s) { mbar_init(x_full(s), 1); mbar_init(x_empty(s), 1); }
    for (int s = 0; s < kAStages; ++s) { mbar_init(a_full(s), 4 * nrb); mbar_init(a_empty(s), 1); }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&s_tmem_base)), "n"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem_base;
  pdl_launch_dependents();

  if (warp == 0) {
    // ================= TMA producer: activation tiles =================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&xmap) : "memory");
      pdl_wait();                                   // x is the previous kernel's output
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % kXStages, use = kb / kXStages;
        if (use > 0) mbar_wait(x_empty(s), (uint32_t)((use - 1) & 1));
        const int k0 = kb < p.nkb_q ? kb * kBK : p.K - p.r + (kb - p.nkb_q) * kBK;
        mbar_expect_tx(x_full(s), kXStageBytes);
        tma_load_2d(xs0 + s * kXStageBytes, &xmap, k0, tok0, x_full(s));
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % kXStages, as = kb % kAStages;
        mbar_wait(x_full(s), (uint32_t)((kb / kXStages) & 1));
        mbar_wait(a_full(as), (uint32_t)((kb / kAStages) & 1));
        tc_fence_after();
#pragma unroll
        for (int k16 = 0; k16 < kBK / 16; ++k16) {
          const uint64_t bdesc = make_b_desc(xs0 + s * kXStageBytes + k16 * 32);
          for (int rb = 0; rb < nrb; ++rb)
            umma_ts_f16(tmem + 128 * rb, tmem + kTmemA0 + 64 * as + 32 * rb + 8 * k16, bdesc, kIdesc,
                        (uint32_t)((kb | k16) != 0));
        }
        tc_commit(x_empty(s));          // both rings are free once these MMAs have read them
        tc_commit(a_empty(as));
      }
      tc_commit(acc_full);
    }
  } else if (warp >= 4) {
    // ================= dequant warps (then epilogue) =================
    const int dw = warp - 4, rb = dw >> 2, quad = dw & 3;
    if (rb < nrb) {
      const int f = n0 + 128 * rb + 32 * quad + lane;            // this lane's output feature
      const uint8_t* qrow = p.qw + (size_t)(f >> 2) * (size_t)(2 * p.K) + (size_t)((f & 3) * 32);
      const __half* owrow = p.ow ? p.ow + (size_t)f * p.r : nullptr;
      const uint32_t lane_taddr = tmem + ((uint32_t)(32 * quad) << 16);
      // Register prefetch ring, kPF k-blocks deep (weights come from L2: ~700 cycles away, one k-block is ~500
      // cycles of MMA), scales one group ahead, the first outlier block a few k-blocks before it is needed.
      constexpr int kPF = 4;
      uint4 ring[kPF][2];
#pragma unroll
      for (int i = 0; i < kPF; ++i) {
        ring[i][0] = ring[i][1] = make_uint4(0, 0, 0, 0);
        if (i < p.nkb_q) {
          ring[i][0] = ldg_nc_v4(qrow + (size_t)i * 128);
          ring[i][1] = ldg_nc_v4(qrow + (size_t)i * 128 + 16);
        }
      }
      const int kb_per_grp = p.G / kBK;
      auto load_scale = [&](int grp, uint32_t& s2o, uint32_t& z2o) {
        const unsigned short sh = ldg_nc_u16(p.scales + (size_t)grp * p.N + f);
        const unsigned short zh = ldg_nc_u16(p.szeros + (size_t)grp * p.N + f);
        s2o = (uint32_t)sh | ((uint32_t)sh << 16);
        z2o = (uint32_t)zh | ((uint32_t)zh << 16);
      };
      uint32_t s2 = 0, z2 = 0, s2n = 0, z2n = 0;
      const int ngrp_q = (p.nkb_q + kb_per_grp - 1) / kb_per_grp;
      if (ngrp_q > 0) load_scale(0, s2n, z2n);
      uint32_t ob[32];                                 // one outlier k-block (64 fp16 of this lane's feature)
      auto load_outlier = [&](int oblk) {
        const __half* src = owrow + (size_t)oblk * kBK;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint4 t4 = ldg_nc_v4(src + 8 * i);
          ob[4 * i + 0] = t4.x; ob[4 * i + 1] = t4.y; ob[4 * i + 2] = t4.z; ob[4 * i + 3] = t4.w;
        }
      };
      const int ob_issue_kb = p.nkb_q > 3 ? p.nkb_q - 3 : 0;     // when the first outlier block's loads are issued
      auto publish = [&](const uint32_t (&v)[32], int kb) {
        const int as = kb % kAStages, use = kb / kAStages;
        if (use > 0) mbar_wait(a_empty(as), (uint32_t)((use - 1) & 1));
        tc_fence_after();
        tmem_st32(lane_taddr + (uint32_t)(kTmemA0 + 64 * as + 32 * rb), v);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a_full(as));
      };
      // ---- int4 k-blocks, kPF per trip so that the ring slots are compile-time registers ----
      for (int kb0 = 0; kb0 < p.nkb_q; kb0 += kPF) {
#pragma unroll
        for (int i = 0; i < kPF; ++i) {
          const int kb = kb0 + i;
          if (kb < p.nkb_q) {
            const uint4 c0 = ring[i][0], c1 = ring[i][1];
            if (kb + kPF < p.nkb_q) {
              ring[i][0] = ldg_nc_v4(qrow + (size_t)(kb + kPF) * 128);
              ring[i][1] = ldg_nc_v4(qrow + (size_t)(kb + kPF) * 128 + 16);
            }
            if (kb % kb_per_grp == 0) {
              s2 = s2n; z2 = z2n;
              const int g1 = kb / kb_per_grp + 1;
              if (g1 < ngrp_q) load_scale(g1, s2n, z2n);
            }
            if (p.ow && kb == ob_issue_kb) load_outlier(0);
            uint32_t v[32];
            const uint32_t w[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                uint32_t hq[4];
                unpack_word_to_half2(w[4 * h + c], hq);          // pairs k = 32 h + 2 c + 8 j (+1), exact 0..15
#pragma unroll
                for (int j = 0; j < 4; ++j) v[16 * h + c + 4 * j] = hfma2_u32(hq[j], s2, z2);   // w = fma(q, s, sz)
              }
            publish(v, kb);
          }
        }
      }
      // ---- outlier k-blocks: fp16 columns, copied unchanged ----
      if (p.ow && p.nkb_q == 0) load_outlier(0);
      for (int kb = p.nkb_q; kb < nkb; ++kb) {
        uint32_t v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = ob[i];
        if (kb + 1 < nkb) load_outlier(kb + 1 - p.nkb_q);
        publish(v, kb);
      }

      // ---- epilogue: y^T tile (lane = feature, columns = tokens) -> + bias -> fp16 -> y[token, feature] ----
      mbar_wait(acc_full, 0);
      tc_fence_after();
      const float bias = p.bias ? __half2float(p.bias[f]) : 0.f;
      __half* stage = reinterpret_cast<__half*>(xs_gen + dw * 2048);         // 32 tokens x 32 features per warp
      const int fw = n0 + 128 * rb + 32 * quad;                               // first feature of this warp
#pragma unroll 1
      for (int tc = 0; tc < kBN / 32; ++tc) {
        uint32_t acc[32];
        tmem_ld32(lane_taddr + (uint32_t)(128 * rb + 32 * tc), acc);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 32; ++i) stage[i * 32 + lane] = __float2half_rn(__uint_as_float(acc[i]) + bias);
        __syncwarp();
        // 32 rows (tokens) of 64 bytes: 4 lanes per row, 8 rows per pass
#pragma unroll
        for (int pass = 0; pass < 4; ++pass) {
          const int row = pass * 8 + (lane >> 2), piece = lane & 3;
          const int tok = tok0 + 32 * tc + row;
          const uint4 val = *reinterpret_cast<const uint4*>(stage + row * 32 + piece * 8);
          if (tok < p.M) *reinterpret_cast<uint4*>(p.y + (size_t)tok * p.N + fw + piece * 8) = val;
        }
        __syncwarp();
      }
      tc_fence_before();
    }
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols) : "memory");
  }
}

// ---- host ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// 2-D map of a row-major fp16 matrix [rows, cols]: box = 64 columns (128 bytes, swizzled) x box_rows rows
int make_tmap_f16_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn fn = encode_tiled();
  if (!fn) return QEFT_E_UNSUPPORTED;
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {cols * 2};
  const cuuint32_t box[2] = {64, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return rc == CUDA_SUCCESS ? QEFT_OK : QEFT_E_UNSUPPORTED;
}

}  // namespace qeft

using namespace qeft;

extern "C" int qeft_gemm_w4(const void* x, const void* qweight, const void* scales, const void* scaled_zeros,
                            const void* oweight, const void* bias, void* y, int M, int N, int K, int r, int G,
                            int dtype, unsigned flags, qeft_stream_t stream) {
  if (!x || !qweight || !scales || !scaled_zeros || !y) return QEFT_E_NULL;
  if (dtype != QEFT_DT_F16) return dtype == QEFT_DT_BF16 ? QEFT_E_UNSUPPORTED : QEFT_E_DTYPE;
  if (G == -1) G = K;
  if (M <= 0 || N <= 0 || K <= 0 || N % 128 != 0 || K % 64 != 0 || G <= 0 || G % 64 != 0 || K % G != 0) return QEFT_E_SHAPE;
  if (r < 0 || r % 64 != 0 || r >= K) return QEFT_E_SHAPE;
  if (r > 0 && !oweight) return QEFT_E_NULL;
  if (!check_align16(x) || !check_align16(qweight) || !check_align16(y) || (r > 0 && !check_align16(oweight))) return QEFT_E_ALIGN;
  CUtensorMap xmap;
  int st = make_tmap_f16_2d(&xmap, x, (uint64_t)M, (uint64_t)K, kBN);
  if (st != QEFT_OK) return st;
  GemmParams prm;
  prm.qw = static_cast<const uint8_t*>(qweight);
  prm.scales = static_cast<const __half*>(scales);
  prm.szeros = static_cast<const __half*>(scaled_zeros);
  prm.ow = r > 0 ? static_cast<const __half*>(oweight) : nullptr;
  prm.bias = static_cast<const __half*>(bias);
  prm.y = static_cast<__half*>(y);
  prm.M = M; prm.N = N; prm.K = K; prm.r = r; prm.G = G;
  prm.nkb_q = (K - r) / kBK;
  prm.nkb = prm.nkb_q + r / kBK;
  const size_t smem = (size_t)kXStages * kXStageBytes + 1024;
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(gemm_w4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)cdiv(M, kBN), (unsigned)cdiv(N, kBM));
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (flags & QEFT_F_PDL) ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_w4_kernel, xmap, prm);
  if (e != cudaSuccess) return (int)e;
  count_launch();
  return QEFT_OK;
}

extern "C" int qeft_gemm_w4_dx(const void*, const void*, const void*, const void*, const void*, void*, int, int, int, int,
                               int, int, unsigned, qeft_stream_t) { return QEFT_E_UNSUPPORTED; }
extern "C" int qeft_dow(const void*, const void*, float*, int, int, int, int, int, int, unsigned, qeft_stream_t) {
  return QEFT_E_UNSUPPORTED;
}
