// Backward of the packed QEFT QuantLinear on tcgen05 tensor cores (sm_100a).
//
//   dx[M, K]  = dy[M, N] . Wdense[N, K]          (contraction over the N output features)
//   dow[N, r] = dy[M, N]^T . x[M, K-r : K]       (gradient of the trainable fp16 outlier columns, fp32)
//
// The math BASELINE.json defines for QuantMatMulQEFT.backward; the reference's own backward
// (qeft/qlinear.py:28-44) calls the forward GEMM on dy and is not usable.
//
// dX (DESIGN.md "Backward"): the contraction index is the ROW of the packed weight, so the dequantised tile is
// used as an MN-major B operand: a 16-byte packed chunk is 32 consecutive k of one row n, i.e. 64 contiguous
// bytes of fp16 in the B tile's row n -- exactly the MN-major SWIZZLE_128B canonical layout (128-byte rows of
// 64 consecutive k).  512 dequant threads each convert one chunk per k-block (lop3 magic number + HFMA2, the
// same arithmetic as the forward) and store it with four 16-byte swizzled stores; dy tiles arrive by TMA
// (K-major A operand).  One CTA computes 256 tokens x 256 input features: two fp32 accumulators of 256 TMEM
// columns (all of tensor memory), so one dequantised tile feeds both token blocks.  The outlier columns
// K-r..K-1 take their B rows from oweight instead of the (dead) int4 columns.
//
// dOW: a plain dense GEMM, both operands MN-major straight from their row-major tensors by TMA
// (A = dy[tok, n0 .. n0+127], B = x_out[tok, 0 .. r-1]), fp32 accumulate, fp32 store / accumulate.
#include "tc_common.cuh"

#include <stdlib.h>

namespace qeft {

// ======================================================================================================
// dX
// ======================================================================================================
constexpr int kDxBK = 64;                 // contraction (n) per k-block
constexpr int kDxBF = 256;                // input features (k) per CTA = UMMA N
constexpr int kDxTB = 2;                  // 128-token blocks per CTA
constexpr int kDxStages = 3;
constexpr int kDxABytes = kDxTB * 128 * kDxBK * 2;     // dy tiles of a stage (32 KB)
constexpr int kDxBBytes = kDxBK * kDxBF * 2;           // dequantised weight tile of a stage (32 KB)
constexpr int kDxStageBytes = kDxABytes + kDxBBytes;
constexpr int kDxDequantWarps = 16;
constexpr int kDxThreads = (4 + kDxDequantWarps) * 32;
constexpr int kDxMaxSplits = 4;           // CTAs that may share a tile (contraction split)
constexpr int kDxSplitCost = 23;          // fixed cost of a split CTA (partial round trip) in k-block times

struct DxParams {
  const uint8_t* qw;
  const __half* scales;
  const __half* szeros;
  const __half* ow;        // [N, r] or null
  __half* dx;              // [M, K]
  int M, N, K, r, G;
  int tbc;                 // 128-token blocks per CTA: 2, or 1 where that fills the SMs' waves better
  // The grid is a 1-D list of CTAs over the tiles (token tile fastest): the first t_main CTAs compute whole tiles, the
  // others come in groups of `splits` that share one of the remaining tiles, each over a contiguous range of the
  // contraction (launches of few tiles: t_main = 0; launches whose last wave is nearly empty, e.g. 13B K = 5120 with 160
  // tiles on 148 SMs: only the tiles of that wave).  fp32 partials [split][M][K] and one arrival counter per tile; the
  // CTA of a tile that arrives last adds the partials in split order (deterministic) and stores dx; the counters reset
  // themselves.
  int ttok, t_main, splits;
  float* ws;
  unsigned* counters;
};

template <bool BF16>
__global__ void __launch_bounds__(kDxThreads, 1)
gemm_w4_dx_kernel(const __grid_constant__ CUtensorMap dymap, const __grid_constant__ CUtensorMap owmap, const DxParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * kDxStages + 1];
  __shared__ uint32_t s_tmem_base;
  __shared__ uint32_t s_last;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t st0 = (smem_addr(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar0 = smem_addr(bars);
  auto full = [&](int s) { return bar0 + 8 * s; };
  auto empty = [&](int s) { return bar0 + 8 * (kDxStages + s); };
  const uint32_t acc_full = bar0 + 8 * (2 * kDxStages);

  int tile = (int)blockIdx.x, zsp = 0, nsplit = 1;
  if (tile >= p.t_main) {
    const int q = tile - p.t_main;
    tile = p.t_main + q / p.splits;
    zsp = q - (q / p.splits) * p.splits;
    nsplit = p.splits;
  }
  const int tok0 = (tile % p.ttok) * (128 * p.tbc);
  const int kf0 = (tile / p.ttok) * kDxBF;
  // this CTA's k-blocks of the contraction: [kb_lo, kb_lo + nkb) (the whole of N unless the launch is split)
  const int nkb_all = p.N / kDxBK;
  const int kb_lo = (nkb_all * zsp) / nsplit;
  const int nkb = (nkb_all * (zsp + 1)) / nsplit - kb_lo;
  const int ntb = (p.M - tok0) > 128 ? p.tbc : 1;   // token blocks with at least one live token

  // The 16 dequant warps form two sets that alternate k-blocks; a thread converts TWO chunks (rows nl, nl + 32) of its
  // set's k-blocks: per k-block the fixed latency of the hand-over (stage-free wait, proxy fence, arrival) is paid by one
  // set while the other converts the next block.  The fp16 outlier columns K-r..K-1 are not converted at all: their
  // rows of `oweight` ARE 128-byte rows of the MN-major swizzled tile, so the TMA producer drops them in place.
  const int o_begin = p.ow != nullptr ? max(kf0, p.K - p.r) : p.K;        // outlier features of this tile: [o_begin, o_end)
  const int o_end = min(kf0 + kDxBF, p.K);
  if (tid == 0) {
    for (int s = 0; s < kDxStages; ++s) { mbar_init(full(s), 1 + kDxDequantWarps / 2); mbar_init(empty(s), 1); }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&s_tmem_base)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem_base;
  pdl_launch_dependents();

  if (warp == 0) {
    // ================= TMA producer: dy tiles (A operand, K-major) =================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&dymap) : "memory");
      pdl_wait();
      for (int i = 0; i < nkb; ++i) {
        const int kb = kb_lo + i;
        const int s = i % kDxStages, use = i / kDxStages;
        if (use > 0) mbar_wait(empty(s), (uint32_t)((use - 1) & 1));
        const int nob = o_begin < o_end ? (o_end - o_begin) / 64 : 0;           // 64-feature boxes of oweight in this tile
        mbar_expect_tx(full(s), (uint32_t)(ntb * 128 * kDxBK * 2 + nob * (kDxBK * 128)));
        for (int tb = 0; tb < ntb; ++tb)
          tma_load_2d(st0 + s * kDxStageBytes + tb * (128 * kDxBK * 2), &dymap, kb * kDxBK, tok0 + 128 * tb, full(s));
        for (int ob = 0; ob < nob; ++ob) {
          const int f = o_begin + 64 * ob;                                  // rows kb * 64 .. + 63 of oweight, columns f - (K - r) .. + 63
          tma_load_2d(st0 + s * kDxStageBytes + kDxABytes + ((f - kf0) >> 6) * (kDxBK * 128), &owmap, f - (p.K - p.r), kb * kDxBK, full(s));
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(kDxBF, false, true) | (BF16 ? ((1u << 7) | (1u << 10)) : 0u);
      // (this one thread issues every MMA of the CTA: its loop is kept free of descriptor arithmetic -- a descriptor's
      // low field is the shared-memory address / 16, so a stage's descriptors are base + constant -- and of runtime trip
      // counts)
      const uint64_t adesc0 = make_sw128_desc(st0, 16, 1024), bdesc0 = make_sw128_desc(st0 + kDxABytes, kDxBK * 128, 1024);
      const bool two = ntb == 2;
      int s = 0;
      uint32_t par = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(full(s), par);
        tc_fence_after();
        const uint64_t a = adesc0 + (uint64_t)((s * kDxStageBytes) >> 4), b = bdesc0 + (uint64_t)((s * kDxStageBytes) >> 4);
#pragma unroll
        for (int k16 = 0; k16 < kDxBK / 16; ++k16) {
          // B: rows = n (k index of the MMA), 16 rows further per UMMA_K; 64-feature chunks 8 KB apart
          const uint32_t accf = (uint32_t)((kb | k16) != 0);
          umma_ss_f16(tmem, a + (uint64_t)((k16 * 32) >> 4), b + (uint64_t)((k16 * 16 * 128) >> 4), idesc, accf);
          if (two) umma_ss_f16(tmem + 256, a + (uint64_t)((128 * kDxBK * 2 + k16 * 32) >> 4), b + (uint64_t)((k16 * 16 * 128) >> 4), idesc, accf);
        }
        tc_commit(empty(s));
        if (++s == kDxStages) { s = 0; par ^= 1u; }
      }
      tc_commit(acc_full);
    }
  } else if (warp >= 4) {
    // ================= dequant warps: one packed chunk (row n, 32 features) per thread and k-block =================
    const int dt = tid - 128;                 // 0..511
    {
      const int set = dt >> 8, d2 = dt & 255;
      // (same quarter-warp composition as below: {both halves of a 64-feature chunk} x {the 4 rows of one qweight row})
      const int kc = (d2 & 1) | (((d2 >> 3) & 3) << 1), nlb = ((d2 >> 1) & 3) | ((d2 >> 5) << 2);   // rows nlb, nlb + 32
      const int kf = kf0 + 32 * kc;
      const bool live = kf < p.K;
      const bool outl = kf >= o_begin && kf < o_end;      // (the TMA producer fills these chunks)
      const int grp = live ? kf / p.G : 0;
      const size_t in_row = (size_t)(kf >> 6) * 128 + (size_t)(((kf >> 5) & 1) * 16);
      const uint32_t dst_row = (uint32_t)((kc >> 1) * (kDxBK * 128) + nlb * 128);
      const int sw = nlb & 7;
      constexpr int kPF2 = 3;                 // the set's k-blocks in flight (= 6 k-blocks of the launch)
      auto load_q = [&](int kb, int h, uint4& q, uint32_t& sz) {
        const int n = kb * kDxBK + nlb + 32 * h;
        q = ldg_nc_v4(p.qw + (size_t)(n >> 2) * (size_t)(2 * p.K) + (size_t)((n & 3) * 32) + in_row);
        sz = (uint32_t)ldg_nc_u16(p.scales + (size_t)grp * p.N + n) | ((uint32_t)ldg_nc_u16(p.szeros + (size_t)grp * p.N + n) << 16);
      };
      uint4 ring[kPF2][2];
      uint32_t rsz[kPF2][2];
#pragma unroll
      for (int i = 0; i < kPF2; ++i)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          ring[i][h] = make_uint4(0, 0, 0, 0); rsz[i][h] = 0;
          if (live && !outl && set + 2 * i < nkb) load_q(kb_lo + set + 2 * i, h, ring[i][h], rsz[i][h]);
        }
      bool done = false;
      for (int i0 = 0; !done; i0 += kPF2) {
#pragma unroll
        for (int u = 0; u < kPF2; ++u) {
          const int kb = set + 2 * (i0 + u);          // (relative to kb_lo)
          if (kb >= nkb) { done = true; break; }
          const int s = kb % kDxStages, use = kb / kDxStages;
          const uint32_t base = st0 + s * kDxStageBytes + kDxABytes + dst_row;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t v[16];
            if (!live) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = 0u;
            } else if (!outl) {
              const uint4 q = ring[u][h];
              const uint32_t sz = rsz[u][h];
              if (kb + 2 * kPF2 < nkb) load_q(kb_lo + kb + 2 * kPF2, h, ring[u][h], rsz[u][h]);
              const uint32_t s2 = (sz & 0xffffu) | (sz << 16), z2 = (sz >> 16) | (sz & 0xffff0000u);
              const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                uint32_t hq[4];
                unpack_word_to_half2(w[c], hq);                        // pairs k = 2 c + 8 j (+1), exact 0..15
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  v[c + 4 * j] = BF16 ? dequant_pair_bf16(hq[j], __half2float(__ushort_as_half((unsigned short)(sz & 0xffffu))),
                                                          __half2float(__ushort_as_half((unsigned short)(sz >> 16))))
                                      : hfma2_u32(hq[j], s2, z2);                        // w = fma(q, s, sz)
              }
            }
            if (h == 0 && use > 0) mbar_wait(empty(s), (uint32_t)((use - 1) & 1));
            if (outl) continue;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint32_t a = base + (uint32_t)(h * 32 * 128) + (uint32_t)(((((kc & 1) << 2) + i) ^ sw) << 4);
              asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v[4 * i]), "r"(v[4 * i + 1]), "r"(v[4 * i + 2]),
                           "r"(v[4 * i + 3]) : "memory");
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(full(s));
        }
      }
    }

    // ---- epilogue: TMEM lanes = tokens, columns = features.  A warp owns 32 tokens x 128 features; it passes 128-byte
    // pieces of its rows through a private 4 KB staging tile (the ring is idle once acc_full has fired; 16-byte slots
    // XOR-swizzled by the row, conflict-free both ways) so that a store instruction writes 4 whole 128-byte lines of
    // dx (8 lanes per token row) instead of 16 bytes in each of 32 rows ----
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int dw = warp - 4, quad = dw & 3, tb = (dw >> 2) & 1, fh = dw >> 3;
    if (tb < ntb) {
      const uint32_t lane_taddr = tmem + ((uint32_t)(32 * quad) << 16) + (uint32_t)(256 * tb + 128 * fh);
      const uint32_t stg = st0 + (uint32_t)dw * 4096u;
      const int trow0 = tok0 + 128 * tb + 32 * quad;           // first token of this warp
      auto put = [&](int slot, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {     // this lane's row, 16-byte slot
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(stg + (uint32_t)(lane * 128) + (uint32_t)(((slot ^ lane) & 7) << 4)),
                     "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
      };
      // 32 rows x 128 bytes of the staging tile -> global rows of `pitch` bytes starting at `dst`
      auto flush = [&](uint8_t* dst, size_t pitch) {
        __syncwarp();
#pragma unroll
        for (int pass = 0; pass < 8; ++pass) {
          const int row = pass * 4 + (lane >> 3), piece = lane & 7;
          uint4 v;
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                       : "r"(stg + (uint32_t)(row * 128) + (uint32_t)(((piece ^ row) & 7) << 4)) : "memory");
          if (trow0 + row < p.M) *reinterpret_cast<uint4*>(dst + (size_t)row * pitch + (size_t)(piece * 16)) = v;
        }
        __syncwarp();
      };
      auto pack2 = [](uint32_t a, uint32_t b) {
        if (BF16) {
          const __nv_bfloat162 v = __floats2bfloat162_rn(__uint_as_float(a), __uint_as_float(b));
          return *reinterpret_cast<const uint32_t*>(&v);
        }
        const __half2 v = __floats2half2_rn(__uint_as_float(a), __uint_as_float(b));
        return *reinterpret_cast<const uint32_t*>(&v);
      };
      if (nsplit > 1) {
        // fp32 partials: one 32-column chunk = 128 bytes per token
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t acc[32];
          tmem_ld32(lane_taddr + 32 * c, acc);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          const int f0 = kf0 + 128 * fh + 32 * c;
          if (f0 >= p.K) break;                                   // (warp-uniform; K % 64 == 0)
#pragma unroll
          for (int j = 0; j < 8; ++j) put(j, acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
          flush(reinterpret_cast<uint8_t*>(p.ws + ((size_t)zsp * (size_t)p.M + (size_t)trow0) * (size_t)p.K + f0),
                (size_t)p.K * sizeof(float));
        }
      } else {
        // fp16 / bf16 result: two 32-column chunks = 128 bytes per token
#pragma unroll 1
        for (int cp = 0; cp < 2; ++cp) {
          const int f0 = kf0 + 128 * fh + 64 * cp;
          if (f0 >= p.K) break;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t acc[32];
            tmem_ld32(lane_taddr + 64 * cp + 32 * h, acc);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 4; ++i)
              put(4 * h + i, pack2(acc[8 * i + 0], acc[8 * i + 1]), pack2(acc[8 * i + 2], acc[8 * i + 3]),
                  pack2(acc[8 * i + 4], acc[8 * i + 5]), pack2(acc[8 * i + 6], acc[8 * i + 7]));
          }
          flush(reinterpret_cast<uint8_t*>(p.dx + (size_t)trow0 * (size_t)p.K + f0), (size_t)p.K * sizeof(__half));
        }
      }
    }
    tc_fence_before();
    if (nsplit > 1) {
      // the split that arrives last at its tile adds the partials (split order) and stores the tile
      __threadfence();
      asm volatile("bar.sync 1, %0;" ::"n"(kDxDequantWarps * 32) : "memory");
      if (tid == 128) s_last = (atomicAdd(p.counters + tile, 1u) == (unsigned)(nsplit - 1)) ? 1u : 0u;
      asm volatile("bar.sync 1, %0;" ::"n"(kDxDequantWarps * 32) : "memory");
      if (s_last) {
        __threadfence();
        const int ntok = min(128 * p.tbc, p.M - tok0), nf4 = min(kDxBF, p.K - kf0) / 4;
        const size_t part = (size_t)p.M * (size_t)p.K;
        // (one SM reads nsplit x 256 KB from L2: 4 positions x up to 4 splits of loads in flight per thread)
        constexpr int kU = 4, kStride = kDxDequantWarps * 32;
        const int total = ntok * (kDxBF / 4);
#pragma unroll 1
        for (int base = tid - 128; base < total; base += kU * kStride) {
          size_t off[kU];
          bool ok[kU];
          float4 v[kU][kDxMaxSplits];
#pragma unroll
          for (int u = 0; u < kU; ++u) {
            const int idx = base + u * kStride;
            const int t = idx / (kDxBF / 4), c4 = idx % (kDxBF / 4);
            ok[u] = idx < total && c4 < nf4;
            off[u] = (size_t)(tok0 + t) * (size_t)p.K + (size_t)(kf0 + 4 * c4);
#pragma unroll
            for (int z = 0; z < kDxMaxSplits; ++z)
              if (ok[u] && z < nsplit) v[u][z] = __ldcg(reinterpret_cast<const float4*>(p.ws + (size_t)z * part + off[u]));
          }
#pragma unroll
          for (int u = 0; u < kU; ++u) {
            if (!ok[u]) continue;
            float4 a = v[u][0];
#pragma unroll
            for (int z = 1; z < kDxMaxSplits; ++z)
              if (z < nsplit) { a.x += v[u][z].x; a.y += v[u][z].y; a.z += v[u][z].z; a.w += v[u][z].w; }
            uint2 o;
            if (BF16) {
              const __nv_bfloat162 lo = __floats2bfloat162_rn(a.x, a.y), hi = __floats2bfloat162_rn(a.z, a.w);
              o.x = *reinterpret_cast<const uint32_t*>(&lo); o.y = *reinterpret_cast<const uint32_t*>(&hi);
            } else {
              const __half2 lo = __floats2half2_rn(a.x, a.y), hi = __floats2half2_rn(a.z, a.w);
              o.x = *reinterpret_cast<const uint32_t*>(&lo); o.y = *reinterpret_cast<const uint32_t*>(&hi);
            }
            *reinterpret_cast<uint2*>(p.dx + off[u]) = o;
          }
        }
        if (tid == 128) p.counters[tile] = 0u;          // ready for the next launch (stream order)
      }
    }
  }
  __syncwarp();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
  }
}

// ======================================================================================================
// dOW
// ======================================================================================================
constexpr int kDwBK = 64;                 // tokens per k-block
constexpr int kDwStages = 4;
constexpr int kDwThreads = 192;           // warp 0 TMA, warp 1 MMA (+ TMEM alloc), warps 2-5 epilogue

struct DowParams {
  float* dow;              // [N, r]
  int M, N, r, accumulate, bf16;
};

// one CTA: 128 output features n0..n0+127 x all r (<= 256) outlier columns, contraction over all M tokens
__global__ void __launch_bounds__(kDwThreads, 1)
dow_kernel(const __grid_constant__ CUtensorMap dymap, const __grid_constant__ CUtensorMap xomap, const DowParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * kDwStages + 1];
  __shared__ uint32_t s_tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t st0 = (smem_addr(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar0 = smem_addr(bars);
  auto full = [&](int s) { return bar0 + 8 * s; };
  auto empty = [&](int s) { return bar0 + 8 * (kDwStages + s); };
  const uint32_t acc_full = bar0 + 8 * (2 * kDwStages);
  const int n0 = blockIdx.x * 128;
  const int nkb = cdiv(p.M, kDwBK);
  const int rchunks = p.r / 64;                                 // 64-column chunks of the outlier block
  const uint32_t a_bytes = 2 * kDwBK * 128;                     // 128 features = two 64-wide chunks of 64 token rows
  const uint32_t stage_bytes = a_bytes + (uint32_t)rchunks * kDwBK * 128;

  if (tid == 0) {
    for (int s = 0; s < kDwStages; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&s_tmem_base)), "n"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem_base;
  pdl_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      pdl_wait();
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % kDwStages, use = kb / kDwStages;
        if (use > 0) mbar_wait(empty(s), (uint32_t)((use - 1) & 1));
        mbar_expect_tx(full(s), stage_bytes);
        const uint32_t base = st0 + s * stage_bytes;
        // rows = tokens (the MMA's k index), 128-byte rows of 64 consecutive features: MN-major tiles
        tma_load_2d(base, &dymap, n0, kb * kDwBK, full(s));
        tma_load_2d(base + kDwBK * 128, &dymap, n0 + 64, kb * kDwBK, full(s));
        for (int c = 0; c < rchunks; ++c) tma_load_2d(base + a_bytes + c * (kDwBK * 128), &xomap, 64 * c, kb * kDwBK, full(s));
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(p.r, true, true) | (p.bf16 ? ((1u << 7) | (1u << 10)) : 0u);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % kDwStages;
        mbar_wait(full(s), (uint32_t)((kb / kDwStages) & 1));
        tc_fence_after();
        const uint32_t a0 = st0 + s * stage_bytes, b0 = a0 + a_bytes;
#pragma unroll
        for (int k16 = 0; k16 < kDwBK / 16; ++k16) {
          const uint64_t adesc = make_sw128_desc(a0 + k16 * 16 * 128, kDwBK * 128, 1024);
          const uint64_t bdesc = make_sw128_desc(b0 + k16 * 16 * 128, kDwBK * 128, 1024);
          umma_ss_f16(tmem, adesc, bdesc, idesc, (uint32_t)((kb | k16) != 0));
        }
        tc_commit(empty(s));
      }
      tc_commit(acc_full);
    }
  } else {
    // epilogue: lane = feature n, columns = outlier column j
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int quad = warp & 3;                    // warps 2..5 -> quadrants 2, 3, 0, 1
    const int n = n0 + 32 * quad + lane;
    const uint32_t lane_taddr = tmem + ((uint32_t)(32 * quad) << 16);
    for (int c = 0; c < p.r / 32; ++c) {
      uint32_t acc[32];
      tmem_ld32(lane_taddr + 32 * c, acc);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      float* dst = p.dow + (size_t)n * p.r + 32 * c;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float4 o = make_float4(__uint_as_float(acc[4 * i]), __uint_as_float(acc[4 * i + 1]), __uint_as_float(acc[4 * i + 2]),
                               __uint_as_float(acc[4 * i + 3]));
        if (p.accumulate) {
          const float4 old = *reinterpret_cast<const float4*>(dst + 4 * i);
          o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
        }
        *reinterpret_cast<float4*>(dst + 4 * i) = o;
      }
    }
    tc_fence_before();
  }
  __syncwarp();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256) : "memory");
  }
}

template <typename Kern>
static int set_smem_once(Kern kern, size_t smem, bool (&done)[64]) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !done[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    if (dev >= 0 && dev < 64) done[dev] = true;
  }
  return QEFT_OK;
}

// Contraction split of a dX launch: cost in k-block times = sum over its waves of (k-blocks per CTA + the fixed cost of a
// CTA); a split must win by 3 %.  Measured on B200: a k-block 0.75 us, an unsplit CTA ~1.5 us on top = 2 k-blocks, a
// split CTA 17 us = 23 (the fp32 partial round trip).  Splitting every tile pays for launches of few tiles (M <= 512:
// 1.4-2.5 x faster); at M = 2048 only the tiles of a nearly empty last wave are split (13B K = 5120: 160 tiles; 7B
// K = 11008: 344).  force_all / force_tail > 1: every tile / the last wave's tiles that many ways (experiments).
// Out: splits (CTAs per shared tile) and t_main (tiles computed whole, first in the grid).
static void dx_plan(int tiles, int nkb, int nsm, size_t mk, int force_all, int force_tail, bool never, int* splits_out, int* t_main_out) {
  int splits = 1, t_main = tiles;
  const int sp_max = nkb / 12 < kDxMaxSplits ? nkb / 12 : kDxMaxSplits;     // (a split keeps at least 12 k-blocks)
  const bool has_tail = tiles > nsm && tiles % nsm != 0;
  if (force_tail > 1) {
    if (has_tail && sp_max >= 2) { splits = force_tail < sp_max ? force_tail : sp_max; t_main = tiles - tiles % nsm; }
  } else if (force_all > 1) {
    if (sp_max >= 2) { splits = force_all < sp_max ? force_all : sp_max; t_main = 0; }
  } else if (!never) {
    long best = (long)cdiv(tiles, nsm) * (nkb + 2) * 100;
    for (int sp = 2; sp <= sp_max; ++sp) {
      const long part = cdiv(nkb, sp) + kDxSplitCost;
      const long all = (long)cdiv(tiles * sp, nsm) * part * 103;
      if (all < best) { best = all * 100 / 103; splits = sp; t_main = 0; }
      if (has_tail) {
        const long tail = ((long)(tiles / nsm) * (nkb + 2) + (long)cdiv((tiles % nsm) * sp, nsm) * part) * 103;
        if (tail < best) { best = tail * 100 / 103; splits = sp; t_main = tiles - tiles % nsm; }
      }
    }
  }
  // (one arrival counter per tile; at most 1 GB of fp32 partials)
  if (tiles > kSplitCounters || (size_t)splits * mk * sizeof(float) > ((size_t)1 << 30)) { splits = 1; t_main = tiles; }
  *splits_out = splits;
  *t_main_out = t_main;
}

}  // namespace qeft

using namespace qeft;

extern "C" int qeft_gemm_w4_dx_plan(int M, int N, int K, int sm_count, int* splits, int* whole_tiles, int* ctas) {
  if (!splits || !whole_tiles || !ctas) return QEFT_E_NULL;
  if (M <= 0 || N <= 0 || K <= 0 || N % 128 != 0 || K % 64 != 0 || sm_count <= 0) return QEFT_E_SHAPE;
  const int tiles = cdiv(M, 256) * cdiv(K, kDxBF);
  dx_plan(tiles, N / kDxBK, sm_count, (size_t)M * (size_t)K, 0, 0, false, splits, whole_tiles);
  *ctas = *whole_tiles + (tiles - *whole_tiles) * *splits;
  return QEFT_OK;
}

extern "C" int qeft_gemm_w4_dx(const void* dy, const void* qweight, const void* scales, const void* scaled_zeros,
                               const void* oweight, void* dx, int M, int N, int K, int r, int G, int dtype,
                               unsigned flags, qeft_stream_t stream) {
  if (!dy || !qweight || !scales || !scaled_zeros || !dx) return QEFT_E_NULL;
  if (dtype != QEFT_DT_F16 && dtype != QEFT_DT_BF16) return QEFT_E_DTYPE;
  if (G == -1) G = K;
  if (M <= 0 || N <= 0 || K <= 0 || N % 128 != 0 || K % 64 != 0 || G <= 0 || G % 64 != 0 || K % G != 0) return QEFT_E_SHAPE;
  if (r < 0 || r % 64 != 0 || r >= K) return QEFT_E_SHAPE;
  if (r > 0 && !oweight) return QEFT_E_NULL;
  if (!check_align16(dy) || !check_align16(qweight) || !check_align16(dx) || (r > 0 && !check_align16(oweight))) return QEFT_E_ALIGN;
  CUtensorMap dymap;
  int st = make_tmap_f16_2d(&dymap, dy, (uint64_t)M, (uint64_t)N, 128);
  if (st != QEFT_OK) return st;
  // oweight [N, r] as 64 x 64 boxes: a box is the 64 rows of a k-block x 64 outlier features, i.e. 128-byte rows in the
  // swizzled MN-major layout of the B tile (r == 0: a valid dummy map over dy, never used)
  CUtensorMap owmap;
  st = r > 0 ? make_tmap_f16_2d(&owmap, oweight, (uint64_t)N, (uint64_t)r, 64) : make_tmap_f16_2d(&owmap, dy, (uint64_t)M, (uint64_t)N, 64);
  if (st != QEFT_OK) return st;
  DxParams prm;
  prm.qw = static_cast<const uint8_t*>(qweight);
  prm.scales = static_cast<const __half*>(scales);
  prm.szeros = static_cast<const __half*>(scaled_zeros);
  prm.ow = r > 0 ? static_cast<const __half*>(oweight) : nullptr;
  prm.dx = static_cast<__half*>(dx);
  prm.M = M; prm.N = N; prm.K = K; prm.r = r; prm.G = G;
  // Tile = 256 features x 2 x 128 tokens, one CTA per SM.  QEFT_DX_TBC=1 halves the token extent of a tile (more,
  // smaller tiles for grids that leave a nearly empty last wave, e.g. 13B K = 5120: 160 tiles on 148 SMs).  Measured on
  // B200 it loses everywhere (5120 x 5120: 272 us against 211; 4096 x 4096: 150 against 92): a 1-block tile costs
  // 0.8-1.0 of a 2-block tile, i.e. the kernel is bound by producing the dequantised B tile (the same work for both),
  // not by the MMAs or their operand traffic.  Kept as an experiment switch only.
  {
    static const int tbc_env = getenv("QEFT_DX_TBC") ? atoi(getenv("QEFT_DX_TBC")) : 0;
    prm.tbc = tbc_env == 1 ? 1 : 2;
  }
  // contraction split (dx_plan above).  QEFT_DX_SPLITS=1: never, =n: every tile n ways; QEFT_DX_TAIL=n: the last wave n ways
  const int ttok = cdiv(M, 128 * prm.tbc), tiles = ttok * cdiv(K, kDxBF);
  int splits = 1, t_main = tiles;
  prm.ws = nullptr; prm.counters = nullptr;
  {
    static const int split_env = getenv("QEFT_DX_SPLITS") ? atoi(getenv("QEFT_DX_SPLITS")) : 0;
    static const int tail_env = getenv("QEFT_DX_TAIL") ? atoi(getenv("QEFT_DX_TAIL")) : 0;
    int dev = 0, nsm = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || nsm <= 0) nsm = 148;
    dx_plan(tiles, N / kDxBK, nsm, (size_t)M * (size_t)K, split_env, tail_env, split_env == 1, &splits, &t_main);
    if (splits > 1) {
      st = split_workspace(static_cast<cudaStream_t>(stream), (size_t)splits * (size_t)M * (size_t)K * sizeof(float), &prm.ws, &prm.counters);
      if (st != QEFT_OK) return st;
    }
  }
  prm.ttok = ttok; prm.t_main = t_main; prm.splits = splits;
  const size_t smem = (size_t)kDxStages * kDxStageBytes + 1024;
  static bool done[2][64] = {};
  const bool bf = dtype == QEFT_DT_BF16;
  st = bf ? set_smem_once(gemm_w4_dx_kernel<true>, smem, done[1]) : set_smem_once(gemm_w4_dx_kernel<false>, smem, done[0]);
  if (st != QEFT_OK) return st;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(t_main + (tiles - t_main) * splits));
  cfg.blockDim = dim3(kDxThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (flags & QEFT_F_PDL) ? 1 : 0;
  cudaError_t e = bf ? cudaLaunchKernelEx(&cfg, gemm_w4_dx_kernel<true>, dymap, owmap, prm)
                     : cudaLaunchKernelEx(&cfg, gemm_w4_dx_kernel<false>, dymap, owmap, prm);
  if (e != cudaSuccess) return (int)e;
  count_launch();
  return QEFT_OK;
}

extern "C" int qeft_dow(const void* dy, const void* x, float* dow, int M, int N, int K, int r, int dtype, int accumulate,
                        unsigned flags, qeft_stream_t stream) {
  if (!dy || !x || !dow) return QEFT_E_NULL;
  if (dtype != QEFT_DT_F16 && dtype != QEFT_DT_BF16) return QEFT_E_DTYPE;
  // x is the [M, K] activation (the last r columns are used) or, with K == r, the compact [M, r] copy
  if (M <= 0 || N <= 0 || N % 128 != 0 || r <= 0 || r % 64 != 0 || r > 256 || K < r || K % 8 != 0) return QEFT_E_SHAPE;
  if (!check_align16(dy) || !check_align16(x) || !check_align16(dow)) return QEFT_E_ALIGN;
  CUtensorMap dymap, xomap;
  int st = make_tmap_f16_2d(&dymap, dy, (uint64_t)M, (uint64_t)N, kDwBK);
  if (st != QEFT_OK) return st;
  // the outlier activations as an [M, r] matrix with row pitch K
  st = make_tmap_f16_2d_pitched(&xomap, static_cast<const __half*>(x) + (K - r), (uint64_t)M, (uint64_t)r, (uint64_t)K, kDwBK);
  if (st != QEFT_OK) return st;
  DowParams prm;
  prm.dow = dow; prm.M = M; prm.N = N; prm.r = r; prm.accumulate = accumulate; prm.bf16 = dtype == QEFT_DT_BF16;
  const size_t smem = (size_t)kDwStages * (2 * kDwBK * 128 + (size_t)(r / 64) * kDwBK * 128) + 1024;
  static bool done[64] = {};
  st = set_smem_once(dow_kernel, 200 * 1024, done);
  if (st != QEFT_OK) return st;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(N / 128));
  cfg.blockDim = dim3(kDwThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (flags & QEFT_F_PDL) ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, dow_kernel, dymap, xomap, prm);
  if (e != cudaSuccess) return (int)e;
  count_launch();
  return QEFT_OK;
}
