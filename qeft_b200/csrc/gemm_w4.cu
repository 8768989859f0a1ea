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
//   * The ACTIVATIONS are the B operand (N_umma = BN tokens): BN x 64 fp16 tiles, K-major, 128-byte swizzle,
//     brought in by TMA (cp.async.bulk.tensor) into a ring of stages of two k-blocks.
//   * One CTA computes 128 features x BN tokens (BN = 256; 64 / 128 for small M): one fp32 accumulator of BN TMEM
//     columns (y^T tile) + a ring of weight blocks in the remaining columns.
//   * The dense fp16 outlier columns are simply the last r/64 k-blocks: same pipeline, the dequant warps copy
//     oweight[f, 64 b .. 64 b + 63] to TMEM unchanged; the dead int4 columns K-r..K-1 are never read.
//   * Warp roles (512 threads): warp 0 TMA producer, warp 1 MMA issuer (one elected lane, tcgen05.commit to
//     mbarriers), warp 2 TMEM allocator, warps 4-15 dequant (three sets of 4 warps that alternate ring stages of two
//     k-blocks; a warp owns 32 features = its TMEM lane quadrant) and then epilogue (tcgen05.ld, + bias, fp16, transposed
//     through shared memory, 16-byte stores; split-K launches: fp32 partials + last-arriver reduction).
//   * Not instantiated: two feature blocks per CTA (NRB = 2, 16 dequant warps) and three k-blocks per ring stage
//     (KPS = 3): both ended in unspecified launch failures in round 1 and were not diagnosed.
#include "tc_common.cuh"

#include <stdlib.h>

#include <mutex>
#include <vector>

namespace qeft {

constexpr int kBK = 64;           // input columns per k-block (one 128-byte swizzle row of fp16)
constexpr int kDequantWarps = 12;
constexpr int kGemmThreads = (4 + kDequantWarps) * 32;
constexpr int kTmemCols = 512;

// NRB: 128-feature blocks (accumulators) per CTA; BN: tokens per CTA (UMMA N)
template <int NRB, int BN>
struct GemmCfg {
  static constexpr int kBM = 128 * NRB;                      // output features per CTA
  static constexpr int kBN = BN;
  static constexpr int kKPS = 2;                             // 64-column k-blocks per ring stage (8 UMMAs per barrier round)
  static constexpr int kXTileBytes = BN * kBK * 2;           // one activation tile (one k-block)
  static constexpr int kXStageBytes = kKPS * kXTileBytes;
  static constexpr int kXStagesMax = (192 * 1024) / kXStageBytes > 8 ? 8 : (192 * 1024) / kXStageBytes;
  static constexpr int kTmemA0 = NRB * BN;                   // first TMEM column of the weight ring
  static constexpr int kABlockCols = 32 * NRB;               // one k-block: 64 fp16 per lane and feature block
  static constexpr int kAStageCols = kKPS * kABlockCols;
  static constexpr int kAStagesMax = (kTmemCols - kTmemA0) / kAStageCols > 8 ? 8 : (kTmemCols - kTmemA0) / kAStageCols;
  // ONE ring: stage s = activation tile s in shared memory + weight block s in tensor memory, one full / one empty
  // barrier per stage (the MMA issuer waits once and commits once per k-block)
  static constexpr int kStages = kXStagesMax < kAStagesMax ? kXStagesMax : kAStagesMax;
  static constexpr int kXStages = kStages, kAStages = kStages;
  static constexpr int kSets = kDequantWarps / (4 * NRB);    // sets of dequant warps that alternate ring stages
  // kind::f16 instruction descriptor: D = F32 (bit 4), A / B format at bits 7 / 10 (0 = F16, 1 = BF16), K-major, N, M = 128
  static constexpr uint32_t kIdescF16 = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  static constexpr uint32_t kIdescBF16 = kIdescF16 | (1u << 7) | (1u << 10);
  static_assert(kTmemA0 + kAStages * kAStageCols <= kTmemCols, "TMEM budget");
  static_assert(kSets >= 1 && kSets * 4 * NRB == kDequantWarps, "every dequant warp belongs to exactly one set");
};

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
  int dbg;                 // QEFT_GEMM_DEBUG bits (bisecting only; results are wrong when set): 1 = no dequant
                           // math, 2 = no activation loads
  // output row pitch and, for the column-sharded prefill (all-gather fused into the epilogue, qeft_gemm_w4_gather):
  // every rank's gathered buffer (offset to THIS rank's columns), the arrival counters, the flag of the launch whose
  // gathered output is this launch's x
  int y_ld;
  int nranks;              // 0: plain launch, y only
  __half* y_peer[QEFT_MAX_RANKS];
  __half* y_mc;            // multicast mapping of the gathered buffer (one store reaches every rank), or null
  uint32_t* done_peer[QEFT_MAX_RANKS];
  const uint32_t* wait_flag;
  const uint32_t* epoch;
  // split-K (small M: gridDim.z CTAs share a tile, each over a contiguous range of ring stages): fp32 partial tiles
  // [split][M][N] and one arrival counter per tile; the CTA that arrives last adds the partials in split order
  float* ws;
  unsigned* counters;
  int stream_k;            // one CTA per SM: 1 = each a contiguous range of the launch's (tile, stage) units (stream-K),
                           // 2 = each a contiguous range of whole tiles (persistent)
};

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

// kind::f16 instruction descriptor (Cfg::kIdesc): D = F32, A = B = F16, both K-major, M = 128, N = BN
// MC: clusters of two CTAs (adjacent feature blocks, same tokens) share every activation tile: each CTA loads one
// half with a multicast TMA, both receive the whole tile -- halves the L2 -> SM activation traffic.
//
// Work of a CTA = a list of SEGMENTS (tile, range of ring stages), walked in lock step by all warp roles:
//   * plain launch: one segment, the whole tile of blockIdx.(x, y);
//   * split-K (gridDim.z > 1): one segment, the blockIdx.z-th part of the tile's stages;
//   * stream-K (p.stream_k; one CTA per SM): the launch's (tile, stage) units, tile-major, are cut into gridDim.x equal
//     contiguous ranges, so every SM finishes at the same time whatever the number of tiles (256 tiles on 148 SMs are
//     1.73 units per CTA instead of two waves); a range covers the tail of one tile, whole tiles, and the head of another.
// A tile whose stages are shared by several segments is reduced through fp32 partial tiles in a workspace: the segment
// that arrives LAST at the tile's counter adds the partials in part order (deterministic), adds the bias and stores.
// TMEM, barriers and the tensor map are set up once per CTA; ring stages and barrier phases run on across segments.
struct GemmSeg { int tok0, n0, tile, sb, se, parts, part; };

template <int NRB, int BN, bool MC, bool BF16>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_w4_kernel(const __grid_constant__ CUtensorMap xmap, const GemmParams p) {
  using Cfg = GemmCfg<NRB, BN>;
  static_assert(NRB == 1, "one 128-feature block per CTA");
  constexpr int kBN = Cfg::kBN, kXStages = Cfg::kXStages, kAStages = Cfg::kAStages;
  constexpr int kXStageBytes = Cfg::kXStageBytes, kTmemA0 = Cfg::kTmemA0, kDequantSets = Cfg::kSets;
  constexpr int kKPS = Cfg::kKPS, kXTileBytes = Cfg::kXTileBytes;
  constexpr uint32_t kIdesc = BF16 ? Cfg::kIdescBF16 : Cfg::kIdescF16;
  constexpr int rb = 0, nrb = 1;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * kXStages + 2];
  __shared__ uint32_t s_tmem_base;
  __shared__ uint32_t s_last;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t xs0 = (smem_addr(smem_raw) + 1023u) & ~1023u;          // 1024-byte aligned stage ring
  // epilogue staging (2 KB per dequant warp) BEHIND the ring: the producer may already be loading the next segment's
  // activation tiles into the ring while a segment's accumulator is drained
  uint8_t* stage_gen = smem_raw + (xs0 - smem_addr(smem_raw)) + (size_t)kXStages * kXStageBytes;
  const uint32_t bar0 = smem_addr(bars);
  auto x_full = [&](int s) { return bar0 + 8 * s; };               // TMA bytes + the dequant warps' arrivals
  auto x_empty = [&](int s) { return bar0 + 8 * (kXStages + s); };  // tcgen05.commit of the k-block's MMAs
  auto a_full = x_full;
  auto a_empty = x_empty;
  const uint32_t acc_full = bar0 + 8 * (2 * kXStages);             // a segment's MMAs are done
  const uint32_t acc_empty = acc_full + 8;                         // ... and its accumulator has been read (12 warps)

  const int nst_all = (p.nkb + kKPS - 1) / kKPS;                   // ring stages of a whole tile
  const int FB = p.N / 128;                                        // feature blocks
  // this CTA's range of (tile, stage) units [u_cur, u_end)
  int u_begin, u_end;            // (units fit 32 bits: at most 8192 tiles x a few hundred stages, times the grid)
  if (p.stream_k == 2) {
    // persistent whole tiles: CTA c takes tiles [T c / G, T (c + 1) / G): TMEM allocation, barrier set-up and the
    // tensor-map prefetch once per CTA, the TMA producer runs ahead into the next tile while this one is drained
    const int T = cdiv(p.M, kBN) * FB;
    u_begin = (int)(((long long)T * blockIdx.x) / gridDim.x) * nst_all;
    u_end = (int)(((long long)T * (blockIdx.x + 1)) / gridDim.x) * nst_all;
  } else if (p.stream_k) {
    const int U = cdiv(p.M, kBN) * FB * nst_all;
    u_begin = (int)(((long long)U * blockIdx.x) / gridDim.x);
    u_end = (int)(((long long)U * (blockIdx.x + 1)) / gridDim.x);
  } else {
    const int nsplit = (int)gridDim.z, split = (int)blockIdx.z;
    const int t0 = ((int)blockIdx.x * FB + (int)blockIdx.y) * nst_all;
    u_begin = t0 + (nst_all * split) / nsplit;
    u_end = t0 + (nst_all * (split + 1)) / nsplit;
  }
  // segment that starts at unit u (tile ids are token-tile major: CTAs that run together share activation tiles in L2)
  auto seg_at = [&](int u) {
    GemmSeg g;
    g.tile = u / nst_all;
    g.sb = u - g.tile * nst_all;
    const int left = u_end - u;
    g.se = left < nst_all - g.sb ? g.sb + left : nst_all;
    g.tok0 = (g.tile / FB) * kBN;
    g.n0 = (g.tile % FB) * 128;
    if (p.stream_k) {            // (a range is at least one tile long: a tile is shared by at most two CTAs)
      g.parts = (g.sb == 0 && g.se == nst_all) ? 1 : 2;
      g.part = g.sb > 0 ? 1 : 0;
    } else {
      g.parts = (int)gridDim.z;
      g.part = (int)blockIdx.z;
    }
    return g;
  };

  if (tid == 0) {
    for (int s = 0; s < kXStages; ++s) { mbar_init(x_full(s), 1 + 4 * nrb); mbar_init(x_empty(s), MC ? 2 : 1); }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, kDequantWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&s_tmem_base)), "n"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  if (MC) cluster_sync_all();        // the peer's barriers are initialised before anything is multicast to them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem_base;
  const uint32_t crank = MC ? cluster_ctarank() : 0;
  pdl_launch_dependents();

  if (warp == 0) {
    // ================= TMA producer: activation tiles =================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&xmap) : "memory");
      pdl_wait();                                   // x is the previous kernel's output
      if (p.wait_flag) {
        // column-sharded chain: x is the gathered output of an earlier launch; every rank's slice has landed once that
        // launch's arrival counter reaches epoch x ranks.  The slices were written by generic-proxy stores (of this and
        // other GPUs) and are read by TMA: order the two proxies after the acquire.
        const uint32_t want = *reinterpret_cast<const volatile uint32_t*>(p.epoch) * (uint32_t)p.nranks * kArrivalsPerLaunch;
        uint32_t got;
        do {
          asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(got) : "l"(p.wait_flag) : "memory");
        } while ((int)(got - want) < 0);
        asm volatile("fence.proxy.async;" ::: "memory");
      }
      int rit = 0;                                  // ring iteration: stages issued so far by this CTA
      for (int u = u_begin; u < u_end;) {
        const GemmSeg g = seg_at(u);
        for (int st = g.sb; st < g.se; ++st, ++rit) {
          const int s = rit % kXStages, use = rit / kXStages;
          if (use > 0) mbar_wait(x_empty(s), (uint32_t)((use - 1) & 1));
          const int nblk = min(kKPS, p.nkb - st * kKPS);
          if (p.dbg & 2) { mbar_arrive(x_full(s)); continue; }
          mbar_expect_tx(x_full(s), (uint32_t)(nblk * kXTileBytes));
          for (int j = 0; j < nblk; ++j) {
            const int kb = st * kKPS + j;
            const int k0 = kb < p.nkb_q ? kb * kBK : p.K - p.r + (kb - p.nkb_q) * kBK;
            const uint32_t dst = xs0 + s * kXStageBytes + j * kXTileBytes;
            if (MC)     // my half of the tokens, to both CTAs (the peer sends the other half)
              tma_load_2d_mc(dst + crank * (kXTileBytes / 2), &xmap, k0, g.tok0 + (int)crank * (kBN / 2), x_full(s), (uint16_t)3);
            else
              tma_load_2d(dst, &xmap, k0, g.tok0, x_full(s));
          }
        }
        u += g.se - g.sb;
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      // (this one thread issues every MMA of the CTA: no descriptor arithmetic -- a descriptor's low field is the shared-
      // memory address / 16, so a stage's descriptors are base + constant --, no divisions, no runtime trip counts)
      const uint64_t bdesc0 = make_b_desc(xs0);
      int s = 0, seg = 0;
      uint32_t par = 0;
      for (int u = u_begin; u < u_end; ++seg) {
        const GemmSeg g = seg_at(u);
        if (seg > 0) {                // the previous segment's accumulator has been read by all epilogue warps
          mbar_wait(acc_empty, (uint32_t)((seg - 1) & 1));
          tc_fence_after();
        }
        for (int st = g.sb; st < g.se; ++st) {
          mbar_wait(x_full(s), par);
          tc_fence_after();
          const uint64_t bs = bdesc0 + (uint64_t)((s * kXStageBytes) >> 4);
          const uint32_t as = tmem + kTmemA0 + Cfg::kAStageCols * s;
          const bool both = (st + 1) * kKPS <= p.nkb;          // (the tile's last stage may hold one k-block)
#pragma unroll
          for (int j = 0; j < kKPS; ++j) {
            if (j == 0 || both) {
#pragma unroll
              for (int k16 = 0; k16 < kBK / 16; ++k16)
                umma_ts_f16(tmem, as + Cfg::kABlockCols * j + 8 * k16, bs + (uint64_t)((j * kXTileBytes + k16 * 32) >> 4), kIdesc,
                            (uint32_t)(((st - g.sb) | j | k16) != 0));
            }
          }
          if (MC) tc_commit_mc(x_empty(s), (uint16_t)3);   // the stage is rewritten by BOTH CTAs' producers
          else tc_commit(x_empty(s));     // the stage (smem tiles + TMEM blocks) is free once these MMAs have read it
          if (++s == kXStages) { s = 0; par ^= 1u; }
        }
        tc_commit(acc_full);
        u += g.se - g.sb;
      }
    }
  } else if (warp >= 4) {
    // ================= dequant warps (then epilogue), segment by segment =================
    // set ws handles the segment's stages ws, ws + kDequantSets, ... (all k-blocks of a stage); inside a set warp dw owns
    // features 32 dw .. 32 dw + 31 of the block (its TMEM lane quadrant is warp % 4)
    const int ws = (warp - 4) / 4, quad = (warp - 4) % 4;
    const uint32_t lane_taddr = tmem + ((uint32_t)(32 * quad) << 16);
    int rit0 = 0, seg = 0;                          // ring iteration of the segment's first stage
    for (int u = u_begin; u < u_end; ++seg) {
      const GemmSeg g = seg_at(u);
      const int tok0 = g.tok0, n0 = g.n0;
      const int kb0 = g.sb * kKPS;                            // the segment's k-blocks [kb0, nkb)
      const int nkb = min(p.nkb, g.se * kKPS);
      const int f = n0 + 32 * quad + lane;                    // this lane's output feature
      const uint8_t* qrow = p.qw + (size_t)(f >> 2) * (size_t)(2 * p.K) + (size_t)((f & 3) * 32);
      const __half* owrow = p.ow ? p.ow + (size_t)f * p.r : nullptr;
      // write one k-block (64 fp16 of this lane's feature) into its stage; the first block of a stage waits for
      // the stage to be free, the last one publishes the stage
      auto publish = [&](const uint32_t (&v)[32], int kb) {
        const int strel = (kb - kb0) / kKPS, j = (kb - kb0) - strel * kKPS;
        const int rit = rit0 + strel;
        const int as = rit % kAStages, use = rit / kAStages;
        if (j == 0 && use > 0) mbar_wait(a_empty(as), (uint32_t)((use - 1) & 1));
        tc_fence_after();
        tmem_st32(lane_taddr + (uint32_t)(kTmemA0 + Cfg::kAStageCols * as + Cfg::kABlockCols * j + 32 * rb), v);
        if (j == kKPS - 1 || kb == nkb - 1) {
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(a_full(as));
        }
      };
      // this set's k-blocks: i-th = block (i % KPS) of the segment's stage ws + (i / KPS) * sets
      auto kb_of = [&](int i) { return kb0 + (ws + (i / kKPS) * kDequantSets) * kKPS + (i % kKPS); };
      // Register prefetch ring over this set's k-blocks, kPF deep (weights come from L2, ~700 cycles away), scales
      // one group ahead.
      constexpr int kPF = 4;
      static_assert(kPF % kKPS == 0, "ring slots must be compile-time registers");
      uint4 ring[kPF][2];
#pragma unroll
      for (int i = 0; i < kPF; ++i) {
        ring[i][0] = ring[i][1] = make_uint4(0, 0, 0, 0);
        const int kb = kb_of(i);
        if (kb < nkb && kb < p.nkb_q) {
          ring[i][0] = ldg_nc_v4(qrow + (size_t)kb * 128);
          ring[i][1] = ldg_nc_v4(qrow + (size_t)kb * 128 + 16);
        }
      }
      const int kb_per_grp = p.G / kBK;
      uint32_t s2 = 0, z2 = 0;
      unsigned short sn = 0, zn = 0;                   // next group's raw halves (first use a whole group later)
      int grp = -1, grp_next = -1;
      const __half* sp = p.scales + f;
      const __half* zp = p.szeros + f;
      {
        const int kb = kb_of(0);
        if (kb < nkb && kb < p.nkb_q) { grp_next = kb / kb_per_grp; sn = ldg_nc_u16(sp + (size_t)grp_next * p.N); zn = ldg_nc_u16(zp + (size_t)grp_next * p.N); }
      }
      bool done = false;
      for (int i0 = 0; !done; i0 += kPF) {
#pragma unroll
        for (int uu = 0; uu < kPF; ++uu) {
          const int i = i0 + uu;
          const int kb = kb_of(i);
          if (kb >= nkb) { done = true; break; }
          uint32_t v[32];
          if (kb < p.nkb_q) {
            const uint4 c0 = ring[uu][0], c1 = ring[uu][1];
            const int kbn = kb_of(i + kPF);
            if (kbn < nkb && kbn < p.nkb_q) {
              ring[uu][0] = ldg_nc_v4(qrow + (size_t)kbn * 128);
              ring[uu][1] = ldg_nc_v4(qrow + (size_t)kbn * 128 + 16);
            }
            const int g_now = kb / kb_per_grp;
            if (g_now != grp) {                        // entered a new group: take the prefetched pair (or load it)
              if (g_now == grp_next) {
                s2 = (uint32_t)sn | ((uint32_t)sn << 16);
                z2 = (uint32_t)zn | ((uint32_t)zn << 16);
              } else {
                const unsigned short a = ldg_nc_u16(sp + (size_t)g_now * p.N), bq = ldg_nc_u16(zp + (size_t)g_now * p.N);
                s2 = (uint32_t)a | ((uint32_t)a << 16);
                z2 = (uint32_t)bq | ((uint32_t)bq << 16);
              }
              grp = g_now;
              // the next group this set will touch: look ahead over its coming k-blocks
              grp_next = -1;
#pragma unroll
              for (int a = 1; a <= 2 * kKPS; ++a) {
                const int kba = kb_of(i + a);
                if (grp_next < 0 && kba < nkb && kba < p.nkb_q && kba / kb_per_grp != g_now) grp_next = kba / kb_per_grp;
              }
              if (grp_next >= 0) { sn = ldg_nc_u16(sp + (size_t)grp_next * p.N); zn = ldg_nc_u16(zp + (size_t)grp_next * p.N); }
            }
            const uint32_t w[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
            if (p.dbg & 1) {
#pragma unroll
              for (int q = 0; q < 32; ++q) v[q] = w[q & 7];
            } else
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                uint32_t hq[4];
                unpack_word_to_half2(w[4 * h + c], hq);          // pairs k = 32 h + 2 c + 8 j (+1), exact 0..15
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  v[16 * h + c + 4 * j] = BF16 ? dequant_pair_bf16(hq[j], __half2float(__ushort_as_half((unsigned short)(s2 & 0xffffu))),
                                                                   __half2float(__ushort_as_half((unsigned short)(z2 & 0xffffu))))
                                               : hfma2_u32(hq[j], s2, z2);                      // w = fma(q, s, sz)
              }
          } else {
            // outlier k-block: fp16 columns, copied unchanged
            const __half* src = owrow + (size_t)(kb - p.nkb_q) * kBK;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const uint4 t4 = ldg_nc_v4(src + 8 * q);
              v[4 * q + 0] = t4.x; v[4 * q + 1] = t4.y; v[4 * q + 2] = t4.z; v[4 * q + 3] = t4.w;
            }
          }
          publish(v, kb);
        }
      }

      // ---- epilogue: y^T tile (lane = feature, columns = tokens) -> + bias -> fp16 -> y[token, feature] ----
      // the sets split the token chunks of 32
      mbar_wait(acc_full, (uint32_t)(seg & 1));
      tc_fence_after();
      const float bias = p.bias ? __half2float(p.bias[f]) : 0.f;
      const int fw = n0 + 32 * quad;                                 // first feature of this warp
      if (g.parts > 1) {
        // ---- shared tile: fp32 partial to the workspace; the segment that arrives last adds all partials of the tile in
        // part order (deterministic), adds the bias and stores fp16.  (semaphore reduce of gemm_cuda.cu:512-585, here
        // without the serialisation: only the last arrival reads.)
        float* wsp = p.ws + (size_t)g.part * (size_t)p.M * (size_t)p.N;
        float* stage_f = reinterpret_cast<float*>(stage_gen + (warp - 4) * 2048);   // 16 tokens x 32 features per warp
#pragma unroll 1
        for (int tc = ws; tc < kBN / 32; tc += kDequantSets) {
          if (tok0 + 32 * tc >= p.M) break;
          uint32_t acc[32];
          tmem_ld32(lane_taddr + (uint32_t)(kBN * rb + 32 * tc), acc);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
            for (int i = 0; i < 16; ++i) stage_f[i * 32 + lane] = __uint_as_float(acc[16 * hh + i]);
            __syncwarp();
            // 16 rows (tokens) of 128 bytes: 8 lanes per row, 4 rows per pass
#pragma unroll
            for (int pass = 0; pass < 4; ++pass) {
              const int row = pass * 4 + (lane >> 3), piece = lane & 7;
              const int tok = tok0 + 32 * tc + 16 * hh + row;
              if (tok < p.M)
                *reinterpret_cast<float4*>(wsp + (size_t)tok * (size_t)p.N + (size_t)(fw + piece * 4)) =
                    *reinterpret_cast<const float4*>(stage_f + row * 32 + piece * 4);
            }
            __syncwarp();
          }
        }
        tc_fence_before();
        if (lane == 0) mbar_arrive(acc_empty);             // (this warp's reads of the accumulator are done)
        __threadfence();
        asm volatile("bar.sync 1, %0;" ::"n"(kDequantWarps * 32) : "memory");
        if (tid == 128) s_last = (atomicAdd(p.counters + g.tile, 1u) == (unsigned)(g.parts - 1)) ? 1u : 0u;
        asm volatile("bar.sync 1, %0;" ::"n"(kDequantWarps * 32) : "memory");
        if (s_last) {
          __threadfence();
          const int ntok = min(kBN, p.M - tok0);
          // (one SM reads parts x 128 KB from L2: 8 positions of loads in flight per thread)
          constexpr int kU = 8, kStride = kDequantWarps * 32;
          const int total = ntok * 32;
          const size_t partsz = (size_t)p.M * (size_t)p.N;
#pragma unroll 1
          for (int base = tid - 128; base < total; base += kU * kStride) {
            size_t off[kU];
            float4 a[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) {
              const int idx = base + u * kStride;
              off[u] = (size_t)(tok0 + (idx >> 5)) * (size_t)p.N + (size_t)(n0 + 4 * (idx & 31));
              a[u] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (idx < total) a[u] = __ldcg(reinterpret_cast<const float4*>(p.ws + off[u]));
            }
            for (int z = 1; z < g.parts; ++z) {
              float4 v[kU];
#pragma unroll
              for (int u = 0; u < kU; ++u) {
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (base + u * kStride < total) v[u] = __ldcg(reinterpret_cast<const float4*>(p.ws + (size_t)z * partsz + off[u]));
              }
#pragma unroll
              for (int u = 0; u < kU; ++u) { a[u].x += v[u].x; a[u].y += v[u].y; a[u].z += v[u].z; a[u].w += v[u].w; }
            }
#pragma unroll
            for (int u = 0; u < kU; ++u) {
              const int idx = base + u * kStride;
              if (idx >= total) continue;
              const int c4 = idx & 31;
              float4 r4 = a[u];
              if (p.bias) {
                const uint2 bb = *reinterpret_cast<const uint2*>(p.bias + n0 + 4 * c4);
                const float2 b01 = half2_bits_to_float2(bb.x), b23 = half2_bits_to_float2(bb.y);
                r4.x += b01.x; r4.y += b01.y; r4.z += b23.x; r4.w += b23.y;
              }
              uint2 o;
              if (BF16) {
                const __nv_bfloat162 lo = __floats2bfloat162_rn(r4.x, r4.y), hi = __floats2bfloat162_rn(r4.z, r4.w);
                o.x = *reinterpret_cast<const uint32_t*>(&lo); o.y = *reinterpret_cast<const uint32_t*>(&hi);
              } else {
                const __half2 lo = __floats2half2_rn(r4.x, r4.y), hi = __floats2half2_rn(r4.z, r4.w);
                o.x = *reinterpret_cast<const uint32_t*>(&lo); o.y = *reinterpret_cast<const uint32_t*>(&hi);
              }
              *reinterpret_cast<uint2*>(p.y + (size_t)(tok0 + (idx >> 5)) * (size_t)p.y_ld + (size_t)(n0 + 4 * c4)) = o;
            }
          }
          if (tid == 128) p.counters[g.tile] = 0u;          // ready for the next launch (stream order)
        }
      } else {
        __half* stage = reinterpret_cast<__half*>(stage_gen + (warp - 4) * 2048);   // 32 tokens x 32 features per warp
#pragma unroll 1
        for (int tc = ws; tc < kBN / 32; tc += kDequantSets) {
          uint32_t acc[32];
          tmem_ld32(lane_taddr + (uint32_t)(kBN * rb + 32 * tc), acc);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float val = __uint_as_float(acc[i]) + bias;
            if (BF16) reinterpret_cast<__nv_bfloat16*>(stage)[i * 32 + lane] = __float2bfloat16_rn(val);
            else stage[i * 32 + lane] = __float2half_rn(val);
          }
          __syncwarp();
          // 32 rows (tokens) of 64 bytes: 4 lanes per row, 8 rows per pass
#pragma unroll
          for (int pass = 0; pass < 4; ++pass) {
            const int row = pass * 8 + (lane >> 2), piece = lane & 3;
            const int tok = tok0 + 32 * tc + row;
            const uint4 val = *reinterpret_cast<const uint4*>(stage + row * 32 + piece * 8);
            if (tok < p.M) {
              const size_t off = (size_t)tok * (size_t)p.y_ld + (size_t)(fw + piece * 8);
              if (p.nranks == 0) {
                *reinterpret_cast<uint4*>(p.y + off) = val;
              } else if (p.y_mc) {
                // one store to the multicast address: the switch replicates it to every rank (multimem.st lowers to
                // this same STG.128 on the multicast mapping)
                asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p.y_mc + off), "r"(val.x),
                             "r"(val.y), "r"(val.z), "r"(val.w) : "memory");
              } else {
                for (int pr = 0; pr < p.nranks; ++pr) *reinterpret_cast<uint4*>(p.y_peer[pr] + off) = val;   // NVLink stores
              }
            }
          }
          __syncwarp();
        }
        tc_fence_before();
        if (lane == 0) mbar_arrive(acc_empty);
      }
      rit0 += g.se - g.sb;
      u += g.se - g.sb;
    }
  }
  __syncwarp();                      // the single-lane roles rejoin their warps
  if (MC) cluster_sync_all();        // no CTA leaves while its peer may still signal its barriers
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols) : "memory");
  }
  if (p.nranks > 0 && tid == 0) {
    // publish (same protocol as the decode GEMV): the CTA's stores happen-before the barrier above, this thread's
    // system-scope fence is cumulative over them and orders them before its relaxed signals; every CTA signals its
    // share of kArrivalsPerLaunch to every rank
    fence_acq_rel_sys();
    const unsigned inc = arrival_share(blockIdx.y * gridDim.x + blockIdx.x, gridDim.x * gridDim.y);
    for (int pr = 0; pr < p.nranks; ++pr)
      asm volatile("red.relaxed.sys.global.add.u32 [%0], %1;" ::"l"(p.done_peer[pr]), "r"(inc) : "memory");
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

// 2-D map of a row-major fp16 matrix [rows, cols] with a row pitch of `pitch` elements: box = 64 columns
// (128 bytes, swizzled) x box_rows rows
int make_tmap_f16_2d_pitched(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch, uint32_t box_rows) {
  EncodeTiledFn fn = encode_tiled();
  if (!fn) return QEFT_E_UNSUPPORTED;
  // the encoder is a driver-API call: it needs the primary context bound to THIS thread, which a thread that has
  // only inherited torch's device bookkeeping (e.g. an autograd worker) may not have yet
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaSetDevice(dev);
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {pitch * 2};
  const cuuint32_t box[2] = {64, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return rc == CUDA_SUCCESS ? QEFT_OK : QEFT_E_UNSUPPORTED;
}
int make_tmap_f16_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  return make_tmap_f16_2d_pitched(map, base, rows, cols, cols, box_rows);
}

// Split-K workspace: fp32 partial tiles and per-tile arrival counters, one set per (device, stream) so that launches on
// different streams never share it; grown on demand.  The first split-K launch on a stream allocates (cudaMalloc), so
// it must happen before that stream is captured into a CUDA graph (a warm-up call, as torch requires anyway).
struct SplitWs { int dev; cudaStream_t stream; float* ws; size_t bytes; unsigned* counters; };
int split_workspace(cudaStream_t stream, size_t bytes, float** ws, unsigned** counters) {
  static std::vector<SplitWs> pool;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  int dev = 0;
  cudaGetDevice(&dev);
  SplitWs* e = nullptr;
  for (auto& w : pool)
    if (w.dev == dev && w.stream == stream) e = &w;
  if (!e) {
    pool.push_back(SplitWs{dev, stream, nullptr, 0, nullptr});
    e = &pool.back();
    cudaError_t rc = cudaMalloc(&e->counters, kSplitCounters * sizeof(unsigned));
    if (rc == cudaSuccess) rc = cudaMemset(e->counters, 0, kSplitCounters * sizeof(unsigned));
    if (rc != cudaSuccess) { pool.pop_back(); return (int)rc; }
  }
  if (e->bytes < bytes) {
    // (work of earlier launches on this stream may still read the old buffer)
    if (e->ws) { cudaStreamSynchronize(stream); cudaFree(e->ws); e->ws = nullptr; e->bytes = 0; }
    const size_t want = bytes < ((size_t)16 << 20) ? ((size_t)16 << 20) : bytes;
    cudaError_t rc = cudaMalloc(&e->ws, want);
    if (rc != cudaSuccess) return (int)rc;
    e->bytes = want;
  }
  *ws = e->ws;
  *counters = e->counters;
  return QEFT_OK;
}

// number of K splits for a small-M launch: minimise (waves of CTAs) x (ring stages per split), one stage of overhead
// per split for the partial round trip; every split keeps at least 3 stages
static int choose_splits(int tiles, int nst, int nsm) {
  int best = 1;
  long best_cost = (long)cdiv(tiles, nsm) * nst + 1;
  for (int sp = 2; sp <= 16 && nst / sp >= 3; ++sp) {
    const long cost = (long)cdiv(tiles * sp, nsm) * cdiv(nst, sp) + 1 + sp / 2;
    if (cost < best_cost) { best_cost = cost; best = sp; }
  }
  return best;
}

template <int NRB, int BN, bool MC, bool BF16>
static int launch_gemm(const void* x, const GemmParams& prm_in, unsigned flags, cudaStream_t stream, bool allow_split = false,
                       bool allow_stream_k = false) {
  GemmParams prm = prm_in;
  using Cfg = GemmCfg<NRB, BN>;
  CUtensorMap xmap;
  int st = make_tmap_f16_2d(&xmap, x, (uint64_t)prm.M, (uint64_t)prm.K, MC ? BN / 2 : BN);
  if (st != QEFT_OK) return st;
  auto kern = gemm_w4_kernel<NRB, BN, MC, BF16>;
  const size_t smem = (size_t)Cfg::kXStages * Cfg::kXStageBytes + (size_t)kDequantWarps * 2048 + 1024;   // ring + epilogue staging
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  cudaLaunchConfig_t cfg = {};
  int splits = 1;
  if (allow_split) {
    static int nsm = 0;
    if (nsm == 0 && (cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || nsm <= 0)) nsm = 148;
    static const int split_env = getenv("QEFT_GEMM_SPLITS") ? atoi(getenv("QEFT_GEMM_SPLITS")) : 0;
    const int tiles = cdiv(prm.M, BN) * cdiv(prm.N, Cfg::kBM), nst = cdiv(prm.nkb, Cfg::kKPS);
    splits = split_env > 0 ? (split_env < nst ? split_env : nst) : choose_splits(tiles, nst, nsm);
    if (tiles > kSplitCounters) splits = 1;
    if (splits > 1) {
      const int rc = split_workspace(stream, (size_t)splits * (size_t)prm.M * (size_t)prm.N * sizeof(float), &prm.ws, &prm.counters);
      if (rc != QEFT_OK) return rc;
    }
  }
  cfg.gridDim = dim3((unsigned)cdiv(prm.M, BN), (unsigned)cdiv(prm.N, Cfg::kBM), (unsigned)splits);
  if (allow_stream_k && splits == 1 && !MC) {
    // stream-K: when the tiles do not fill whole waves (256 tiles on 148 SMs), one CTA per SM takes an equal contiguous
    // share of the launch's (tile, stage) units; QEFT_GEMM_STREAMK=1 turns it on
    static int nsm2 = 0;
    if (nsm2 == 0 && (cudaDeviceGetAttribute(&nsm2, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || nsm2 <= 0)) nsm2 = 148;
    static const int sk_env = getenv("QEFT_GEMM_STREAMK") ? atoi(getenv("QEFT_GEMM_STREAMK")) : 0;   // (measured slower: DESIGN.md 3.2)
    const int tiles = cdiv(prm.M, BN) * cdiv(prm.N, Cfg::kBM);
    const int waves = cdiv(tiles, nsm2);
    static const int persist_env = getenv("QEFT_GEMM_PERSIST") ? atoi(getenv("QEFT_GEMM_PERSIST")) : 1;   // (0: one CTA per tile)
    if (persist_env && prm.nranks == 0 && tiles > nsm2) {     // (sharded launches: no gain measured, left one CTA per tile)
      prm.stream_k = 2;
      cfg.gridDim = dim3((unsigned)nsm2, 1, 1);
    } else if (sk_env && prm.nranks == 0 && tiles > nsm2 && tiles <= kSplitCounters && (long)waves * nsm2 * 100 > (long)tiles * 105) {
      const int rc = split_workspace(stream, (size_t)2 * (size_t)prm.M * (size_t)prm.N * sizeof(float), &prm.ws, &prm.counters);
      if (rc != QEFT_OK) return rc;
      prm.stream_k = 1;
      cfg.gridDim = dim3((unsigned)nsm2, 1, 1);
    }
  }
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (MC) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 1; attr[na].val.clusterDim.y = 2; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (flags & QEFT_F_PDL) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, xmap, prm);
  if (e != cudaSuccess) return (int)e;
  count_launch();
  return QEFT_OK;
}

}  // namespace qeft

using namespace qeft;

static int gemm_entry(const void* x, const void* qweight, const void* scales, const void* scaled_zeros,
                      const void* oweight, const void* bias, void* y, int M, int N, int K, int r, int G,
                      int dtype, unsigned flags, const qeft_gather_t* gat, qeft_stream_t stream) {
  if (!x || !qweight || !scales || !scaled_zeros || (!y && !gat)) return QEFT_E_NULL;
  if (dtype != QEFT_DT_F16 && dtype != QEFT_DT_BF16) return QEFT_E_DTYPE;
  if (G == -1) G = K;
  if (M <= 0 || N <= 0 || K <= 0 || N % 128 != 0 || K % 64 != 0 || G <= 0 || G % 64 != 0 || K % G != 0) return QEFT_E_SHAPE;
  if (r < 0 || r % 64 != 0 || r >= K) return QEFT_E_SHAPE;
  if (r > 0 && !oweight) return QEFT_E_NULL;
  if (!check_align16(x) || !check_align16(qweight) || (y && !check_align16(y)) || (r > 0 && !check_align16(oweight)))
    return QEFT_E_ALIGN;
  GemmParams prm = {};
  prm.qw = static_cast<const uint8_t*>(qweight);
  prm.scales = static_cast<const __half*>(scales);
  prm.szeros = static_cast<const __half*>(scaled_zeros);
  prm.ow = r > 0 ? static_cast<const __half*>(oweight) : nullptr;
  prm.bias = static_cast<const __half*>(bias);
  prm.y = static_cast<__half*>(y);
  prm.M = M; prm.N = N; prm.K = K; prm.r = r; prm.G = G;
  prm.nkb_q = (K - r) / kBK;
  prm.nkb = prm.nkb_q + r / kBK;
  prm.y_ld = N;
  if (gat) {
    if (gat->nranks < 1 || gat->nranks > QEFT_MAX_RANKS || !gat->epoch) return QEFT_E_SHAPE;
    if (gat->y_ld < N || gat->y_ld % 8 != 0) return QEFT_E_SHAPE;
    prm.nranks = gat->nranks;
    prm.y_ld = gat->y_ld;
    prm.y_mc = static_cast<__half*>(gat->y_mc[0]);
    if (prm.y_mc && !check_align16(prm.y_mc)) return QEFT_E_ALIGN;
    for (int pr = 0; pr < gat->nranks; ++pr) {
      if ((!prm.y_mc && !gat->y_peer[pr][0]) || !gat->done_peer[pr]) return QEFT_E_NULL;
      if (!check_align16(gat->y_peer[pr][0])) return QEFT_E_ALIGN;
      prm.y_peer[pr] = static_cast<__half*>(gat->y_peer[pr][0]);
      prm.done_peer[pr] = gat->done_peer[pr];
    }
    prm.wait_flag = gat->wait_flag;
    prm.epoch = gat->epoch;
  }
  static const int dbg_env = getenv("QEFT_GEMM_DEBUG") ? atoi(getenv("QEFT_GEMM_DEBUG")) : 0;
  prm.dbg = dbg_env;
  static const int cfg_env = getenv("QEFT_GEMM_CFG") ? atoi(getenv("QEFT_GEMM_CFG")) : 0;
  cudaStream_t cs = static_cast<cudaStream_t>(stream);
  // QEFT_GEMM_CFG=3: 2-CTA clusters multicast every activation tile (halves the L2 reads; measured 2-5 % slower than
  // the plain launch on B200 because L2 bandwidth is not the limiter at these sizes, so it is not the default)
  // bf16: x, oweight and y are bf16 (bias, scales, scaled zeros stay fp16 as in the checkpoint); the int4 columns are
  // dequantised in fp32 with one rounding to bf16
  // Small M (the reference's split-K tile configurations, gemm_cuda.cu:952-1004): a narrow token tile so that the MMA
  // and the activation ring do not work on padding, and K split over several CTAs per tile so that all SMs dequantise
  // (a 4096 x 11008 layer has 32 feature blocks for 148 SMs).  QEFT_GEMM_SMALLM=0 turns it off.
  static const int smallm_env = getenv("QEFT_GEMM_SMALLM") ? atoi(getenv("QEFT_GEMM_SMALLM")) : 1;
  if (smallm_env && !gat && M <= 128) {
    if (dtype == QEFT_DT_BF16)
      return M <= 64 ? launch_gemm<1, 64, false, true>(x, prm, flags, cs, true) : launch_gemm<1, 128, false, true>(x, prm, flags, cs, true);
    return M <= 64 ? launch_gemm<1, 64, false, false>(x, prm, flags, cs, true) : launch_gemm<1, 128, false, false>(x, prm, flags, cs, true);
  }
  // a few hundred tokens on a narrow layer still leave SMs idle (4096 x 11008 at M = 256: 32 tiles): K is split there too
  const bool split_ok = smallm_env && !gat && cdiv(M, 256) * (N / 128) <= 74;     // (at most half a wave of tiles)
  if (dtype == QEFT_DT_BF16) return launch_gemm<1, 256, false, true>(x, prm, flags, cs, split_ok, true);
  if (cfg_env == 3 && N % 256 == 0 && !gat) return launch_gemm<1, 256, true, false>(x, prm, flags, cs);
  return launch_gemm<1, 256, false, false>(x, prm, flags, cs, split_ok, true);
}

extern "C" int qeft_gemm_w4(const void* x, const void* qweight, const void* scales, const void* scaled_zeros,
                            const void* oweight, const void* bias, void* y, int M, int N, int K, int r, int G,
                            int dtype, unsigned flags, qeft_stream_t stream) {
  return gemm_entry(x, qweight, scales, scaled_zeros, oweight, bias, y, M, N, K, r, G, dtype, flags, nullptr, stream);
}

extern "C" int qeft_gemm_w4_gather(const void* x, const void* qweight, const void* scales, const void* scaled_zeros,
                                   const void* oweight, const void* bias, int M, int N, int K, int r, int G, int dtype,
                                   unsigned flags, const qeft_gather_t* gather, qeft_stream_t stream) {
  if (!gather) return QEFT_E_NULL;
  return gemm_entry(x, qweight, scales, scaled_zeros, oweight, bias, nullptr, M, N, K, r, G, dtype, flags, gather, stream);
}
