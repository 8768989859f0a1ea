// Persistent multi-stage decode kernel for the packed QEFT QuantLinear (sm_100a): "decode programs".
//
// Replaces the reference's one-kernel-per-projection decode path (qeft/qlinear.py:251-263 ->
// gemv_kernel_qeft, qeft/kernel/quantization_new/gemv/gemv_cuda_qeft.cu:75-222, launcher :392-513) for a whole
// chain of dependent GEMVs: one decoder block (qkv -> o -> gate/up -> down) or a whole token (4 x layers stages)
// is ONE cooperative launch.  A stage is what qeft_gemv_w4_multi computes (up to 4 projections sharing x, fp16
// outlier columns, bias, the o_proj gather of qlinear.py:273-275), optionally with the elementwise glue of a
// Llama block folded in (RMSNorm on the way in -- kernel/layernorm/layernorm.cu:25-51 --, SiLU(gate)*up or a
// residual add on the way out).
//
// Why (round-1 measurements): a decode token is 128 dependent launches of 10-50 MB; each launch
// paid ~4.5 us of fixed serial cost (kernel boundary, x staging, ring fill, reduce) against 1.5-7 us of DRAM time,
// and the old kernel needed ~155 warp instructions per KB of weights, i.e. it could not consume faster than HBM
// delivers, so nothing was ever caught up.  Here:
//   * one CTA of 16 warps per SM stays resident for the whole program; stage boundaries are a gpu-scope arrival
//     counter (red.release.gpu / ld.acquire.gpu), never a kernel boundary;
//   * the weight stream never stops at a boundary: every warp owns a private ring of D 1-KB units filled with
//     cp.async (LDGSTS, 16 B per lane, placed so that the consumer's LDS.128 are conflict-free and lane-private),
//     and the ring's prefetch cursor runs ahead through the NEXT stages' weights (they never depend on x) while
//     the warp waits for the barrier and for x: ~170 KB per SM stay in flight across the boundary;
//   * the inner loop is ~45 warp instructions per KB: int8 tensor-core dot products (IMMA.16832, two AND masks per
//     packed word, exact s32 accumulation) against x held as four signed-byte digits of a 30-bit block fixed point
//     with ONE exponent per activation row (so the digit weights leave the loop), scales applied once per
//     128-column group in fp32;
//   * x is converted once per stage and CTA from L2 (redundantly per CTA: one L2 round trip, no second
//     publish/subscribe hop).
//
// Work split: the stage's qweight rows (4 output rows each) are split evenly over the CTAs (balance to one
// qweight row); a CTA cuts its rows into 16-row tiles and each tile into units of one 128-column int4 step
// (1 KB) or 32 fp16 outlier columns (1 KB); the CTA's (tile, unit) sequence is split into 16 contiguous runs,
// one per warp.  Partial sums meet in shared memory in a fixed order (deterministic results).
#include "common.cuh"

#include <stdlib.h>

#include <vector>

namespace qeft {

constexpr int kDWarps = 16;
constexpr int kDThreads = kDWarps * 32;       // consumer threads; one more warp only fills the ring
constexpr int kDBlock = kDThreads + 32;
#ifndef QEFT_DEC_MAXREG
#define QEFT_DEC_MAXREG 96                      // 17 warps are allocated like 20: 640 x 96 = 61440 registers (104 and 120 do not launch)
#endif
#ifndef QEFT_DEC_STAGGER
#define QEFT_DEC_STAGGER 16                     // qweight-row areas 16 bytes apart (mod 128): conflict-free ldmatrix (measured: bulk copies do not care)
#endif
constexpr int kKB = 32;                       // 128-column steps per tile-block
constexpr int kQArea = kKB * 256 + QEFT_DEC_STAGGER;   // slot bytes per qweight row: 8 KB of packed words (+ optional stagger)
constexpr int kSlotW = 4 * kQArea;            // the four qweight rows of a 16-row tile
constexpr int kSideSteps = kKB * 16;           // side bytes per qweight row and block: 16 per step (+ 8 r for the outlier columns)
constexpr size_t kDSmemMax = 227 * 1024;

struct DecPart {
  const uint8_t* qw;
  const __half* scales;
  const __half* szeros;
  const __half* ow;       // plain [N, r]
  const __half* bias;
  __half* y;              // [m, N]
  uint2* y_ll;            // data-flow copy of y or null: [m * N / 2] words {two fp16 results, epoch of the run}
  int ll_consumer;        // first stage that reads y_ll (the words are written only when that stage is part of the launch)
  int pad1;
  const uint8_t* side;    // decode side table (built at program creation), per qweight row: [steps][scales of its 4 rows |
                          // scaled zeros of its 4 rows] then the fp16 outlier columns in MMA-fragment order
  int side_q;             // bytes per qweight row of `side`
  int N;
  int q_begin;            // first qweight row of this part in the stage-wide numbering (plain stages)
};

// Builds one part's side table: side[q][s] = {scales[grp(s)][4q..4q+3], szeros[grp(s)][4q..4q+3]} (16 bytes per 128-column
// step s), followed by the outlier columns of rows 4q..4q+3 as HMMA m16n8k16 A fragments: unit u (16 columns), lane-row
// gi (rows 4q+2gi, 4q+2gi+1), lane-column t: {row a cols 2t,2t+1 | row b same | row a cols 8+2t,9+2t | row b same} (16 bytes).
// A one-time re-layout of 5 % + 11 % of the layer's bytes, the counterpart of the reference's `oweight_interleaved`
// (qeft/qlinear.py:70-79, 213): every block of the kernel then needs ONE contiguous side copy per qweight row.
__global__ void dec_build_side_kernel(const __half* scales, const __half* szeros, const __half* ow, uint8_t* side, int N, int r,
                                      int g128, int nsteps, int side_q) {
  const int q = blockIdx.x;
  uint8_t* dst = side + (size_t)q * side_q;
  for (int i = threadIdx.x; i < nsteps * 8; i += blockDim.x) {
    const int s = i >> 3, e = i & 7;
    const int grp = g128 == 1 ? s : (g128 == 0 ? 0 : s / g128);
    const __half* src = (e < 4 ? scales : szeros) + (size_t)grp * N + 4 * q + (e & 3);
    reinterpret_cast<__half*>(dst)[s * 8 + e] = *src;
  }
  __half* o = reinterpret_cast<__half*>(dst + (size_t)nsteps * 16);
  for (int i = threadIdx.x; i < (r >> 4) * 64; i += blockDim.x) {
    const int u = i >> 6, w = i & 63;                        // 64 halves per unit
    const int gi = w >> 5, t = (w >> 3) & 3, e = w & 7;      // e: {a0.lo, a0.hi, a1.lo, a1.hi, a2.lo, a2.hi, a3.lo, a3.hi}
    const int row = 4 * q + 2 * gi + ((e >> 1) & 1);
    const int col = 16 * u + 2 * t + (e & 1) + ((e >> 2) & 1) * 8;
    o[i] = ow[(size_t)row * r + col];
  }
}

struct DecStage {
  DecPart part[QEFT_GEMV_MAX_PARTS];
  const __half* x;            // [m, K]
  const int32_t* gather;      // [K] or null: x[:, gather[k]] is column k
  const __half* norm_w;       // [K] or null: RMSNorm weight applied to x on the way in
  const __half* residual;     // [m, N] or null: added to the (fp16-rounded) result of part 0
  const uint2* x_ll;          // data-flow copy of x (the y_ll of the stage that produces it) or null
  const uint2* res_ll;        // the same for the residual
  int x_src, res_src;         // producing stages (the copies are valid only when those ran in THIS launch)
  int force_barrier;          // a dependency on an earlier stage that is not linked by data-flow words
  int nx_ll, nx_src;          // x_ll != null && !force_barrier / x_src of the NEXT stage (whether it waits at a barrier)
  float norm_eps;
  int nparts;
  int K, r;
  int g128;                   // G / 128, 0 for per-channel scales
  int nsteps, nchunks, nou;   // int4 steps, live 32-column chunks, outlier units (r / 32)
  int total_q;                // qweight rows of all parts
  int epilogue;               // QEFT_EPI_*
};

struct DecLayout {            // shared-memory carve-up (bytes from the start of dynamic shared memory)
  int nslots, slot;           // ring at offset 0: nslots slots of `slot` bytes: [4 x kQArea packed words][4 x sarea side bytes]
  int slot_s, sarea;
  int xdig, xdig_bytes, xsum, xo, part, part_bytes, misc;
  int debug;                  // QEFT_DECODE_DEBUG bit mask (bisecting switches; results are WRONG when set)
  unsigned long long* stamps; // debug (QEFT_DECODE_STAMPS): [stage][4 CTAs][8] globaltimer values, or null
};

__device__ __forceinline__ void dec_stamp(const DecLayout& L, int s, int i) {
  if (L.stamps && threadIdx.x == 0) {
    const int c = blockIdx.x == 0 ? 0 : (blockIdx.x == gridDim.x - 1 ? 1 : (blockIdx.x == gridDim.x / 2 ? 2 : (blockIdx.x == 1 ? 3 : -1)));
    if (c >= 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      L.stamps[((size_t)s * 4 + c) * 8 + i] = t;
    }
  }
}

// ---- small PTX helpers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t d_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void d_cp16(uint32_t dst, const void* src, uint32_t nbytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
}
__device__ __forceinline__ void d_cp8(uint32_t dst, const void* src, uint32_t nbytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
}
__device__ __forceinline__ void d_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void d_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint4 d_lds128(uint32_t a) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a) : "memory");
  return r;
}
// predicated: lanes with p == 0 do not touch shared memory and keep the previous register contents
__device__ __forceinline__ void d_lds128_if(uint4& r, uint32_t a, int p) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t@q ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];\n\t}"
               : "+r"(r.x), "+r"(r.y), "+r"(r.z), "+r"(r.w) : "r"(a), "r"(p) : "memory");
}
__device__ __forceinline__ uint32_t d_lds32(uint32_t a) {
  uint32_t r;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(a) : "memory");
  return r;
}
__device__ __forceinline__ float d_ldsf(uint32_t a) {
  float r;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(a) : "memory");
  return r;
}
__device__ __forceinline__ void d_sts128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ uint32_t d_prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}
// coherent (L2) loads for data written by other CTAs of the same launch
__device__ __forceinline__ uint4 d_ldcg128(const void* p) {
  uint4 r;
  asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ unsigned short d_ldcg16(const void* p) {
  unsigned short r;
  asm volatile("ld.global.cg.u16 %0, [%1];" : "=h"(r) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void d_imma(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                       uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void d_imma0(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                        uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
               : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "r"(0));
}

// stage-wide qweight row -> (part, part-local qweight row)
__device__ __forceinline__ void dec_locate(const DecStage* S, int vq, int& pi, int& lq) {
  if (S->epilogue == QEFT_EPI_SWIGLU) { pi = vq & 1; lq = vq >> 1; return; }
  pi = 0;
#pragma unroll
  for (int i = 1; i < QEFT_GEMV_MAX_PARTS; ++i)
    if (i < S->nparts && vq >= S->part[i].q_begin) pi = i;
  lq = vq - S->part[pi].q_begin;
}

// ---- mbarrier / bulk-copy / ldmatrix helpers ---------------------------------------------------------------
__device__ __forceinline__ void d_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void d_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void d_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\t"
               "bra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ bool d_mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// 1-D bulk copy global -> shared by the TMA engine (UBLKCP): no LSU issue slots, completion on the mbarrier
__device__ __forceinline__ void d_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// arrive on the mbarrier when all cp.async issued so far by this thread have landed (no pending-count increment)
__device__ __forceinline__ void d_cp_async_arrive(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void d_cp16p(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void d_cp8p(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void d_ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}

// barrier of the 16 consumer warps (the producer warp never joins)
__device__ __forceinline__ void d_consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kDThreads) : "memory"); }

struct DecTiles { int qa, nq, ntiles; };

// this CTA's qweight rows of a stage (balance to one qweight row), cut into 16-row tiles
__device__ __forceinline__ DecTiles dec_tiles(const DecStage* S, int cta, int ncta) {
  DecTiles R;
  const int al = S->epilogue == QEFT_EPI_SWIGLU ? 2 : 1;     // SwiGLU: gate / up qweight rows alternate, CTAs own pairs
  const unsigned Qa = (unsigned)(S->total_q / al);           // (total_q x grid < 2^31 is checked at program creation)
  R.qa = al * (int)((Qa * (unsigned)cta) / (unsigned)ncta);
  R.nq = al * (int)((Qa * (unsigned)(cta + 1)) / (unsigned)ncta) - R.qa;
  R.ntiles = (R.nq + 3) >> 2;
  return R;
}

// M: batch rows (1 or 2); the B fragment's 8 columns are M x 4 digit columns.
//
// Data path.  The CTA's work is a sequence of TILE-BLOCKS: 16 output rows (4 qweight rows) x up to kKB = 32 int4 steps
// of 128 columns (32 KB), the last block of a tile also carrying the tile's fp16 outlier columns.  A ring of L.nslots
// slots holds them; a slot is filled by
//   * 4 bulk copies (one per qweight row, up to 8 KB each, TMA engine; destination rows staggered by 16 bytes),
//   * cp.async copies of the block's scale / scaled-zero rows (8 bytes per qweight row and step) and outlier columns,
// all completing on the slot's mbarrier.  All 16 warps consume a block together (warp w: steps w and w + 16, A fragments
// by conflict-free ldmatrix straight from the copied bytes), and the LAST warp to finish a block refills its slot with
// the block nslots ahead in the sequence -- which may belong to a later stage: the stream does not stop at a stage
// boundary, ~150 KB per SM stay in flight while the CTAs meet at the barrier and convert the next x.
// LL: the launch uses data-flow words (stages ordered by polling their inputs); false compiles every such path out.
template <int M, bool LL>
__global__ void __maxnreg__(QEFT_DEC_MAXREG)
decode_w4_kernel(const DecStage* __restrict__ stages, int s_begin, int s_end, unsigned* sync, const DecLayout L,
                 int nbar_total, int uses_ll) {
  extern __shared__ __align__(128) uint8_t dsm[];
  constexpr int NCOLS = 4 * M;
  constexpr uint32_t XSTEP = 128u * NCOLS;                   // digit bytes per 128-column step: [2 nibble halves][NCOLS][4 t][16 B]
  constexpr uint32_t XHALF = 64u * NCOLS;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int cta = blockIdx.x, ncta = gridDim.x;
  const int NS = L.nslots;

  const uint32_t ring = d_smem_u32(dsm);
  const uint32_t xdig = d_smem_u32(dsm + L.xdig);
  const uint32_t xsum = d_smem_u32(dsm + L.xsum);            // [nsteps][2] fp32 group sums of x
  const uint32_t xo = d_smem_u32(dsm + L.xo);                // [M][r] fp16 outlier activations
  float* part = reinterpret_cast<float*>(dsm + L.part);      // [tile][warp][M][16 rows]
  float* red = reinterpret_cast<float*>(dsm + L.misc);       // [16 warps][4] staging reductions
  float* coef = red + kDWarps * 4;                           // [8] flush weights of a quad's accumulator columns
  const uint32_t bars = d_smem_u32(dsm + L.misc + 512);      // [nslots] "slot filled" mbarriers
  const uint32_t ebars = bars + 64;                          // [nslots] "slot consumed" mbarriers (16 warp arrivals)
  DecStage* pcache = reinterpret_cast<DecStage*>(dsm + L.misc + 768);     // descriptor of the producer's stage
  DecStage* ccache2 = reinterpret_cast<DecStage*>(dsm + L.misc + 1536);   // descriptors of the stage being consumed / the next one
  static_assert(sizeof(DecStage) <= 512 && sizeof(DecStage) % 4 == 0, "descriptor cache slots are 512 bytes");
  constexpr int kStageWords = (int)(sizeof(DecStage) / 4);

  if (tid == 0) {
    for (int i = 0; i < NS; ++i) {
      d_mbar_init(bars + 8 * i, 1);                          // one arrive.expect_tx per fill; the bytes arrive by bulk copies
      d_mbar_init(ebars + 8 * i, kDWarps);                   // one arrival per consumer warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // (stage descriptors live in global memory; every fence of the stage barrier drops them from L1, and a chain of
  // dependent L2 round trips per block is what made the first versions of this kernel slow: they are read from a
  // shared-memory copy instead)
  if (tid >= 32 && tid < 32 + kStageWords)
    reinterpret_cast<uint32_t*>(ccache2)[tid - 32] = reinterpret_cast<const uint32_t*>(stages + s_begin)[tid - 32];
  // the barrier counter only grows; `base` is its value when every CTA of this launch has started
  const unsigned base = *reinterpret_cast<volatile unsigned*>(sync + 1);
  // data-flow words carry the epoch of the run that wrote them (sync[2] = epoch of the last run that used them)
  const unsigned epoch = *reinterpret_cast<volatile unsigned*>(sync + 2) + 1u;
  int nbar = 0;                                              // stage barriers passed so far
  __syncthreads();

  // the fourth digit column of every batch row is never written: it must read as zero (its accumulator column is unused)
  for (int i = tid; i < L.xdig_bytes / 16; i += kDBlock) d_sts128(xdig + (uint32_t)i * 16, 0u, 0u, 0u, 0u);
  __syncthreads();

  // debug counters (QEFT_DECODE_STAMPS): warp 0 and warp 15 of CTA 0 and the producer warp, clock64 cycles
  long long dbg_wait = 0, dbg_math = 0, dbg_issue = 0, dbg_fill = 0, dbg_prev = 0;
  int dbg_nissue = 0, dbg_nwaited = 0, dbg_nblocks = 0;
  const bool dbg = L.stamps != nullptr && cta == 0 && (warp == 0 || warp >= kDWarps - 1);

  // =================================== the producer warp ========================================================
  // Walks the CTA's tile-blocks through ALL stages of the launch and fills the ring: it only ever waits for a slot to
  // be consumed, never for the stage barrier, so the weight stream runs ahead across stage boundaries.
  if (warp == kDWarps) {
    int pslot = 0;
    uint32_t ppar = 1;                                       // (first pass: "consumed" phases count as complete)
#pragma unroll 1
    for (int st = s_begin; st < s_end; ++st) {
      __syncwarp();
      for (int i = lane; i < kStageWords; i += 32)
        reinterpret_cast<uint32_t*>(pcache)[i] = reinterpret_cast<const uint32_t*>(stages + st)[i];
      __syncwarp();
      const DecStage* S = pcache;
      const DecTiles R = dec_tiles(S, cta, ncta);
      const int K = S->K, r = S->r, ns = S->nsteps, g128 = S->g128;
      const int KB = (ns + kKB - 1) / kKB;
      const size_t row_bytes = (size_t)(2 * K);
#pragma unroll 1
      for (int j = 0; j < R.ntiles; ++j) {
        // this lane's qweight row of the tile (lane & 3)
        const int q = lane & 3, qi = 4 * j + q;
        const bool own = qi < R.nq;
        int pi, lq;
        dec_locate(S, R.qa + (own ? qi : 0), pi, lq);
        const DecPart& P = S->part[pi];
        const unsigned own4 = __ballot_sync(0xffffffffu, own) & 0xfu;
        const uint8_t* wsrc = P.qw + (size_t)lq * row_bytes;
        const uint8_t* dsrc = P.side + (size_t)lq * (size_t)P.side_q;
#pragma unroll 1
        for (int kb = 0; kb < KB; ++kb) {
          const long long t0 = dbg ? clock64() : 0;
          const uint32_t sbase = ring + (uint32_t)pslot * (uint32_t)L.slot, bar = bars + 8 * pslot;
          d_mbar_wait(ebars + 8 * pslot, ppar);              // all 16 warps are done with the slot's previous block
          const int nsb = ns - kb * kKB < kKB ? ns - kb * kKB : kKB;
          const size_t off = (size_t)kb * (size_t)(kKB * 256);
          const uint32_t want = (uint32_t)(nsb * 256);
          const uint32_t nbytes = (row_bytes - off) < want ? (uint32_t)(row_bytes - off) : want;   // (last step of a K % 128 == 64 row)
          // side bytes of the block: its steps' scales / scaled zeros and, on the tile's last block, the outlier columns
          // (they follow the last step's scales in the table, so it is one contiguous copy)
          const uint32_t sbytes = (uint32_t)(nsb * 16 + (kb == KB - 1 ? 8 * r : 0));
          if (lane == 0) d_mbar_expect_tx(bar, (uint32_t)__popc(own4) * (((L.debug & 4) ? 0u : nbytes) + ((L.debug & 2) ? 0u : sbytes)));
          __syncwarp();
          // rows this CTA does not own are NOT copied: their slot bytes are stale, their (independent) MMA rows are never stored
          if (lane < 4 && own && !(L.debug & 4)) d_bulk_g2s(sbase + (uint32_t)(q * kQArea), wsrc + off, nbytes, bar);
          if (lane < 4 && own && !(L.debug & 2))
            d_bulk_g2s(sbase + (uint32_t)(L.slot_s + q * L.sarea), dsrc + (size_t)kb * (size_t)kSideSteps, sbytes, bar);
          if (++pslot == NS) { pslot = 0; ppar ^= 1u; }
          if (dbg) { dbg_issue += clock64() - t0; ++dbg_nissue; }
        }
      }
    }
    if (dbg && lane == 0) {
      unsigned long long* o = L.stamps + (size_t)(s_end - s_begin) * 32 + 16;
      o[0] = (unsigned long long)dbg_issue; o[1] = (unsigned long long)dbg_nissue;
    }
    return;
  }

  // =================================== the 16 consumer warps ====================================================
  int cslot = 0;
  uint32_t cpar = 0;
  const int has_col = g < NCOLS;
  // batch 1: lane t = 2 of every quad accumulates the zero-point term (its accumulator columns are no digits)
  const bool zlane = M == 1 && t == 2;
  const int gx = g < M ? g : 0;
  // ldmatrix row addresses: lane l supplies row (l & 7) of matrix (l >> 3): matrices 0 / 2 = tile rows 2 i (MMA rows 0..7),
  // 1 / 3 = tile rows 2 i + 1 (MMA rows 8..15); matrices 2, 3 are the second 16-byte chunk
  const int lrow = 2 * (lane & 7) + ((lane >> 3) & 1);
  const uint32_t laneA = (uint32_t)((lrow >> 2) * kQArea + (lrow & 3) * 32 + (lane >> 4) * 16);
  // side area of this lane's qweight row (g >> 1): per step 16 bytes {scales of its 4 rows | scaled zeros}; rows 2g, 2g+1
  // are the pair (g & 1) of the row; the outlier fragments follow the block's last step
  const uint32_t laneS = (uint32_t)(L.slot_s + (g >> 1) * L.sarea + (g & 1) * 4 + (zlane ? 8 : 0));
  const uint32_t laneO = (uint32_t)(L.slot_s + (g >> 1) * L.sarea + (g & 1) * 64 + t * 16);
  const uint32_t xdig_lane = xdig + (uint32_t)(g * 64 + t * 16);

#pragma unroll 1
  for (int s = s_begin; s < s_end; ++s) {
    const DecStage* S = reinterpret_cast<const DecStage*>(reinterpret_cast<const uint8_t*>(ccache2) + ((s - s_begin) & 1) * 512);

    // x produced by an earlier stage of this launch and linked by data-flow words: the stage polls the words it reads,
    // no barrier.  Every other stage after the first waits until all CTAs have stored their rows of the previous one.
    // (The descriptor copy of stage s was written before the previous stage's post-consume barrier.)
    const bool ll_x = LL && S->x_ll != nullptr && S->x_src >= s_begin && !S->force_barrier;
    if (s > s_begin && !ll_x) {
      ++nbar;
      if (tid == 0) {
        const unsigned want = base + (unsigned)nbar * (unsigned)ncta;
        // (polling with relaxed loads and one acquire fence at the end was measured slower: +0.4 us per stage)
        unsigned got;
        do {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(got) : "l"(sync) : "memory");
        } while ((int)(got - want) < 0);
        if (cta == 0 && nbar == 1)              // every CTA has read `base`: publish the next launch's base
          *reinterpret_cast<volatile unsigned*>(sync + 1) = base + (unsigned)nbar_total * (unsigned)ncta;
      }
      d_consumer_sync();
    }
    const DecTiles R = dec_tiles(S, cta, ncta);
    const int K = S->K, r = S->r, ns = S->nsteps;

    dec_stamp(L, s, 0);
    // ---- x: block fixed point per 128-column step, three signed-byte digits ------------------------------------------
    //   x_k ~= X_k 2^(e-22),  X_k = d0 + 256 d1 + 65536 d2,  d_i in [-128, 127],  2^e > max |x| of the step
    // (exact for every element within 12 binades of the step's maximum; fp16 has 11 significant bits).  One exponent per
    // STEP keeps the staging free of a CTA-wide reduction (the stage boundary's critical path: measured 0.8 us for the
    // first version's single exponent per row); the price is one multiply per row pair and step in the main loop.
    {
      const __half* xg = S->x;
      // o_proj's gather (qlinear.py:275): copy x to shared memory first (coalesced, one L2 round trip; the buffer aliases
      // the partial-sum slices, idle between two stages), then gather from there instead of 16 scattered 2-byte L2 loads
      const bool xraw_ok = S->gather != nullptr && !ll_x && (size_t)M * (size_t)K * 2 <= (size_t)L.part_bytes;
      const uint32_t xraw = d_smem_u32(dsm + L.part);
      if (xraw_ok) {
        for (int i = tid; i < M * (K >> 3); i += kDThreads) {
          const uint4 v = d_ldcg128(xg + (size_t)i * 8);
          d_sts128(xraw + (uint32_t)i * 16, v.x, v.y, v.z, v.w);
        }
        d_consumer_sync();
      }
      // data-flow input: x is read from the 8-byte words {two fp16 values, epoch} the producing stage stores, each load
      // repeated until the word carries this run's epoch (no barrier, no fence: the word is its own flag)
      const uint2* xll = ll_x ? S->x_ll : nullptr;
      auto ll_word = [&](const uint2* base_ll, size_t idx) -> uint32_t {
        uint2 w;
        do {
          asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(w.x), "=r"(w.y) : "l"(base_ll + idx) : "memory");
        } while (w.y != epoch);
        return w.x;
      };
      auto ldx16 = [&](const __half* row, int b, int col) -> unsigned short {
        if (xll) {
          const uint32_t w = ll_word(xll, ((size_t)b * K + col) >> 1);
          return (unsigned short)((col & 1) ? (w >> 16) : (w & 0xffffu));
        }
        if (xraw_ok) {
          unsigned short r16;
          asm volatile("ld.shared.u16 %0, [%1];" : "=h"(r16) : "r"(xraw + (uint32_t)((b * K + col) * 2)) : "memory");
          return r16;
        }
        return d_ldcg16(row + col);
      };
      const int32_t* gat = S->gather;
      const __half* nw = S->norm_w;
      const int live_k = S->nchunks * 32;
      const int nitems = M * ns * 8;
      const int npass = (nitems + kDThreads - 1) / kDThreads;
      // item = (batch row b, step, chunk tt, hs): the 16 columns k0 .. k0+7 and k0+16 .. k0+23, k0 = 128 step + 32 tt + 8 hs;
      // the 8 items of a step sit in 8 adjacent lanes, which agree on the step's sum and maximum by shuffles
      auto load_item = [&](int it, uint4& v0, uint4& v1, int& b, int& st, bool& live) {
        const int sb = it >> 3;
        b = sb / ns;
        st = sb - b * ns;
        const int k0 = st * 128 + ((it >> 1) & 3) * 32 + (it & 1) * 8;
        live = it < nitems && k0 < live_k;
        v0 = v1 = make_uint4(0u, 0u, 0u, 0u);
        if (live) {
          const __half* xr = xg + (size_t)b * K;
          if (gat) {
            const int4 i0 = __ldg(reinterpret_cast<const int4*>(gat + k0)), i1 = __ldg(reinterpret_cast<const int4*>(gat + k0 + 4));
            const int4 i2 = __ldg(reinterpret_cast<const int4*>(gat + k0 + 16)), i3 = __ldg(reinterpret_cast<const int4*>(gat + k0 + 20));
            auto pk = [&](int a, int c) { return (uint32_t)ldx16(xr, b, a) | ((uint32_t)ldx16(xr, b, c) << 16); };
            v0 = make_uint4(pk(i0.x, i0.y), pk(i0.z, i0.w), pk(i1.x, i1.y), pk(i1.z, i1.w));
            v1 = make_uint4(pk(i2.x, i2.y), pk(i2.z, i2.w), pk(i3.x, i3.y), pk(i3.z, i3.w));
          } else if (xll) {
            // 8 + 8 values = 4 + 4 words: all loads first (one round trip when the data is there), then re-poll stragglers
            const uint2* p0 = xll + (((size_t)b * K + k0) >> 1);
            uint2 w[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(w[j].x), "=r"(w[j].y) : "l"(p0 + (j < 4 ? j : j + 4)) : "memory");
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (w[j].y != epoch) w[j].x = ll_word(p0, (size_t)(j < 4 ? j : j + 4));
            v0 = make_uint4(w[0].x, w[1].x, w[2].x, w[3].x);
            v1 = make_uint4(w[4].x, w[5].x, w[6].x, w[7].x);
          } else {
            v0 = d_ldcg128(xr + k0);
            v1 = d_ldcg128(xr + k0 + 16);
          }
        }
      };
      // the loads of the first two passes and of the outlier activations are in flight together
      uint4 k0a, k0b, k1a, k1b;
      int kb0, kb1, ks0, ks1;
      bool kl0, kl1;
      load_item(tid, k0a, k0b, kb0, ks0, kl0);
      load_item(kDThreads + tid, k1a, k1b, kb1, ks1, kl1);
      const int nxo = M * (r >> 3);
      uint4 xo_v = make_uint4(0u, 0u, 0u, 0u);
      const int xo_tid = kDThreads - 1 - tid;              // the threads the digit items use least
      if (xo_tid < nxo) {
        const int b = xo_tid / (r >> 3), jj = xo_tid - b * (r >> 3);
        const __half* xr = xg + (size_t)b * K;
        if (gat) {
          const int4 a = __ldg(reinterpret_cast<const int4*>(gat + K - r + 8 * jj)), c = __ldg(reinterpret_cast<const int4*>(gat + K - r + 8 * jj + 4));
          auto pk = [&](int i0, int i1) { return (uint32_t)ldx16(xr, b, i0) | ((uint32_t)ldx16(xr, b, i1) << 16); };
          xo_v = make_uint4(pk(a.x, a.y), pk(a.z, a.w), pk(c.x, c.y), pk(c.z, c.w));
        } else if (xll) {
          const size_t i0 = ((size_t)b * K + K - r + 8 * jj) >> 1;
          xo_v = make_uint4(ll_word(xll, i0), ll_word(xll, i0 + 1), ll_word(xll, i0 + 2), ll_word(xll, i0 + 3));
        } else {
          xo_v = d_ldcg128(xr + K - r + 8 * jj);
        }
      }
      float rs0 = 1.f, rs1 = 1.f;
      if (nw) {
        // RMSNorm on the way in (HF LlamaRMSNorm / kernel/layernorm/layernorm.cu:25-51): the row's sum of squares needs the
        // whole row: one CTA-wide reduction (only for stages with a norm)
        float ss0 = 0.f, ss1 = 0.f;
        auto sumsq = [&](const uint4& v0, const uint4& v1, int b) {
          const uint32_t w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
          float ss = 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) { const float2 f = half2_bits_to_float2(w[j]); ss = fmaf(f.x, f.x, fmaf(f.y, f.y, ss)); }
          if (b == 0) ss0 += ss; else ss1 += ss;
        };
        if (kl0) sumsq(k0a, k0b, kb0);
        if (kl1) sumsq(k1a, k1b, kb1);
        for (int q = 2; q < npass; ++q) {
          uint4 v0, v1; int b, st; bool live;
          load_item(q * kDThreads + tid, v0, v1, b, st, live);
          if (live) sumsq(v0, v1, b);
        }
        if (xo_tid < nxo) sumsq(xo_v, make_uint4(0u, 0u, 0u, 0u), xo_tid / (r >> 3));
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
          ss0 += __shfl_xor_sync(0xffffffffu, ss0, o);
          ss1 += __shfl_xor_sync(0xffffffffu, ss1, o);
        }
        if (lane == 0) { red[warp * 4 + 2] = ss0; red[warp * 4 + 3] = ss1; }
        d_consumer_sync();
        ss0 = ss1 = 0.f;
#pragma unroll
        for (int w = 0; w < kDWarps; ++w) { ss0 += red[w * 4 + 2]; ss1 += red[w * 4 + 3]; }
        rs0 = rsqrtf(ss0 / (float)K + S->norm_eps);
        rs1 = rsqrtf(ss1 / (float)K + S->norm_eps);
      }
      dec_stamp(L, s, 4);
      // the reference's two roundings: (x * rs).to(fp16), then * weight in fp16
      auto normed = [&](float v, int col, float rs) { return __half2float(__hmul(nw[col], __float2half_rn(v * rs))); };
      for (int q = 0; q < npass; ++q) {
        const int it = q * kDThreads + tid;
        uint4 v0, v1; int b, st; bool live;
        if (q == 0) { v0 = k0a; v1 = k0b; b = kb0; st = ks0; live = kl0; }
        else if (q == 1) { v0 = k1a; v1 = k1b; b = kb1; st = ks1; live = kl1; }
        else load_item(it, v0, v1, b, st, live);
        const bool valid = it < nitems;
        const int tt = (it >> 1) & 3, hs = it & 1;
        const uint32_t w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
        float2 f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = half2_bits_to_float2(w[j]);
        if (nw && live) {
          const int k0 = st * 128 + tt * 32 + hs * 8;
          const float rs = b ? rs1 : rs0;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int k = k0 + (j < 4 ? 2 * j : 16 + 2 * (j - 4));
            const int c0 = gat ? gat[k] : k, c1 = gat ? gat[k + 1] : k + 1;
            f[j] = make_float2(normed(f[j].x, c0, rs), normed(f[j].y, c1, rs));
          }
        }
        float sum = 0.f, mx = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          sum += f[j].x + f[j].y;
          mx = fmaxf(mx, fmaxf(fabsf(f[j].x), fabsf(f[j].y)));
        }
#pragma unroll
        for (int o = 4; o >= 1; o >>= 1) {
          sum += __shfl_xor_sync(0xffffffffu, sum, o);
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        if (valid) {
          // 2^e > mx;  X = rint(x 2^(22-e)) by the magic-number add (|X| < 2^22): bits(fma(x, sc, 1.5 2^23)) = 0x4B400000 + X.
          // Z = X + 0x808080 has unsigned bytes b_i with X = sum (b_i - 128) 256^i: the signed digits are the bytes of Z ^ 0x808080.
          const int e = mx > 0.f ? (int)((__float_as_uint(mx) >> 23) & 0xff) - 126 : -100;
          const float sc = e > -100 ? __uint_as_float((uint32_t)(127 + 22 - e) << 23) : 0.f;
          auto digits = [&](float v) {
            const int bits = __float_as_int(fmaf(v, sc, 12582912.f));
            return (uint32_t)(bits + (0x00808080 - 0x4B400000)) ^ 0x00808080u;
          };

          // Word c of digit column d = bytes {first[2c], second[2c], first[2c+1], second[2c+1]} (first = k0 + ., second =
          // k0 + 16 + .): the B-fragment register of lane t = c for this item's 32-column chunk tt = 2 T + h and nibble
          // position hs.  Lane t's 16-byte row of (hs, column) holds its four chunks' words, index tt.
          const uint32_t dst = xdig + (uint32_t)st * XSTEP + (uint32_t)hs * XHALF + (uint32_t)(4 * b) * 64 + (uint32_t)tt * 4;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            // (f[c] = first[2c], first[2c+1]; f[4 + c] = second[2c], second[2c+1]: four digit words live at a time)
            const uint32_t da = digits(f[c].x), db = digits(f[4 + c].x), dc = digits(f[c].y), dd = digits(f[4 + c].y);
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              const uint32_t sel = 0x0040u + 0x11u * (uint32_t)d;
              const uint32_t ww = d_prmt(d_prmt(da, db, sel), d_prmt(dc, dd, sel), 0x5410u);
              asm volatile("st.shared.u32 [%0], %1;" ::"r"(dst + (uint32_t)(d * 64 + c * 16)), "r"(ww) : "memory");
            }
          }
          if ((it & 7) == 0) {
            // per step and batch row: {sum of x, weight of digit 0 = 2^(e-22) / 16 (the nibble trick's 16 q)}
            const float cg = e > -100 ? __uint_as_float((uint32_t)(127 + e - 22 - 4) << 23) : 0.f;
            asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(xsum + (uint32_t)(st * 16 + b * 8)), "f"(sum), "f"(cg) : "memory");
          }
        }
      }
      dec_stamp(L, s, 6);
      if (xo_tid < nxo) {
        const int b = xo_tid / (r >> 3), jj = xo_tid - b * (r >> 3);
        if (nw) {
          const float rs = b ? rs1 : rs0;
          uint32_t w[4] = {xo_v.x, xo_v.y, xo_v.z, xo_v.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int k = K - r + 8 * jj + 2 * j;
            const int c0 = gat ? gat[k] : k, c1 = gat ? gat[k + 1] : k + 1;
            const float2 f = half2_bits_to_float2(w[j]);
            const __half2 h = __halves2half2(__hmul(nw[c0], __float2half_rn(f.x * rs)), __hmul(nw[c1], __float2half_rn(f.y * rs)));
            w[j] = *reinterpret_cast<const uint32_t*>(&h);
          }
          xo_v = make_uint4(w[0], w[1], w[2], w[3]);
        }
        d_sts128(xo + (uint32_t)((b * r + 8 * jj) * 2), xo_v.x, xo_v.y, xo_v.z, xo_v.w);
      }
      d_consumer_sync();
    }

    // the next stage's descriptor, into the other buffer: every warp is past its last read of the stage before this one,
    // and the copy lands long before the post-consume barrier after which the next stage reads it
    static_assert(sizeof(DecStage) % 16 == 0, "the descriptor is copied in 16-byte pieces");
    if (s + 1 < s_end && tid >= 32 && tid < 32 + (int)(sizeof(DecStage) / 16))
      d_cp16p(d_smem_u32(reinterpret_cast<uint8_t*>(ccache2) + ((s + 1 - s_begin) & 1) * 512) + (uint32_t)(tid - 32) * 16,
              reinterpret_cast<const uint8_t*>(stages + s + 1) + (size_t)(tid - 32) * 16);
    dec_stamp(L, s, 1);
    // ---- the tile-blocks of the stage -------------------------------------------------------------------------
    {
      const int KB = (ns + kKB - 1) / kKB;
      const uint32_t xo_lane = xo + (uint32_t)((gx * r + 2 * t) * 2);
      uint4 xe = make_uint4(0u, 0u, 0u, 0u), xq = xe, xe2 = xe, xq2 = xe;
      float2 xs0 = make_float2(0.f, 0.f), xs1 = xs0;
      // B fragments (digit bytes), group sum and digit weight of one step
      auto load_x = [&](int gs, uint4& xe_, uint4& xq_, float2& xs) {
        const uint32_t xc = xdig_lane + (uint32_t)gs * XSTEP;
        d_lds128_if(xe_, xc, has_col);
        d_lds128_if(xq_, xc + XHALF, has_col);
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(xs.x), "=f"(xs.y) : "r"(xsum + (uint32_t)(gs * 16 + (M == 2 ? (t >> 1) * 8 : 0))) : "memory");
      };
      // single-k-block stages (K <= 4096 + r): a warp works on the same two steps of every tile: their B operands are
      // loaded once per stage and stay in registers
      const bool bcache = KB == 1;
      if (bcache) {
        if (warp < ns) load_x(warp, xe, xq, xs0);
        if (warp + kDWarps < ns) load_x(warp + kDWarps, xe2, xq2, xs1);
      }
#pragma unroll 1
      for (int j = 0; j < R.ntiles; ++j) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};    // rows 2g, 2g+1 x accumulator columns 2t, 2t+1: sum over groups of scale * P
        float zacc0 = 0.f, zacc1 = 0.f;         // M = 2: rows 2g, 2g+1, sum over groups of scaled zero * X_g, batch row t >> 1
        float yo[4] = {0.f, 0.f, 0.f, 0.f};     // outlier columns: rows 2g, 2g+1 x batch rows 2t, 2t+1
#pragma unroll 1
        for (int kb = 0; kb < KB; ++kb) {
          const uint32_t sbase = ring + (uint32_t)cslot * (uint32_t)L.slot;
          // wait for the block (lane 0 polls, so that the loop is warp-uniform); meanwhile help refilling free slots
          const long long tw0 = dbg ? clock64() : 0;
          if (dbg && dbg_prev) dbg_fill += tw0 - dbg_prev;   // (block-to-block period inside a stage)
          dbg_prev = tw0;
          const bool ready0 = dbg ? d_mbar_test(bars + 8 * cslot, cpar) : true;
          d_mbar_wait(bars + 8 * cslot, cpar);               // the block's bytes have landed (every lane acquires them)
          const long long tw1 = dbg ? clock64() : 0;
          if (dbg) { dbg_wait += tw1 - tw0; ++dbg_nblocks; if (!ready0) ++dbg_nwaited; }
          const int nsb = ns - kb * kKB < kKB ? ns - kb * kKB : kKB;
          // one 128-column int4 step of 16 rows: A fragments by ldmatrix from the copied bytes, two AND masks per word
          // (low nibbles q, high nibbles 16 q: both valid u8), 4 IMMA with exact s32 accumulation; 16 lo + hi = 16 sum(q X).
          // A warp's two steps of the block (warp, warp + 16) are loaded together and then computed: twice the loads
          // in flight per warp.
          auto load_step = [&](int st, uint32_t (&a0)[4], uint32_t (&a1)[4], uint4& xe_, uint4& xq_, uint32_t& sw, uint32_t& zw, float2& xs) {
            d_ldmatrix_x4(a0, sbase + laneA + (uint32_t)(st * 256));
            d_ldmatrix_x4(a1, sbase + laneA + (uint32_t)(st * 256 + 128));
            if (!bcache) load_x(kb * kKB + st, xe_, xq_, xs);
            sw = d_lds32(sbase + laneS + (uint32_t)(st * 16));
            zw = M == 2 ? d_lds32(sbase + laneS + (uint32_t)(st * 16 + 8)) : 0u;
          };
          auto math_step = [&](const uint32_t (&a0)[4], const uint32_t (&a1)[4], const uint4& xe_, const uint4& xq_, uint32_t sw, uint32_t zw, float2 xs2) {
            const float xs = xs2.x;
            constexpr uint32_t kLoM = 0x0f0f0f0fu, kHiM = 0xf0f0f0f0u;
            int lo[4], hi[4];
            d_imma0(lo, a0[0] & kLoM, a0[1] & kLoM, a0[2] & kLoM, a0[3] & kLoM, xe_.x, xe_.y);
            d_imma0(hi, a0[0] & kHiM, a0[1] & kHiM, a0[2] & kHiM, a0[3] & kHiM, xq_.x, xq_.y);
            d_imma(lo, a1[0] & kLoM, a1[1] & kLoM, a1[2] & kLoM, a1[3] & kLoM, xe_.z, xe_.w);
            d_imma(hi, a1[0] & kHiM, a1[1] & kHiM, a1[2] & kHiM, a1[3] & kHiM, xq_.z, xq_.w);
            float2 sc = half2_bits_to_float2(sw);
            float f0 = (float)(lo[0] * 16 + hi[0]), f2 = (float)(lo[2] * 16 + hi[2]);
            const float f1 = (float)(lo[1] * 16 + hi[1]), f3 = (float)(lo[3] * 16 + hi[3]);
            if (M == 1) {
              if (zlane) { f0 = xs; f2 = xs; }               // scaled zero x group sum of x in the zero-point lane
              else { sc.x *= xs2.y; sc.y *= xs2.y; }         // digit lanes: scale x the step's digit weight
            } else {
              sc.x *= xs2.y; sc.y *= xs2.y;
              const float2 zz = half2_bits_to_float2(zw);
              zacc0 = fmaf(zz.x, xs, zacc0);
              zacc1 = fmaf(zz.y, xs, zacc1);
            }
            acc[0] = fmaf(sc.x, f0, acc[0]);
            acc[1] = fmaf(sc.x, f1, acc[1]);
            acc[2] = fmaf(sc.y, f2, acc[2]);
            acc[3] = fmaf(sc.y, f3, acc[3]);
          };
          if (warp < nsb && !(L.debug & 1)) {
            uint32_t a0[4], a1[4], b0[4], b1[4], sw0, sw1 = 0, zw0, zw1 = 0;
            const bool two = warp + kDWarps < nsb;
            load_step(warp, a0, a1, xe, xq, sw0, zw0, xs0);
            if (two) load_step(warp + kDWarps, b0, b1, xe2, xq2, sw1, zw1, xs1);
            math_step(a0, a1, xe, xq, sw0, zw0, xs0);
            if (two) math_step(b0, b1, xe2, xq2, sw1, zw1, xs1);
          }
          if (kb == KB - 1 && r > 0 && !(L.debug & 1)) {
            // 16 fp16 outlier columns of 16 rows per unit: one HMMA; units go to the warps from the top
#pragma unroll 1
            for (int u = kDWarps - 1 - warp; u < (r >> 4); u += kDWarps) {
              const uint4 a4 = d_lds128(sbase + laneO + (uint32_t)(nsb * 16 + u * 128));
              const uint32_t a[4] = {a4.x, a4.y, a4.z, a4.w};
              const uint32_t b0 = d_lds32(xo_lane + (uint32_t)(u * 32)), b1 = d_lds32(xo_lane + (uint32_t)(u * 32 + 16));
              mma_m16n8k16_f16f32(yo, a[0], a[1], a[2], a[3], b0, b1);
            }
          }
          // this warp is done with the slot
          __syncwarp();
          if (dbg) dbg_math += clock64() - tw1;
          if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ebars + 8 * cslot) : "memory");
          if (++cslot == NS) { cslot = 0; cpar ^= 1u; }
        }
        {
          // end of the tile: accumulator columns -> one value per (row, batch row), to this warp's slice of the tile
          // digit columns of a batch row weigh 1, 256, 65536 (the fourth column is unused); M = 1: lane t = 2 is the zero-point lane
          const float c0 = (M == 1 ? (t == 0 ? 1.f : (t == 1 ? 65536.f : (t == 2 ? 1.f : 0.f))) : ((t & 1) ? 65536.f : 1.f));
          const float c1 = (M == 1 ? (t == 0 ? 256.f : 0.f) : ((t & 1) ? 0.f : 256.f));
          float v1 = fmaf(c0, acc[0], c1 * acc[1]), v2 = fmaf(c0, acc[2], c1 * acc[3]);
          v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
          v2 += __shfl_xor_sync(0xffffffffu, v2, 1);
          float* dst = part + (size_t)((j * kDWarps + warp) * M) * 16;
          if (M == 1) {
            v1 += __shfl_xor_sync(0xffffffffu, v1, 2);
            v2 += __shfl_xor_sync(0xffffffffu, v2, 2);
            if (t == 0) { dst[2 * g] = v1 + yo[0]; dst[2 * g + 1] = v2 + yo[2]; }
          } else {
            // batch row 1's outlier sums live in lane t = 0 of the quad (accumulator column 1)
            const float o1 = __shfl_sync(0xffffffffu, yo[1], lane & ~3), o3 = __shfl_sync(0xffffffffu, yo[3], lane & ~3);
            if (t == 0) { dst[2 * g] = v1 + zacc0 + yo[0]; dst[2 * g + 1] = v2 + zacc1 + yo[2]; }
            if (t == 2) { dst[16 + 2 * g] = v1 + zacc0 + o1; dst[16 + 2 * g + 1] = v2 + zacc1 + o3; }
          }
        }
      }
    }
    dbg_prev = 0;
    asm volatile("cp.async.wait_all;" ::: "memory");        // (the next stage's descriptor)
    d_consumer_sync();
    dec_stamp(L, s, 2);

    // ---- add the warps' slices in a fixed order, epilogue, store the rows this CTA owns -------------------------
    {
      const int epi = S->epilogue;
      const int nitems = R.ntiles * 16 * M;
      const uint2* rll = (LL && S->res_ll != nullptr && S->res_src >= s_begin) ? S->res_ll : nullptr;
      for (int i0 = warp * 32; i0 < nitems; i0 += kDThreads) {          // whole warps: the pair exchange below is a shuffle
        const int i = i0 + lane;
        const int b = i % M, rr = (i / M) & 15, j = i / (16 * M);
        const int qi = 4 * j + (rr >> 2);
        // (SwiGLU: up rows are consumed by their gate rows)
        const bool valid = i < nitems && qi < R.nq && !(epi == QEFT_EPI_SWIGLU && (rr & 4));
        auto tile_sum = [&](int row) {
          const float* src = part + (size_t)(j * kDWarps * M + b) * 16 + row;
          float a = 0.f;
#pragma unroll
          for (int w = 0; w < kDWarps; ++w) a += src[w * M * 16];
          return a;
        };
        __half h = __float2half_rn(0.f);
        int n = 0;
        const DecPart* Pp = &S->part[0];
        if (valid) {
          int pi, lq;
          dec_locate(S, R.qa + qi, pi, lq);
          Pp = &S->part[pi];
          n = 4 * lq + (rr & 3);
          float a = tile_sum(rr);
          if (Pp->bias) a += __half2float(Pp->bias[n]);
          h = __float2half_rn(a);
          if (epi == QEFT_EPI_SWIGLU) {
            // silu(gate) * up, both rounded to fp16 first like the unfused linears (HF LlamaMLP: act_fn(gate_proj(x)) * up_proj(x))
            const DecPart& Pu = S->part[1];
            float u = tile_sum(rr + 4);
            if (Pu.bias) u += __half2float(Pu.bias[n]);
            const float gf = __half2float(h);
            const __half sg = __float2half_rn(gf / (1.f + __expf(-gf)));
            h = __hmul(sg, __float2half_rn(u));
          } else if (S->residual) {
            __half res;
            if (rll) {
              // the residual was produced in this launch: read it from its data-flow words
              uint2 w;
              do {
                asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(w.x), "=r"(w.y)
                             : "l"(rll + (((size_t)b * Pp->N + n) >> 1)) : "memory");
              } while (w.y != epoch);
              res = __ushort_as_half((unsigned short)((n & 1) ? (w.x >> 16) : (w.x & 0xffffu)));
            } else {
              res = __ushort_as_half(d_ldcg16(S->residual + (size_t)b * Pp->N + n));
            }
            h = __hadd(res, h);
          }
          Pp->y[(size_t)b * Pp->N + n] = h;
        }
        // data-flow copy for the stages of this launch that read y: rows n (even) and n + 1 sit M lanes apart
        const unsigned hb = (unsigned)__half_as_ushort(h);
        const unsigned hn = LL ? __shfl_down_sync(0xffffffffu, hb, M) : 0u;
        if (LL && valid && !(n & 1) && uses_ll && Pp->y_ll != nullptr && Pp->ll_consumer < s_end) {
          uint2 w = make_uint2(hb | (hn << 16), epoch);
          asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(Pp->y_ll + (((size_t)b * Pp->N + n) >> 1)), "r"(w.x), "r"(w.y) : "memory");
        }
      }
    }
    if (s + 1 < s_end && !(LL && S->nx_ll && S->nx_src >= s_begin)) {
      // the next stage waits at a barrier (one signal per CTA: one signal per warp was measured slower -- 16 x 148
      // atomics on one address per stage)
      d_consumer_sync();
      if (tid == 0) asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(sync), "r"(1u) : "memory");
    }
    dec_stamp(L, s, 3);
  }
  // the run's epoch becomes the base of the next run's (every CTA read it at its start: this CTA could only get here
  // after consuming words of all the others)
  if (LL && uses_ll && cta == 0 && tid == 0) *reinterpret_cast<volatile unsigned*>(sync + 2) = epoch;
  if (dbg && lane == 0) {
    unsigned long long* o = L.stamps + (size_t)(s_end - s_begin) * 32 + (warp == 0 ? 0 : 8);
    o[0] = (unsigned long long)dbg_wait; o[1] = (unsigned long long)dbg_math; o[2] = 0;
    o[3] = (unsigned long long)dbg_fill; o[4] = 0; o[5] = (unsigned long long)dbg_nwaited;
    o[6] = (unsigned long long)dbg_nblocks; o[7] = (unsigned long long)clock64();
  }
}

// ----------------------------------------------------------------------------------------------------
struct DecProgram {
  DecStage* d_stages = nullptr;
  unsigned* d_sync = nullptr;
  unsigned long long* d_stamps = nullptr;
  std::vector<void*> side_tables;
  std::vector<DecStage> h_stages;
  int m = 1;
  int device = 0;
  int nsm = 0;
};

static int dec_env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

template <int M, bool LL>
static int dec_launch(const DecProgram* p, int s0, int s1, const DecLayout& L, size_t smem, int grid, cudaStream_t stream,
                      int nbar_total, int uses_ll) {
  auto kern = decode_w4_kernel<M, LL>;
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDSmemMax);
    if (e != cudaSuccess) return (int)e;
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kDBlock);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;     // all CTAs co-resident: they wait for one another at stage boundaries
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (s1 - s0 > 1) ? 1 : 0;        // (barriers or data-flow polling between stages: CTAs wait for one another)
  const DecStage* st = p->d_stages;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, st, s0, s1, p->d_sync, L, nbar_total, uses_ll);
  if (e != cudaSuccess) return (int)e;
  count_launch();
  return QEFT_OK;
}

}  // namespace qeft

using namespace qeft;

extern "C" int qeft_decode_program_create(const qeft_decode_stage_t* stages, int nstages, int m, qeft_decode_program_t** out) {
  if (!stages || !out) return QEFT_E_NULL;
  if (nstages < 1) return QEFT_E_SHAPE;
  if (m < 1 || m > 2) return QEFT_E_BATCH;
  DecProgram* p = new DecProgram();
  p->m = m;
  p->h_stages.resize(nstages);
  for (int s = 0; s < nstages; ++s) {
    const qeft_decode_stage_t& q = stages[s];
    DecStage& d = p->h_stages[s];
    d = DecStage{};
    int G = q.G == -1 ? q.K : q.G;
    const int K = q.K, r = q.r;
    int st = QEFT_OK;
    if (!q.x) st = QEFT_E_NULL;
    else if (q.nparts < 1 || q.nparts > QEFT_GEMV_MAX_PARTS) st = QEFT_E_SHAPE;
    else if (K <= 0 || K % 64 != 0 || G <= 0 || K % G != 0 || (G % 128 != 0 && G != K)) st = QEFT_E_SHAPE;
    else if (r < 0 || r % 32 != 0 || r >= K) st = QEFT_E_SHAPE;
    else if (!check_align16(q.x) || (q.x_gather && !check_align16(q.x_gather))) st = QEFT_E_ALIGN;
    else if (q.epilogue != QEFT_EPI_NONE && q.epilogue != QEFT_EPI_SWIGLU && q.epilogue != QEFT_EPI_RESIDUAL) st = QEFT_E_DTYPE;
    else if (q.epilogue == QEFT_EPI_SWIGLU && (q.nparts != 2 || q.parts[0].N != q.parts[1].N)) st = QEFT_E_SHAPE;
    else if (q.epilogue == QEFT_EPI_RESIDUAL && (q.nparts != 1 || !q.residual)) st = QEFT_E_NULL;
    int total_q = 0;
    for (int i = 0; st == QEFT_OK && i < q.nparts; ++i) {
      const qeft_gemv_part_t& a = q.parts[i];
      const bool need_y = !(q.epilogue == QEFT_EPI_SWIGLU && i == 1);
      if (!a.qweight || !a.scales || !a.scaled_zeros || (need_y && !a.y) || (r > 0 && !a.oweight)) st = QEFT_E_NULL;
      else if (a.N <= 0 || a.N % 8 != 0) st = QEFT_E_SHAPE;
      else if (!check_align16(a.qweight) || !check_align16(a.scales) || !check_align16(a.scaled_zeros) ||
               (r > 0 && !check_align16(a.oweight)) || (r % 8 != 0))
        st = QEFT_E_ALIGN;
      if (st != QEFT_OK) break;
      DecPart& dp = d.part[i];
      dp.qw = static_cast<const uint8_t*>(a.qweight);
      dp.scales = static_cast<const __half*>(a.scales);
      dp.szeros = static_cast<const __half*>(a.scaled_zeros);
      dp.ow = r > 0 ? static_cast<const __half*>(a.oweight) : static_cast<const __half*>(a.scales);
      dp.bias = static_cast<const __half*>(a.bias);
      dp.y = static_cast<__half*>(a.y);
      dp.y_ll = nullptr;
      dp.ll_consumer = 1 << 30;
      dp.side = nullptr;
      dp.side_q = cdiv(K - r, 128) * 16 + 8 * r;
      dp.N = a.N;
      dp.q_begin = total_q;
      total_q += a.N / 4;
    }
    if (st == QEFT_OK && (long)total_q * 1024 >= (1L << 31)) st = QEFT_E_UNSUPPORTED;   // 32-bit row arithmetic in the kernel
    if (st != QEFT_OK) { delete p; return st; }
    d.x = static_cast<const __half*>(q.x);
    d.gather = q.x_gather;
    d.norm_w = static_cast<const __half*>(q.norm_weight);
    d.norm_eps = q.norm_eps;
    d.residual = q.epilogue == QEFT_EPI_RESIDUAL ? static_cast<const __half*>(q.residual) : nullptr;
    d.nparts = q.nparts;
    d.K = K; d.r = r;
    d.g128 = (G == K) ? 0 : G / 128;
    d.nsteps = cdiv(K - r, 128);
    d.nchunks = (K - r) / 32;
    d.nou = r / 32;
    d.total_q = total_q;
    d.epilogue = q.epilogue;
  }
  // Data-flow links: a stage whose x (or residual) IS the y of an earlier stage's projection reads it from that
  // projection's data-flow words (allocated here, owned by the program) instead of waiting at a barrier.
  const int ll_env = dec_env_int("QEFT_DECODE_LL", 0);      // (read at every creation: tests build both kinds of program)
  for (int s = 0; s < nstages; ++s) {
    DecStage& d = p->h_stages[s];
    d.x_ll = nullptr; d.res_ll = nullptr; d.x_src = -1; d.res_src = -1; d.force_barrier = 0; d.nx_ll = 0; d.nx_src = -1;
  }
  auto link = [&](const void* ptr, int width, int s, const uint2*& out_ll, int& out_src) -> int {
    // the latest earlier stage with a projection whose output buffer is exactly `ptr` ([m, width])
    for (int ps = s - 1; ps >= 0; --ps) {
      DecStage& pd = p->h_stages[ps];
      const int nout = pd.epilogue == QEFT_EPI_SWIGLU ? 1 : pd.nparts;
      for (int i = 0; i < nout; ++i) {
        DecPart& pp = pd.part[i];
        if (pp.y != ptr) continue;
        if (pp.N != width || (width & 1)) return -1;            // produced in the program, but not linkable: barrier
        if (!pp.y_ll) {
          void* buf = nullptr;
          if (cudaMalloc(&buf, (size_t)m * pp.N * 4) != cudaSuccess) return -2;
          cudaMemset(buf, 0, (size_t)m * pp.N * 4);
          p->side_tables.push_back(buf);
          pp.y_ll = static_cast<uint2*>(buf);
        }
        if (s < pp.ll_consumer) pp.ll_consumer = s;
        out_ll = pp.y_ll;
        out_src = ps;
        return 1;
      }
    }
    return 0;                                                   // not produced by this program: external input
  };
  if (ll_env) {
    for (int s = 1; s < nstages; ++s) {
      DecStage& d = p->h_stages[s];
      const int rx = link(d.x, d.K, s, d.x_ll, d.x_src);
      int rr = 0;
      if (d.residual) rr = link(d.residual, d.part[0].N, s, d.res_ll, d.res_src);
      if (rx == -2 || rr == -2) {
        for (void* b : p->side_tables) cudaFree(b);
        delete p;
        return (int)cudaErrorMemoryAllocation;
      }
      if (rx < 0 || rr < 0) d.force_barrier = 1;
    }
    for (int s = 0; s + 1 < nstages; ++s) {
      const DecStage& nx = p->h_stages[s + 1];
      p->h_stages[s].nx_ll = (nx.x_ll != nullptr && !nx.force_barrier) ? 1 : 0;
      p->h_stages[s].nx_src = nx.x_src;
    }
  }
  // the decode side tables (see dec_build_side_kernel): one per projection, owned by the program
  for (int s = 0; s < nstages; ++s) {
    DecStage& d = p->h_stages[s];
    for (int i = 0; i < d.nparts; ++i) {
      DecPart& dp = d.part[i];
      void* buf = nullptr;
      cudaError_t e = cudaMalloc(&buf, (size_t)(dp.N / 4) * (size_t)dp.side_q);
      if (e != cudaSuccess) {
        for (void* b : p->side_tables) cudaFree(b);
        delete p;
        return (int)e;
      }
      p->side_tables.push_back(buf);
      dec_build_side_kernel<<<dp.N / 4, 128>>>(dp.scales, dp.szeros, dp.ow, static_cast<uint8_t*>(buf), dp.N, d.r, d.g128, d.nsteps, dp.side_q);
      dp.side = static_cast<const uint8_t*>(buf);
    }
  }
  {
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
      for (void* b : p->side_tables) cudaFree(b);
      delete p;
      return (int)e;
    }
  }
  cudaGetDevice(&p->device);
  if (cudaDeviceGetAttribute(&p->nsm, cudaDevAttrMultiProcessorCount, p->device) != cudaSuccess || p->nsm <= 0) p->nsm = 148;
  cudaError_t e = cudaMalloc(&p->d_stages, sizeof(DecStage) * (size_t)nstages);
  if (e == cudaSuccess) e = cudaMalloc(&p->d_sync, 256);
  if (e == cudaSuccess) e = cudaMemcpy(p->d_stages, p->h_stages.data(), sizeof(DecStage) * (size_t)nstages, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemset(p->d_sync, 0, 256);
  if (e != cudaSuccess) {
    if (p->d_stages) cudaFree(p->d_stages);
    if (p->d_sync) cudaFree(p->d_sync);
    for (void* b : p->side_tables) cudaFree(b);
    delete p;
    return (int)e;
  }
  *out = reinterpret_cast<qeft_decode_program_t*>(p);
  return QEFT_OK;
}

extern "C" int qeft_decode_program_destroy(qeft_decode_program_t* prog) {
  if (!prog) return QEFT_E_NULL;
  DecProgram* p = reinterpret_cast<DecProgram*>(prog);
  cudaFree(p->d_stages);
  cudaFree(p->d_sync);
  if (p->d_stamps) cudaFree(p->d_stamps);
  for (void* b : p->side_tables) cudaFree(b);
  delete p;
  return QEFT_OK;
}

// debug: host_out[nstages][4 CTAs][4] globaltimer stamps of the last run (QEFT_DECODE_STAMPS=1)
extern "C" __attribute__((visibility("default"))) int qeft_decode_debug_stamps(qeft_decode_program_t* prog, unsigned long long* host_out) {
  if (!prog || !host_out) return QEFT_E_NULL;
  DecProgram* p = reinterpret_cast<DecProgram*>(prog);
  if (!p->d_stamps) return QEFT_E_UNSUPPORTED;
  cudaError_t e = cudaMemcpy(host_out, p->d_stamps, (p->h_stages.size() * 32 + 24) * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  return e == cudaSuccess ? QEFT_OK : (int)e;
}

extern "C" int qeft_decode_program_num_stages(const qeft_decode_program_t* prog) {
  return prog ? (int)reinterpret_cast<const DecProgram*>(prog)->h_stages.size() : QEFT_E_NULL;
}

extern "C" int qeft_decode_program_run(qeft_decode_program_t* prog, int stage_begin, int stage_end, unsigned flags,
                                       qeft_stream_t stream) {
  (void)flags;
  if (!prog) return QEFT_E_NULL;
  DecProgram* p = reinterpret_cast<DecProgram*>(prog);
  const int n = (int)p->h_stages.size();
  if (stage_begin < 0 || stage_end > n || stage_begin >= stage_end) return QEFT_E_SHAPE;
  static const int grid_env = dec_env_int("QEFT_DECODE_GRID", 0);
  static const int slots_env = dec_env_int("QEFT_DECODE_SLOTS", 0);
  const int grid = grid_env > 0 ? grid_env : p->nsm;
  const int m = p->m;
  // shared memory: the largest stage of the range sizes the x buffers and the partial-sum slices
  int max_steps = 0, max_r = 0, max_tiles = 1;
  for (int s = stage_begin; s < stage_end; ++s) {
    const DecStage& d = p->h_stages[s];
    max_steps = d.nsteps > max_steps ? d.nsteps : max_steps;
    max_r = d.r > max_r ? d.r : max_r;
    const int al = d.epilogue == QEFT_EPI_SWIGLU ? 2 : 1;
    const int nq_max = al * cdiv(d.total_q / al, grid);
    const int T = cdiv(nq_max, 4);
    max_tiles = T > max_tiles ? T : max_tiles;
  }
  DecLayout L;
  L.slot_s = kSlotW;
  L.sarea = kSideSteps + 8 * max_r + 16;      // + 16: the four qweight rows' scale words fall into different banks
  L.slot = (int)(((size_t)(kSlotW + 4 * L.sarea) + 127) & ~(size_t)127);
  const size_t xdig = (size_t)max_steps * 128 * 4 * m;
  const size_t xsum = (size_t)max_steps * 16 + 16;
  const size_t xo = (size_t)m * max_r * 2 + 16;
  const size_t part = (size_t)max_tiles * kDWarps * m * 16 * sizeof(float);
  const size_t misc = 3072;
  const size_t fixed = ((xdig + 127) & ~(size_t)127) + ((xsum + 127) & ~(size_t)127) + ((xo + 127) & ~(size_t)127) +
                       ((part + 127) & ~(size_t)127) + misc;
  if (fixed + 2 * (size_t)L.slot > kDSmemMax) return QEFT_E_UNSUPPORTED;
  int nslots = (int)((kDSmemMax - fixed) / (size_t)L.slot);
  if (nslots > 6) nslots = 6;
  if (slots_env > 0 && nslots > slots_env) nslots = slots_env;
  L.nslots = nslots;
  size_t off = (size_t)nslots * (size_t)L.slot;
  L.xdig = (int)off; off += (xdig + 127) & ~(size_t)127;
  L.xdig_bytes = (int)((xdig + 127) & ~(size_t)127);
  L.xsum = (int)off; off += (xsum + 127) & ~(size_t)127;
  L.xo = (int)off; off += (xo + 127) & ~(size_t)127;
  L.part = (int)off; off += (part + 127) & ~(size_t)127;
  L.part_bytes = (int)part;
  L.misc = (int)off; off += misc;
  static const int debug_env = dec_env_int("QEFT_DECODE_DEBUG", 0);
  L.debug = debug_env;
  static const int stamps_env = dec_env_int("QEFT_DECODE_STAMPS", 0);
  if (stamps_env && !p->d_stamps) {
    const size_t bytes = ((size_t)n * 32 + 24) * sizeof(unsigned long long);
    if (cudaMalloc(&p->d_stamps, bytes) == cudaSuccess) cudaMemset(p->d_stamps, 0, bytes);
    else p->d_stamps = nullptr;
  }
  L.stamps = p->d_stamps;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
// which stages of the range wait at a barrier, and whether any reads data-flow words written in this launch
  int nbar_total = 0, uses_ll = 0;
  for (int s = stage_begin + 1; s < stage_end; ++s) {
    const DecStage& d = p->h_stages[s];
    const bool ll = d.x_ll != nullptr && d.x_src >= stage_begin && !d.force_barrier;
    nbar_total += ll ? 0 : 1;
    uses_ll |= ll ? 1 : 0;
    if (d.res_ll != nullptr && d.res_src >= stage_begin) uses_ll = 1;
  }
  if (uses_ll)
    return m == 1 ? dec_launch<1, true>(p, stage_begin, stage_end, L, off, grid, st, nbar_total, uses_ll)
                  : dec_launch<2, true>(p, stage_begin, stage_end, L, off, grid, st, nbar_total, uses_ll);
  return m == 1 ? dec_launch<1, false>(p, stage_begin, stage_end, L, off, grid, st, nbar_total, uses_ll)
                : dec_launch<2, false>(p, stage_begin, stage_end, L, off, grid, st, nbar_total, uses_ll);
}
