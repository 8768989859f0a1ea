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

#include <type_traits>
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
  const void* y_ll;       // non-null: a later stage of the program reads y by data-flow (polls it: see the kernel's header)
  int ll_consumer;        // first stage that does
  int pad1;
  const uint8_t* side;    // decode side table (built at program creation), per qweight row: [steps][scales of its 4 rows |
                          // scaled zeros of its 4 rows] then the fp16 outlier columns in MMA-fragment order
  int side_q;             // bytes per qweight row of `side`
  int N;
  int q_begin;            // first qweight row of this part in the stage-wide numbering (plain stages)
  // column-sharded programs (qeft_decode_program_shard): y is this rank's slice [N] of a gathered [nranks x N] row that
  // every rank holds; the rows are stored into EVERY rank's copy (peer-mapped pointers over NVLink, y_peer[p] = this
  // rank's slice in rank p's buffer); y_full = the local copy's base (what a later stage reads as its x)
  int nranks;
  __half* y_peer[QEFT_MAX_RANKS];
  const void* y_full;
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
  const void* x_ll;           // non-null: x is the y of an earlier stage of the program and may be read by data-flow
  const void* res_ll;         // the same for the residual
  int x_src, res_src;         // producing stages (data-flow applies only when those run in THIS launch)
  int force_barrier;          // a dependency on an earlier stage that cannot be read by data-flow
  int nx_ll, nx_src;          // x_ll != null && !force_barrier / x_src of the NEXT stage (whether it waits at a barrier)
  float norm_eps;
  int nparts;
  int K, r;
  int g128;                   // G / 128, 0 for per-channel scales
  int nsteps, nchunks, nou;   // int4 steps, live 32-column chunks, outlier units (r / 32)
  int total_q;                // qweight rows of all parts
  int epilogue;               // QEFT_EPI_*
};

// Data-flow outputs are re-armed at the start of every launch: one entry per projection whose y a later stage polls.
struct DecReset { unsigned short* y; int n; int prod, cons; int pad; };   // [m x N] halves, producing / first consuming stage

// Column-sharded programs: one counter per rank (peer-mapped), advanced by every rank at the launch's two rank barriers.
struct DecRanks { int nranks, rank; unsigned* bar_peer[QEFT_MAX_RANKS]; };

struct DecLayout {            // shared-memory carve-up (bytes from the start of dynamic shared memory)
  int nslots, slot;           // ring at offset 0: nslots slots of `slot` bytes: [4 x kQArea packed words][4 x sarea side bytes]
  int slot_s, sarea;
  int xdig, xdig_bytes, xsum, xo, part, part_bytes, misc;
  int poll_ns;                // back-off between two polls of a data-flow input (QEFT_DECODE_POLL_NS)
  int debug;                  // QEFT_DECODE_DEBUG bit mask (bisecting switches; results are WRONG when set)
  unsigned long long* stamps; // debug (QEFT_DECODE_STAMPS): [stage][4 CTAs][8] globaltimer values, or null
  unsigned long long* stamps_all;   // debug: [stage][160 CTAs][4]: stage entered, x staged, blocks consumed, rows stored
};

template <bool DBG>
__device__ __forceinline__ void dec_stamp(const DecLayout& L, int s, int i) {
  if (DBG && L.stamps_all && threadIdx.x == 0 && i < 4 && blockIdx.x < 160) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    L.stamps_all[((size_t)s * 160 + blockIdx.x) * 4 + i] = t;
  }
  if (DBG && L.stamps && threadIdx.x == 0) {
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
// ---- data-flow by sentinel ---------------------------------------------------------------------------------------
// A y that a later stage of the same launch reads is its own "ready" flag: the launch first overwrites it with the fp16
// bit pattern 0xFFFF (a NaN no result ever has: results that are NaN are stored as 0x7E00), the producing CTAs store
// plain results, and a consumer re-reads any 16-byte piece in which an 0xFFFF is left.  No fence, no counter, no second
// copy of the data: a value is usable the moment it is visible, and every fp16 element is its own flag, so no ordering
// between different elements is needed (relaxed gpu-scope accesses).
constexpr unsigned kDfEmpty = 0xFFFFu;
__device__ __forceinline__ bool d_has_empty(const uint4& v) {
  return (__vcmpeq2(v.x, 0xFFFFFFFFu) | __vcmpeq2(v.y, 0xFFFFFFFFu) | __vcmpeq2(v.z, 0xFFFFFFFFu) | __vcmpeq2(v.w, 0xFFFFFFFFu)) != 0u;
}
__device__ __forceinline__ uint4 d_ldrelaxed128(const void* p) {
  uint4 r;
#ifdef QEFT_DEC_POLL_CG
  asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
#else
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
#endif
  return r;
}
__device__ __forceinline__ unsigned short d_ldrelaxed16(const void* p) {
  unsigned short r;
  asm volatile("ld.relaxed.gpu.global.u16 %0, [%1];" : "=h"(r) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void d_strelaxed16(void* p, unsigned short v) {
#ifdef QEFT_DEC_POLL_CG
  *reinterpret_cast<volatile unsigned short*>(p) = v;
#else
  asm volatile("st.relaxed.gpu.global.u16 [%0], %1;" ::"l"(p), "h"(v) : "memory");
#endif
}
// 16 bytes of x: polled until every element has arrived when the producer runs in this launch
__device__ __forceinline__ uint4 d_ldx128(const void* p, bool poll, unsigned ns) {
  if (!poll) return d_ldcg128(p);
  uint4 v = d_ldrelaxed128(p);
  while (d_has_empty(v)) { __nanosleep(ns); v = d_ldrelaxed128(p); }
  return v;
}
__device__ __forceinline__ unsigned short d_ldx16(const void* p, bool poll, unsigned ns) {
  if (!poll) return d_ldcg16(p);
  unsigned short v = d_ldrelaxed16(p);
  while (v == kDfEmpty) { __nanosleep(ns); v = d_ldrelaxed16(p); }
  return v;
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
// DBG: the instrumented build of the kernel (QEFT_DECODE_STAMPS / QEFT_DECODE_DEBUG): in-kernel timers and the bisecting
// switches; compiled out of the shipping instances (they cost ~10 % of the consumers' issue slots when merely predicated).
template <int M, bool LL, bool DBG>
__global__ void __maxnreg__(QEFT_DEC_MAXREG)
decode_w4_kernel(const DecStage* __restrict__ stages, int s_begin, int s_end, unsigned* sync, const DecLayout L,
                 int nbar_total, int uses_ll, const DecReset* __restrict__ resets, int nresets, const DecRanks RK) {
  extern __shared__ __align__(128) uint8_t dsm[];
  constexpr int NCOLS = 4 * M;
  constexpr uint32_t XSTEP = 128u * NCOLS;                   // digit bytes per 128-column step: [2 nibble halves][NCOLS][4 t][16 B]
  constexpr uint32_t XHALF = 64u * NCOLS;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int cta = blockIdx.x, ncta = gridDim.x;
  const int NS = L.nslots;
  // (Column-sharded programs poll at gpu scope too: a peer's NVLink stores to THIS GPU's memory are performed in this
  // GPU's L2, the point of coherence of its HBM, and every element validates itself.  System-scope polling loads were
  // measured 6 % slower per 70B token at 2 ranks.)
  const unsigned pns = (unsigned)L.poll_ns;

  const uint32_t ring = d_smem_u32(dsm);
  const uint32_t xdig = d_smem_u32(dsm + L.xdig);
  const uint32_t xsum = d_smem_u32(dsm + L.xsum);            // [nsteps][2] fp32 group sums of x
  const uint32_t xo = d_smem_u32(dsm + L.xo);                // [M][r] fp16 outlier activations
  float* part = reinterpret_cast<float*>(dsm + L.part);      // [tile][warp][M][16 rows]
  float* red = reinterpret_cast<float*>(dsm + L.misc);       // [16 warps][4] staging reductions
  float* coef = red + kDWarps * 4;                           // [8] flush weights of a quad's accumulator columns
  const uint32_t bars = d_smem_u32(dsm + L.misc + 512);      // [nslots] "slot filled" mbarriers
  const uint32_t ebars = bars + 64;                          // [nslots] "slot consumed" mbarriers (16 warp arrivals)
  DecStage* pcache = reinterpret_cast<DecStage*>(dsm + L.misc + 1024);    // descriptor of the producer's stage
  DecStage* ccache2 = reinterpret_cast<DecStage*>(dsm + L.misc + 2048);   // descriptors of the stage being consumed / the next one
  static_assert(sizeof(DecStage) <= 1024 && sizeof(DecStage) % 4 == 0, "descriptor cache slots are 1024 bytes");
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
  int nbar = 0;                                              // stage barriers passed so far
  __syncthreads();

  // (digit bytes of steps / columns a stage does not write read as zero)
  for (int i = tid; i < L.xdig_bytes / 16; i += kDBlock) d_sts128(xdig + (uint32_t)i * 16, 0u, 0u, 0u, 0u);
  __syncthreads();

  // debug counters (QEFT_DECODE_STAMPS): warp 0 and warp 15 of CTA 0 and the producer warp, clock64 cycles
  long long dbg_wait = 0, dbg_math = 0, dbg_issue = 0, dbg_fill = 0, dbg_prev = 0;
  int dbg_nissue = 0, dbg_nwaited = 0, dbg_nblocks = 0;
  const bool dbg = DBG && L.stamps != nullptr && cta == 0 && (warp == 0 || warp >= kDWarps - 1);
  const int dbgsw = DBG ? L.debug : 0;                       // bisecting switches (results are wrong when set)

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
          if (lane == 0) d_mbar_expect_tx(bar, (uint32_t)__popc(own4) * (((dbgsw & 4) ? 0u : nbytes) + ((dbgsw & 2) ? 0u : sbytes)));
          __syncwarp();
          // rows this CTA does not own are NOT copied: their slot bytes are stale, their (independent) MMA rows are never stored
          if (lane < 4 && own && !(dbgsw & 4)) d_bulk_g2s(sbase + (uint32_t)(q * kQArea), wsrc + off, nbytes, bar);
          if (lane < 4 && own && !(dbgsw & 2))
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
  // tile-end weights of this lane's two accumulator columns (four digit columns of a batch row weigh 256^i) and its
  // 8 bytes in the warp's partial-sum slice [tile][warp][M][16 rows]
  const float dc0 = (M == 1 ? (t == 0 ? 1.f : (t == 1 ? 65536.f : (t == 2 ? 1.f : 0.f))) : ((t & 1) ? 65536.f : 1.f));
  const float dc1 = (M == 1 ? (t == 0 ? 256.f : (t == 1 ? 16777216.f : 0.f)) : ((t & 1) ? 16777216.f : 256.f));
  const uint32_t part_w = d_smem_u32(dsm + L.part) + (uint32_t)(warp * M * 64 + g * 8);

  // grid barrier of the consumer warps: one release per CTA on a counter that only grows, one polling thread per CTA
  // (one signal per warp was measured slower -- 16 x 148 atomics on one address per stage; polling with relaxed loads and
  // one acquire fence at the end: +0.4 us per stage)
  auto barrier_arrive = [&]() {
    d_consumer_sync();
    if (tid == 0) asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(sync), "r"(1u) : "memory");
  };
  auto barrier_wait = [&]() {
    ++nbar;
    if (tid == 0) {
      const unsigned want = base + (unsigned)nbar * (unsigned)ncta;
      unsigned got;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(got) : "l"(sync) : "memory");
      } while ((int)(got - want) < 0);
      if (cta == 0 && nbar == 1)              // every CTA has read `base`: publish the next launch's base
        *reinterpret_cast<volatile unsigned*>(sync + 1) = base + (unsigned)nbar_total * (unsigned)ncta;
    }
    d_consumer_sync();
  };
  // Column-sharded programs: barrier over all ranks.  Every CTA orders its earlier stores (local resets, peer stores) with
  // a system-scope fence and arrives at the local grid barrier; CTA 0 then adds one to every rank's counter and waits until
  // its own has been advanced by every rank; a second grid barrier holds the other CTAs until then.
  unsigned rk_done = 0;                                      // rank barriers this program has passed before this launch
  int rk_now = 0;
  if (RK.nranks > 1 && cta == 0 && tid == 0) rk_done = *reinterpret_cast<volatile unsigned*>(sync + 3);
  auto rank_barrier = [&]() {
    d_consumer_sync();
    if (tid == 0) {
      asm volatile("fence.acq_rel.sys;" ::: "memory");
      asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(sync), "r"(1u) : "memory");
    }
    barrier_wait();
    if (cta == 0 && tid == 0) {
      ++rk_now;
      asm volatile("fence.acq_rel.sys;" ::: "memory");
      for (int pr = 0; pr < RK.nranks; ++pr)
        asm volatile("red.relaxed.sys.global.add.u32 [%0], %1;" ::"l"(RK.bar_peer[pr]), "r"(1u) : "memory");
      const unsigned want = (rk_done + (unsigned)rk_now) * (unsigned)RK.nranks;
      unsigned got;
      do {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(got) : "l"(RK.bar_peer[RK.rank]) : "memory");
      } while ((int)(got - want) < 0);
    }
    barrier_arrive();
    barrier_wait();
  };
  if (LL && uses_ll) {
    // re-arm the data-flow outputs of this launch (one list entry per thread, this CTA's share of its elements), then
    // one grid barrier: nobody polls a buffer that still holds the previous run's results
    for (int i = tid; i < nresets; i += kDThreads) {
      const DecReset e = resets[i];
      if (e.prod >= s_begin && e.prod < s_end && e.cons < s_end) {
        const int i0 = (int)(((long long)e.n * cta) / ncta), i1 = (int)(((long long)e.n * (cta + 1)) / ncta);
        for (int k = i0; k < i1; ++k) d_strelaxed16(e.y + k, (unsigned short)kDfEmpty);
      }
    }
    if (RK.nranks > 1) {
      rank_barrier();         // no rank stores into a peer's buffer before that peer has re-armed it
    } else {
      barrier_arrive();
      barrier_wait();
    }
  }

#pragma unroll 1
  for (int s = s_begin; s < s_end; ++s) {
    const DecStage* S = reinterpret_cast<const DecStage*>(reinterpret_cast<const uint8_t*>(ccache2) + ((s - s_begin) & 1) * 1024);

    // x produced by an earlier stage of this launch and readable by data-flow: the stage polls the elements it reads,
    // no barrier.  Every other stage after the first waits until all CTAs have stored their rows of the previous one.
    // (The descriptor copy of stage s was written before the previous stage's post-consume barrier.)
    const bool ll_x = LL && S->x_ll != nullptr && S->x_src >= s_begin && !S->force_barrier;
    const int K = S->K, r = S->r, ns = S->nsteps;
    // o_proj's gather indices of this thread's first staging item (and of its outlier activations) do not depend on x:
    // they are loaded BEFORE the barrier, like the weights (one L2 round trip less on the boundary's critical path)
    int4 gi0 = make_int4(0, 0, 0, 0), gi1 = gi0, go0 = gi0, go1 = gi0;
    if (S->gather != nullptr) {
      const int sb = tid >> 4, b = sb / ns, st = sb - b * ns, k0 = st * 128 + (tid & 15) * 8;
      if (tid < M * ns * 16 && k0 < S->nchunks * 32) {
        gi0 = __ldg(reinterpret_cast<const int4*>(S->gather + k0));
        gi1 = __ldg(reinterpret_cast<const int4*>(S->gather + k0 + 4));
      }
      const int xt = kDThreads - 1 - tid;
      if (xt < M * (r >> 3)) {
        const int jj = xt % (r >> 3);
        go0 = __ldg(reinterpret_cast<const int4*>(S->gather + K - r + 8 * jj));
        go1 = __ldg(reinterpret_cast<const int4*>(S->gather + K - r + 8 * jj + 4));
      }
    }
    if (s > s_begin && !ll_x) barrier_wait();
    const DecTiles R = dec_tiles(S, cta, ncta);

    dec_stamp<DBG>(L, s, 0);
    // ---- x: block fixed point per 128-column step, signed-byte digits ------------------------------------------------
    //   x_k ~= X_k 2^(e-22),  X_k = rint(x_k 2^(22-e)) a 23-bit integer,  2^e > max |x| of the step
    // (exact for every element within 12 binades of the step's maximum; fp16 has 11 significant bits); what is stored are
    // four base-256 digits d_i in [-128, 127] of 16 X_k (low-nibble columns) or X_k - 16 X_partner (high-nibble columns).
    // One exponent per
    // STEP keeps the staging free of a CTA-wide reduction (the stage boundary's critical path: measured 0.8 us for the
    // first version's single exponent per row); the price is one multiply per row pair and step in the main loop.
    {
      const __half* xg = S->x;
      // o_proj's gather (qlinear.py:275): copy x to shared memory first (coalesced, one L2 round trip; the buffer aliases
      // the partial-sum slices, idle between two stages), then gather from there instead of 16 scattered 2-byte L2 loads
#ifdef QEFT_DEC_LEAN_EXPERIMENT
      const bool xraw_ok = false;
#else
      const bool xraw_ok = S->gather != nullptr && (size_t)M * (size_t)K * 2 <= (size_t)L.part_bytes;
#endif
      const uint32_t xraw = d_smem_u32(dsm + L.part);
      if (xraw_ok) {
        // (without a barrier in front of this stage, warps may still be adding the previous stage's slices)
        if (ll_x) d_consumer_sync();
        for (int i = tid; i < M * (K >> 3); i += kDThreads) {
          const uint4 v = d_ldx128(xg + (size_t)i * 8, ll_x, pns);
          d_sts128(xraw + (uint32_t)i * 16, v.x, v.y, v.z, v.w);
        }
        d_consumer_sync();
      }
      // data-flow input (ll_x): every load of x is repeated until the elements it covers have arrived (d_ldx128 / d_ldx16)
      auto ldx16 = [&](const __half* row, int b, int col) -> unsigned short {
        if (xraw_ok) {
          unsigned short r16;
          asm volatile("ld.shared.u16 %0, [%1];" : "=h"(r16) : "r"(xraw + (uint32_t)((b * K + col) * 2)) : "memory");
          return r16;
        }
        return d_ldx16(row + col, ll_x, pns);
      };
#ifdef QEFT_DEC_LEAN_EXPERIMENT   // (code-size experiment: no gather, no norm -- results are wrong for stages that have them)
      const int32_t* const gat = nullptr;
      const __half* const nw = nullptr;
#else
      const int32_t* gat = S->gather;
      const __half* nw = S->norm_w;
#endif
      const int live_k = S->nchunks * 32;
      const int nitems = M * ns * 16;
      const int npass = (nitems + kDThreads - 1) / kDThreads;
      // item = (batch row b, step, sub): the 8 columns k0 .. k0+7, k0 = 128 step + 8 sub (one 16-byte load, lanes contiguous);
      // sub = 4 tt + 2 half + hs: 32-column chunk tt, first / second 16 columns of the chunk, low / high nibble position.
      // The 16 items of a step sit in 16 adjacent lanes, which agree on the step's sum and maximum by shuffles.
      auto load_item = [&](int it, uint4& v0, int& b, int& st, bool& live) {
        const int sb = it >> 4;
        b = sb / ns;
        st = sb - b * ns;
        const int k0 = st * 128 + (it & 15) * 8;
        live = it < nitems && k0 < live_k;
        v0 = make_uint4(0u, 0u, 0u, 0u);
        if (live) {
          const __half* xr = xg + (size_t)b * K;
          if (gat) {
            int4 i0 = gi0, i1 = gi1;                           // (the first pass's indices were loaded before the barrier)
            if (it != tid) {
              i0 = __ldg(reinterpret_cast<const int4*>(gat + k0));
              i1 = __ldg(reinterpret_cast<const int4*>(gat + k0 + 4));
            }
            auto pk = [&](int a, int c) { return (uint32_t)ldx16(xr, b, a) | ((uint32_t)ldx16(xr, b, c) << 16); };
            v0 = make_uint4(pk(i0.x, i0.y), pk(i0.z, i0.w), pk(i1.x, i1.y), pk(i1.z, i1.w));
          } else {
            v0 = d_ldx128(xr + k0, ll_x, pns);
          }
        }
      };
      // the loads of the first two passes and of the outlier activations are in flight together
      uint4 k0a, k1a;
      int kb0, kb1, ks0, ks1;
      bool kl0, kl1;
      load_item(tid, k0a, kb0, ks0, kl0);
      load_item(kDThreads + tid, k1a, kb1, ks1, kl1);
      const int nxo = M * (r >> 3);
      uint4 xo_v = make_uint4(0u, 0u, 0u, 0u);
      const int xo_tid = kDThreads - 1 - tid;              // the threads the digit items use least
      if (xo_tid < nxo) {
        const int b = xo_tid / (r >> 3), jj = xo_tid - b * (r >> 3);
        const __half* xr = xg + (size_t)b * K;
        if (gat) {
          const int4 a = go0, c = go1;
          auto pk = [&](int i0, int i1) { return (uint32_t)ldx16(xr, b, i0) | ((uint32_t)ldx16(xr, b, i1) << 16); };
          xo_v = make_uint4(pk(a.x, a.y), pk(a.z, a.w), pk(c.x, c.y), pk(c.z, c.w));
        } else {
          xo_v = d_ldx128(xr + K - r + 8 * jj, ll_x, pns);
        }
      }
      float rs0 = 1.f, rs1 = 1.f;
      if (nw) {
        // RMSNorm on the way in (HF LlamaRMSNorm / kernel/layernorm/layernorm.cu:25-51): the row's sum of squares needs the
        // whole row: one CTA-wide reduction (only for stages with a norm)
        float ss0 = 0.f, ss1 = 0.f;
        auto sumsq = [&](const uint4& v0, int b) {
          const uint32_t w[4] = {v0.x, v0.y, v0.z, v0.w};
          float ss = 0.f;
#pragma unroll
          for (int j = 0; j < 4; ++j) { const float2 f = half2_bits_to_float2(w[j]); ss = fmaf(f.x, f.x, fmaf(f.y, f.y, ss)); }
          if (b == 0) ss0 += ss; else ss1 += ss;
        };
        if (kl0) sumsq(k0a, kb0);
        if (kl1) sumsq(k1a, kb1);
        for (int q = 2; q < npass; ++q) {
          uint4 v0; int b, st; bool live;
          load_item(q * kDThreads + tid, v0, b, st, live);
          if (live) sumsq(v0, b);
        }
        if (xo_tid < nxo) sumsq(xo_v, xo_tid / (r >> 3));
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
      dec_stamp<DBG>(L, s, 4);
      // the reference's two roundings: (x * rs).to(fp16), then * weight in fp16
      auto normed = [&](float v, int col, float rs) { return __half2float(__hmul(nw[col], __float2half_rn(v * rs))); };
      // one pass = one item per thread.  (Passes in pairs -- both loads first, then two interleaved conversion chains -- were
      // measured: 1 % faster on one GPU, 17 % SLOWER per sharded 70B token at 2 ranks, 7.81 ms against 6.70.)
      auto digit_pass = [&](int it, uint4 v0, int b, int st, bool live) {
        const bool valid = it < nitems;
        const int sub = it & 15, tt = sub >> 2, half = (sub >> 1) & 1, hs = sub & 1;
        const uint32_t w[4] = {v0.x, v0.y, v0.z, v0.w};
        float2 f[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) f[j] = half2_bits_to_float2(w[j]);
        if (nw && live) {
          const int k0 = st * 128 + sub * 8;
          const float rs = b ? rs1 : rs0;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int k = k0 + 2 * j;
            const int c0 = gat ? gat[k] : k, c1 = gat ? gat[k + 1] : k + 1;
            f[j] = make_float2(normed(f[j].x, c0, rs), normed(f[j].y, c1, rs));
          }
        }
        float sum = 0.f, mx = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          sum += f[j].x + f[j].y;
          mx = fmaxf(mx, fmaxf(fabsf(f[j].x), fabsf(f[j].y)));
        }
#pragma unroll
        for (int o = 8; o >= 1; o >>= 1) {
          sum += __shfl_xor_sync(0xffffffffu, sum, o);
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        // 2^e > mx;  X = rint(x 2^(22-e)) by the magic-number add (|X| < 2^22): bits(fma(x, sc, 1.5 2^23)) = 0x4B400000 + X.
        // A byte of a packed word holds TWO weights: lo (this byte position of the nibble-half hs = 0 item) and hi (the same
        // position of the hs = 1 item, the lane next door), byte = lo + 16 hi.  The consumers multiply the RAW byte with the
        // digits of 16 X_lo and the masked byte 16 hi with the digits of X_hi - 16 X_lo:
        //   sum byte * 16 X_lo + sum 16 hi * (X_hi - 16 X_lo) = 16 (sum lo X_lo + sum hi X_hi),
        // exactly, in ONE s32 accumulator chain and with one AND mask per word instead of two.  Both operands are below
        // 2^27 in magnitude: four signed-byte digits V = sum d_i 256^i, the bytes of (V + 0x80808080) ^ 0x80808080.
        // (all lanes take part: the exchanges with the neighbouring items are shuffles)
        const int e = mx > 0.f ? (int)((__float_as_uint(mx) >> 23) & 0xff) - 126 : -100;
        const float sc = e > -100 ? __uint_as_float((uint32_t)(127 + 22 - e) << 23) : 0.f;
        uint32_t D[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int xa = __float_as_int(fmaf(f[j].x, sc, 12582912.f)) - 0x4B400000;
          const int xb = __float_as_int(fmaf(f[j].y, sc, 12582912.f)) - 0x4B400000;
          const int pa = __shfl_xor_sync(0xffffffffu, xa, 1), pb = __shfl_xor_sync(0xffffffffu, xb, 1);
          const int va = hs ? xa - 16 * pa : 16 * xa, vb = hs ? xb - 16 * pb : 16 * xb;
          D[2 * j] = ((uint32_t)va + 0x80808080u) ^ 0x80808080u;
          D[2 * j + 1] = ((uint32_t)vb + 0x80808080u) ^ 0x80808080u;
        }
        // Word c (= B-fragment register of lane t = c) of digit column d holds the bytes {first[2c], second[2c], first[2c+1],
        // second[2c+1]}, first / second = the chunk's columns 0..7 / 16..23 (+ 8 hs): the `half` = 0 item of a pair has the
        // first[] values, its neighbour two lanes on the second[] ones.  They swap four values each; the half = 0 lane then
        // writes words c = 0, 1 and the half = 1 lane words c = 2, 3.
        uint32_t Fv[4], Sv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t send = half ? D[j] : D[4 + j];
          const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 2);
          Fv[j] = half ? recv : D[j];
          Sv[j] = half ? D[4 + j] : recv;
        }
        if (valid) {
          // lane t's 16-byte row of (hs, column) holds its four chunks' words, index tt
          const uint32_t dst = xdig + (uint32_t)st * XSTEP + (uint32_t)hs * XHALF + (uint32_t)(4 * b) * 64 + (uint32_t)tt * 4 + (uint32_t)(half * 32);
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const uint32_t da = Fv[2 * cc], db = Sv[2 * cc], dc = Fv[2 * cc + 1], dd = Sv[2 * cc + 1];
#pragma unroll
            for (int d = 0; d < 4; ++d) {
              const uint32_t sel = 0x0040u + 0x11u * (uint32_t)d;
              const uint32_t ww = d_prmt(d_prmt(da, db, sel), d_prmt(dc, dd, sel), 0x5410u);
              asm volatile("st.shared.u32 [%0], %1;" ::"r"(dst + (uint32_t)(d * 64 + cc * 16)), "r"(ww) : "memory");
            }
          }
          if (sub == 0) {
            // per step and batch row: {sum of x, weight of digit 0 = 2^(e-22) / 16 (both operands carry a factor 16)}
            const float cg = e > -100 ? __uint_as_float((uint32_t)(127 + e - 22 - 4) << 23) : 0.f;
            asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(xsum + (uint32_t)(st * 16 + b * 8)), "f"(sum), "f"(cg) : "memory");
          }
        }
      };
      for (int q = 0; q < npass; ++q) {
        const int it = q * kDThreads + tid;
        uint4 v0; int b, st; bool live;
        if (q == 0) { v0 = k0a; b = kb0; st = ks0; live = kl0; }
        else if (q == 1) { v0 = k1a; b = kb1; st = ks1; live = kl1; }
        else load_item(it, v0, b, st, live);
        digit_pass(it, v0, b, st, live);
      }
      dec_stamp<DBG>(L, s, 6);
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
      d_cp16p(d_smem_u32(reinterpret_cast<uint8_t*>(ccache2) + ((s + 1 - s_begin) & 1) * 1024) + (uint32_t)(tid - 32) * 16,
              reinterpret_cast<const uint8_t*>(stages + s + 1) + (size_t)(tid - 32) * 16);
    dec_stamp<DBG>(L, s, 1);
    // ---- the tile-blocks of the stage -------------------------------------------------------------------------
    // The consumers are bound by instruction issue (4 warps per scheduler, every warp the same ~150-instruction chain per
    // block), so the loop is specialised at compile time: BC = single-k-block stage (K <= 4096 + r) whose B operands stay in
    // registers for the whole stage; the instrumented paths exist only in the DBG instances.
    {
      const int KB = (ns + kKB - 1) / kKB;
      const uint32_t xo_lane = xo + (uint32_t)((gx * r + 2 * t) * 2);
      // B fragments (digit bytes), group sum and digit weight of one step
      auto load_x = [&](int gs, uint4& xe_, uint4& xq_, float2& xs) {
        const uint32_t xc = xdig_lane + (uint32_t)gs * XSTEP;
        d_lds128_if(xe_, xc, has_col);
        d_lds128_if(xq_, xc + XHALF, has_col);
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(xs.x), "=f"(xs.y) : "r"(xsum + (uint32_t)(gs * 16 + (M == 2 ? (t >> 1) * 8 : 0))) : "memory");
      };
      const int nou16 = r >> 4;
      auto run_tiles = [&](auto bc_tag) {
        constexpr bool BC = decltype(bc_tag)::value;
        uint4 xe = make_uint4(0u, 0u, 0u, 0u), xq = xe, xe2 = xe, xq2 = xe;
        float2 xs0 = make_float2(0.f, 0.f), xs1 = xs0;
        if (BC) {
          if (warp < ns) load_x(warp, xe, xq, xs0);
          if (warp + kDWarps < ns) load_x(warp + kDWarps, xe2, xq2, xs1);
        }
        uint32_t sbase = ring + (uint32_t)cslot * (uint32_t)L.slot;
        uint32_t pdst = part_w;                              // this warp's slice of tile 0
#pragma unroll 1
        for (int j = 0; j < R.ntiles; ++j) {
          float acc[4] = {0.f, 0.f, 0.f, 0.f};    // rows 2g, 2g+1 x accumulator columns 2t, 2t+1: sum over groups of scale * P
          float zacc0 = 0.f, zacc1 = 0.f;         // M = 2: rows 2g, 2g+1, sum over groups of scaled zero * X_g, batch row t >> 1
          float yo[4] = {0.f, 0.f, 0.f, 0.f};     // outlier columns: rows 2g, 2g+1 x batch rows 2t, 2t+1
#pragma unroll 1
          for (int kb = 0; kb < KB; ++kb) {
            // wait for the block's bytes (every lane acquires them)
            long long tw0 = 0, tw1 = 0;
            if constexpr (DBG) {
              if (dbg) {
                tw0 = clock64();
                if (dbg_prev) dbg_fill += tw0 - dbg_prev;      // (block-to-block period inside a stage)
                dbg_prev = tw0;
                if (!d_mbar_test(bars + 8 * cslot, cpar)) ++dbg_nwaited;
              }
            }
            d_mbar_wait(bars + 8 * cslot, cpar);
            if constexpr (DBG) {
              if (dbg) { tw1 = clock64(); dbg_wait += tw1 - tw0; ++dbg_nblocks; }
            }
            const int nsb = BC ? ns : (ns - kb * kKB < kKB ? ns - kb * kKB : kKB);
            // one 128-column int4 step of 16 rows: A fragments by ldmatrix from the copied bytes; the raw bytes (lo + 16 hi)
            // against the digits of 16 X_lo, the masked bytes (16 hi) against the digits of X_hi - 16 X_lo (see the staging):
            // one AND per word, 4 IMMA chained into one exact s32 accumulator = 16 sum(q X) per digit column.
            // A warp's two steps of the block (warp, warp + 16) are loaded together and then computed: twice the loads
            // in flight per warp.
            auto load_step = [&](int st, uint32_t (&a0)[4], uint32_t (&a1)[4], uint4& xe_, uint4& xq_, uint32_t& sw, uint32_t& zw, float2& xs) {
              d_ldmatrix_x4(a0, sbase + laneA + (uint32_t)(st * 256));
              d_ldmatrix_x4(a1, sbase + laneA + (uint32_t)(st * 256 + 128));
              if (!BC) load_x(kb * kKB + st, xe_, xq_, xs);
              sw = d_lds32(sbase + laneS + (uint32_t)(st * 16));
              zw = M == 2 ? d_lds32(sbase + laneS + (uint32_t)(st * 16 + 8)) : 0u;
            };
            constexpr uint32_t kHiM = 0xf0f0f0f0u;
            // the step's s32 sums (one per accumulator column) -> fp32, times scale x digit weight, into the tile's sums
            auto post_step = [&](const int (&ac)[4], uint32_t sw, uint32_t zw, float2 xs2) {
              const float xs = xs2.x;
              float2 sc = half2_bits_to_float2(sw);
              float f0 = (float)ac[0], f2 = (float)ac[2];
              const float f1 = (float)ac[1], f3 = (float)ac[3];
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
            if (warp < nsb && !(dbgsw & 1)) {
              uint32_t a0[4], a1[4], sw0, zw0;
              int ac[4];
              load_step(warp, a0, a1, xe, xq, sw0, zw0, xs0);
              if (warp + kDWarps < nsb) {
                // two steps: their IMMA chains are issued alternately (each chain is 4 dependent instructions)
                uint32_t b0[4], b1[4], sw1, zw1;
                int bc[4];
                load_step(warp + kDWarps, b0, b1, xe2, xq2, sw1, zw1, xs1);
                d_imma0(ac, a0[0], a0[1], a0[2], a0[3], xe.x, xe.y);
                d_imma0(bc, b0[0], b0[1], b0[2], b0[3], xe2.x, xe2.y);
                d_imma(ac, a0[0] & kHiM, a0[1] & kHiM, a0[2] & kHiM, a0[3] & kHiM, xq.x, xq.y);
                d_imma(bc, b0[0] & kHiM, b0[1] & kHiM, b0[2] & kHiM, b0[3] & kHiM, xq2.x, xq2.y);
                d_imma(ac, a1[0], a1[1], a1[2], a1[3], xe.z, xe.w);
                d_imma(bc, b1[0], b1[1], b1[2], b1[3], xe2.z, xe2.w);
                d_imma(ac, a1[0] & kHiM, a1[1] & kHiM, a1[2] & kHiM, a1[3] & kHiM, xq.z, xq.w);
                d_imma(bc, b1[0] & kHiM, b1[1] & kHiM, b1[2] & kHiM, b1[3] & kHiM, xq2.z, xq2.w);
                post_step(ac, sw0, zw0, xs0);
                post_step(bc, sw1, zw1, xs1);
              } else {
                d_imma0(ac, a0[0], a0[1], a0[2], a0[3], xe.x, xe.y);
                d_imma(ac, a0[0] & kHiM, a0[1] & kHiM, a0[2] & kHiM, a0[3] & kHiM, xq.x, xq.y);
                d_imma(ac, a1[0], a1[1], a1[2], a1[3], xe.z, xe.w);
                d_imma(ac, a1[0] & kHiM, a1[1] & kHiM, a1[2] & kHiM, a1[3] & kHiM, xq.z, xq.w);
                post_step(ac, sw0, zw0, xs0);
              }
            }
            if ((BC || kb == KB - 1) && r > 0 && !(dbgsw & 1)) {
              // 16 fp16 outlier columns of 16 rows per unit: one HMMA; units go to the warps from the top
#pragma unroll 1
              for (int u = kDWarps - 1 - warp; u < nou16; u += kDWarps) {
                const uint4 a4 = d_lds128(sbase + laneO + (uint32_t)(nsb * 16 + u * 128));
                const uint32_t a[4] = {a4.x, a4.y, a4.z, a4.w};
                const uint32_t b0 = d_lds32(xo_lane + (uint32_t)(u * 32)), b1 = d_lds32(xo_lane + (uint32_t)(u * 32 + 16));
                mma_m16n8k16_f16f32(yo, a[0], a[1], a[2], a[3], b0, b1);
              }
            }
            // this warp is done with the slot
            __syncwarp();
            if constexpr (DBG) {
              if (dbg) dbg_math += clock64() - tw1;
            }
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ebars + 8 * cslot) : "memory");
            sbase += (uint32_t)L.slot;
            if (++cslot == NS) { cslot = 0; cpar ^= 1u; sbase = ring; }
          }
          {
            // end of the tile: accumulator columns -> one value per (row, batch row), to this warp's slice of the tile
            // (digit columns of a batch row weigh 256^i; M = 1: lane t = 2 is the zero-point lane)
            float v1 = fmaf(dc0, acc[0], dc1 * acc[1]), v2 = fmaf(dc0, acc[2], dc1 * acc[3]);
            v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
            v2 += __shfl_xor_sync(0xffffffffu, v2, 1);
            if (M == 1) {
              v1 += __shfl_xor_sync(0xffffffffu, v1, 2);
              v2 += __shfl_xor_sync(0xffffffffu, v2, 2);
              if (t == 0) asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(pdst), "f"(v1 + yo[0]), "f"(v2 + yo[2]) : "memory");
            } else {
              // batch row 1's outlier sums live in lane t = 0 of the quad (accumulator column 1)
              const float o1 = __shfl_sync(0xffffffffu, yo[1], lane & ~3), o3 = __shfl_sync(0xffffffffu, yo[3], lane & ~3);
              if (t == 0) asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(pdst), "f"(v1 + zacc0 + yo[0]), "f"(v2 + zacc1 + yo[2]) : "memory");
              if (t == 2) asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(pdst + 64u), "f"(v1 + zacc0 + o1), "f"(v2 + zacc1 + o3) : "memory");
            }
            pdst += (uint32_t)(kDWarps * M * 64);
          }
        }
      };
      if (KB == 1) run_tiles(std::true_type{});
      else run_tiles(std::false_type{});
    }
    dbg_prev = 0;
    asm volatile("cp.async.wait_all;" ::: "memory");        // (the next stage's descriptor)
    d_consumer_sync();
    dec_stamp<DBG>(L, s, 2);

    // ---- add the warps' slices in a fixed order, epilogue, store the rows this CTA owns -------------------------
    {
      const int epi = S->epilogue;
      const int nitems = R.ntiles * 16 * M;
      const bool res_poll = LL && S->res_ll != nullptr && S->res_src >= s_begin;   // the residual is produced in this launch
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
            const __half res = __ushort_as_half(d_ldx16(S->residual + (size_t)b * Pp->N + n, res_poll, pns));
            h = __hadd(res, h);
          }
          if (!LL) Pp->y[(size_t)b * Pp->N + n] = h;
        }
        if (LL) {
          // a y that later stages of this launch poll: 0xFFFF means "not yet written", so a NaN result is stored as 0x7E00.
          // Batch 1: the four rows of a qweight row (four adjacent lanes, valid together) leave as ONE 8-byte store --
          // a quarter of the transactions, which is what matters for the stores that cross NVLink.
          unsigned hb = (unsigned)__half_as_ushort(h);
          if ((hb & 0x7FFFu) > 0x7C00u) hb = 0x7E00u;
          const unsigned h1 = __shfl_down_sync(0xffffffffu, hb, 1), h2 = __shfl_down_sync(0xffffffffu, hb, 2),
                         h3 = __shfl_down_sync(0xffffffffu, hb, 3);
          const bool vec = M == 1 && (reinterpret_cast<uintptr_t>(Pp->y) & 7u) == 0;
          if (valid && vec) {
            if ((rr & 3) == 0) {
              const unsigned w0 = hb | (h1 << 16), w1 = h2 | (h3 << 16);
              if (Pp->nranks > 1) {
                // column-sharded: this rank's slice of the gathered row, into every rank's copy (NVLink stores)
                for (int pr = 0; pr < Pp->nranks; ++pr)
                  asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(Pp->y_peer[pr] + n), "r"(w0), "r"(w1) : "memory");
              } else {
                asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(Pp->y + n), "r"(w0), "r"(w1) : "memory");
              }
            }
          } else if (valid) {
            if (Pp->nranks > 1) {
              for (int pr = 0; pr < Pp->nranks; ++pr)
                asm volatile("st.relaxed.sys.global.u16 [%0], %1;" ::"l"(Pp->y_peer[pr] + n), "h"((unsigned short)hb) : "memory");
            } else {
              d_strelaxed16(Pp->y + (size_t)b * Pp->N + n, (unsigned short)hb);
            }
          }
        }
      }
    }
    dec_stamp<DBG>(L, s, 5);
    if (s + 1 < s_end && !(LL && S->nx_ll && S->nx_src >= s_begin)) barrier_arrive();   // the next stage waits at a barrier
    dec_stamp<DBG>(L, s, 3);
  }
  if (LL && RK.nranks > 1) {
    // every rank's slices of the last outputs have landed everywhere when the launch ends
    rank_barrier();
    if (cta == 0 && tid == 0) *reinterpret_cast<volatile unsigned*>(sync + 3) = rk_done + (unsigned)rk_now;
  }
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
  DecReset* d_resets = nullptr;
  int nresets = 0;
  int ll_env = -1;              // QEFT_DECODE_LL at creation: 1 / 0 = data-flow ordering on / off, unset = only for sharded programs
  bool use_ll = false;
  bool dirty = false;           // the description changed since the last upload (qeft_decode_program_shard)
  DecRanks ranks = {};          // column-sharded programs: the ranks' barrier counters
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

template <int M, bool LL, bool DBG>
static int dec_launch(const DecProgram* p, int s0, int s1, const DecLayout& L, size_t smem, int grid, cudaStream_t stream,
                      int nbar_total, int uses_ll) {
  auto kern = decode_w4_kernel<M, LL, DBG>;
  const DecReset* resets = p->d_resets;
  const int nresets = p->nresets;
  const DecRanks ranks = p->ranks;
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
  cfg.numAttrs = (s1 - s0 > 1 || p->ranks.nranks > 1) ? 1 : 0;        // (barriers or data-flow polling between stages: CTAs wait for one another)
  const DecStage* st = p->d_stages;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, st, s0, s1, p->d_sync, L, nbar_total, uses_ll, resets, nresets, ranks);
  if (e != cudaSuccess) return (int)e;
  count_launch();
  return QEFT_OK;
}

// Data-flow links: a stage whose x (or residual) IS the y of an earlier stage's projection polls that buffer instead of
// waiting at a barrier (the kernel's "data-flow by sentinel").  Only a buffer written by exactly ONE projection of the
// program can be a flag for itself: anything else (a scratch buffer reused by several stages, a width mismatch) makes
// the consuming stage wait at a barrier.  On one GPU barriers are faster, so data-flow is used by default
// only by column-sharded programs, where it IS the exchange; QEFT_DECODE_LL=1 / 0 at creation forces it on / off.  For a column-sharded projection
// (qeft_decode_program_shard) "the y" is the local copy of the gathered row.  Recomputed and uploaded whenever the
// program's description changes.
static int dec_link(DecProgram* p) {
  const int nstages = (int)p->h_stages.size(), m = p->m;
  auto out_of = [](const DecPart& pp) -> const void* { return pp.nranks > 1 ? pp.y_full : static_cast<const void*>(pp.y); };
  auto width_of = [](const DecPart& pp) { return pp.nranks > 1 ? pp.nranks * pp.N : pp.N; };
  for (int s = 0; s < nstages; ++s) {
    DecStage& d = p->h_stages[s];
    d.x_ll = nullptr; d.res_ll = nullptr; d.x_src = -1; d.res_src = -1; d.force_barrier = 0; d.nx_ll = 0; d.nx_src = -1;
    for (int i = 0; i < d.nparts; ++i) { d.part[i].y_ll = nullptr; d.part[i].ll_consumer = 1 << 30; }
  }
  auto writers = [&](const void* ptr) {
    int n = 0;
    for (int ps = 0; ps < nstages; ++ps) {
      const DecStage& pd = p->h_stages[ps];
      const int nout = pd.epilogue == QEFT_EPI_SWIGLU ? 1 : pd.nparts;
      for (int i = 0; i < nout; ++i) n += out_of(pd.part[i]) == ptr ? 1 : 0;
    }
    return n;
  };
  auto link = [&](const void* ptr, int width, int s, const void*& out_ll, int& out_src, bool slice_ok) -> int {
    // the latest earlier stage with a projection whose output buffer is exactly `ptr` ([m, width]); slice_ok (residuals,
    // which are polled element by element): or, batch 1, contains [ptr, ptr + width) -- a rank's own slice of a gathered row
    for (int ps = s - 1; ps >= 0; --ps) {
      DecStage& pd = p->h_stages[ps];
      const int nout = pd.epilogue == QEFT_EPI_SWIGLU ? 1 : pd.nparts;
      for (int i = 0; i < nout; ++i) {
        DecPart& pp = pd.part[i];
        const char* base = static_cast<const char*>(out_of(pp));
        const char* q = static_cast<const char*>(ptr);
        const bool inside = slice_ok && m == 1 && base && q >= base && q + 2 * (size_t)width <= base + 2 * (size_t)width_of(pp);
        if (out_of(pp) != ptr && !inside) continue;
        if (inside && writers(out_of(pp)) == 1) {
          pp.y_ll = out_of(pp);
          if (s < pp.ll_consumer) pp.ll_consumer = s;
          out_ll = out_of(pp);
          out_src = ps;
          return 1;
        }
        if (width_of(pp) != width || writers(ptr) != 1) return -1;   // produced in the program, but not linkable: barrier
        pp.y_ll = ptr;
        if (s < pp.ll_consumer) pp.ll_consumer = s;
        out_ll = ptr;
        out_src = ps;
        return 1;
      }
    }
    return 0;                                                   // not produced by this program: external input
  };
  std::vector<DecReset> resets;
  p->use_ll = p->ll_env == 1 || (p->ll_env < 0 && p->ranks.nranks > 1);
  if (p->use_ll) {
    for (int s = 1; s < nstages; ++s) {
      DecStage& d = p->h_stages[s];
      const int rx = link(d.x, d.K, s, d.x_ll, d.x_src, false);
      int rr = 0;
      if (d.residual) rr = link(d.residual, d.part[0].N, s, d.res_ll, d.res_src, true);
      if (rx < 0 || rr < 0) d.force_barrier = 1;
    }
    for (int s = 0; s + 1 < nstages; ++s) {
      const DecStage& nx = p->h_stages[s + 1];
      p->h_stages[s].nx_ll = (nx.x_ll != nullptr && !nx.force_barrier) ? 1 : 0;
      p->h_stages[s].nx_src = nx.x_src;
    }
    for (int s = 0; s < nstages; ++s) {
      const DecStage& d = p->h_stages[s];
      for (int i = 0; i < d.nparts; ++i)
        if (d.part[i].y_ll)
          resets.push_back(DecReset{reinterpret_cast<unsigned short*>(const_cast<void*>(d.part[i].y_ll)), m * width_of(d.part[i]), s,
                                    d.part[i].ll_consumer, 0});
    }
  }
  cudaError_t e = cudaMemcpy(p->d_stages, p->h_stages.data(), sizeof(DecStage) * (size_t)nstages, cudaMemcpyHostToDevice);
  if (p->d_resets) { cudaFree(p->d_resets); p->d_resets = nullptr; }
  p->nresets = (int)resets.size();
  if (e == cudaSuccess && p->nresets > 0) {
    e = cudaMalloc(&p->d_resets, sizeof(DecReset) * resets.size());
    if (e == cudaSuccess) e = cudaMemcpy(p->d_resets, resets.data(), sizeof(DecReset) * resets.size(), cudaMemcpyHostToDevice);
  }
  p->dirty = false;
  return (int)e;
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
  if (e == cudaSuccess) e = cudaMemset(p->d_sync, 0, 256);
  p->ll_env = dec_env_int("QEFT_DECODE_LL", -1);            // (read at every creation: tests build both kinds of program)
  if (e == cudaSuccess) e = (cudaError_t)dec_link(p);
  if (e != cudaSuccess) {
    if (p->d_stages) cudaFree(p->d_stages);
    if (p->d_sync) cudaFree(p->d_sync);
    if (p->d_resets) cudaFree(p->d_resets);
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
  if (p->d_resets) cudaFree(p->d_resets);
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
  cudaError_t e = cudaMemcpy(host_out, p->d_stamps, (p->h_stages.size() * (32 + 640) + 24) * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  return e == cudaSuccess ? QEFT_OK : (int)e;
}

extern "C" int qeft_decode_program_set_ranks(qeft_decode_program_t* prog, int nranks, int rank, void* const* barrier_peer) {
  if (!prog || !barrier_peer) return QEFT_E_NULL;
  if (nranks < 1 || nranks > QEFT_MAX_RANKS || rank < 0 || rank >= nranks) return QEFT_E_SHAPE;
  DecProgram* p = reinterpret_cast<DecProgram*>(prog);
  if (p->m != 1) return QEFT_E_BATCH;
  p->ranks.nranks = nranks;
  p->ranks.rank = rank;
  for (int i = 0; i < nranks; ++i) {
    if (!barrier_peer[i]) return QEFT_E_NULL;
    p->ranks.bar_peer[i] = static_cast<unsigned*>(barrier_peer[i]);
  }
  p->dirty = true;
  return QEFT_OK;
}

extern "C" int qeft_decode_program_shard(qeft_decode_program_t* prog, int stage, int part, void* const* y_full_peer) {
  if (!prog || !y_full_peer) return QEFT_E_NULL;
  DecProgram* p = reinterpret_cast<DecProgram*>(prog);
  const int P = p->ranks.nranks, rank = p->ranks.rank;
  if (P < 2) return QEFT_E_UNSUPPORTED;                      // qeft_decode_program_set_ranks first
  if (stage < 0 || stage >= (int)p->h_stages.size()) return QEFT_E_SHAPE;
  DecStage& d = p->h_stages[stage];
  const int nout = d.epilogue == QEFT_EPI_SWIGLU ? 1 : d.nparts;
  if (part < 0 || part >= nout) return QEFT_E_SHAPE;
  DecPart& dp = d.part[part];
  for (int i = 0; i < P; ++i) {
    if (!y_full_peer[i]) return QEFT_E_NULL;
    if (!check_align16(y_full_peer[i])) return QEFT_E_ALIGN;
    dp.y_peer[i] = static_cast<__half*>(y_full_peer[i]) + (size_t)rank * dp.N;
  }
  dp.nranks = P;
  dp.y_full = y_full_peer[rank];
  dp.y = dp.y_peer[rank];
  p->dirty = true;
  return QEFT_OK;
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
  if (p->dirty) {                                            // (re-links and uploads: not inside a stream capture)
    const int rc = dec_link(p);
    if (rc != QEFT_OK) return rc;
  }
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
  const size_t misc = 4096;
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
  static const int poll_env = dec_env_int("QEFT_DECODE_POLL_NS", 0);
  L.poll_ns = poll_env;
  static const int stamps_env = dec_env_int("QEFT_DECODE_STAMPS", 0);
  if (stamps_env && !p->d_stamps) {
    const size_t bytes = ((size_t)n * (32 + 640) + 24) * sizeof(unsigned long long);
    if (cudaMalloc(&p->d_stamps, bytes) == cudaSuccess) cudaMemset(p->d_stamps, 0, bytes);
    else p->d_stamps = nullptr;
  }
  L.stamps = p->d_stamps;
  L.stamps_all = p->d_stamps ? p->d_stamps + (size_t)n * 32 + 24 : nullptr;
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
  const bool sharded = p->ranks.nranks > 1;
  if (sharded && !p->use_ll) return QEFT_E_UNSUPPORTED;      // the exchange between ranks IS the data-flow protocol
  if (uses_ll) nbar_total += sharded ? 2 : 1;                // the barrier after re-arming the data-flow outputs
  if (sharded) nbar_total += 2;                              // the rank barrier that ends the launch
  const bool instrumented = stamps_env != 0 || debug_env != 0;
#define QEFT_DEC_GO(MM, LLV, DBGV) dec_launch<MM, LLV, DBGV>(p, stage_begin, stage_end, L, off, grid, st, nbar_total, uses_ll)
  if (instrumented) {
    if (uses_ll || sharded) return m == 1 ? QEFT_DEC_GO(1, true, true) : QEFT_DEC_GO(2, true, true);
    return m == 1 ? QEFT_DEC_GO(1, false, true) : QEFT_DEC_GO(2, false, true);
  }
  if (uses_ll || sharded) return m == 1 ? QEFT_DEC_GO(1, true, false) : QEFT_DEC_GO(2, true, false);
  return m == 1 ? QEFT_DEC_GO(1, false, false) : QEFT_DEC_GO(2, false, false);
#undef QEFT_DEC_GO
}
