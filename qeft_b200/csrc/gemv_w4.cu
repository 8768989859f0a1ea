// Decode-path dequant + GEMV for the packed QEFT QuantLinear (sm_100a).
//
// Replaces gemv_kernel / gemv_kernel_qeft of the reference
// (qeft/kernel/quantization_new/gemv/gemv_cuda.cu:73-204, gemv_cuda_qeft.cu:75-222).
//
// Design (see DESIGN.md "GEMV"; the measurements behind it are in profiles/r01_microbench_stream.md):
//   * persistent grid, one 8-warp CTA per SM.  The qweight rows (4 output rows each) of all projections of the
//     launch are split evenly over the CTAs (balance to one qweight row: 4096 rows over 148 SMs is 6 or 7 per
//     CTA), so every SM streams the same number of bytes and the whole chip finishes together.
//   * a CTA walks its rows as 16-row tiles (4 qweight rows; tiles are aligned to 16 rows of the projection, a
//     CTA streams only the qweight rows it owns of a boundary tile).  A tile's work is cut into units of two
//     128-column steps (int4) or two 32-column steps (fp16 outlier columns); unit u goes to warp u % 8.
//   * every lane prefetches exactly the 16-byte chunks it will itself consume with cp.async (LDGSTS) into a
//     lane-private shared-memory ring, D units deep (~90 KB in flight per SM): no producer warp, no barrier on
//     the data path, no registers tied up by loads in flight.  (1-D bulk copies were measured first: they need
//     >= 8 KB per operation to reach HBM speed, this layout's contiguous pieces are 256 B.)  The ring is filled
//     BEFORE griddepcontrol.wait, so under programmatic dependent launch the next kernel's weights stream in
//     while this kernel computes; only x waits for the previous kernel.
//   * a lane turns every nibble pair into an fp16 pair with ONE lop3 (the nibble is OR-ed into the mantissa of
//     1024.0, giving 1024+q or 1024+16q exactly).  These are, without any shuffle, the A fragment of
//     mma.m16n8k16 (the packed order IS that fragment order); x is the B fragment (batch m <= 8 columns);
//     accumulation is fp32 in two chains, one per nibble position, and the 1024 bias is removed per
//     128-column step with the step's x sums:
//         sum(q x) = acc_lo + acc_hi / 16 - (1024 sum_lo(x) + 64 sum_hi(x))
//         y += s * sum(q x) + sz * sum(x)                       (fp32, once per step)
//   * the fp16 outlier columns go through the same MMA (either layout: plain [N, r] or the reference's
//     row-pair interleaved [N/2, 2r], regrouped with byte permutes).  Accumulator row "g" of a lane is tile row
//     8 (g/4) + g%4 and row "g+8" is that + 4, which is the interleaved layout's own row pairing.
//   * each warp keeps a tile's partial sums in registers and writes them once to its own shared-memory slice;
//     after a CTA barrier the slices are added in a fixed order (deterministic), bias is added, and the CTA's
//     owned rows are stored as fp16.
#include "common.cuh"

#include <stdlib.h>

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
  int q_begin;            // first qweight row of this part in the launch-wide numbering
};

struct GemvParams;
typedef GemvParams GemvParamsFwd;
struct GemvParams {
  GemvPart part[QEFT_GEMV_MAX_PARTS];
  int nparts;
  const __half* x;        // [m, K]
  const int32_t* gather;  // [K] or null
  int m, K, r;
  int g128;               // G / 128 (1 for the common G = 128), 0 for per-channel scales (G == K)
  int ow_layout;
  int nsteps;             // ceil((K - r) / 128)
  int nchunks;            // (K - r) / 32 live 32-column chunks
  int nku;                // int4 units per tile = ceil(nsteps / 2)
  int nou;                // outlier units per tile = ceil(r / 64)
  int xstride;            // halves between batch rows of the staged x (128 nsteps + 32: rows start 16 banks apart)
  int q_lo, q_hi;         // window of launch-wide qweight rows this launch covers
  int max_tiles;          // upper bound of tiles per CTA (sizes the partial-sum slices)
  unsigned long long* stamps;   // debug (QEFT_GEMV_STAMPS): [launch slot][8] globaltimer values of CTA 0, or null
  unsigned long long* cta_stamps;   // debug: [512 CTAs][4] start / waited / staged / stored of this launch, or null
  // fused all-gather (column-sharded decode): results go to every rank's gathered buffer
  int nranks;                   // 0: plain launch (part.y)
  int y_ld;
  __half* y_peer[QEFT_MAX_RANKS][QEFT_GEMV_MAX_PARTS];
  uint32_t* done_peer[QEFT_MAX_RANKS];
  const uint32_t* wait_flag;
  int flag_only;          // 1: the arrival counter alone orders this launch after its input (no grid-completion wait)
  const uint32_t* epoch;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void stamp(const GemvParamsFwd& p, int i);
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint4 lds_v4(uint32_t a) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a) : "memory");
  return r;
}
// predicated forms: no branch, lanes with p == false do not touch shared memory (their outputs are undefined)
__device__ __forceinline__ uint4 lds_v4_if(uint32_t a, int p) {
  uint4 r;
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t@q ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];\n\t}"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a), "r"(p) : "memory");
  return r;
}
__device__ __forceinline__ uint2 lds_v2_if(uint32_t a, int p) {
  uint2 r;
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %3, 0;\n\t@q ld.shared.v2.u32 {%0,%1}, [%2];\n\t}"
               : "=r"(r.x), "=r"(r.y) : "r"(a), "r"(p) : "memory");
  return r;
}
__device__ __forceinline__ void cp_async16_if(uint32_t dst, const void* src, int p) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q cp.async.cg.shared.global [%0], [%1], 16;\n\t}"
               ::"r"(dst), "l"(src), "r"(p) : "memory");
}
__device__ __forceinline__ uint2 lds_v2(uint32_t a) {
  uint2 r;
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(a) : "memory");
  return r;
}
__device__ __forceinline__ float lds_h(uint32_t a) {
  unsigned short h;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(a) : "memory");
  return __half2float(__ushort_as_half(h));
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}


// Activation loads.  When x is the gathered buffer of an earlier launch of a column-sharded chain (wait_flag set), peers
// may still be storing into it while this kernel is already resident (programmatic dependent launch): the read-only
// (.nc) path requires data that is constant for the kernel's lifetime, so those launches read x through L2 (.cg) after
// the acquire of the arrival counter.  Everything else keeps the read-only path.
__device__ __forceinline__ uint4 ldx_v4(const void* p, bool coherent) {
  uint4 r;
  if (coherent) asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  else r = ldg_nc_v4(p);
  return r;
}
__device__ __forceinline__ uint2 ldx_v2(const void* p, bool coherent) {
  uint2 r;
  if (coherent) asm volatile("ld.global.cg.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p) : "memory");
  else r = ldg_nc_v2(p);
  return r;
}
__device__ __forceinline__ unsigned short ldx_u16(const void* p, bool coherent) {
  unsigned short r;
  if (coherent) asm volatile("ld.global.cg.u16 %0, [%1];" : "=h"(r) : "l"(p) : "memory");
  else r = ldg_nc_u16(p);
  return r;
}
__device__ __forceinline__ __half ldx_h(const __half* p, bool coherent) { return __ushort_as_half(ldx_u16(p, coherent)); }

__device__ __forceinline__ void stamp(const GemvParams& p, int i) {
  if (p.stamps && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (blockIdx.x == 0) p.stamps[i] = t;
    if (p.cta_stamps && blockIdx.x < 512 && (i == 0 || i == 2 || i == 3 || i == 5)) {
      const int j = i == 0 ? 0 : (i == 2 ? 1 : (i == 3 ? 2 : 3));
      p.cta_stamps[blockIdx.x * 4 + j] = t;
    }
  }
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

constexpr int kStepBytes = 256;                 // one qweight row's bytes of a 128-column step
constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kSlotBytes = 4 * 512 + 128;       // per warp and unit: 4 chunks of 16 B per lane + 2 steps x (16 s | 16 z)

// tiles [t0, t0 + nt) of part `part` (ordinals cum .. cum + nt in this CTA); the CTA owns the part-local qweight
// rows [pa, pb)
struct Seg { int part, t0, nt, pa, pb, cum; };

// D: ring depth in units per warp.  XS: x staged in shared memory (otherwise read through L1/L2 at every use).
// I8 (batch m <= 2, needs XS): the int4 columns run on the int8 tensor path -- see "int8 path" below.
template <int D, bool XS, bool I8>
__global__ void __launch_bounds__(kThreads, 2)
gemv_w4_kernel(const GemvParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ Seg segs[QEFT_GEMV_MAX_PARTS];
  __shared__ int s_nseg, s_ntiles;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int ra = 8 * (g >> 2) + (g & 3), rb = ra + 4;   // tile rows of this lane's two accumulator rows
  const int K = p.K, r = p.r, m = p.m;
  const bool xcoh = p.wait_flag != nullptr;     // x is a gathered buffer written during this chain: coherent loads
  const int nsteps = p.nsteps, nchunks = p.nchunks;
  const int nku = p.nku, upt = p.nku + p.nou;
  const bool inter = p.ow_layout == QEFT_OW_INTERLEAVED;

  // ---- shared memory carve-up -----------------------------------------------------------
  uint8_t* ring = smem_raw;                                                      // [warps][D][kSlotBytes]
  // int8 path: an accumulator column is (batch row b, digit d) = 3 b + d; the outlier sums get m more columns
  const int ncols = I8 ? 3 * m : m;
  const int pc = I8 ? 4 * m : m;                                                 // columns of a partial-sum row
  float* part = reinterpret_cast<float*>(ring + (size_t)kWarps * D * kSlotBytes);   // [max_tiles][warps][16][pc]
  float4* sums = reinterpret_cast<float4*>(part + (size_t)p.max_tiles * kWarps * 16 * pc);
                                     // [nsteps][4]: columns 2t, 2t+1: {sum x, sum x, c, c}   (I8: {coef, coef, sum x, sum x})
  __half* xs = reinterpret_cast<__half*>(sums + nsteps * 4);                     // staged x [m][xstride], dead columns zeroed
  uint8_t* xdig = reinterpret_cast<uint8_t*>(xs);                                // I8: [nsteps][ncols][144 B] digit bytes
  __half* xo = I8 ? reinterpret_cast<__half*>(xdig + (size_t)nsteps * ncols * 144)
                  : xs + (XS ? (size_t)m * p.xstride : 0);                       // staged outlier activations [m][r]
  int* xexp = reinterpret_cast<int*>(xo + (size_t)m * r);                        // I8: [nsteps][m] block exponents
  const uint32_t ring_u32 = smem_u32(ring) + (uint32_t)(warp * D * kSlotBytes);

  if (tid == 0) {
    // the CTA's share of the launch's qweight rows, cut into per-part tile runs
    const long span = p.q_hi - p.q_lo;
    const int qa = p.q_lo + (int)(((long)blockIdx.x * span) / gridDim.x);
    const int qb = p.q_lo + (int)(((long)(blockIdx.x + 1) * span) / gridDim.x);
    int ns = 0, cum = 0;
    for (int i = 0; i < p.nparts; ++i) {
      const int b = p.part[i].q_begin, e = b + (p.part[i].N >> 2);
      const int lo = max(qa, b), hi = min(qb, e);
      if (lo < hi) {
        Seg s;
        s.part = i; s.pa = lo - b; s.pb = hi - b;
        s.t0 = s.pa >> 2; s.nt = ((s.pb + 3) >> 2) - s.t0; s.cum = cum;
        cum += s.nt;
        segs[ns++] = s;
      }
    }
    s_nseg = ns; s_ntiles = cum;
  }
  __syncthreads();
  stamp(p, 0);
  pdl_launch_dependents();
  const int nseg = s_nseg, ntiles = s_ntiles;
  const int nunits = ntiles * upt;
  // this warp's contiguous run of units [u0, u1) of the CTA's (tile, unit-in-tile) sequence
  const int u0 = (int)(((long)warp * nunits) / kWarps), u1 = (int)(((long)(warp + 1) * nunits) / kWarps);

  // ---- prefetch side -------------------------------------------------------------------------------
  // cached per tile: global sources of this lane
  const uint8_t* pf_w = nullptr;     // row ra's chunk t of step 0 (row rb: + one qweight row)
  const __half* pf_sc = nullptr;     // lanes 0..7: scale / scaled-zero source of step (lane / 4) at group 0
  const __half* pf_ow = nullptr;     // outlier source of this lane at column 0
  int pf_ownA = 0, pf_ownB = 0, pf_sc_ok = 0;
  size_t pf_sc_step = 0;             // halves between consecutive steps' scale rows (G = 128: N)
  const uint8_t* pf_wb = nullptr;    // pf_w + one qweight row
  const __half* pf_scl = nullptr;    // pf_sc + (lane / 4) steps (this lane's step of unit 0)
  size_t pf_sc_step2 = 0;            // two steps
  auto pf_set_tile = [&](int tord) {
    int sg = 0;
#pragma unroll
    for (int j = 1; j < QEFT_GEMV_MAX_PARTS; ++j)
      if (j < nseg && tord >= segs[j].cum) sg = j;
    const Seg S = segs[sg];
    const GemvPart& P = p.part[S.part];
    const int T = S.t0 + tord - S.cum;
    const int qA = 4 * T + 2 * (g >> 2);
    pf_ownA = qA >= S.pa && qA < S.pb;
    pf_ownB = qA + 1 >= S.pa && qA + 1 < S.pb;
    pf_w = P.qw + (size_t)qA * (size_t)(2 * K) + (size_t)((t >> 1) * 128 + (g & 3) * 32 + (t & 1) * 16);
    const int which = (lane >> 1) & 1, half8 = lane & 1;
    pf_sc = (which ? P.szeros : P.scales) + 16 * T + 8 * half8;
    pf_sc_ok = lane < 8 && (16 * T + 8 * half8) < P.N;
    pf_sc_step = (size_t)P.N;
    pf_wb = pf_w + (size_t)(2 * K);
    pf_scl = pf_sc + (size_t)(lane >> 2) * (size_t)P.N;
    pf_sc_step2 = 2 * (size_t)P.N;
    if (inter)
      pf_ow = P.ow + (size_t)(8 * T + 4 * (g >> 2) + (g & 3)) * (size_t)(2 * r) + 8 * t;
    else
      pf_ow = P.ow + (size_t)(16 * T + ra) * (size_t)r + 8 * t;
  };
  // issue the copies of unit `su` of the cached tile into `slot` (predicated copies, no divergent branches)
  const int pf_nfull = p.g128 == 1 ? (nchunks >> 3) : 0;     // units whose two steps are fully live (and G = 128)
  auto issue = [&](uint32_t slot, int su) {
    const uint32_t mine = slot + lane * 16;
    if (su < pf_nfull) {
      // the common case: no liveness tests, no group arithmetic -- five predicated copies
      const uint8_t* a = pf_w + (uint32_t)su * (uint32_t)(2 * kStepBytes);
      const uint8_t* b = pf_wb + (uint32_t)su * (uint32_t)(2 * kStepBytes);
      cp_async16_if(mine, a, pf_ownA);
      cp_async16_if(mine + 1024, a + kStepBytes, pf_ownA);
      cp_async16_if(mine + 512, b, pf_ownB);
      cp_async16_if(mine + 1536, b + kStepBytes, pf_ownB);
      cp_async16_if(slot + 2048 + lane * 16, pf_scl + (size_t)su * pf_sc_step2, pf_sc_ok);
    } else if (su < nku) {
      const uint8_t* a = pf_w + (size_t)su * (2 * kStepBytes);
      const uint8_t* b = a + (size_t)(2 * K);
      const int c0 = 8 * su + t;                    // this lane's 32-column chunk of the unit's first step
      const int l0 = c0 < nchunks, l1 = c0 + 4 < nchunks;
      cp_async16_if(mine, a, pf_ownA & l0);
      cp_async16_if(mine + 512, b, pf_ownB & l0);
      cp_async16_if(mine + 1024, a + kStepBytes, pf_ownA & l1);
      cp_async16_if(mine + 1536, b + kStepBytes, pf_ownB & l1);
      const int s = 2 * su + (lane >> 2);
      int grp = s;
      if (p.g128 != 1) grp = p.g128 == 0 ? 0 : s / p.g128;       // uniform
      cp_async16_if(slot + 2048 + lane * 16, pf_sc + (size_t)grp * pf_sc_step, pf_sc_ok & (int)(s < nsteps));
    } else {
      const int c0 = 64 * (su - nku);               // first outlier column of the unit
      if (inter) {
        // one interleaved row holds rows ra and rb: 16 bytes = 4 columns of both
        const __half* a = pf_ow + 2 * c0;
        const int own = pf_ownA | pf_ownB;
#pragma unroll
        for (int j = 0; j < 4; ++j) cp_async16_if(mine + j * 512, a + 32 * j, own & (int)(c0 + 16 * j < r));
      } else {
        const __half* a = pf_ow + c0;
        const __half* b = a + (size_t)(4 * r);
#pragma unroll
        for (int si = 0; si < 2; ++si) {
          const int l = c0 + 32 * si < r;
          cp_async16_if(mine + (2 * si) * 512, a + 32 * si, pf_ownA & l);
          cp_async16_if(mine + (2 * si + 1) * 512, b + 32 * si, pf_ownB & l);
        }
      }
    }
  };

  // fill the ring: the first D units of this warp (weights do not depend on the previous kernel)
  int pu = u0;                         // next unit to prefetch
  int ptord = u0 / upt, psu = u0 - ptord * upt;
  if (u0 < u1) pf_set_tile(ptord);
#pragma unroll 1                       // (one copy of the issue code: the prologue runs once, instruction-cache cold)
  for (int d = 0; d < D; ++d) {
    if (pu < u1) {
      issue(ring_u32 + d * kSlotBytes, psu);
      ++pu;
      if (++psu == upt) { psu = 0; ++ptord; if (pu < u1) pf_set_tile(ptord); }
    }
    cp_async_commit();                 // always: the group count is the clock
  }

  // zero the partial-sum slices (a warp without a unit in a tile contributes zero)
  for (int i = tid; i < ntiles * kWarps * 16 * pc; i += kThreads) part[i] = 0.f;

  stamp(p, 1);
  // x (and y as a reused buffer) belong to the previous kernel until here.  In the column-sharded chain the input is
  // the gathered buffer of an earlier launch, complete exactly when that launch's arrival counter says so (every rank,
  // this one included, stores its slice and then signals with release semantics): the counter alone orders the two
  // launches, and waiting for the previous GRID to drain as well (its system-scope fences and NVLink stores included)
  // only adds the kernel-boundary latency to every launch of the chain.
  if (!p.flag_only) pdl_wait();
  if (p.wait_flag) {
    // column-sharded chain: the launch this one depends on has finished on THIS rank (stream order); wait until every
    // rank's slice of its output has landed here: its arrival counter reaches epoch x ranks
    if (tid == 0) {
      const uint32_t want = *reinterpret_cast<const volatile uint32_t*>(p.epoch) * (uint32_t)p.nranks * kArrivalsPerLaunch;
      uint32_t got;
      do {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(got) : "l"(p.wait_flag) : "memory");
      } while ((int)(got - want) < 0);
    }
    __syncthreads();
  }
  stamp(p, 2);

  // ---- x: staged copy and per-step sums ---------------------------------------------------------------
  const __half* xg = p.x;
  if constexpr (I8) {
    // ---- int8 path: x of every 128-column step as block fixed point, three signed-byte digits -------------
    //   x_k ~= X_k 2^(e-22),  X_k = d0 + 256 d1 + 65536 d2,  d_i in [-128, 127],  2^e > max |x| of the step
    // (exact for every element within 12 binades of the step's maximum; fp16 has 11 significant bits).  The digit
    // bytes are stored in the byte order of the masked weight words, so that they are B fragments of
    // mma.m16n8k32.u8.s8 as they lie: for the 32-column chunk t of a step, word c of a weight chunk masked with
    // 0x0f0f0f0f holds k = {2c, 2c+16, 2c+1, 2c+17} and masked with 0xf0f0f0f0 holds 16 x {2c+8, 2c+24, 2c+9, 2c+25}.
    // One pass: item = (batch row, step, chunk t, nibble position h) = 16 values (columns k0..k0+7 and k0+16..k0+23);
    // the 8 items of a step sit in 8 adjacent lanes, which agree on the step's sum and maximum by shuffles.
    const int live_k = nchunks * 32;
    float4* sums4 = sums;
    const int nitems = m * nsteps * 8;
    // all global loads of up to kPre passes are issued before the first use (one L2 round trip, not one per pass);
    // with an o_proj gather (qlinear.py:275) up to kPreG passes: the 16 indices of an item as four vector loads,
    // then the 16 scattered halves -- two dependent round trips for the whole staging instead of two per pass
    constexpr int kPre = 4, kPreG = 2;
    uint4 pre[kPre][2];
    if (!p.gather) {
#pragma unroll
      for (int q = 0; q < kPre; ++q) {
        const int it = q * kThreads + tid;
        pre[q][0] = pre[q][1] = make_uint4(0u, 0u, 0u, 0u);
        if (it < nitems) {
          const int sb = it >> 3, b = sb / nsteps, s = sb - b * nsteps;
          const int k0 = s * 128 + ((it >> 1) & 3) * 32 + 8 * (it & 1);
          if (k0 < live_k) {
            pre[q][0] = ldx_v4(xg + (size_t)b * K + k0, xcoh);
            pre[q][1] = ldx_v4(xg + (size_t)b * K + k0 + 16, xcoh);
          }
        }
      }
    } else {
      int4 gi[kPreG][4];
      bool glive[kPreG];
      const unsigned short* grow[kPreG];
#pragma unroll
      for (int q = 0; q < kPreG; ++q) {
        const int it = q * kThreads + tid;
        const int sb = it < nitems ? (it >> 3) : 0, b = sb / nsteps, s = sb - b * nsteps;
        const int k0 = s * 128 + ((it >> 1) & 3) * 32 + 8 * (it & 1);
        glive[q] = it < nitems && k0 < live_k;
        grow[q] = reinterpret_cast<const unsigned short*>(xg + (size_t)b * K);
#pragma unroll
        for (int j = 0; j < 4; ++j) gi[q][j] = make_int4(0, 0, 0, 0);
        if (glive[q]) {
          const int4* ip = reinterpret_cast<const int4*>(p.gather + k0);
          gi[q][0] = __ldg(ip); gi[q][1] = __ldg(ip + 1); gi[q][2] = __ldg(ip + 4); gi[q][3] = __ldg(ip + 5);
        }
      }
#pragma unroll
      for (int q = 0; q < kPreG; ++q) {
        pre[q][0] = pre[q][1] = make_uint4(0u, 0u, 0u, 0u);
        if (glive[q]) {
          const unsigned short* xr = grow[q];
          auto pk = [&](int a, int b2) { return (uint32_t)ldx_u16(xr + a, xcoh) | ((uint32_t)ldx_u16(xr + b2, xcoh) << 16); };
          pre[q][0] = make_uint4(pk(gi[q][0].x, gi[q][0].y), pk(gi[q][0].z, gi[q][0].w), pk(gi[q][1].x, gi[q][1].y), pk(gi[q][1].z, gi[q][1].w));
          pre[q][1] = make_uint4(pk(gi[q][2].x, gi[q][2].y), pk(gi[q][2].z, gi[q][2].w), pk(gi[q][3].x, gi[q][3].y), pk(gi[q][3].z, gi[q][3].w));
        }
      }
    }
    const int npre = p.gather ? kPreG : kPre;
    // the outlier activations (fp16, legacy MMA): loaded here, stored after the digit passes, so that their round
    // trip overlaps the conversion work instead of following it
    const int nxo = m * (r >> 3);
    const bool xo_early = nxo <= kThreads;
    uint4 xo_v = make_uint4(0u, 0u, 0u, 0u);
    if (xo_early && tid < nxo) {
      const int b = tid / (r >> 3), jj = tid - b * (r >> 3);
      const __half* xrow = xg + (size_t)b * K;
      if (p.gather) {
        const int4* ip = reinterpret_cast<const int4*>(p.gather + K - r + 8 * jj);
        const int4 a = __ldg(ip), c = __ldg(ip + 1);
        const unsigned short* xr = reinterpret_cast<const unsigned short*>(xrow);
        auto pk = [&](int i0, int i1) { return (uint32_t)ldx_u16(xr + i0, xcoh) | ((uint32_t)ldx_u16(xr + i1, xcoh) << 16); };
        xo_v = make_uint4(pk(a.x, a.y), pk(a.z, a.w), pk(c.x, c.y), pk(c.z, c.w));
      } else {
        xo_v = ldx_v4(xrow + K - r + 8 * jj, xcoh);
      }
    }
#pragma unroll 1
    for (int base = 0, pass = 0; base < nitems; base += kThreads, ++pass) {   // whole warps enter together (full-mask shuffles)
      const int it = base + tid;
      const bool valid = it < nitems;
      const int h = it & 1, tt = (it >> 1) & 3;
      const int sb = valid ? (it >> 3) : 0;                        // (batch row, step) index
      const int b = sb / nsteps, s = sb - b * nsteps;
      const int k0 = s * 128 + tt * 32 + 8 * h;
      const bool live = valid && k0 < live_k;                      // live_k is a multiple of 32: whole chunks
      const __half* xrow = xg + (size_t)b * K;
      uint32_t w[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) w[j] = 0u;
      if (live) {
        if (pass < npre) {
          // select the preloaded pair without dynamic register indexing
          uint4 v0 = pre[0][0], v1 = pre[0][1];
#pragma unroll
          for (int q = 1; q < kPre; ++q)
            if (pass == q) { v0 = pre[q][0]; v1 = pre[q][1]; }
          w[0] = v0.x; w[1] = v0.y; w[2] = v0.z; w[3] = v0.w; w[4] = v1.x; w[5] = v1.y; w[6] = v1.z; w[7] = v1.w;
        } else if (p.gather) {
          __half tmp[16];
#pragma unroll
          for (int j = 0; j < 8; ++j) { tmp[j] = ldx_h(xrow + p.gather[k0 + j], xcoh); tmp[8 + j] = ldx_h(xrow + p.gather[k0 + 16 + j], xcoh); }
#pragma unroll
          for (int j = 0; j < 8; ++j) w[j] = reinterpret_cast<uint32_t*>(tmp)[j];
        } else {
          const uint4 v0 = ldx_v4(xrow + k0, xcoh), v1 = ldx_v4(xrow + k0 + 16, xcoh);
          w[0] = v0.x; w[1] = v0.y; w[2] = v0.z; w[3] = v0.w; w[4] = v1.x; w[5] = v1.y; w[6] = v1.z; w[7] = v1.w;
        }
      }
      float2 f[8];
      float sum = 0.f, mx = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        f[j] = half2_bits_to_float2(w[j]);
        sum += f[j].x + f[j].y;
        mx = fmaxf(mx, fmaxf(fabsf(f[j].x), fabsf(f[j].y)));
      }
#pragma unroll
      for (int o = 4; o >= 1; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      }
      // 2^e > mx: e = exponent field - 126 (fp16 inputs are normal floats or zero)
      const int e = mx > 0.f ? (int)((__float_as_uint(mx) >> 23) & 0xff) - 126 : -100;
      const float sc = e > -100 ? __uint_as_float((uint32_t)(127 + 22 - e) << 23) : 0.f;          // 2^(22-e)
      if (valid) {
        // word c of a digit row = digits of {first[2c], second[2c], first[2c+1], second[2c+1]}.
        // X = rint(x 2^(22-e)) by the magic-number add (|X| < 2^22): bits(fma(x, sc, 1.5 2^23)) = 0x4B400000 + X.
        // Z = X + 0x808080 has unsigned bytes b_i with X = sum (b_i - 128) 256^i, so the signed digits are the
        // bytes of Z ^ 0x808080: three instructions per value, then byte permutes gather each digit row.
        uint32_t dw[3][4];
        auto digits = [&](float v) {
          const int bits = __float_as_int(fmaf(v, sc, 12582912.f));
          return (uint32_t)(bits + (0x00808080 - 0x4B400000)) ^ 0x00808080u;
        };
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t d0 = digits(f[c].x), d1 = digits(f[4 + c].x), d2 = digits(f[c].y), d3 = digits(f[4 + c].y);
          dw[0][c] = prmt(prmt(d0, d1, 0x0040), prmt(d2, d3, 0x0040), 0x5410);
          dw[1][c] = prmt(prmt(d0, d1, 0x0051), prmt(d2, d3, 0x0051), 0x5410);
          dw[2][c] = prmt(prmt(d0, d1, 0x0062), prmt(d2, d3, 0x0062), 0x5410);
        }
#pragma unroll
        for (int d = 0; d < 3; ++d)
          *reinterpret_cast<uint4*>(xdig + ((size_t)s * ncols + 3 * b + d) * 144 + tt * 32 + h * 16) =
              make_uint4(dw[d][0], dw[d][1], dw[d][2], dw[d][3]);
        // the step's table: lane (tt, h) of the 8 writes accumulator columns ... one float4 {coef(2j), coef(2j+1), X(2j), X(2j+1)}
        // per column pair j; columns of this batch row are 3b .. 3b+2 (digit d: coef 2^(e-22+8d)/16, X only for d = 0)
        if ((it & 7) == 0) {
          float* row = reinterpret_cast<float*>(sums4 + s * 4);
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            const int col = 3 * b + d;
            float* dst = row + (col >> 1) * 4 + (col & 1);
            dst[0] = e > -100 ? __uint_as_float((uint32_t)(127 + e - 22 + 8 * d - 4) << 23) : 0.f;
            dst[2] = d == 0 ? sum : 0.f;
          }
          if (b == m - 1)                                          // the columns no batch row owns
            for (int col = 3 * m; col < 8; ++col) {
              float* dst = row + (col >> 1) * 4 + (col & 1);
              dst[0] = 0.f; dst[2] = 0.f;
            }
        }
      }
    }
    if (xo_early) {
      if (tid < nxo) {
        const int b = tid / (r >> 3), jj = tid - b * (r >> 3);
        *reinterpret_cast<uint4*>(xo + (size_t)b * r + 8 * jj) = xo_v;
      }
    } else if (r > 0) {                                            // the outlier activations (fp16, legacy MMA)
      for (int j = tid; j < m * (r >> 3); j += kThreads) {
        const int b = j / (r >> 3), jj = j - b * (r >> 3);
        const __half* xrow = xg + (size_t)b * K;
        uint4 v;
        if (p.gather) {
          __half tmp[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) tmp[e] = ldx_h(xrow + p.gather[K - r + 8 * jj + e], xcoh);
          v = *reinterpret_cast<uint4*>(tmp);
        } else {
          v = ldx_v4(xrow + K - r + 8 * jj, xcoh);
        }
        *reinterpret_cast<uint4*>(xo + (size_t)b * r + 8 * jj) = v;
      }
    }
  } else {
    // units of 16 halves; 8 consecutive units = one 128-column step.  Inside a unit the first 8 halves sit in
    // "low nibble" k-slots (k % 16 < 8) and the last 8 in "high nibble" slots.
    const int upr = nsteps * 8;                             // units per batch row that the int4 steps touch
    const int live_k = nchunks * 32;
    float* sums_f = reinterpret_cast<float*>(sums);
    for (int b = 0; b < m; ++b) {
      const __half* xrow = xg + (size_t)b * K;
      for (int u = tid; u < ((upr + 31) & ~31); u += kThreads) {   // whole warps enter together (full-mask shuffles)
        const int k = u * 16;
        float lo = 0.f, hi = 0.f;
        if (u < upr) {
          uint4 v0 = make_uint4(0u, 0u, 0u, 0u), v1 = v0;
          if (k < live_k) {                                  // live_k is a multiple of 32: whole units
            if (XS && p.gather) {
              __half tmp[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) tmp[j] = ldx_h(xrow + p.gather[k + j], xcoh);
              v0 = *reinterpret_cast<uint4*>(tmp);
              v1 = *reinterpret_cast<uint4*>(tmp + 8);
            } else {
              v0 = ldx_v4(xrow + k, xcoh);
              v1 = ldx_v4(xrow + k + 8, xcoh);
            }
            const float2 a0 = half2_bits_to_float2(v0.x), a1 = half2_bits_to_float2(v0.y);
            const float2 a2 = half2_bits_to_float2(v0.z), a3 = half2_bits_to_float2(v0.w);
            const float2 c0 = half2_bits_to_float2(v1.x), c1 = half2_bits_to_float2(v1.y);
            const float2 c2 = half2_bits_to_float2(v1.z), c3 = half2_bits_to_float2(v1.w);
            lo = ((a0.x + a0.y) + (a1.x + a1.y)) + ((a2.x + a2.y) + (a3.x + a3.y));
            hi = ((c0.x + c0.y) + (c1.x + c1.y)) + ((c2.x + c2.y) + (c3.x + c3.y));
          }
          if (XS) {
            // dead columns of the last step are staged as zeros.  Inside a step the 8-column piece j of 32-column
            // chunk t sits at halves 32 j + 8 t, so that the four lanes t of one fragment load read 64 contiguous bytes.
            const int tt = (k & 127) >> 5, j0 = (k & 31) >> 3;
            __half* d = xs + (size_t)b * p.xstride + (k & ~127) + 8 * tt;
            *reinterpret_cast<uint4*>(d + 32 * j0) = v0;
            *reinterpret_cast<uint4*>(d + 32 * (j0 + 1)) = v1;
          }
        }
        lo += __shfl_xor_sync(0xffffffffu, lo, 4); hi += __shfl_xor_sync(0xffffffffu, hi, 4);
        lo += __shfl_xor_sync(0xffffffffu, lo, 2); hi += __shfl_xor_sync(0xffffffffu, hi, 2);
        lo += __shfl_xor_sync(0xffffffffu, lo, 1); hi += __shfl_xor_sync(0xffffffffu, hi, 1);
        if ((u & 7) == 0 && u < upr) {
          // float4 slot b/2 of the step: {X(2t), X(2t+1), C(2t), C(2t+1)}
          float* d = sums_f + ((u >> 3) * 4 + (b >> 1)) * 4 + (b & 1);
          d[0] = lo + hi;
          d[2] = fmaf(1024.f, lo, 64.f * hi);
        }
      }
      if (XS && r > 0) {                                     // the outlier activations
        for (int j = tid; j < (r >> 3); j += kThreads) {
          uint4 v;
          if (p.gather) {
            __half tmp[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) tmp[e] = ldx_h(xrow + p.gather[K - r + 8 * j + e], xcoh);
            v = *reinterpret_cast<uint4*>(tmp);
          } else {
            v = ldx_v4(xrow + K - r + 8 * j, xcoh);
          }
          *reinterpret_cast<uint4*>(xo + (size_t)b * r + 8 * j) = v;
        }
      }
    }
    for (int i = tid; i < nsteps * 16; i += kThreads) {      // batch columns >= m
      const int col = 2 * ((i >> 2) & 3) + (i & 1);
      if (col >= m) sums_f[i] = 0.f;
    }
  }
  __syncthreads();

  stamp(p, 3);
  // ---- the units of this warp -----------------------------------------------------------------------
  // Batch column n of the B fragment only feeds output column n, so lanes of batch rows >= m simply re-read
  // row 0 (a broadcast; their output columns are never stored): no divergence, no zero fill.
  const int gx = g < m ? g : 0;
  const uint32_t x_lane = smem_u32(xs) + (uint32_t)((gx * p.xstride + t * 8) * 2);
  const uint32_t xo_lane = smem_u32(xo) + (uint32_t)((gx * r) * 2);
  const __half* xg_lane = xg + (size_t)gx * K;
  const uint32_t sums_lane = smem_u32(sums) + 16 * t;
  // int8 path: digit column g of the B fragment (columns >= ncols re-read the last one: never stored)
  const uint32_t xdig_lane = smem_u32(xdig) + (uint32_t)((g < ncols ? g : ncols - 1) * 144 + t * 32);
  const uint32_t xdig_step = (uint32_t)(ncols * 144);

  // B fragments of step s: x[gx][128 s + 32 t .. +32], natural order
  auto load_x = [&](uint32_t (&xb)[16], int s) {
    if (XS) {
      const uint32_t a = x_lane + (uint32_t)(s * 256);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 v = lds_v4(a + 64 * j);
        xb[4 * j + 0] = v.x; xb[4 * j + 1] = v.y; xb[4 * j + 2] = v.z; xb[4 * j + 3] = v.w;
      }
    } else {
      const bool live = (4 * s + t) < nchunks;
      const __half* xp = xg_lane + s * 128 + t * 32;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 v = live ? ldx_v4(xp + 8 * j, xcoh) : make_uint4(0u, 0u, 0u, 0u);
        xb[4 * j + 0] = v.x; xb[4 * j + 1] = v.y; xb[4 * j + 2] = v.z; xb[4 * j + 3] = v.w;
      }
    }
  };

  float ya[4] = {0.f, 0.f, 0.f, 0.f};      // rows ra, rb x accumulator columns 2t, 2t+1 of the current tile
  float yo[4] = {0.f, 0.f, 0.f, 0.f};      // I8: the outlier units' sums (columns = batch rows 2t, 2t+1)
  auto flush = [&](int tord) {
    float* dst = part + ((size_t)tord * kWarps + warp) * 16 * pc;
    if (2 * t < ncols) { dst[ra * pc + 2 * t] = ya[0]; dst[rb * pc + 2 * t] = ya[2]; }
    if (2 * t + 1 < ncols) { dst[ra * pc + 2 * t + 1] = ya[1]; dst[rb * pc + 2 * t + 1] = ya[3]; }
    ya[0] = ya[1] = ya[2] = ya[3] = 0.f;
    if (I8) {
      if (2 * t < m) { dst[ra * pc + ncols + 2 * t] = yo[0]; dst[rb * pc + ncols + 2 * t] = yo[2]; }
      if (2 * t + 1 < m) { dst[ra * pc + ncols + 2 * t + 1] = yo[1]; dst[rb * pc + ncols + 2 * t + 1] = yo[3]; }
      yo[0] = yo[1] = yo[2] = yo[3] = 0.f;
    }
  };

  // one 128-column int4 step: 8 words (4 of row ra, 4 of row rb) -> 8 mma in two chains.  xc[4j + w] is the
  // natural-order half2 (k = 8j + 2w, +1) of this lane's 32-column chunk, i.e. the k-pair that word w's j-th
  // half2 multiplies.  mma (j, cc) takes k-slots (2t, 2t+1) from word 2cc and (2t+8, 2t+9) from word 2cc+1, so
  // its B registers are the adjacent pair xc[4j + 2cc], xc[4j + 2cc + 1]; even j (low nibbles, 1024+q)
  // accumulate in `lo`, odd j (high nibbles, 1024+16q) in `hi`.
  // `sp`: shared address of the step's scale rows (16 scales | 16 scaled zeros) + 2 ra.
  auto int4_step = [&](const uint4& va, const uint4& vb, int s, uint32_t sp) {
    if constexpr (I8) {
      // int8 path: two AND masks per word (low nibbles: q, high nibbles: 16 q, both valid u8), 4 IMMA per step,
      // exact s32 accumulation; 16 lo + hi = 16 sum(q X); y += s * coef/16 * (16 lo + hi) + sz * sum(x)
      const uint32_t xa = xdig_lane + (uint32_t)s * xdig_step;
      const uint4 xe = lds_v4(xa), xq = lds_v4(xa + 16);
      float4 sm;
      asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(sm.x), "=f"(sm.y), "=f"(sm.z), "=f"(sm.w)
                   : "r"(sums_lane + (uint32_t)(s * 64)) : "memory");
      constexpr uint32_t kLoM = 0x0f0f0f0fu, kHiM = 0xf0f0f0f0u;
      int lo[4] = {0, 0, 0, 0}, hi[4] = {0, 0, 0, 0};
      auto imma = [](int (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
        asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      };
      imma(lo, va.x & kLoM, vb.x & kLoM, va.y & kLoM, vb.y & kLoM, xe.x, xe.y);
      imma(hi, va.x & kHiM, vb.x & kHiM, va.y & kHiM, vb.y & kHiM, xq.x, xq.y);
      imma(lo, va.z & kLoM, vb.z & kLoM, va.w & kLoM, vb.w & kLoM, xe.z, xe.w);
      imma(hi, va.z & kHiM, vb.z & kHiM, va.w & kHiM, vb.w & kHiM, xq.z, xq.w);
      const float sa = lds_h(sp), sbv = lds_h(sp + 8), za = lds_h(sp + 32), zb = lds_h(sp + 40);
      const float f0 = (float)(lo[0] * 16 + hi[0]), f1 = (float)(lo[1] * 16 + hi[1]);
      const float f2 = (float)(lo[2] * 16 + hi[2]), f3 = (float)(lo[3] * 16 + hi[3]);
      ya[0] = fmaf(sa * sm.x, f0, fmaf(za, sm.z, ya[0]));
      ya[1] = fmaf(sa * sm.y, f1, fmaf(za, sm.w, ya[1]));
      ya[2] = fmaf(sbv * sm.x, f2, fmaf(zb, sm.z, ya[2]));
      ya[3] = fmaf(sbv * sm.y, f3, fmaf(zb, sm.w, ya[3]));
      return;
    }
    uint32_t xc[16];
    load_x(xc, s);
    float4 sm;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(sm.x), "=f"(sm.y), "=f"(sm.z), "=f"(sm.w)
                 : "r"(sums_lane + (uint32_t)(s * 64)) : "memory");
    const uint32_t wa_[4] = {va.x, va.y, va.z, va.w};
    const uint32_t wb_[4] = {vb.x, vb.y, vb.z, vb.w};
    float lo[4], hi[4];
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
          mma_m16n8k16_zero(acc, a0[j], b0[j], a1[j], b1[j], xc[4 * j + 2 * cc], xc[4 * j + 2 * cc + 1]);
        else
          mma_m16n8k16_f16f32(acc, a0[j], b0[j], a1[j], b1[j], xc[4 * j + 2 * cc], xc[4 * j + 2 * cc + 1]);
      }
    }
    // y += s * (lo + hi/16 - c) + z * X
    const float sa = lds_h(sp), sbv = lds_h(sp + 8), za = lds_h(sp + 32), zb = lds_h(sp + 40);
    ya[0] = fmaf(sa, fmaf(hi[0], 0.0625f, lo[0]) - sm.z, fmaf(za, sm.x, ya[0]));
    ya[1] = fmaf(sa, fmaf(hi[1], 0.0625f, lo[1]) - sm.w, fmaf(za, sm.y, ya[1]));
    ya[2] = fmaf(sbv, fmaf(hi[2], 0.0625f, lo[2]) - sm.z, fmaf(zb, sm.x, ya[2]));
    ya[3] = fmaf(sbv, fmaf(hi[3], 0.0625f, lo[3]) - sm.w, fmaf(zb, sm.y, ya[3]));
  };

  // the consume / refill protocol of one unit: wait for the oldest group, read the slot (the caller's `body`
  // uses w0..w3 and the scale rows), then refill the slot that was read ONE unit earlier.  A lane only reads its
  // own 16-byte chunks (no hazard); the scale rows are shared by the warp: the __syncwarp after the wait orders
  // the previous unit's reads of them before this unit's refill, and this unit's copies before its reads.
  int ctord = u0 / upt, csu = u0 - ctord * upt;
  uint32_t slot = ring_u32, prev_slot = ring_u32 + (D - 1) * kSlotBytes;
  const uint32_t ring_end = ring_u32 + D * kSlotBytes;
  const int nfull = nchunks >> 3;            // units of a tile whose two steps are completely live
  bool first = true;
#pragma unroll 1
  for (int u = u0; u < u1; ++u) {
    const uint32_t mine = slot + lane * 16;
    cp_async_wait<D - 2>();
    __syncwarp();
    if (!first) {
      if (pu < u1) {
        issue(prev_slot, psu);
        ++pu;
        if (++psu == upt) { psu = 0; ++ptord; if (pu < u1) pf_set_tile(ptord); }
      }
      cp_async_commit();
    }
    first = false;
    const uint4 w0 = lds_v4(mine), w1 = lds_v4(mine + 512), w2 = lds_v4(mine + 1024), w3 = lds_v4(mine + 1536);
    const uint32_t sp = slot + 2048 + 2 * ra;

    if (csu < nfull) {
      // ---- two complete 128-column int4 steps, 8 mma each (the common case: no branch inside) ----
      int4_step(w0, w1, 2 * csu, sp);
      int4_step(w2, w3, 2 * csu + 1, sp + 64);
    } else if (csu < nku) {
      int4_step(w0, w1, 2 * csu, sp);
      if (2 * csu + 1 < nsteps) int4_step(w2, w3, 2 * csu + 1, sp + 64);
    } else {
      // ---- outlier unit: up to 64 fp16 columns of the tile's 16 rows, 4 mma ----
      const int c0 = 64 * (csu - nku);
      if (inter) {
        const uint4 wv[4] = {w0, w1, w2, w3};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (c0 + 16 * j < r) {                  // uniform
            const uint4 u4 = wv[j];               // {ra, rb} at columns c0 + 16 j + 4t .. + 3
            uint2 xv;
            if (XS) xv = lds_v2(xo_lane + (uint32_t)((c0 + 16 * j + 4 * t) * 2));
            else xv = ldx_v2(xg_lane + K - r + c0 + 16 * j + 4 * t, xcoh);
            mma_m16n8k16_f16f32(I8 ? yo : ya, prmt(u4.x, u4.y, 0x5410), prmt(u4.x, u4.y, 0x7632), prmt(u4.z, u4.w, 0x5410),
                                prmt(u4.z, u4.w, 0x7632), xv.x, xv.y);
          }
        }
      } else {
#pragma unroll
        for (int si = 0; si < 2; ++si) {
          if (c0 + 32 * si < r) {                 // uniform
            const uint4 ua = si ? w2 : w0, ub = si ? w3 : w1;   // rows ra / rb, columns c0 + 32 si + 8t .. + 7
            uint4 xv;
            if (XS) xv = lds_v4(xo_lane + (uint32_t)((c0 + 32 * si + 8 * t) * 2));
            else xv = ldx_v4(xg_lane + K - r + c0 + 32 * si + 8 * t, xcoh);
            mma_m16n8k16_f16f32(I8 ? yo : ya, ua.x, ub.x, ua.y, ub.y, xv.x, xv.y);
            mma_m16n8k16_f16f32(I8 ? yo : ya, ua.z, ub.z, ua.w, ub.w, xv.z, xv.w);
          }
        }
      }
    }
    if (++csu == upt) { flush(ctord); csu = 0; ++ctord; }
    prev_slot = slot;
    slot += kSlotBytes;
    if (slot == ring_end) slot = ring_u32;
  }
  if (csu != 0) flush(ctord);
  cp_async_wait<0>();
  stamp(p, 4);

  // ---- add the warps' slices in a fixed order, round, store the rows this CTA owns -------------------------
  __syncthreads();
  for (int i = tid; i < ntiles * 16 * m; i += kThreads) {
    const int b = i % m, rr = (i / m) & 15, tord = i / (16 * m);
    int sg = 0;
#pragma unroll
    for (int j = 1; j < QEFT_GEMV_MAX_PARTS; ++j)
      if (j < nseg && tord >= segs[j].cum) sg = j;
    const Seg S = segs[sg];
    const int row = 16 * (S.t0 + tord - S.cum) + rr;
    const int q = row >> 2;
    if (q >= S.pa && q < S.pb) {
      const GemvPart& P = p.part[S.part];
      const float* src = part + (size_t)tord * kWarps * 16 * pc + rr * pc;
      float acc = 0.f;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) {
        const float* sw = src + w * 16 * pc;
        if (I8) acc += ((sw[3 * b] + sw[3 * b + 1]) + sw[3 * b + 2]) + sw[ncols + b];
        else acc += sw[b];
      }
      if (P.bias) acc += __half2float(P.bias[row]);
      const __half h = __float2half_rn(acc);
      if (p.nranks == 0) {
        P.y[(size_t)b * P.N + row] = h;
      } else {
        for (int pr = 0; pr < p.nranks; ++pr) p.y_peer[pr][S.part][(size_t)b * p.y_ld + row] = h;   // NVLink stores
      }
    }
  }
  if (p.nranks > 0) {
    // publish: the CTA's peer stores happen-before the barrier; thread 0's system-scope fence is cumulative over them
    // and orders them before its relaxed signals (fence-based release).  Every CTA signals every rank directly with
    // its share of kArrivalsPerLaunch (the shares of a launch add up to exactly that), so a launch costs each CTA ONE
    // fence round trip: no CTA counting, no second fence in a last CTA, no release (= another fence) per signal.
    __syncthreads();
    if (tid == 0) {
      fence_acq_rel_sys();
      const unsigned inc = arrival_share(blockIdx.x, gridDim.x);
      for (int pr = 0; pr < p.nranks; ++pr)
        asm volatile("red.relaxed.sys.global.add.u32 [%0], %1;" ::"l"(p.done_peer[pr]), "r"(inc) : "memory");
    }
  }
  stamp(p, 5);
  if (p.stamps && threadIdx.x == 0) {          // debug: the last CTA's finish time
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    atomicMax(p.stamps + 6, t);
  }
}

// Consumer side of the arrival-counter protocol for code that is not one of the chain's kernels (a copy to the host,
// any other stream work): returns once every rank's slice of the launch has landed in this rank's gathered buffer.
__global__ void gather_wait_kernel(const uint32_t* flag, const uint32_t* epoch, int nranks) {
  if (threadIdx.x == 0) {
    const uint32_t want = *reinterpret_cast<const volatile uint32_t*>(epoch) * (uint32_t)nranks * kArrivalsPerLaunch;
    uint32_t got;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(got) : "l"(flag) : "memory");
    } while ((int)(got - want) < 0);
  }
}

// ----------------------------------------------------------------------------------------------------
constexpr size_t kStageXMaxBytes = 64 * 1024;   // stage x in shared memory when it is at most this big
constexpr size_t kSmemPerSm = 227 * 1024;
constexpr size_t kSmemCtaOverhead = 1024 + 256; // per-CTA reservation + static shared
constexpr size_t kPartMaxBytes = 24 * 1024;     // partial-sum slices per CTA; larger launches are split by rows

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

template <int D, bool XS, bool I8>
static int launch_one(const GemvParams& prm, int grid, size_t smem, unsigned flags, cudaStream_t stream) {
  auto kern = gemv_w4_kernel<D, XS, I8>;
  static bool attr_set[64] = {};    // per instantiation and device
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSmemPerSm - 1024));
    if (e != cudaSuccess) return (int)e;
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kThreads);
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

// debug timeline: QEFT_GEMV_STAMPS=1 allocates [4096][8] u64; launch i writes slot i % 4096 (read with qeft_gemv_debug_stamps)
static unsigned long long* g_stamps = nullptr;
static unsigned long long g_stamp_launch = 0;

template <bool XS, bool I8>
static int launch_gemv(GemvParams& prm, int total_q, unsigned flags, cudaStream_t stream) {
  static const int stamps_env = env_int("QEFT_GEMV_STAMPS", 0);
  if (stamps_env && !g_stamps) {
    const size_t bytes = (4096 * 8 + 64 * 512 * 4) * sizeof(unsigned long long);
    if (cudaMalloc(&g_stamps, bytes) != cudaSuccess) g_stamps = nullptr;
    else cudaMemset(g_stamps, 0, bytes);
  }
  prm.cta_stamps = g_stamps ? g_stamps + 4096 * 8 + (g_stamp_launch % 64) * 512 * 4 : nullptr;
  prm.stamps = g_stamps ? g_stamps + (g_stamp_launch++ % 4096) * 8 : nullptr;
  static const int depth_env = env_int("QEFT_GEMV_DEPTH", 0);
  static const int cps_env = env_int("QEFT_GEMV_CTAS_PER_SM", 0);
  const int m = prm.m;
  const size_t xbytes = I8 ? (size_t)prm.nsteps * 3 * m * 144 + sizeof(__half) * (size_t)m * prm.r + sizeof(int) * (size_t)prm.nsteps * m + 16
                           : (XS ? sizeof(__half) * (size_t)m * (size_t)(prm.xstride + prm.r) : 0);
  const size_t sums = sizeof(float) * 16 * (size_t)prm.nsteps;
  // the kernel is bound by instruction issue, not by bytes in flight: wide launches run two CTAs per SM (16 warps)
  // once every CTA still gets at least 8 qweight rows; narrow ones keep one CTA per SM so that the next launch's
  // CTA can be co-resident under programmatic dependent launch
  const int cps_want = cps_env > 0 ? cps_env : (total_q >= 2 * 8 * num_sms() ? 2 : 1);
  // rows per launch: the partial-sum slices of a CTA must fit kPartMaxBytes
  const size_t tile_part = sizeof(float) * kWarps * 16 * (size_t)(I8 ? 4 * m : m);
  int tiles_fit = (int)(kPartMaxBytes / tile_part);
  if (tiles_fit < 3) tiles_fit = 3;
  const int q_per_cta_max = 4 * (tiles_fit - 2 * prm.nparts > 1 ? tiles_fit - 2 * prm.nparts : 1);
  const size_t half_sm = kSmemPerSm / 2 - kSmemCtaOverhead;
  for (int q_lo = 0; q_lo < total_q;) {
    int cps = cps_want, q_hi = total_q, grid = 0;
    size_t fixed = 0;
    for (;;) {
      // several CTAs per SM only if they are really co-resident (each within 1/cps of the shared memory with a ring
      // of at least 3 units per warp): otherwise the extra CTAs would run as a second wave
      const int sms = num_sms() * cps;
      q_hi = total_q;
      grid = sms < (q_hi - q_lo) ? sms : (q_hi - q_lo);
      if ((long)grid * q_per_cta_max < (long)(q_hi - q_lo)) q_hi = q_lo + grid * q_per_cta_max;
      grid = sms < (q_hi - q_lo) ? sms : (q_hi - q_lo);
      const int qmax = cdiv(q_hi - q_lo, grid);
      prm.max_tiles = cdiv(qmax, 4) + 2 * prm.nparts;
      fixed = (size_t)prm.max_tiles * tile_part + sums + xbytes + 128;
      if (cps > 1 && fixed + (size_t)kWarps * 3 * kSlotBytes > kSmemPerSm / cps - kSmemCtaOverhead) { --cps; continue; }
      break;
    }
    if (prm.nranks > 0 && (q_lo != 0 || q_hi != total_q)) return QEFT_E_UNSUPPORTED;   // the fused gather signals once per launch
    prm.q_lo = q_lo; prm.q_hi = q_hi;
    // one CTA per SM: half of the SM when the ring still gets at least 4 units per warp (so that the next launch's CTA
    // can be co-resident under programmatic dependent launch), else the whole SM; cps CTAs per SM: 1/cps each
    size_t budget = cps > 1 ? kSmemPerSm / cps - kSmemCtaOverhead : half_sm;
    if (cps == 1 && fixed + (size_t)kWarps * 4 * kSlotBytes > budget) budget = kSmemPerSm - kSmemCtaOverhead;
    if (fixed + (size_t)kWarps * 2 * kSlotBytes > budget) return QEFT_E_UNSUPPORTED;
    int depth = (int)((budget - fixed) / ((size_t)kWarps * kSlotBytes));
    if (depth_env > 0 && depth > depth_env) depth = depth_env;
    int st;
#define QEFT_GEMV_LAUNCH(DD) st = launch_one<DD, XS, I8>(prm, grid, fixed + (size_t)kWarps * DD * kSlotBytes, flags, stream)
    if (depth >= 10) QEFT_GEMV_LAUNCH(10);
    else if (depth >= 8) QEFT_GEMV_LAUNCH(8);
    else if (depth >= 6) QEFT_GEMV_LAUNCH(6);
    else if (depth >= 5) QEFT_GEMV_LAUNCH(5);
    else if (depth >= 4) QEFT_GEMV_LAUNCH(4);
    else if (depth >= 3) QEFT_GEMV_LAUNCH(3);
    else QEFT_GEMV_LAUNCH(2);
#undef QEFT_GEMV_LAUNCH
    if (st != QEFT_OK) return st;
    q_lo = q_hi;
  }
  return QEFT_OK;
}

}  // namespace qeft

using namespace qeft;

static int gemv_entry(const void* x, const qeft_gemv_part_t* parts, int nparts, int ow_layout, const int32_t* x_gather,
                      int m, int K, int r, int G, unsigned flags, const qeft_gather_t* gat, qeft_stream_t stream) {
  if (!x || !parts) return QEFT_E_NULL;
  if (nparts < 1 || nparts > QEFT_GEMV_MAX_PARTS) return QEFT_E_SHAPE;
  if (m < 1 || m > 8) return QEFT_E_BATCH;
  if (G == -1) G = K;
  if (K <= 0 || K % 64 != 0 || G <= 0 || K % G != 0 || (G % 128 != 0 && G != K)) return QEFT_E_SHAPE;
  if (r < 0 || r % 32 != 0 || r >= K) return QEFT_E_SHAPE;
  if (r > 0 && ow_layout != QEFT_OW_PLAIN && ow_layout != QEFT_OW_INTERLEAVED) return QEFT_E_DTYPE;
  if (r == 0) ow_layout = QEFT_OW_NONE;
  if (!check_align16(x) || (x_gather && !check_align16(x_gather))) return QEFT_E_ALIGN;
  GemvParams prm = {};
  int total_q = 0;
  for (int i = 0; i < nparts; ++i) {
    const qeft_gemv_part_t& q = parts[i];
    if (!q.qweight || !q.scales || !q.scaled_zeros || (!q.y && !gat)) return QEFT_E_NULL;
    if (r > 0 && !q.oweight) return QEFT_E_NULL;
    if (q.N <= 0 || q.N % 8 != 0) return QEFT_E_SHAPE;
    if (!check_align16(q.qweight) || !check_align16(q.scales) || !check_align16(q.scaled_zeros) ||
        (r > 0 && !check_align16(q.oweight)))
      return QEFT_E_ALIGN;
    GemvPart& d = prm.part[i];
    d.qw = static_cast<const uint8_t*>(q.qweight);
    d.scales = static_cast<const __half*>(q.scales);
    d.szeros = static_cast<const __half*>(q.scaled_zeros);
    d.ow = static_cast<const __half*>(q.oweight);
    d.bias = static_cast<const __half*>(q.bias);
    d.y = static_cast<__half*>(q.y);
    d.N = q.N;
    d.q_begin = total_q;
    total_q += q.N / 4;
  }
  prm.nparts = nparts;
  prm.x = static_cast<const __half*>(x);
  prm.gather = x_gather;
  prm.m = m; prm.K = K; prm.r = r;
  prm.g128 = (G == K) ? 0 : G / 128;
  prm.ow_layout = ow_layout;
  prm.nsteps = cdiv(K - r, 128);
  prm.nchunks = (K - r) / 32;
  prm.nku = cdiv(prm.nsteps, 2);
  prm.nou = cdiv(r, 64);
  prm.xstride = 128 * prm.nsteps + 32;
  if (gat) {
    if (gat->nranks < 1 || gat->nranks > QEFT_MAX_RANKS || !gat->epoch) return QEFT_E_SHAPE;
    prm.nranks = gat->nranks;
    prm.y_ld = gat->y_ld;
    for (int pr = 0; pr < gat->nranks; ++pr) {
      if (!gat->done_peer[pr]) return QEFT_E_NULL;
      prm.done_peer[pr] = gat->done_peer[pr];
      for (int i = 0; i < nparts; ++i) {
        if (!gat->y_peer[pr][i]) return QEFT_E_NULL;
        prm.y_peer[pr][i] = static_cast<__half*>(gat->y_peer[pr][i]);
      }
    }
    prm.wait_flag = gat->wait_flag;
    static const int gridwait_env = getenv("QEFT_GATHER_GRIDWAIT") ? atoi(getenv("QEFT_GATHER_GRIDWAIT")) : 0;
    prm.flag_only = (gat->wait_flag != nullptr && (flags & QEFT_F_PDL) && !gridwait_env) ? 1 : 0;
    prm.epoch = gat->epoch;
  }   // the int4 steps' columns (dead columns of the last step staged as zeros)
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool stage = x_gather != nullptr || (size_t)m * (size_t)(K + 136) * 2 <= kStageXMaxBytes;
  // batch 1-2 with staged x: the int8 tensor path (half the MMAs, 16 instead of 40 unpack instructions per step)
  static const int i8_env = env_int("QEFT_GEMV_I8", 1);
  const size_t i8_bytes = (size_t)prm.nsteps * 3 * m * 144 + (size_t)m * r * 2;
  if (stage && m <= 2 && i8_env != 0 && i8_bytes <= kStageXMaxBytes) return launch_gemv<true, true>(prm, total_q, flags, st);
  return stage ? launch_gemv<true, false>(prm, total_q, flags, st) : launch_gemv<false, false>(prm, total_q, flags, st);
}

extern "C" int qeft_gemv_w4_multi(const void* x, const qeft_gemv_part_t* parts, int nparts, int ow_layout,
                                  const int32_t* x_gather, int m, int K, int r, int G, unsigned flags,
                                  qeft_stream_t stream) {
  return gemv_entry(x, parts, nparts, ow_layout, x_gather, m, K, r, G, flags, nullptr, stream);
}

extern "C" int qeft_gemv_w4_multi_gather(const void* x, const qeft_gemv_part_t* parts, int nparts, int ow_layout,
                                         const int32_t* x_gather, int m, int K, int r, int G, unsigned flags,
                                         const qeft_gather_t* gather, qeft_stream_t stream) {
  if (!gather) return QEFT_E_NULL;
  return gemv_entry(x, parts, nparts, ow_layout, x_gather, m, K, r, G, flags, gather, stream);
}

extern "C" int qeft_gather_wait(const uint32_t* arrival_counter, const uint32_t* epoch, int nranks, qeft_stream_t stream) {
  if (!arrival_counter || !epoch) return QEFT_E_NULL;
  if (nranks < 1 || nranks > QEFT_MAX_RANKS) return QEFT_E_SHAPE;
  gather_wait_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(arrival_counter, epoch, nranks);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  count_launch();
  return QEFT_OK;
}

extern "C" __attribute__((visibility("default"))) int qeft_gemv_debug_stamps(unsigned long long* host_out, int nslots) {
  if (!g_stamps || !host_out || nslots <= 0 || nslots > 4096) return QEFT_E_NULL;
  cudaError_t e = cudaMemcpy(host_out, g_stamps, (size_t)nslots * 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  return e == cudaSuccess ? QEFT_OK : (int)e;
}

// per-CTA stamps of launch slot `slot` (launch number % 64): host_out[512][4]
extern "C" __attribute__((visibility("default"))) int qeft_gemv_debug_cta_stamps(unsigned long long* host_out, int slot) {
  if (!g_stamps || !host_out || slot < 0 || slot >= 64) return QEFT_E_NULL;
  cudaError_t e = cudaMemcpy(host_out, g_stamps + 4096 * 8 + (size_t)slot * 512 * 4, 512 * 4 * sizeof(unsigned long long),
                             cudaMemcpyDeviceToHost);
  return e == cudaSuccess ? QEFT_OK : (int)e;
}

extern "C" int qeft_gemv_w4(const void* x, const void* qweight, const void* scales, const void* scaled_zeros,
                            const void* oweight, int ow_layout, const void* bias, const int32_t* x_gather,
                            void* y, int m, int N, int K, int r, int G, unsigned flags, qeft_stream_t stream) {
  qeft_gemv_part_t part = {qweight, scales, scaled_zeros, oweight, bias, y, N};
  return qeft_gemv_w4_multi(x, &part, 1, ow_layout, x_gather, m, K, r, G, flags, stream);
}
