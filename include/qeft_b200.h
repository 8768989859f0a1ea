/*
 * qeft_b200.h -- C ABI of the B200-native packed QuantLinear kernels.
 *
 * This is the drop-in boundary for the reference's `qeft_cuda` extension
 * (qeft/kernel/qeft_cuda.cpp:10-27 in xvyaward/qeft).  Every entry point takes raw
 * DEVICE pointers, plain ints and a CUDA stream handle, never allocates, never
 * synchronises, and returns an int status:
 *      0   success
 *     <0   invalid argument (QEFT_E_*), nothing was launched
 *     >0   a cudaError_t from the launch
 * There is no CPU fallback: without a CUDA device every compute entry returns
 * a cudaError_t.
 *
 * Packed layout (bit-exact with qeft/qlinear.py:70-121,180-215 of the reference):
 *   qweight       int16 [N/4, K]   4 rows x 64 columns per 128-byte tile, AWQ-v2 nibble order
 *   scales        fp16  [K/G, N]
 *   scaled_zeros  fp16  [K/G, N]   = -(zero * scale);  w = fma(q, scale, scaled_zero)
 *   oweight       fp16  [N, r]     dense outlier ("weak") columns = input columns K-r .. K-1
 *   oweight_interleaved fp16 [N/2, 2r]  row-pair interleaved copy used by the reference GEMV
 *   bias          fp16  [N] or NULL
 * The last r int4 columns of qweight are dead (they carry the zero point) and are
 * never read by these kernels.
 */
#ifndef QEFT_B200_H_
#define QEFT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QEFT_B200_ABI_VERSION 1

#if defined(_WIN32)
#define QEFT_API
#else
#define QEFT_API __attribute__((visibility("default")))
#endif

/* status codes (negative = argument errors) */
#define QEFT_OK 0
#define QEFT_E_NULL (-1)     /* a required pointer is NULL */
#define QEFT_E_SHAPE (-2)    /* N, K, r, G or m violate the layout's divisibility rules */
#define QEFT_E_BATCH (-3)    /* GEMV batch outside 1..8 (reference: "Unsupported batch size for gemv kernel.") */
#define QEFT_E_DTYPE (-4)    /* unknown dtype / layout enum */
#define QEFT_E_ALIGN (-5)    /* pointer not 16-byte aligned */
#define QEFT_E_UNSUPPORTED (-6)

/* activation / output element type of the GEMM-class entries */
#define QEFT_DT_F16 0
#define QEFT_DT_BF16 1

/* layout of the outlier weights handed to the GEMV */
#define QEFT_OW_NONE 0         /* r == 0 (reference gemv_4bit) */
#define QEFT_OW_PLAIN 1        /* oweight [N, r] */
#define QEFT_OW_INTERLEAVED 2  /* oweight_interleaved [N/2, 2r] (reference gemv_4bit_qeft) */

/* launch flags */
#define QEFT_F_PDL 1u          /* launch with programmatic dependent launch: weight prefetch of this
                                  kernel overlaps the tail of the previous kernel in the stream */

typedef void* qeft_stream_t;   /* cudaStream_t */

QEFT_API int qeft_abi_version(void);
/* static string: build arch, compiler */
QEFT_API const char* qeft_build_info(void);
/* number of kernels this library has launched since load (all entry points); bench.py's gpu_launches */
QEFT_API uint64_t qeft_launch_count(void);
QEFT_API const char* qeft_status_string(int status);

/*
 * Decode GEMV:  y[m, N] = x[m, K] . Wdense^T (+ bias),  m in 1..8.
 * Replaces  gemv_4bit       (qeft/kernel/quantization_new/gemv/gemv_cuda.cu:358-437)   with ow_layout = NONE
 *      and  gemv_4bit_qeft  (qeft/kernel/quantization_new/gemv/gemv_cuda_qeft.cu:392-513) with INTERLEAVED.
 * The outlier columns REPLACE the int4 columns (gemv_cuda_qeft.cu:168-176).
 * x_gather (int32 [K] or NULL): when given, x[:, x_gather[k]] is read in place of x[:, k]; this fuses
 * the o_proj `index_select(x, -1, reorder_ids)` of qeft/qlinear.py:273-275.
 * Requires N % 8 == 0, K % 64 == 0, G % 128 == 0 (or G == K), K % G == 0, r % 32 == 0, r < K.
 */
QEFT_API int qeft_gemv_w4(const void* x, const void* qweight, const void* scales, const void* scaled_zeros,
                 const void* oweight, int ow_layout, const void* bias, const int32_t* x_gather,
                 void* y, int m, int N, int K, int r, int G, unsigned flags, qeft_stream_t stream);

/*
 * Several GEMVs that share one x in ONE launch (q/k/v, gate/up): part p writes y[p][m, N[p]].
 * No reference counterpart (the reference launches one kernel per projection, qlinear.py:253-263).
 */
typedef struct {
  const void* qweight;
  const void* scales;
  const void* scaled_zeros;
  const void* oweight;
  const void* bias;
  void* y;
  int N;
} qeft_gemv_part_t;

#define QEFT_GEMV_MAX_PARTS 4
QEFT_API int qeft_gemv_w4_multi(const void* x, const qeft_gemv_part_t* parts, int nparts, int ow_layout,
                       const int32_t* x_gather, int m, int K, int r, int G, unsigned flags,
                       qeft_stream_t stream);

/*
 * Column-sharded decode (no reference counterpart: the reference has no multi-GPU code): the same
 * multi-projection GEMV, but every rank computes its slice of the output features and the kernel's epilogue stores
 * that slice DIRECTLY into the gathered buffer of every rank (peer-mapped pointers over NVLink) -- an all-gather
 * fused into the GEMV, no collective call.  Completion is signalled per launch: every CTA, after ONE system-scope
 * fence that orders its peer stores, adds its share of QEFT_ARRIVALS_PER_LAUNCH (relaxed, system scope) to every
 * rank's arrival counter `done_peer[p]`; the shares of a launch add up to exactly QEFT_ARRIVALS_PER_LAUNCH whatever
 * its grid.  A dependent launch passes that counter as `wait_flag` and spins (acquire, system scope) until it reaches
 * `*epoch * nranks * QEFT_ARRIVALS_PER_LAUNCH` before it reads its input; with QEFT_F_PDL that counter is the ONLY
 * ordering between the two launches (no grid-completion wait), so the chain never pays a kernel boundary.
 * `epoch` is a device-resident step counter the caller increments once per decode step (graph-replay friendly).
 * x_gather, when given, must be 16-byte aligned.
 * A launch with `wait_flag` reads x through L2 (coherent loads), never through the read-only path.
 * Contract of QEFT_F_PDL for every entry point: only x (and y as a reused buffer) are ordered after the previous
 * kernel in the stream; the packed tensors (qweight, scales, scaled_zeros, oweight, bias) are prefetched BEFORE the
 * dependency wait and must not be written by the kernel immediately before the launch (launch without the flag
 * after a pack / dequant / cast that produces them).
 */
#define QEFT_ARRIVALS_PER_LAUNCH 65536u
#define QEFT_MAX_RANKS 8
typedef struct {
  int nranks;
  int y_ld;                                            /* elements between batch rows of the gathered buffers */
  void* y_peer[QEFT_MAX_RANKS][QEFT_GEMV_MAX_PARTS];   /* rank p's buffer for part i, offset to THIS rank's columns */
  uint32_t* done_peer[QEFT_MAX_RANKS];                 /* rank p's arrival counter of this launch */
  const uint32_t* wait_flag;                           /* local arrival counter of the launch depended on, or NULL */
  const uint32_t* epoch;
  /* qeft_gemm_w4_gather only, optional: the MULTICAST mapping (NVLS; torch symmetric memory's multicast_ptr) of the
   * gathered buffer of part i, offset to this rank's columns.  When set, every tile is stored once to this address
   * and the NVSwitch replicates it to all ranks (egress bytes / (nranks - 1)); y_peer is then unused. */
  void* y_mc[QEFT_GEMV_MAX_PARTS];
} qeft_gather_t;

QEFT_API int qeft_gemv_w4_multi_gather(const void* x, const qeft_gemv_part_t* parts, int nparts, int ow_layout,
                              const int32_t* x_gather, int m, int K, int r, int G, unsigned flags,
                              const qeft_gather_t* gather, qeft_stream_t stream);

/*
 * Blocks `stream` (a one-thread kernel) until the launch whose local arrival counter is `arrival_counter` has received
 * every rank's slice: for consumers of a gathered buffer that are not chain kernels (a copy to the host).
 */
QEFT_API int qeft_gather_wait(const uint32_t* arrival_counter, const uint32_t* epoch, int nranks, qeft_stream_t stream);

/*
 * Decode programs: a chain of dependent decode GEMVs (one decoder block, or a whole token) as ONE cooperative launch
 * of a persistent kernel (csrc/decode_w4.cu).  No reference counterpart as an entry point: it replaces the
 * reference's sequence of one `gemv_4bit_qeft` launch per projection (qeft/qlinear.py:251-263, driven per token by
 * qeft/main.py:356-366) plus, optionally, the elementwise glue between them.
 * A stage is what qeft_gemv_w4_multi computes (batch m <= 2, oweight in the PLAIN [N, r] layout), with
 *   norm_weight != NULL : x is first RMS-normalised like HF LlamaRMSNorm / the reference's FT layernorm
 *                         (qeft/kernel/layernorm/layernorm.cu:25-51): weight * (x * rsqrt(mean(x^2) + eps)).to(fp16)
 *   QEFT_EPI_SWIGLU     : parts = {gate, up} (equal N); writes y0 = silu(fp16(gate)) * fp16(up)  (parts[1].y unused)
 *   QEFT_EPI_RESIDUAL   : one part; writes y = residual + fp16(linear)
 * Stage i+1 may read what stage i wrote (a gpu-scope barrier separates consecutive stages of a launch).
 * qeft_decode_program_run launches stages [begin, end) on `stream`; the program owns only its descriptor copy and
 * its barrier words, never the tensors.  One run of a program at a time (runs on one stream are ordered).
 */
#define QEFT_EPI_NONE 0
#define QEFT_EPI_SWIGLU 1
#define QEFT_EPI_RESIDUAL 2
typedef struct {
  qeft_gemv_part_t parts[QEFT_GEMV_MAX_PARTS];
  int nparts;
  int K, r, G;
  const void* x;               /* fp16 [m, K] */
  const int32_t* x_gather;     /* int32 [K] or NULL (o_proj reorder, qeft/qlinear.py:273-275) */
  const void* norm_weight;     /* fp16 [K] or NULL */
  float norm_eps;
  int epilogue;                /* QEFT_EPI_* */
  const void* residual;        /* fp16 [m, N] (QEFT_EPI_RESIDUAL) */
} qeft_decode_stage_t;
typedef struct qeft_decode_program qeft_decode_program_t;
QEFT_API int qeft_decode_program_create(const qeft_decode_stage_t* stages, int nstages, int m,
                                        qeft_decode_program_t** out);
QEFT_API int qeft_decode_program_run(qeft_decode_program_t* prog, int stage_begin, int stage_end, unsigned flags,
                                     qeft_stream_t stream);
QEFT_API int qeft_decode_program_num_stages(const qeft_decode_program_t* prog);
/*
 * Column-sharded programs (batch 1).  Every rank builds the SAME program over its own row slab of every
 * projection (N = the slab's rows) and then declares which outputs are exchanged:
 *   qeft_decode_program_set_ranks: barrier_peer[p] = rank p's barrier word (one uint32 per rank, zero-initialised,
 *     peer-mapped: symmetric memory), used by the two rank barriers of every launch;
 *   qeft_decode_program_shard(stage, part, y_full_peer): the part's y is this rank's slice [rank * N, (rank + 1) * N) of
 *     a gathered row [nranks * N] that every rank holds; y_full_peer[p] = the base of rank p's copy (peer-mapped).  The
 *     kernel stores the slice into every rank's copy straight from its epilogue (NVLink stores), and a later stage whose
 *     x is the local copy (x == y_full_peer[rank], K == nranks * N) starts on the elements as they arrive: the all-gather
 *     is the data-flow protocol of the kernel (each fp16 element is its own arrival flag), there is no collective
 *     call, no fence and no counter per stage.  The copies must not be written by anything else.
 */
QEFT_API int qeft_decode_program_set_ranks(qeft_decode_program_t* prog, int nranks, int rank, void* const* barrier_peer);
QEFT_API int qeft_decode_program_shard(qeft_decode_program_t* prog, int stage, int part, void* const* y_full_peer);
QEFT_API int qeft_decode_program_destroy(qeft_decode_program_t* prog);

/*
 * Prefill / fine-tune GEMM:  y[M, N] = x[M, K] . Wdense^T (+ bias)  on tcgen05 tensor cores.
 * Replaces  gemm_4bit (qeft/kernel/quantization_new/gemm/gemm_cuda.cu:929-1033)  PLUS the separate
 * `y += F.linear(x[..., -r:], oweight)` and `y + bias` of qeft/qlinear.py:264-268 in one kernel.
 * oweight is the PLAIN [N, r] tensor (may be NULL when r == 0).  GEMV semantics for the outlier
 * columns (the dead int4 columns are skipped).  Requires N % 128 == 0, K % 64 == 0, r % 64 == 0, G % 64 == 0.
 * dtype: QEFT_DT_F16 (x, oweight, y fp16; w = fma.rn.f16(q, s, sz) like the reference) or QEFT_DT_BF16 (x, oweight,
 * y bf16; the int4 columns are dequantised in fp32 with one rounding to bf16).  scales, scaled_zeros and bias are
 * always the checkpoint's fp16.  The same holds for qeft_gemm_w4_dx and qeft_dow.
 */
QEFT_API int qeft_gemm_w4(const void* x, const void* qweight, const void* scales, const void* scaled_zeros,
                 const void* oweight, const void* bias, void* y, int M, int N, int K, int r, int G,
                 int dtype, unsigned flags, qeft_stream_t stream);

/*
 * Column-sharded prefill: the same GEMM on this rank's N output features (a row slab of the packed
 * layer), with the all-gather fused into the epilogue: every output tile is stored into the gathered [M, y_ld] buffer
 * of EVERY rank at this rank's column offset (gather->y_peer[p][0], peer-mapped pointers over NVLink), tile by tile
 * while other tiles are still computing.  Signalling as in qeft_gemv_w4_multi_gather: every CTA adds its share
 * of QEFT_ARRIVALS_PER_LAUNCH to gather->done_peer[p] of every rank; a launch whose x is a gathered buffer passes that
 * launch's local counter as gather->wait_flag (its TMA producer acquires it before the first load).
 */
QEFT_API int qeft_gemm_w4_gather(const void* x, const void* qweight, const void* scales, const void* scaled_zeros,
                        const void* oweight, const void* bias, int M, int N, int K, int r, int G, int dtype,
                        unsigned flags, const qeft_gather_t* gather, qeft_stream_t stream);

/*
 * Backward wrt the input:  dx[M, K] = dy[M, N] . Wdense   (same packed bytes, contraction over N).
 * The math BASELINE.json defines for QuantMatMulQEFT.backward (the reference's qlinear.py:28-44 is
 * not usable).
 * Launches that would leave SMs idle (few tiles, or a nearly empty last wave) are split along the contraction: fp32
 * partial tiles go through a workspace owned by the (device, stream) pair -- the same one as qeft_gemm_w4's split-K --
 * and the last CTA to arrive at a tile adds them in split order (results do not depend on scheduling; two launches are
 * bit-equal).  The first split launch on a stream allocates the workspace (cudaMalloc): run one before capturing that
 * stream into a CUDA graph.
 */
QEFT_API int qeft_gemm_w4_dx(const void* dy, const void* qweight, const void* scales, const void* scaled_zeros,
                    const void* oweight, void* dx, int M, int N, int K, int r, int G,
                    int dtype, unsigned flags, qeft_stream_t stream);

/*
 * Host-only: the launch plan qeft_gemm_w4_dx uses for [M, N, K] on a device with sm_count SMs (no GPU needed; tests and
 * capacity planning).  *splits = CTAs that share a split tile (1: no tile is split), *whole_tiles = tiles computed by one
 * CTA each (they come first in the grid), *ctas = the grid size.  Workspace of the launch: splits * M * K * 4 bytes when
 * splits > 1.
 */
QEFT_API int qeft_gemm_w4_dx_plan(int M, int N, int K, int sm_count, int* splits, int* whole_tiles, int* ctas);

/*
 * Gradient of the trainable outlier columns:  dow[N, r] (fp32) (+)= dy[M, N]^T . x[M, K-r:K].
 * accumulate != 0 adds into dow (gradient accumulation into the fp32 master grad).
 */
QEFT_API int qeft_dow(const void* dy, const void* x, float* dow, int M, int N, int K, int r,
             int dtype, int accumulate, unsigned flags, qeft_stream_t stream);

/*
 * Device packers (bit-exact with qeft/qlinear.py:70-121).
 *   qeft_pack_w4     : intweight int32 [N, K] (values 0..15; not clamped, like the reference) -> qweight int16 [N/4, K]
 *   qeft_unpack_w4   : the inverse, int32 [N, K]
 *   qeft_dequant_w4  : dense fp16 [N, K] = fma(q, s, sz); when oweight != NULL the last r columns are oweight
 *   qeft_interleave_oweight : oweight fp16 [N, r] -> oweight_interleaved [N/2, 2r] (pack_oweight); used to
 *                      refresh the GEMV copy after a fine-tuning step changed oweight
 */
QEFT_API int qeft_pack_w4(const int32_t* intweight, void* qweight, int N, int K, qeft_stream_t stream);
QEFT_API int qeft_unpack_w4(const void* qweight, int32_t* intweight, int N, int K, qeft_stream_t stream);
QEFT_API int qeft_dequant_w4(const void* qweight, const void* scales, const void* scaled_zeros, const void* oweight,
                    void* w_dense, int N, int K, int r, int G, int dtype, qeft_stream_t stream);
QEFT_API int qeft_interleave_oweight(const void* oweight, void* oweight_interleaved, int N, int r, int src_fp32,
                            qeft_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* QEFT_B200_H_ */
