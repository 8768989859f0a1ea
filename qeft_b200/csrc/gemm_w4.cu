// placeholder until the tcgen05 GEMM lands (replaced in a later commit)
#include "common.cuh"
extern "C" int qeft_gemm_w4(const void*, const void*, const void*, const void*, const void*, const void*, void*, int, int,
                            int, int, int, int, unsigned, qeft_stream_t) { return QEFT_E_UNSUPPORTED; }
extern "C" int qeft_gemm_w4_dx(const void*, const void*, const void*, const void*, const void*, void*, int, int, int, int,
                               int, int, unsigned, qeft_stream_t) { return QEFT_E_UNSUPPORTED; }
extern "C" int qeft_dow(const void*, const void*, float*, int, int, int, int, int, int, unsigned, qeft_stream_t) {
  return QEFT_E_UNSUPPORTED;
}
