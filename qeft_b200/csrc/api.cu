// Library-level entry points of the C ABI (include/qeft_b200.h).
#include "common.cuh"

extern "C" int qeft_abi_version(void) { return QEFT_B200_ABI_VERSION; }

extern "C" const char* qeft_build_info(void) {
  return "qeft_b200 sm_100a nvcc " __DATE__ " " __TIME__;
}

extern "C" uint64_t qeft_launch_count(void) { return (uint64_t)qeft::g_launch_count; }

extern "C" const char* qeft_status_string(int status) {
  switch (status) {
    case QEFT_OK: return "ok";
    case QEFT_E_NULL: return "required pointer is NULL";
    case QEFT_E_SHAPE: return "shape violates the packed layout's divisibility rules";
    case QEFT_E_BATCH: return "Unsupported batch size for gemv kernel.";
    case QEFT_E_DTYPE: return "unknown dtype / layout enum";
    case QEFT_E_ALIGN: return "pointer is not 16-byte aligned";
    case QEFT_E_UNSUPPORTED: return "configuration not supported by this build";
    default: return status > 0 ? cudaGetErrorString((cudaError_t)status) : "unknown status";
  }
}
