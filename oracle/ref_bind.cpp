// Binding for the UNMODIFIED reference kernels, test infrastructure only (see oracle/build_ref.py).
//
// The reference's own module file (qeft/kernel/qeft_cuda.cpp:10-27) also binds its attention and layernorm kernels,
// which are outside this path; this file declares the three functions of the path exactly as the reference's headers
// do (quantization_new/gemv/gemv_cuda.h, quantization_new/gemm/gemm_cuda.h) and exposes them under the same names.
// The kernels themselves are compiled from the sources where they lie under /root/reference; nothing is copied.
#include <torch/extension.h>

torch::Tensor gemm_4bit(torch::Tensor in_feats, torch::Tensor kernel, torch::Tensor scales, torch::Tensor zeros);
torch::Tensor gemv_4bit(torch::Tensor in_feats, torch::Tensor kernel, torch::Tensor scaling_factors, torch::Tensor zeros,
                        int m, int n, int k, int group_size);
torch::Tensor gemv_4bit_qeft(torch::Tensor in_feats, torch::Tensor kernel, torch::Tensor scaling_factors,
                             torch::Tensor zeros, torch::Tensor oweight, int m, int n, int k, int group_size);

PYBIND11_MODULE(qeft_cuda_ref, m) {
  m.def("gemm_4bit", &gemm_4bit, "reference gemm/gemm_cuda.cu");
  m.def("gemv_4bit", &gemv_4bit, "reference gemv/gemv_cuda.cu");
  m.def("gemv_4bit_qeft", &gemv_4bit_qeft, "reference gemv/gemv_cuda_qeft.cu");
}
