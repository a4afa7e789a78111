// INTEGRATION.md section B made concrete: the pybind11 / torch-extension binding a maintainer of the reference would keep,
// with the body of each entry function replaced by a call into this repository's C ABI (include/omniserve_b200.h).
// Two modules' worth of functions are bound here with the reference's exact Python-visible signatures:
//   qgemm_w4a8_per_chn.gemm_forward_cuda   (kernels/csrc/qgemm/w4a8_per_chn/pybind.cpp, gemm_cuda.cu:601-657)
//   fused_kernels.invoke_quant_fuse_sum     (kernels/csrc/fused.cpp:52-76, fused_kernels.cu:255-271)
// Built in-tree by integration/build_stub.py (g++ against the torch headers, linked to libomniserve_b200.so) and exercised
// by tests/test_gpu_pybind_stub.py.  Nothing here is product code: it documents and proves the integration path.
#include <ATen/cuda/CUDAContext.h>
#include <torch/extension.h>

#include "omniserve_b200.h"

static void gemm_forward_cuda(torch::Tensor in_feats, torch::Tensor kernel, torch::Tensor wscales, torch::Tensor ascales,
                              torch::Tensor w_szs, torch::Tensor a_ssums, torch::Tensor out_feats) {
  const int M = in_feats.size(0), K = in_feats.size(1), N = out_feats.size(-1);
  const int rc = ob_w4a8_gemm_per_chn(in_feats.data_ptr<int8_t>(), kernel.data_ptr<int8_t>(), wscales.data_ptr(),
                                      ascales.data_ptr(), w_szs.data_ptr(), a_ssums.data_ptr(), out_feats.data_ptr(), M, N, K,
                                      (int)out_feats.stride(-2), at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(rc == 0, "ob_w4a8_gemm_per_chn: ", ob_error_string(rc));
}

static void invoke_quant_fuse_sum(torch::Tensor out, torch::Tensor input, torch::Tensor input_sum, torch::Tensor scale) {
  const int H = input.size(-1);
  const int T = input.numel() / H;
  const int rc = ob_invoke_quant_fuse_sum(out.data_ptr<int8_t>(), input.data_ptr(), input_sum.data_ptr(), scale.data_ptr(), T, H,
                                          at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(rc == 0, "ob_invoke_quant_fuse_sum: ", ob_error_string(rc));
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("gemm_forward_cuda", &gemm_forward_cuda, "W4A8 per-channel GEMM (B200 C ABI behind the reference's binding)");
  m.def("invoke_quant_fuse_sum", &invoke_quant_fuse_sum, "per-token INT8 quant + sum (B200 C ABI behind the reference's binding)");
}
