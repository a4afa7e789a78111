"""omniserve_backend.qgemm_w8a8 (reference: kernels/csrc/qgemm/w8a8/{pybind.cpp,w8a8_gemm_cuda.cu:537-600}) -- the GEMM of
LServe's released W8A8KV8 setting (scripts/lserve_benchmark/launch.sh:6-7)."""
from .. import _lib as L


def w8a8_gemm_forward_cuda(in_feats, kernel, wscales, ascales, out_feats):
    """out_feats[M,N] (fp16, written in place; may be a row-slice view) = (in_feats[M,K] . kernel[N,K]^T) * wscales[n] *
    ascales[m]  (INT8 x INT8 -> INT32, epilogue of w8a8_gemm_cuda.cu:515-530)."""
    L.require_cuda(in_feats, kernel, wscales, ascales, out_feats)
    M, K = in_feats.shape[0], in_feats.shape[1]
    N = out_feats.shape[-1]
    if in_feats.stride(-1) != 1 or (M > 1 and in_feats.stride(0) != K) or out_feats.stride(-1) != 1:
        raise RuntimeError("w8a8_gemm_forward_cuda: in_feats must be row-contiguous and out_feats unit-stride in N")
    if not kernel.is_contiguous() or tuple(kernel.shape) != (N, K):
        raise RuntimeError("w8a8_gemm_forward_cuda: kernel must be a contiguous int8 [N, K] tensor")
    ldc = out_feats.stride(-2) if out_feats.dim() >= 2 else N
    L.check(L.lib().ob_w8a8_gemm(L.ptr(in_feats), L.ptr(kernel), L.ptr(wscales), L.ptr(ascales), L.ptr(out_feats), M, N, K, ldc,
                                 L.stream()), "qgemm_w8a8.w8a8_gemm_forward_cuda")
