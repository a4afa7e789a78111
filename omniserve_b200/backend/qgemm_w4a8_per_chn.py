"""omniserve_backend.qgemm_w4a8_per_chn (reference: kernels/csrc/qgemm/w4a8_per_chn/{pybind.cpp,gemm_cuda.cu:601-657})."""
from .. import _lib as L


def gemm_forward_cuda(in_feats, kernel, wscales, ascales, w_szs, a_ssums, out_feats):
    """out_feats[M,N] (fp16, written in place; may be a row-slice view) =
    (in_feats[M,K] . W^T) * wscales[n] * ascales[m] - w_szs[n] * a_ssums[m]."""
    L.require_cuda(in_feats, kernel, wscales, ascales, w_szs, a_ssums, out_feats)
    M, K = in_feats.shape[0], in_feats.shape[1]
    N = out_feats.shape[-1]
    if in_feats.stride(-1) != 1 or (M > 1 and in_feats.stride(0) != K) or out_feats.stride(-1) != 1:
        raise RuntimeError("gemm_forward_cuda: in_feats must be row-contiguous and out_feats unit-stride in N")
    ldc = out_feats.stride(-2) if out_feats.dim() >= 2 else N
    L.check(
        L.lib().ob_w4a8_gemm_per_chn(
            L.ptr(in_feats), L.ptr(kernel), L.ptr(wscales), L.ptr(ascales), L.ptr(w_szs), L.ptr(a_ssums),
            L.ptr(out_feats), M, N, K, ldc, L.stream()),
        "qgemm_w4a8_per_chn.gemm_forward_cuda")
