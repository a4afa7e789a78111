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


def gemm_forward_cuda_add_norm_quant(in_feats, kernel, wscales, ascales, w_szs, a_ssums, out_feats, hidden_in, hidden_out,
                                     norm_weight, norm_out, norm_sum, norm_scale, epsilon):
    """Extension (decode-sized M): gemm_forward_cuda followed IN THE SAME LAUNCH by
    hidden_out = hidden_in + out_feats and rms_norm_general(_fuse_sum)(norm_out, hidden_out, norm_weight, norm_sum,
    norm_scale, epsilon, True) -- the `residual + proj -> layernorm -> int8` step of llama_w4a8_unpad.py:425-431.
    norm_sum may be None; ascales / a_ssums may alias norm_scale / norm_sum.  Returns False (nothing launched) when the
    shape is outside the fused path (M > 256 or N > 4096): call the three ops separately then."""
    L.require_cuda(in_feats, kernel, wscales, ascales, w_szs, a_ssums, out_feats, hidden_in, hidden_out, norm_weight, norm_out,
                   norm_sum, norm_scale)
    M, K = in_feats.shape[0], in_feats.shape[1]
    N = out_feats.shape[-1]
    if M > 256 or N > 4096 or not (hidden_in.is_contiguous() and hidden_out.is_contiguous() and norm_out.is_contiguous()):
        return False
    ldc = out_feats.stride(-2) if out_feats.dim() >= 2 else N
    L.check(
        L.lib().ob_w4a8_gemm_add_norm_quant(
            0, L.ptr(in_feats), L.ptr(kernel), 0, 0, L.ptr(wscales), L.ptr(ascales), L.ptr(w_szs), L.ptr(a_ssums),
            L.ptr(out_feats), M, N, K, ldc, L.ptr(hidden_in), L.ptr(hidden_out), L.ptr(norm_weight), L.ptr(norm_out),
            L.ptr(norm_sum), L.ptr(norm_scale), float(epsilon), L.stream()),
        "qgemm_w4a8_per_chn.gemm_forward_cuda_add_norm_quant")
    return True
