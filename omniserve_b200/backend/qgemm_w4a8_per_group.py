"""omniserve_backend.qgemm_w4a8_per_group (reference: kernels/csrc/qgemm/w4a8_per_group/gemm_cuda.cu:635-707)."""
from .. import _lib as L


def gemm_forward_cuda(in_feats, kernel, zeros, scales_i8, wscales, ascales, out_feats):
    """Two-level (g128) W4A8 GEMM; out_feats written in place."""
    L.require_cuda(in_feats, kernel, zeros, scales_i8, wscales, ascales, out_feats)
    M, K = in_feats.shape[0], in_feats.shape[1]
    N = out_feats.shape[-1]
    if in_feats.stride(-1) != 1 or (M > 1 and in_feats.stride(0) != K) or out_feats.stride(-1) != 1:
        raise RuntimeError("gemm_forward_cuda: in_feats must be row-contiguous and out_feats unit-stride in N")
    ldc = out_feats.stride(-2) if out_feats.dim() >= 2 else N
    L.check(
        L.lib().ob_w4a8_gemm_per_group(
            L.ptr(in_feats), L.ptr(kernel), L.ptr(zeros), L.ptr(scales_i8), L.ptr(wscales), L.ptr(ascales),
            L.ptr(out_feats), M, N, K, ldc, L.stream()),
        "qgemm_w4a8_per_group.gemm_forward_cuda")


def gemm_forward_cuda_add_norm_quant(in_feats, kernel, zeros, scales_i8, wscales, ascales, out_feats, hidden_in, hidden_out,
                                     norm_weight, norm_out, norm_scale, epsilon):
    """Extension: per-group GEMM + `hidden_in + out -> rms_norm_general -> int8` in one launch (see
    qgemm_w4a8_per_chn.gemm_forward_cuda_add_norm_quant).  Returns False when the shape is outside the fused path."""
    L.require_cuda(in_feats, kernel, zeros, scales_i8, wscales, ascales, out_feats, hidden_in, hidden_out, norm_weight, norm_out,
                   norm_scale)
    M, K = in_feats.shape[0], in_feats.shape[1]
    N = out_feats.shape[-1]
    if M > 256 or N > 4096 or not (hidden_in.is_contiguous() and hidden_out.is_contiguous() and norm_out.is_contiguous()):
        return False
    ldc = out_feats.stride(-2) if out_feats.dim() >= 2 else N
    L.check(
        L.lib().ob_w4a8_gemm_add_norm_quant(
            1, L.ptr(in_feats), L.ptr(kernel), L.ptr(zeros), L.ptr(scales_i8), L.ptr(wscales), L.ptr(ascales), 0, 0,
            L.ptr(out_feats), M, N, K, ldc, L.ptr(hidden_in), L.ptr(hidden_out), L.ptr(norm_weight), L.ptr(norm_out), 0,
            L.ptr(norm_scale), float(epsilon), L.stream()),
        "qgemm_w4a8_per_group.gemm_forward_cuda_add_norm_quant")
    return True
