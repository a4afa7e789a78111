"""omniserve_backend.layernorm_ops (reference: kernels/csrc/layernorm.cpp:52-76, layernorm_kernels.cu:409-513)."""
from .. import _lib as L


def _rows(t):
    h = t.shape[-1]
    return t.numel() // h, h


def rms_norm(out, input, weight, epsilon, use_quant=False):
    if use_quant:
        raise NotImplementedError("rms_norm(use_quant=True): static int8 path is not used by the W4A8 models")
    L.require_cuda(out, input, weight)
    L.require_contiguous(out, input, weight)
    T, H = _rows(input)
    L.check(L.lib().ob_rms_norm(L.ptr(out), L.ptr(input), L.ptr(weight), float(epsilon), T, H, L.stream()), "rms_norm")


def rms_norm_general(out, input, weight, scaling, epsilon, use_per_token_quant=False):
    if not use_per_token_quant:
        raise NotImplementedError("rms_norm_general: per-tensor scaling is not on the W4A8 path")
    L.require_cuda(out, input, weight, scaling)
    L.require_contiguous(out, input, weight, scaling)
    T, H = _rows(input)
    L.check(
        L.lib().ob_rms_norm_general(L.ptr(out), L.ptr(input), L.ptr(weight), L.ptr(scaling), float(epsilon), T, H,
                                    L.stream()),
        "rms_norm_general")


def rms_norm_general_fuse_sum(out, input, weight, input_sum, scaling, epsilon, use_per_token_quant=False):
    if not use_per_token_quant:
        raise NotImplementedError("rms_norm_general_fuse_sum: per-tensor branch asserts false in the reference too")
    L.require_cuda(out, input, weight, input_sum, scaling)
    L.require_contiguous(out, input, weight, input_sum, scaling)
    T, H = _rows(input)
    L.check(
        L.lib().ob_rms_norm_general_fuse_sum(L.ptr(out), L.ptr(input), L.ptr(weight), L.ptr(input_sum),
                                             L.ptr(scaling), float(epsilon), T, H, L.stream()),
        "rms_norm_general_fuse_sum")


def add_rms_norm_general(out, hidden_in, delta, hidden_out, weight, input_sum, scaling, epsilon):
    """Extension (not in the reference): hidden_out = hidden_in + delta (fp16, == torch.add), then
    rms_norm_general(_fuse_sum)(out, hidden_out, ...) with per-token quant; input_sum may be None."""
    L.require_cuda(out, hidden_in, delta, hidden_out, weight, input_sum, scaling)
    L.require_contiguous(out, hidden_in, delta, hidden_out, weight, input_sum, scaling)
    T, H = _rows(hidden_in)
    L.check(
        L.lib().ob_add_rms_norm_general(L.ptr(out), L.ptr(hidden_in), L.ptr(delta), L.ptr(hidden_out), L.ptr(weight),
                                        L.ptr(input_sum), L.ptr(scaling), float(epsilon), T, H, L.stream()),
        "add_rms_norm_general")


def add_rms_norm(out, hidden_in, delta, weight, epsilon):
    """Extension: rms_norm(out, hidden_in + delta, weight, eps) with the add fused (fp16 out)."""
    L.require_cuda(out, hidden_in, delta, weight)
    L.require_contiguous(out, hidden_in, delta, weight)
    T, H = _rows(hidden_in)
    L.check(L.lib().ob_add_rms_norm(L.ptr(out), L.ptr(hidden_in), L.ptr(delta), L.ptr(weight), float(epsilon), T, H,
                                    L.stream()), "add_rms_norm")


def peer_add_rms_norm_general(out, hidden_in, peer, hidden_out, weight, input_sum, scaling, epsilon):
    """Extension for tensor parallelism: like add_rms_norm_general, but `delta` is the SUM over all ranks of the partial
    results each rank's row-parallel GEMM left in `peer`'s symmetric buffer -- the all-reduce is done by this kernel over
    NVLink peer memory (omniserve_b200/peer.py:PeerBuffer).  Collective: every rank must call it with the same shape."""
    L.require_cuda(out, hidden_in, hidden_out, weight, input_sum, scaling)
    L.require_contiguous(out, hidden_in, hidden_out, weight, input_sum, scaling)
    T, H = _rows(hidden_in)
    L.check(
        L.lib().ob_peer_add_rms_norm_general(L.ptr(out), L.ptr(hidden_in), peer.ctx_ref(), L.ptr(hidden_out), L.ptr(weight),
                                             L.ptr(input_sum), L.ptr(scaling), float(epsilon), T, H, L.stream()),
        "peer_add_rms_norm_general")


def peer_add_rms_norm(out, hidden_in, peer, weight, epsilon):
    """Extension: rms_norm(out, hidden_in + all_reduce(partials in peer memory), weight, eps), fp16 out."""
    L.require_cuda(out, hidden_in, weight)
    L.require_contiguous(out, hidden_in, weight)
    T, H = _rows(hidden_in)
    L.check(L.lib().ob_peer_add_rms_norm(L.ptr(out), L.ptr(hidden_in), peer.ctx_ref(), L.ptr(weight), float(epsilon), T, H,
                                         L.stream()), "peer_add_rms_norm")


def invoke_dequant_add_residual_rms_norm_quant(*a, **k):
    raise NotImplementedError("legacy W8A8 op, not on the W4A8KV4 path")
