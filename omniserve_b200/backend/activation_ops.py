"""omniserve_backend.activation_ops (reference: kernels/csrc/activation.cpp, activation_kernels.cu:84-97)."""
from .. import _lib as L


def silu_and_mul(out, input):
    L.require_cuda(out, input)
    L.require_contiguous(out, input)
    d = input.shape[-1] // 2
    T = input.numel() // input.shape[-1]
    L.check(L.lib().ob_silu_and_mul(L.ptr(out), L.ptr(input), T, d, L.stream()), "silu_and_mul")


def silu_and_mul_quant(out, input, input_sum, scale):
    """Extension (not in the reference): silu_and_mul fused with invoke_quant(_fuse_sum); input_sum may be None."""
    L.require_cuda(out, input, input_sum, scale)
    L.require_contiguous(out, input, input_sum, scale)
    d = input.shape[-1] // 2
    T = input.numel() // input.shape[-1]
    L.check(
        L.lib().ob_silu_and_mul_quant(L.ptr(out), L.ptr(input), L.ptr(input_sum), L.ptr(scale), T, d, L.stream()),
        "silu_and_mul_quant")


def gelu_new(*a, **k):
    raise NotImplementedError("gelu_new is not used by the Llama W4A8 path")


def gelu_fast(*a, **k):
    raise NotImplementedError("gelu_fast is not used by the Llama W4A8 path")


def invoke_dequant_silu_and_mul_quant(*a, **k):
    raise NotImplementedError("legacy W8A8 op, not on the W4A8KV4 path")
