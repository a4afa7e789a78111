"""omniserve_backend.fused_kernels (reference: kernels/csrc/fused.cpp:52-76, fused_kernels.cu:218-271)."""
import torch

from .. import _lib as L


def _rows(t):
    h = t.shape[-1]
    return t.numel() // h, h


def invoke_quant(out, input, scale):
    """Per-token INT8 quant: scale must be a fp16 tensor [num_tokens] (the at::Half scalar overload of the
    reference is the static-scale path that the W4A8 models never call)."""
    if not isinstance(scale, torch.Tensor):
        raise NotImplementedError("invoke_quant(at::Half scale): static-scale overload is not on the W4A8 path")
    L.require_cuda(out, input, scale)
    assert input.is_contiguous() and out.is_contiguous()
    T, H = _rows(input)
    L.check(L.lib().ob_invoke_quant(L.ptr(out), L.ptr(input), L.ptr(scale), T, H, L.stream()), "invoke_quant")


def invoke_quant_fuse_sum(out, input, input_sum, scale):
    if not isinstance(scale, torch.Tensor):
        raise NotImplementedError("invoke_quant_fuse_sum(at::Half): static-scale overload is not on the W4A8 path")
    L.require_cuda(out, input, input_sum, scale)
    assert input.is_contiguous() and out.is_contiguous()
    T, H = _rows(input)
    L.check(
        L.lib().ob_invoke_quant_fuse_sum(L.ptr(out), L.ptr(input), L.ptr(input_sum), L.ptr(scale), T, H, L.stream()),
        "invoke_quant_fuse_sum")


def invoke_dequant(*a, **k):
    raise NotImplementedError("invoke_dequant is legacy (unused by the w4a8 models, SURVEY.md section 2a)")


def invoke_dequant_add_residual(*a, **k):
    raise NotImplementedError("invoke_dequant_add_residual is legacy (unused by the w4a8 models)")
