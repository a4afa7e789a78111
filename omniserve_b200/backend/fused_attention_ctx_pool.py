"""omniserve_backend.fused_attention_ctx_pool (reference: sparse_utils/ContextPool/context_pool_kernel.cu:145-213,
pybind.cpp; caller omniserve/modeling/layers/ctx_attn/ctx_update_kv.py)."""
import torch

from .. import _lib as L


def paged_min_max_pool(input, retrieval_kv_pointers, cu_seqlens, pooling_heads_idx, max_seqlen, pooling_size,
                       page_size, size_per_retrieval_token, kv_cache_with_zeros):
    """Channel-wise max / min of the post-RoPE keys of every `pooling_size`-token sub-chunk, written into the kmax /
    kmin area of the K pages.  input: fp16 [total_tokens, num_heads, head_dim] contiguous; in place into the pages."""
    L.require_cuda(input, cu_seqlens, pooling_heads_idx, retrieval_kv_pointers)
    # the checks of context_pool_kernel.cu:156-170
    if input.dtype != torch.float16:
        raise RuntimeError("context pooling only support fp16 for input")
    if cu_seqlens.dtype != torch.int32:
        raise RuntimeError("context pooling only support int32 for cu_seqlens")
    if pooling_heads_idx.dtype != torch.int32:
        raise RuntimeError("context pooling only support int32 for pooling_heads_idx")
    for name, x in (("input", input), ("cu_seqlens", cu_seqlens), ("pooling_heads_idx", pooling_heads_idx)):
        if not x.is_contiguous():
            raise RuntimeError(f"{name} must be contiguous")
    if retrieval_kv_pointers is None:
        raise RuntimeError("paged_min_max_pool needs the retrieval page table")
    if not retrieval_kv_pointers.is_contiguous():
        raise RuntimeError("retrieval_kv_pointers must be contiguous")
    L.check(
        L.lib().ob_paged_min_max_pool(
            L.ptr(input), L.ptr(retrieval_kv_pointers), L.ptr(cu_seqlens), L.ptr(pooling_heads_idx),
            input.stride(-3), input.stride(-2), retrieval_kv_pointers.shape[-1], cu_seqlens.numel() - 1,
            pooling_heads_idx.numel(), input.shape[2], int(max_seqlen), int(pooling_size), int(page_size),
            int(size_per_retrieval_token), 1 if kv_cache_with_zeros else 0, L.stream()),
        "paged_min_max_pool")
