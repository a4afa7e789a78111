"""omniserve_backend.fused_attention_selector (reference: sparse_utils/KVPageSelector/fused_kv_page_selector.cpp:171-334;
caller omniserve/modeling/layers/decoding_attention.py:88-143)."""
import ctypes as C

import torch

from .. import _lib as L
from . import _attn_common as A


def single_query_page_selector(q, k, v, retrieval_kv_pointers, streaming_kv_pointers, retrieval_head_flags,
                               head_rank_table, dynamic_sparse_page_idxes, length_per_sample_, alibi_slopes_,
                               memory_max_seqlen, tokens_per_block, size_per_retrieval_token,
                               size_per_streaming_token, sink_token_num, local_token_num, sink_block_num,
                               local_block_num, num_retrieval_kv_heads, num_streaming_kv_heads, timestep,
                               rotary_embedding_dim, rotary_base, rotary_embedding_scale, neox_rotary_style,
                               int4_kv_cache, kv_cache_with_zeros, tokens_per_sub_chunk,
                               hidden_dim_per_retrieval_token, multiblock_switch):
    """-> fp16 [B, Hq, padded_sub_chunks]: per sub-chunk upper bound sum_d max(q_d kmax_d, q_d kmin_d) of q.k
    for retrieval heads (rows of streaming heads are zero).  k, v, the streaming arguments, the page-index list and
    multiblock_switch are unused, as in the reference."""
    A._check_qkv(q, k, v)
    A._require_kv4(int4_kv_cache, kv_cache_with_zeros)
    if alibi_slopes_ is not None or not neox_rotary_style:
        raise NotImplementedError("alibi / GPT-J rotary are not used by the Llama path")
    if retrieval_kv_pointers is None:
        raise RuntimeError("single_query_page_selector needs the retrieval page table")
    L.require_cuda(retrieval_kv_pointers, retrieval_head_flags, head_rank_table, length_per_sample_)
    if length_per_sample_ is not None and (length_per_sample_.dtype != torch.int32 or not length_per_sample_.is_contiguous()):
        raise RuntimeError("length_per_sample must be a contiguous int32 tensor")
    B, Hq, Dh = q.shape
    group = tokens_per_block // tokens_per_sub_chunk
    n_sub = (int(timestep) + tokens_per_sub_chunk - 1) // tokens_per_sub_chunk   # fused_kv_page_selector.cpp:274-277
    padded = (n_sub + group - 1) // group * group
    out = torch.empty((B, Hq, padded), dtype=q.dtype, device=q.device)  # zeroed by the op (reference: torch::zeros)
    a = L.PageSelectorArgs()
    a.q, a.q_batch_stride, a.out = L.ptr(q), q.stride(0), L.ptr(out)
    a.retrieval_kv_pointers, a.r_max_pages = L.ptr(retrieval_kv_pointers), retrieval_kv_pointers.shape[-1]
    a.length_per_sample = L.ptr(length_per_sample_)
    a.retrieval_head_flags, a.head_rank_table = L.ptr(retrieval_head_flags), L.ptr(head_rank_table)
    a.batch, a.num_heads, a.num_kv_heads, a.head_dim, a.tokens_per_block = B, Hq, k.shape[1], Dh, tokens_per_block
    a.size_per_retrieval_token, a.num_retrieval_kv_heads = int(size_per_retrieval_token), int(num_retrieval_kv_heads)
    a.timestep = int(timestep)
    a.rotary_embedding_dim, a.rotary_base, a.rotary_scale = int(rotary_embedding_dim), float(rotary_base), float(rotary_embedding_scale)
    a.tokens_per_sub_chunk, a.hidden_dim_per_retrieval_token = int(tokens_per_sub_chunk), int(hidden_dim_per_retrieval_token)
    if padded:
        L.check(L.lib().ob_kv4_page_selector(C.byref(a), L.stream()), "single_query_page_selector")
    return out


def page_topk(stats, sub_chunks_per_page, k_out):
    """Extension: the page choice of decoding_attention.py:132-141 on the device.
    stats: fp16 [B, Hq, padded_sub_chunks] (output of single_query_page_selector).  Returns int32 [B, Hq, k_out]: the k_out-1
    best pages among all but the newest (page score = max over its sub-chunks), then the newest page."""
    import torch
    L.require_cuda(stats)
    L.require_contiguous(stats)
    B, Hq, padded = stats.shape
    total = padded // sub_chunks_per_page
    out = torch.empty((B, Hq, k_out), dtype=torch.int32, device=stats.device)
    L.check(L.lib().ob_kv4_page_topk(L.ptr(stats), L.ptr(out), B * Hq, padded, sub_chunks_per_page, total, k_out, L.stream()),
            "page_topk")
    return out
