"""Shared argument marshalling for the fused_attention_* mirrors."""
import ctypes as C

import torch

from .. import _lib as L


def _check_qkv(q, k, v):
    L.require_cuda(q, k, v)
    if q.dtype != torch.float16:
        raise RuntimeError("single_query_attention: only fp16 is supported (reference: Dh=128 fp16 only)")
    if k.stride(2) != 1 or k.stride(1) != k.shape[-1] or v.stride(2) != 1 or v.stride(1) != v.shape[-1]:
        raise RuntimeError("k / v must have unit inner stride and head stride == head_dim")  # fused_attention.cpp:178-180
    if q.stride(2) != 1 or q.stride(1) != q.shape[-1]:
        raise RuntimeError("q must have unit inner stride and head stride == head_dim")


def single_query(q, k, v, r_tab, s_tab, flags, rank, dyn, lengths, tokens_per_block, n_r_heads, n_s_heads,
                 sink, local, sink_blk, local_blk, timestep, rot_dim, rot_base, rot_scale, force_split=0,
                 tokens_per_sub_chunk=0, hidden_dim_per_retrieval_token=0, quant=None, history_is_stable=False,
                 kv8_scales=None):
    """quant = (out_i8 [B, Hq*Dh], scale fp16 [B], sum fp16 [B] or None): also quantise the output row per token.
    kv8_scales = (kv_scale_quant_orig, kv_scale_orig_quant) float32 [2] CUDA tensors: per-tensor INT8 pages."""
    _check_qkv(q, k, v)
    B, Hq, Dh = q.shape
    Hkv = k.shape[1]
    out = torch.empty((B, Hq, Dh), dtype=q.dtype, device=q.device)  # torch::empty_like(q) is contiguous here
    a = L.KV4DecodeArgs()
    a.q, a.k, a.v = L.ptr(q), L.ptr(k), L.ptr(v)
    a.q_batch_stride, a.k_batch_stride, a.v_batch_stride = q.stride(0), k.stride(0), v.stride(0)
    a.out = L.ptr(out)
    a.retrieval_kv_pointers = L.ptr(r_tab)
    a.streaming_kv_pointers = L.ptr(s_tab)
    a.r_max_pages = 0 if r_tab is None else r_tab.shape[-1]
    a.s_max_pages = 0 if s_tab is None else s_tab.shape[-1]
    if lengths is not None:
        if lengths.dtype != torch.int32 or not lengths.is_contiguous():
            raise RuntimeError("length_per_sample must be a contiguous int32 tensor")  # fused_attention.cpp:185-190
    a.length_per_sample = L.ptr(lengths)
    a.retrieval_head_flags = L.ptr(flags)
    a.head_rank_table = L.ptr(rank)
    a.dynamic_sparse_page_idxes = L.ptr(dyn)
    a.num_dynamic_sparse_pages = 0 if dyn is None else dyn.shape[-1]
    a.batch, a.num_heads, a.num_kv_heads, a.head_dim, a.tokens_per_block = B, Hq, Hkv, Dh, tokens_per_block
    a.num_retrieval_kv_heads, a.num_streaming_kv_heads = n_r_heads, n_s_heads
    a.sink_token_num, a.local_token_num, a.sink_block_num, a.local_block_num = sink, local, sink_blk, local_blk
    a.timestep = int(timestep)
    a.rotary_embedding_dim, a.rotary_base, a.rotary_scale = int(rot_dim), float(rot_base), float(rot_scale)
    a.force_split = force_split
    a.tokens_per_sub_chunk, a.hidden_dim_per_retrieval_token = int(tokens_per_sub_chunk), int(hidden_dim_per_retrieval_token)
    a.history_is_stable = 1 if history_is_stable else 0
    if kv8_scales is not None:
        sqo, soq = (_kv8_scale(t) for t in kv8_scales)
        a.kv_scale_quant_orig, a.kv_scale_orig_quant = L.ptr(sqo), L.ptr(soq)
    if quant is not None:
        qo, qs, qsum = quant
        L.require_cuda(qo, qs, qsum)
        if qo.dtype != torch.int8 or not qo.is_contiguous() or qo.numel() != B * Hq * Dh:
            raise RuntimeError("fused quant output must be a contiguous int8 [B, Hq*Dh] tensor")
        a.quant_out, a.quant_scale, a.quant_sum = L.ptr(qo), L.ptr(qs), L.ptr(qsum)
    L.check(L.lib().ob_kv4_single_query_attention(C.byref(a), L.stream()), "single_query_attention")
    return out


def _kv8_scale(t):
    """fused_attention_per_tensor/.../fused_attention.cpp: float32 tensor with the K and V scale."""
    L.require_cuda(t)
    if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() < 2:
        raise RuntimeError("kv_scale_quant_orig / kv_scale_orig_quant must be contiguous float32 tensors with 2 elements (K, V)")
    return t


def apply_rope_update_kv(qkv, seq_lens, padding_offset, r_tab, s_tab, flags, rank, head_num, kv_head_num, seq_len,
                         n_r_heads, n_s_heads, sink, local, sink_blk, local_blk, rot_dim, rot_base, rot_scale,
                         pool_sub_chunk=0, kv8_scale_orig_quant=None):
    """pool_sub_chunk > 0 (extension): also write the kmax / kmin page statistics of the retrieval heads in the same pass
    (ob_kv4_apply_rope_update_kv_cache_pool)."""
    L.require_cuda(qkv, seq_lens, padding_offset)
    if not qkv.is_contiguous():
        raise RuntimeError("qkv must be contiguous")
    a = L.KV4PrefillArgs()
    a.qkv = L.ptr(qkv)
    a.seq_lens = L.ptr(seq_lens)
    a.padding_offset = L.ptr(padding_offset)
    a.max_seq_len = int(seq_len)
    a.retrieval_kv_pointers, a.streaming_kv_pointers = L.ptr(r_tab), L.ptr(s_tab)
    a.r_max_pages = 0 if r_tab is None else r_tab.shape[-1]
    a.s_max_pages = 0 if s_tab is None else s_tab.shape[-1]
    a.retrieval_head_flags, a.head_rank_table = L.ptr(flags), L.ptr(rank)
    a.num_tokens, a.batch, a.num_heads, a.num_kv_heads = qkv.shape[0], seq_lens.shape[0], head_num, kv_head_num
    a.num_retrieval_kv_heads, a.num_streaming_kv_heads = n_r_heads, n_s_heads
    a.sink_token_num, a.local_token_num, a.sink_block_num, a.local_block_num = sink, local, sink_blk, local_blk
    a.rotary_embedding_dim, a.rotary_base, a.rotary_scale = int(rot_dim), float(rot_base), float(rot_scale)
    if kv8_scale_orig_quant is not None:
        a.kv_scale_orig_quant = L.ptr(_kv8_scale(kv8_scale_orig_quant))
    if pool_sub_chunk:
        L.check(L.lib().ob_kv4_apply_rope_update_kv_cache_pool(C.byref(a), int(pool_sub_chunk), L.stream()),
                "apply_bias_rope_update_kv_cache + paged_min_max_pool (fused)")
    else:
        L.check(L.lib().ob_kv4_apply_rope_update_kv_cache(C.byref(a), L.stream()), "apply_bias_rope_update_kv_cache")


def compute_padding_offsets(cu_seqlens, max_seqlen, tot_num_tokens):
    L.require_cuda(cu_seqlens)
    out = torch.empty((tot_num_tokens,), dtype=torch.int32, device=cu_seqlens.device)
    L.check(
        L.lib().ob_compute_padding_offsets(L.ptr(out), L.ptr(cu_seqlens), cu_seqlens.shape[0] - 1, int(max_seqlen),
                                           L.stream()),
        "compute_padding_offsets")
    return out


def _require_kv4(int4_kv_cache, kv_cache_with_zeros):
    if not (int4_kv_cache and kv_cache_with_zeros):
        raise NotImplementedError("only the KV4 (INT4 + zero point) cache of the north-star path is implemented")
