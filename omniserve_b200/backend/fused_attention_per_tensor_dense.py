"""omniserve_backend.fused_attention_per_tensor_dense -- static per-tensor KV8 pages
(reference: kernels/csrc/fused_attention/fused_attention_per_tensor/dense_attention/fused_attention.h:17-47,
per_tensor_common/update_kv_cache.h:17-47; callers decoding_attention.py:185-236, ctx_update_kv.py:49-92).

Pages hold INT8 codes [H_pool][64 tokens][128] (size_per_token = H_pool * 128; the scale area the cache engine reserves after
them is unused); code = cvt.rni.sat.s8(x * kv_scale_orig_quant[K|V]); value = code * kv_scale_quant_orig[K|V]."""
from . import _attn_common as A

compute_padding_offsets = A.compute_padding_offsets


def _require_kv8(int4_kv_cache, kv_cache_with_zeros, *sizes_and_heads):
    if int4_kv_cache or kv_cache_with_zeros:
        raise NotImplementedError("the per-tensor path implements INT8 pages without zero points (arg_utils.py:499-503: per_tensor "
                                  "has no zero point); KV4 is the fine_grained path")
    for size, heads in zip(sizes_and_heads[::2], sizes_and_heads[1::2]):
        if size != heads * 128:
            raise RuntimeError("size_per_token must be num_kv_heads * 128 bytes for INT8 pages")


def single_query_attention(q, k, v, kv_scale_quant_orig_, kv_scale_orig_quant_, retrieval_kv_pointers, streaming_kv_pointers,
                           retrieval_head_flags, head_rank_table, length_per_sample_, alibi_slopes_, memory_max_seqlen,
                           tokens_per_block, size_per_retrieval_token, size_per_streaming_token, sink_token_num, local_token_num,
                           sink_block_num, local_block_num, num_retrieval_kv_heads, num_streaming_kv_heads, timestep,
                           rotary_embedding_dim, rotary_base, rotary_embedding_scale, neox_rotary_style, int4_kv_cache,
                           kv_cache_with_zeros, multiblock_switch):
    _require_kv8(int4_kv_cache, kv_cache_with_zeros, size_per_retrieval_token, num_retrieval_kv_heads, size_per_streaming_token,
                 num_streaming_kv_heads)
    if alibi_slopes_ is not None or not neox_rotary_style:
        raise NotImplementedError("alibi / GPT-J rotary are not used by the Llama path")
    if kv_scale_quant_orig_ is None or kv_scale_orig_quant_ is None:
        raise RuntimeError("per-tensor KV8 attention needs kv_scale_quant_orig and kv_scale_orig_quant")
    return A.single_query(q, k, v, retrieval_kv_pointers, streaming_kv_pointers, retrieval_head_flags, head_rank_table, None,
                          length_per_sample_, tokens_per_block, num_retrieval_kv_heads, num_streaming_kv_heads, sink_token_num,
                          local_token_num, sink_block_num, local_block_num, timestep, rotary_embedding_dim, rotary_base,
                          rotary_embedding_scale, kv8_scales=(kv_scale_quant_orig_, kv_scale_orig_quant_))


def apply_bias_rope_update_kv_cache(qkv, kv_scale_orig_quant, retrieval_seq_lens, streaming_seq_lens, padding_offset,
                                    retrieval_kv_pointers, streaming_kv_pointers, retrieval_head_flags, head_rank_table, head_num,
                                    kv_head_num, seq_len, tokens_per_block, size_per_retrieval_token, size_per_streaming_token,
                                    sink_token_num, local_token_num, sink_block_num, local_block_num, num_retrieval_kv_heads,
                                    num_streaming_kv_heads, rotary_embedding_dim, rotary_embedding_base, rotary_embedding_scale,
                                    rotary_embedding_max_positions, neox_rotary_style, int4_kv_cache, kv_cache_with_zeros):
    """Prefill: in-place NeoX RoPE of q, k in the packed qkv buffer + INT8 quantise-and-write of k, v pages."""
    _require_kv8(int4_kv_cache, kv_cache_with_zeros, size_per_retrieval_token, num_retrieval_kv_heads, size_per_streaming_token,
                 num_streaming_kv_heads)
    if tokens_per_block != 64 or not neox_rotary_style:
        raise NotImplementedError("tokens_per_block must be 64 and rotary NeoX-style")
    if kv_scale_orig_quant is None:
        raise RuntimeError("per-tensor KV8 writer needs kv_scale_orig_quant")
    A.apply_rope_update_kv(qkv, retrieval_seq_lens, padding_offset, retrieval_kv_pointers, streaming_kv_pointers,
                           retrieval_head_flags, head_rank_table, head_num, kv_head_num, seq_len, num_retrieval_kv_heads,
                           num_streaming_kv_heads, sink_token_num, local_token_num, sink_block_num, local_block_num,
                           rotary_embedding_dim, rotary_embedding_base, rotary_embedding_scale,
                           kv8_scale_orig_quant=kv_scale_orig_quant)
