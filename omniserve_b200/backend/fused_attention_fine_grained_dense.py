"""omniserve_backend.fused_attention_fine_grained_dense
(reference: .../fused_attention_fine_grained/dense_attention/fused_attention.cpp, fine_grained_common/update_kv_cache.cu:27-136)."""
from . import _attn_common as A

compute_padding_offsets = A.compute_padding_offsets


def single_query_attention(q, k, v, retrieval_kv_pointers, streaming_kv_pointers, retrieval_head_flags,
                           head_rank_table, length_per_sample_, alibi_slopes_, memory_max_seqlen, tokens_per_block,
                           size_per_retrieval_token, size_per_streaming_token, sink_token_num, local_token_num,
                           sink_block_num, local_block_num, num_retrieval_kv_heads, num_streaming_kv_heads, timestep,
                           rotary_embedding_dim, rotary_base, rotary_embedding_scale, neox_rotary_style, int4_kv_cache,
                           kv_cache_with_zeros, multiblock_switch):
    A._require_kv4(int4_kv_cache, kv_cache_with_zeros)
    if alibi_slopes_ is not None or not neox_rotary_style:
        raise NotImplementedError("alibi / GPT-J rotary are not used by the Llama path")
    return A.single_query(q, k, v, retrieval_kv_pointers, streaming_kv_pointers, retrieval_head_flags,
                          head_rank_table, None, length_per_sample_, tokens_per_block, num_retrieval_kv_heads,
                          num_streaming_kv_heads, sink_token_num, local_token_num, sink_block_num, local_block_num,
                          timestep, rotary_embedding_dim, rotary_base, rotary_embedding_scale)


def apply_bias_rope_update_kv_cache(qkv, retrieval_seq_lens, streaming_seq_lens, padding_offset,
                                    retrieval_kv_pointers, streaming_kv_pointers, retrieval_head_flags,
                                    head_rank_table, head_num, kv_head_num, seq_len, tokens_per_block,
                                    size_per_retrieval_token, size_per_streaming_token, sink_token_num,
                                    local_token_num, sink_block_num, local_block_num, num_retrieval_kv_heads,
                                    num_streaming_kv_heads, rotary_embedding_dim, rotary_embedding_base,
                                    rotary_embedding_scale, rotary_embedding_max_positions, neox_rotary_style,
                                    int4_kv_cache, kv_cache_with_zeros):
    """Prefill: in-place NeoX RoPE of q,k in the packed qkv buffer + KV4 quantise-and-write of k,v pages."""
    A._require_kv4(int4_kv_cache, kv_cache_with_zeros)
    if tokens_per_block != 64 or not neox_rotary_style:
        raise NotImplementedError("tokens_per_block must be 64 and rotary NeoX-style")
    A.apply_rope_update_kv(qkv, retrieval_seq_lens, padding_offset, retrieval_kv_pointers, streaming_kv_pointers,
                           retrieval_head_flags, head_rank_table, head_num, kv_head_num, seq_len,
                           num_retrieval_kv_heads, num_streaming_kv_heads, sink_token_num, local_token_num,
                           sink_block_num, local_block_num, rotary_embedding_dim, rotary_embedding_base,
                           rotary_embedding_scale)


def apply_bias_rope_update_kv_cache_pool(qkv, retrieval_seq_lens, streaming_seq_lens, padding_offset, retrieval_kv_pointers,
                                         streaming_kv_pointers, retrieval_head_flags, head_rank_table, head_num, kv_head_num,
                                         seq_len, tokens_per_block, size_per_retrieval_token, size_per_streaming_token,
                                         sink_token_num, local_token_num, sink_block_num, local_block_num,
                                         num_retrieval_kv_heads, num_streaming_kv_heads, rotary_embedding_dim,
                                         rotary_embedding_base, rotary_embedding_scale, rotary_embedding_max_positions,
                                         neox_rotary_style, int4_kv_cache, kv_cache_with_zeros, tokens_per_sub_chunk=16):
    """Extension: apply_bias_rope_update_kv_cache fused with
    fused_attention_ctx_pool.paged_min_max_pool of the rotated keys of the retrieval heads -- one pass over the chunk instead
    of two (ctx_update_kv.py:104-178).  Same arguments as the unfused op + the sub-chunk size; bit-identical pages."""
    A._require_kv4(int4_kv_cache, kv_cache_with_zeros)
    if tokens_per_block != 64 or not neox_rotary_style:
        raise NotImplementedError("tokens_per_block must be 64 and rotary NeoX-style")
    A.apply_rope_update_kv(qkv, retrieval_seq_lens, padding_offset, retrieval_kv_pointers, streaming_kv_pointers,
                           retrieval_head_flags, head_rank_table, head_num, kv_head_num, seq_len,
                           num_retrieval_kv_heads, num_streaming_kv_heads, sink_token_num, local_token_num,
                           sink_block_num, local_block_num, rotary_embedding_dim, rotary_embedding_base,
                           rotary_embedding_scale, pool_sub_chunk=tokens_per_sub_chunk)
