"""omniserve_backend.fused_attention_fine_grained_sparse
(reference: .../fused_attention_fine_grained/sparse_attention/fused_attention.cpp:198-377)."""
from . import _attn_common as A
from .fused_attention_fine_grained_dense import apply_bias_rope_update_kv_cache  # noqa: F401 (same symbol in the reference)

compute_padding_offsets = A.compute_padding_offsets


def single_query_attention(q, k, v, retrieval_kv_pointers, streaming_kv_pointers, retrieval_head_flags,
                           head_rank_table, dynamic_sparse_page_idxes, length_per_sample_, alibi_slopes_,
                           memory_max_seqlen, tokens_per_block, size_per_retrieval_token, size_per_streaming_token,
                           sink_token_num, local_token_num, sink_block_num, local_block_num, num_retrieval_kv_heads,
                           num_streaming_kv_heads, timestep, rotary_embedding_dim, rotary_base, rotary_embedding_scale,
                           neox_rotary_style, int4_kv_cache, kv_cache_with_zeros, tokens_per_sub_chunk,
                           hidden_dim_per_retrieval_token, multiblock_switch):
    A._require_kv4(int4_kv_cache, kv_cache_with_zeros)
    if alibi_slopes_ is not None or not neox_rotary_style:
        raise NotImplementedError("alibi / GPT-J rotary are not used by the Llama path")
    return A.single_query(q, k, v, retrieval_kv_pointers, streaming_kv_pointers, retrieval_head_flags,
                          head_rank_table, dynamic_sparse_page_idxes, length_per_sample_, tokens_per_block,
                          num_retrieval_kv_heads, num_streaming_kv_heads, sink_token_num, local_token_num,
                          sink_block_num, local_block_num, timestep, rotary_embedding_dim, rotary_base, rotary_embedding_scale,
                          tokens_per_sub_chunk=tokens_per_sub_chunk,
                          hidden_dim_per_retrieval_token=hidden_dim_per_retrieval_token)
