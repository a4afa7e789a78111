"""omniserve_backend.fused_attention_per_tensor_sparse -- static per-tensor KV8 pages with LServe's dynamic page choice
(reference: fused_attention_per_tensor/sparse_attention/fused_attention.h:17-51, caller
decoding_attention.py:239-304).  Same kernel as the dense per-tensor op + `dynamic_sparse_page_idxes`; the kmax / kmin page
statistics sit behind the (unused) scale area exactly as in KV4 pages, so fused_attention_ctx_pool.paged_min_max_pool and
fused_attention_selector.single_query_page_selector work on these pages unchanged (they take size_per_retrieval_token)."""
from . import _attn_common as A
from .fused_attention_per_tensor_dense import _require_kv8, apply_bias_rope_update_kv_cache  # noqa: F401

compute_padding_offsets = A.compute_padding_offsets


def single_query_attention(q, k, v, kv_scale_quant_orig_, kv_scale_orig_quant_, retrieval_kv_pointers, streaming_kv_pointers,
                           retrieval_head_flags, head_rank_table, dynamic_sparse_page_idxes, length_per_sample_, alibi_slopes_,
                           memory_max_seqlen, tokens_per_block, size_per_retrieval_token, size_per_streaming_token,
                           sink_token_num, local_token_num, sink_block_num, local_block_num, num_retrieval_kv_heads,
                           num_streaming_kv_heads, timestep, rotary_embedding_dim, rotary_base, rotary_embedding_scale,
                           neox_rotary_style, int4_kv_cache, kv_cache_with_zeros, tokens_per_sub_chunk,
                           hidden_dim_per_retrieval_token, multiblock_switch):
    _require_kv8(int4_kv_cache, kv_cache_with_zeros, size_per_retrieval_token, num_retrieval_kv_heads, size_per_streaming_token,
                 num_streaming_kv_heads)
    if alibi_slopes_ is not None or not neox_rotary_style:
        raise NotImplementedError("alibi / GPT-J rotary are not used by the Llama path")
    if kv_scale_quant_orig_ is None or kv_scale_orig_quant_ is None:
        raise RuntimeError("per-tensor KV8 attention needs kv_scale_quant_orig and kv_scale_orig_quant")
    return A.single_query(q, k, v, retrieval_kv_pointers, streaming_kv_pointers, retrieval_head_flags, head_rank_table,
                          dynamic_sparse_page_idxes, length_per_sample_, tokens_per_block, num_retrieval_kv_heads,
                          num_streaming_kv_heads, sink_token_num, local_token_num, sink_block_num, local_block_num, timestep,
                          rotary_embedding_dim, rotary_base, rotary_embedding_scale, tokens_per_sub_chunk=tokens_per_sub_chunk,
                          hidden_dim_per_retrieval_token=hidden_dim_per_retrieval_token,
                          kv8_scales=(kv_scale_quant_orig_, kv_scale_orig_quant_))
