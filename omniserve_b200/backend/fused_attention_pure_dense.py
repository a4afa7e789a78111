"""omniserve_backend.fused_attention_pure_dense
(reference: kernels/csrc/fused_attention/fused_attention_pure_dense/fused_attention.cpp:150-256)."""
from . import _attn_common as A

compute_padding_offsets = A.compute_padding_offsets


def single_query_attention(q, k, v, kv_pointers, length_per_sample_, alibi_slopes_, memory_max_seqlen,
                           tokens_per_block, size_per_token, timestep, rotary_embedding_dim, rotary_base,
                           neox_rotary_style, int4_kv_cache, kv_cache_with_zeros):
    """QServe dense decode attention over KV4 pages; appends the new token's K/V; returns fp16 [B,Hq,Dh]."""
    A._require_kv4(int4_kv_cache, kv_cache_with_zeros)
    if alibi_slopes_ is not None:
        raise NotImplementedError("alibi is not used by the Llama path")
    if not neox_rotary_style:
        raise NotImplementedError("only NeoX-style rotary embedding (reference callers pass True)")
    Hkv, Dh = k.shape[1], k.shape[-1]
    if size_per_token != Hkv * Dh // 2:
        raise RuntimeError("size_per_token must be num_kv_heads * head_dim / 2 for KV4")
    return A.single_query(q, k, v, kv_pointers, None, None, None, None, length_per_sample_, tokens_per_block, Hkv, 0,
                          0, 0, 0, 0, timestep, rotary_embedding_dim, rotary_base, 1.0)


def single_query_attention_quant(q, k, v, kv_pointers, length_per_sample_, alibi_slopes_, memory_max_seqlen,
                                 tokens_per_block, size_per_token, timestep, rotary_embedding_dim, rotary_base,
                                 neox_rotary_style, int4_kv_cache, kv_cache_with_zeros, quant_out, quant_sum, quant_scale,
                                 history_is_stable=False):
    """Extension (not in the reference): single_query_attention followed by fused_kernels.invoke_quant(_fuse_sum) of
    its [B, Hq*Dh] output (llama_w4a8_unpad.py:351-354) in ONE launch -- the last CTA of each sequence to finish
    quantises the row.  quant_sum may be None.  Bit-identical to the two-op chain.  history_is_stable: see
    include/omniserve_b200.h (decode loops pass True; the default waits for the stream dependency first)."""
    A._require_kv4(int4_kv_cache, kv_cache_with_zeros)
    if alibi_slopes_ is not None or not neox_rotary_style:
        raise NotImplementedError("alibi / GPT-J rotary are not used by the Llama path")
    Hkv, Dh = k.shape[1], k.shape[-1]
    if size_per_token != Hkv * Dh // 2:
        raise RuntimeError("size_per_token must be num_kv_heads * head_dim / 2 for KV4")
    return A.single_query(q, k, v, kv_pointers, None, None, None, None, length_per_sample_, tokens_per_block, Hkv, 0,
                          0, 0, 0, 0, timestep, rotary_embedding_dim, rotary_base, 1.0,
                          quant=(quant_out, quant_scale, quant_sum), history_is_stable=history_is_stable)


def apply_bias_rope_update_kv_cache(qkv, seq_lens, padding_offset, kv_pointers, head_num, kv_head_num, seq_len,
                                    tokens_per_block, size_per_token, rotary_embedding_dim, rotary_embedding_base,
                                    rotary_embedding_max_positions, neox_rotary_style, int4_kv_cache,
                                    kv_cache_with_zeros):
    """The 15-argument op of fused_attention_pure_dense/update_kv_cache.cu:20-36: one pool holding every KV head, no
    rotary scaling.  In-place NeoX RoPE of q, k in the packed qkv buffer + KV4 quantise-and-write of the K / V pages."""
    A._require_kv4(int4_kv_cache, kv_cache_with_zeros)
    if tokens_per_block != 64 or not neox_rotary_style:
        raise NotImplementedError("tokens_per_block must be 64 and rotary NeoX-style")
    if kv_pointers is None:
        raise RuntimeError("kv_pointers is required")
    if size_per_token != kv_head_num * rotary_embedding_dim // 2:
        raise RuntimeError("size_per_token must be num_kv_heads * head_dim / 2 for KV4")
    A.apply_rope_update_kv(qkv, seq_lens, padding_offset, kv_pointers, None, None, None, head_num, kv_head_num, seq_len,
                           kv_head_num, 0, 0, 0, 0, 0, rotary_embedding_dim, rotary_embedding_base, 1.0)
