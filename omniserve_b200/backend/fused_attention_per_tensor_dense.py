"""omniserve_backend.fused_attention_per_tensor_* -- static per-tensor KV8 mode, outside the KV4 north-star path
(OOS for v1; section 8b: "stub OK in v1")."""


def single_query_attention(*a, **k):
    raise NotImplementedError("per-tensor KV8 attention is outside the W4A8KV4 path (SURVEY.md section 2a)")


def apply_bias_rope_update_kv_cache(*a, **k):
    raise NotImplementedError("per-tensor KV8 KV writer is outside the W4A8KV4 path")


def compute_padding_offsets(cu_seqlens, max_seqlen, tot_num_tokens):
    from ._attn_common import compute_padding_offsets as f
    return f(cu_seqlens, max_seqlen, tot_num_tokens)
