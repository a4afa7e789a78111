"""Host side of LServe's dynamic sparse decode on top of the `omniserve_backend`-compatible ops.

Mirrors (same argument meaning, same page choice):
  * DecodingAttentionWrapper.dynamic_select_topk_pages   omniserve/modeling/layers/decoding_attention.py:88-143
  * DecodingAttentionWrapper.forward_w_dynamic_sparse_fine_grained   .../decoding_attention.py (selector reuse every
    `selector_update_interval` steps, then fused_attention_fine_grained_sparse.single_query_attention)
  * PagedMinMaxPoolWrapper.forward   omniserve/modeling/layers/ctx_update_kv.py:158-178
PyTorch is used for the top-k over page scores exactly as in the reference; the scores, the statistics and the
attention go through the C ABI.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

from .backend import fused_attention_ctx_pool, fused_attention_fine_grained_sparse, fused_attention_selector


@dataclass
class SparseDecodeConfig:
    head_dim: int = 128
    tokens_per_block: int = 64
    sub_chunk_size: int = 16              # tokens per sub-chunk (64 / sub_chunk_per_block)
    dynamic_sparse_token_budget: int = 4096
    selector_update_interval: int = 4
    memory_max_len: int = 1 << 20
    rotary_base: float = 500000.0
    rope_scaling_factor: float = 1.0
    multiblock_switch: int = 2048
    device_topk: bool = True              # page choice on the device (csrc/lserve_ops.cu: page_topk_kernel) instead of torch ops


def paged_min_max_pool(keys, retrieval_block_table, cu_seqlens, max_seq_len, pooling_heads_idx, num_retrieval_kv_heads,
                       cfg: SparseDecodeConfig):
    """PagedMinMaxPoolWrapper.forward (ctx_update_kv.py:158-178): keys fp16 [T, Hkv, Dh] contiguous, post-RoPE."""
    size_per_retrieval_token = num_retrieval_kv_heads * cfg.head_dim // 2
    fused_attention_ctx_pool.paged_min_max_pool(keys, retrieval_block_table, cu_seqlens, pooling_heads_idx, max_seq_len,
                                                cfg.sub_chunk_size, cfg.tokens_per_block, size_per_retrieval_token, True)


def dynamic_select_topk_pages(q, k, v, retrieval_block_table, streaming_block_table, retrieval_head_flags,
                              head_rank_table, lengths_per_sample, sink_size, local_size, sink_blocks, local_blocks,
                              num_retrieval_kv_heads, num_streaming_kv_heads, timestep: int, cfg: SparseDecodeConfig):
    """decoding_attention.py:88-143.  -> int32 [B, Hq, P]; the newest page is always last."""
    tpb = cfg.tokens_per_block
    if timestep <= cfg.dynamic_sparse_token_budget:
        n = timestep // tpb + 1
        return torch.arange(n, device=q.device, dtype=torch.int32).view(1, 1, -1).expand(q.shape[0], q.shape[1], -1).contiguous()
    budget = min(cfg.dynamic_sparse_token_budget, timestep)
    size_r = num_retrieval_kv_heads * cfg.head_dim // 2
    size_s = num_streaming_kv_heads * cfg.head_dim // 2
    stats = fused_attention_selector.single_query_page_selector(
        q, k, v, retrieval_block_table, streaming_block_table, retrieval_head_flags, head_rank_table, None,
        lengths_per_sample, None, cfg.memory_max_len, tpb, size_r, size_s, sink_size, local_size, sink_blocks,
        local_blocks, num_retrieval_kv_heads, num_streaming_kv_heads, timestep, cfg.head_dim, cfg.rotary_base,
        cfg.rope_scaling_factor, True, True, True, cfg.sub_chunk_size, num_retrieval_kv_heads * cfg.head_dim, 1000000)
    group = tpb // cfg.sub_chunk_size
    total = stats.shape[-1] // group
    k_out = min(max(3, budget // tpb), total)
    if cfg.device_topk:
        # one kernel instead of the reference's view / max / topk / cat / to(int32) chain (decoding_attention.py:132-141)
        return fused_attention_selector.page_topk(stats, group, k_out)
    stats = stats.view(q.shape[0], q.shape[1], -1, group)
    stats = torch.max(stats, dim=-1).values
    _, idx = stats[:, :, :-1].topk(k=k_out - 1, dim=-1)
    idx = torch.cat([idx, torch.ones_like(idx[..., :1]) * (total - 1)], dim=-1).contiguous()
    return idx.to(torch.int32)


def sparse_decode_attention(q, k, v, retrieval_block_table, streaming_block_table, retrieval_head_flags,
                            head_rank_table, lengths_per_sample, sink_size, local_size, sink_blocks, local_blocks,
                            num_retrieval_kv_heads, num_streaming_kv_heads, timestep: int, cfg: SparseDecodeConfig,
                            cached_page_idx: Optional[torch.Tensor] = None):
    """forward_w_dynamic_sparse_fine_grained: reuse the cached page list unless `timestep` is a multiple of the
    update interval, then run the sparse KV4 attention.  Returns (attn_output [B,Hq,Dh], page_idx)."""
    if (timestep % cfg.selector_update_interval != 0) and cached_page_idx is not None:
        page_idx = cached_page_idx
    else:
        page_idx = dynamic_select_topk_pages(q, k, v, retrieval_block_table, streaming_block_table,
                                             retrieval_head_flags, head_rank_table, lengths_per_sample, sink_size,
                                             local_size, sink_blocks, local_blocks, num_retrieval_kv_heads,
                                             num_streaming_kv_heads, timestep, cfg)
    size_r = num_retrieval_kv_heads * cfg.head_dim // 2
    size_s = num_streaming_kv_heads * cfg.head_dim // 2
    out = fused_attention_fine_grained_sparse.single_query_attention(
        q, k, v, retrieval_block_table, streaming_block_table, retrieval_head_flags, head_rank_table, page_idx,
        lengths_per_sample, None, cfg.memory_max_len, cfg.tokens_per_block, size_r, size_s, sink_size, local_size,
        sink_blocks, local_blocks, num_retrieval_kv_heads, num_streaming_kv_heads, timestep, cfg.head_dim,
        cfg.rotary_base, cfg.rope_scaling_factor, True, True, True, cfg.sub_chunk_size,
        num_retrieval_kv_heads * cfg.head_dim, cfg.multiblock_switch)
    return out, page_idx
