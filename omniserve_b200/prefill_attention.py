"""Prefill (context-stage) attention.  In the reference this is third-party FP16 FlashAttention
(`flash_attn_varlen_func` / `block_sparse_attn`, omniserve/modeling/layers/ctx_attn/ctx_attn_func.py:40-87) and is
NOT part of the quantised hot path (out of scope).  We call PyTorch's fused SDPA, which is
library code in the same sense."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def make(cu_seqlens: torch.Tensor, max_len: int, hq: int, hkv: int, dh: int):
    lens = (cu_seqlens[1:] - cu_seqlens[:-1]).tolist()
    offs = cu_seqlens.tolist()
    uniform = len(set(lens)) == 1
    g = hq // hkv

    def attn(q, k, v):  # q [T,hq,dh], k/v [T,hkv,dh] (strided views) -> [T,hq,dh]
        T = q.shape[0]
        if uniform:
            B, L = len(lens), lens[0]
            qq = q.reshape(B, L, hq, dh).transpose(1, 2)
            kk = k.reshape(B, L, hkv, dh).transpose(1, 2)
            vv = v.reshape(B, L, hkv, dh).transpose(1, 2)
            o = F.scaled_dot_product_attention(qq, kk, vv, is_causal=True, enable_gqa=(g > 1))
            return o.transpose(1, 2).reshape(T, hq, dh)
        out = torch.empty((T, hq, dh), dtype=q.dtype, device=q.device)
        for b, L in enumerate(lens):
            s = offs[b]
            qq = q[s:s + L].transpose(0, 1).unsqueeze(0)
            kk = k[s:s + L].transpose(0, 1).unsqueeze(0)
            vv = v[s:s + L].transpose(0, 1).unsqueeze(0)
            o = F.scaled_dot_product_attention(qq, kk, vv, is_causal=True, enable_gqa=(g > 1))
            out[s:s + L] = o.squeeze(0).transpose(0, 1)
        return out

    return attn
