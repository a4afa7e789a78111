"""CPU restatement of the reference's static per-tensor INT8 KV cache ("KV8", kv_quant_granularity = per_tensor).

TEST INFRASTRUCTURE -- see oracle/__init__.py.  numpy only, loops per (sequence, head): keep cases small.
PARITY PINNING: checked against the reference's own compiled kernels (oracle/_ref fused_attention_per_tensor_dense) in
tests/test_gpu_kv8.py when oracle/_ref is present; the reference's Python tests hold no golden vectors for this path.

Reference files restated (under /root/reference/kernels/csrc/fused_attention):
  * page layout: common/kvCacheUtils.h (same addressing as KV4), omniserve/worker/cache_engine.py:73-88 with
    element size 1: K / V page = [H][64 tokens][Dh] int8 | (reserved, unused) fp16 [2][H][64] | (sparse only) kmax / kmin
  * quantise: common/decoderMaskedMultiheadAttentionUtils.h store_8bits_kv_cache_vec(int8_t*, vec, idx, float scale):
    code = cvt.rni.sat.s8(float(x) * kv_scale_orig_quant)   (no zero point: arg_utils.py:499-503)
  * dequantise: ...Utils.h convert_from_8bit_kv_cache: fp16(float(code) * kv_scale_quant_orig)
  * the rest (RoPE, softmax, new token handled in fp16 / fp32) as in oracle/kv4.py
    (fused_attention_per_tensor/dense_attention/decoderMaskedMultiheadAttentionTemplate.hpp:928-1001,1349-1377,1559-1717,
    1947-2057,2140-2162)
"""
from __future__ import annotations

import numpy as np

from .kv4 import TOKENS_PER_BLOCK, f16, f32, f64, rope_neox


class PagedKV8:
    """A K pool and a V pool of per-tensor INT8 pages, byte-identical to the reference's layout."""

    def __init__(self, num_pages: int, n_kv_heads: int, head_dim: int = 128, k_stats_subchunks: int = 0):
        self.H, self.Dh, self.P = n_kv_heads, head_dim, num_pages
        self.data_bytes = n_kv_heads * TOKENS_PER_BLOCK * head_dim
        self.sz_bytes = TOKENS_PER_BLOCK * n_kv_heads * 4               # reserved by the cache engine, unused here
        self.stats_bytes = 2 * k_stats_subchunks * n_kv_heads * head_dim * 2
        self.k_page_bytes = self.data_bytes + self.sz_bytes + self.stats_bytes
        self.v_page_bytes = self.data_bytes + self.sz_bytes
        self.k_pool = np.zeros((num_pages, self.k_page_bytes), np.uint8)
        self.v_pool = np.zeros((num_pages, self.v_page_bytes), np.uint8)
        self.n_sub = k_stats_subchunks

    def data(self, which, page):  # [H, 64, Dh] int8 view
        pool = self.k_pool if which == "k" else self.v_pool
        return pool[page, : self.data_bytes].view(np.int8).reshape(self.H, TOKENS_PER_BLOCK, self.Dh)

    def kstats(self, page):
        o = self.data_bytes + self.sz_bytes
        n = self.n_sub * self.H * self.Dh * 2
        kmax = self.k_pool[page, o: o + n].view(f16).reshape(self.n_sub, self.H * self.Dh)
        kmin = self.k_pool[page, o + n: o + 2 * n].view(f16).reshape(self.n_sub, self.H * self.Dh)
        return kmax, kmin


def kv8_quant(x_f16, scale_orig_quant: float) -> np.ndarray:
    """cvt.rni.sat.s8(float(x) * scale) -- round to nearest even, saturate."""
    y = (np.asarray(x_f16, f16).astype(f32) * f32(scale_orig_quant)).astype(f32)
    return np.clip(np.rint(y), -128, 127).astype(np.int8)


def kv8_dequant_f16(code: np.ndarray, scale_quant_orig: float) -> np.ndarray:
    return (code.astype(f32) * f32(scale_quant_orig)).astype(f32).astype(f16)


def kv8_dequant_exact(code: np.ndarray, scale_quant_orig: float) -> np.ndarray:
    return code.astype(f64) * f64(f32(scale_quant_orig))


def prefill_write(cache: PagedKV8, block_table, qkv_f16, seq_lens, n_q_heads: int, rotary_dim: int, base: float,
                  scale_orig_quant, scale: float = 1.0):
    """apply_bias_rope_update_kv_cache of fused_attention_per_tensor_dense (per_tensor_common/applyBiasRopeUpdateKVCache.h):
    in-place NeoX RoPE of q, k; INT8 codes of post-RoPE k (scale_orig_quant[0]) and raw v (scale_orig_quant[1])."""
    H, Dh = cache.H, cache.Dh
    t0 = 0
    for b, L in enumerate(seq_lens):
        blk = qkv_f16[t0: t0 + L]
        pos = np.arange(L)
        qv = blk[:, : n_q_heads * Dh].reshape(L, n_q_heads, Dh)
        kv = blk[:, n_q_heads * Dh: (n_q_heads + H) * Dh].reshape(L, H, Dh)
        vv = blk[:, (n_q_heads + H) * Dh:].reshape(L, H, Dh)
        qv[:] = rope_neox(qv, pos[:, None], rotary_dim, base, scale)
        kv[:] = rope_neox(kv, pos[:, None], rotary_dim, base, scale)
        for t in range(L):
            page, slot = int(block_table[b, t // TOKENS_PER_BLOCK]), t % TOKENS_PER_BLOCK
            cache.data("k", page)[:, slot, :] = kv8_quant(kv[t], scale_orig_quant[0])
            cache.data("v", page)[:, slot, :] = kv8_quant(vv[t], scale_orig_quant[1])
        t0 += L


def decode_attention(q, k, v, cache: PagedKV8, block_table, lengths, rotary_dim: int, base: float, scale_quant_orig,
                     scale_orig_quant, scale: float = 1.0, mimic: bool = False, positions_fn=None, head_rank=None,
                     append: bool = True, update_stats_sub_chunk: int = 0):
    """single_query_attention of fused_attention_per_tensor_{dense,sparse}.  Arguments as oracle/kv4.py:decode_attention
    + the two per-tensor scale pairs (K, V).  mimic=True rounds the dequantised values to fp16 like the reference
    (convert_from_8bit_kv_cache) and the probabilities to fp16; mimic=False is exact arithmetic on the stored codes."""
    q, k, v = (np.asarray(t, f16) for t in (q, k, v))
    B, Hq, Dh = q.shape
    Hkv = k.shape[1]
    g = Hq // Hkv
    out = np.zeros((B, Hq, Dh), f16)
    for b in range(B):
        tl = int(lengths[b]) - 1
        q_r = rope_neox(q[b], tl, rotary_dim, base, scale)
        k_r = rope_neox(k[b], tl, rotary_dim, base, scale)
        page, slot = int(block_table[b, tl // TOKENS_PER_BLOCK]), tl % TOKENS_PER_BLOCK
        for hq in range(Hq):
            hk = hq // g
            rank = hk if head_rank is None else int(head_rank[hk])
            pos = np.arange(tl) if positions_fn is None else np.asarray(positions_fn(b, hq, tl))
            n = len(pos)
            kc = np.zeros((n, Dh), np.int8)
            vc = np.zeros((n, Dh), np.int8)
            for j, t in enumerate(pos):
                pg = int(block_table[b, int(t) // TOKENS_PER_BLOCK])
                kc[j] = cache.data("k", pg)[rank, int(t) % TOKENS_PER_BLOCK]
                vc[j] = cache.data("v", pg)[rank, int(t) % TOKENS_PER_BLOCK]
            if mimic:
                kd, vd = kv8_dequant_f16(kc, scale_quant_orig[0]).astype(f64), kv8_dequant_f16(vc, scale_quant_orig[1]).astype(f64)
            else:
                kd, vd = kv8_dequant_exact(kc, scale_quant_orig[0]), kv8_dequant_exact(vc, scale_quant_orig[1])
            logits = np.empty(n + 1, f64)
            logits[:n] = kd @ q_r[hq].astype(f64)
            logits[n] = (q_r[hq].astype(f64) * k_r[hk].astype(f64)).sum()
            logits *= 1.0 / np.sqrt(Dh)
            e = np.exp(logits - logits.max())
            p = e / e.sum()
            if mimic:
                p = p.astype(f16).astype(f64)
            o = p[n] * v[b, hk].astype(f64)
            if n:
                o = o + p[:n] @ vd
            out[b, hq] = o.astype(f16)
        if append:
            for hk in range(Hkv):
                rank = hk if head_rank is None else int(head_rank[hk])
                cache.data("k", page)[rank, slot] = kv8_quant(k_r[hk], scale_orig_quant[0])
                cache.data("v", page)[rank, slot] = kv8_quant(v[b, hk], scale_orig_quant[1])
                if update_stats_sub_chunk:
                    kmax, kmin = cache.kstats(page)
                    sub = slot // update_stats_sub_chunk
                    sl = slice(rank * Dh, (rank + 1) * Dh)
                    kmax[sub, sl] = np.maximum(kmax[sub, sl], k_r[hk])
                    kmin[sub, sl] = np.minimum(kmin[sub, sl], k_r[hk])
    return out
