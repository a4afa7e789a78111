"""CPU restatement of the reference's KV4 page format, KV quantiser, RoPE and decode attention.

TEST INFRASTRUCTURE -- see oracle/__init__.py.  numpy only; loops are per (sequence, head), so
keep cases small (the GPU tests use ctx <= ~300 against this oracle and properties at full size).

Reference files restated (under /root/reference/kernels/csrc/fused_attention):
  * page addressing / layout: fused_attention_pure_dense/kvCacheUtils.h:47-126,
    omniserve/worker/cache_engine.py:73-88  (K page = [H][64][Dh/2] bytes | scales f16 [H][64] |
    zeros f16 [H][64] | (sparse only) kmax f16 [4][H*Dh] | kmin f16 [4][H*Dh]; V page = same w/o stats)
  * per-token/per-head asymmetric 4-bit quant: .../decoderMaskedMultiheadAttentionTemplate.hpp:1045-1114,1221-1257
    and fine_grained_common/applyBiasRopeUpdateKVCache.h:303-345
  * nibble packing: .../decoderMaskedMultiheadAttentionUtils.h:1838-1852 (``cvt.rni.sat.u8`` then ``& 0xF``: no clamp to 15)
  * dequant (fp16 ``__hfma2(v, s, -s*z)``): ...Utils.h:2125-2213
  * RoPE (NeoX pairs (i, i+Dh/2), fp32 pow/cos/sin): ...Utils.h:1147-1167
  * cached-key q.k in fp16x2 FMA: ...Template.hpp:450-467; softmax :1757-1831; P.V :1901-1977;
    cross-group fp16 smem reduce :2139-2160; output :2162-2221
"""
from __future__ import annotations

import numpy as np

f16, f32, f64 = np.float16, np.float32, np.float64
TOKENS_PER_BLOCK = 64


# ----------------------------------------------------------------------------------------------
# page pool
# ----------------------------------------------------------------------------------------------
class PagedKV4:
    """A K pool and a V pool of KV4 pages, byte-identical to the reference's layout."""

    def __init__(self, num_pages: int, n_kv_heads: int, head_dim: int = 128, k_stats_subchunks: int = 0):
        self.H, self.Dh, self.P = n_kv_heads, head_dim, num_pages
        self.data_bytes = n_kv_heads * TOKENS_PER_BLOCK * head_dim // 2  # == mBytesPerSeq
        self.sz_bytes = TOKENS_PER_BLOCK * n_kv_heads * 4               # scales + zeros, fp16 each
        self.stats_bytes = 2 * k_stats_subchunks * n_kv_heads * head_dim * 2
        self.k_page_bytes = self.data_bytes + self.sz_bytes + self.stats_bytes
        self.v_page_bytes = self.data_bytes + self.sz_bytes
        self.k_pool = np.zeros((num_pages, self.k_page_bytes), np.uint8)
        self.v_pool = np.zeros((num_pages, self.v_page_bytes), np.uint8)
        self.n_sub = k_stats_subchunks

    def _pool(self, which):
        return self.k_pool if which == "k" else self.v_pool

    def data(self, which, page):  # [H, 64, Dh/2] uint8 view
        return self._pool(which)[page, : self.data_bytes].reshape(self.H, TOKENS_PER_BLOCK, self.Dh // 2)

    def scales(self, which, page):  # [H, 64] fp16 view
        o = self.data_bytes
        return self._pool(which)[page, o : o + self.sz_bytes // 2].view(f16).reshape(self.H, TOKENS_PER_BLOCK)

    def zeros(self, which, page):
        o = self.data_bytes + self.sz_bytes // 2
        return self._pool(which)[page, o : o + self.sz_bytes // 2].view(f16).reshape(self.H, TOKENS_PER_BLOCK)

    def kstats(self, page):  # (kmax, kmin) each [n_sub, H*Dh] fp16 views
        o = self.data_bytes + self.sz_bytes
        n = self.n_sub * self.H * self.Dh * 2
        kmax = self.k_pool[page, o : o + n].view(f16).reshape(self.n_sub, self.H * self.Dh)
        kmin = self.k_pool[page, o + n : o + 2 * n].view(f16).reshape(self.n_sub, self.H * self.Dh)
        return kmax, kmin


# ----------------------------------------------------------------------------------------------
# quant / dequant of one (token, head) vector
# ----------------------------------------------------------------------------------------------
def kv4_quant(x_f16: np.ndarray):
    """x fp16 [..., Dh] -> (q uint8 [..., Dh] in 0..15, scale fp16 [...], zero fp16 [...]).

    scale = f16((max-min)/15); zero = f16(-15*min/(max-min)); q = rni.sat.u8(x*(1/scale) + zero) & 0xF
    (Template.hpp:1063-1081 + Utils.h:1838-1852).  A value that rounds to 16 wraps to 0 -- reproduced.
    """
    x = np.asarray(x_f16, f16).astype(f32)
    mx = x.max(axis=-1)
    mn = x.min(axis=-1)
    with np.errstate(divide="ignore", invalid="ignore"):
        scale = ((mx - mn) / f32(15.0)).astype(f16)
        zero = (f32(-15.0) * mn / (mx - mn)).astype(f16)
        inv = f32(1.0) / scale.astype(f32)
        # mul + add contracted to one fma by nvcc (default -fmad=true): single rounding
        t = (x.astype(f64) * inv[..., None].astype(f64) + zero[..., None].astype(f64)).astype(f32)
    t = np.nan_to_num(t, nan=0.0, posinf=255.0, neginf=0.0)
    q = np.clip(np.rint(t), 0, 255).astype(np.uint8) & 0xF
    return q, scale, zero


def pack_nibbles(q: np.ndarray) -> np.ndarray:
    """[..., Dh] u4 -> [..., Dh/2] bytes: byte j = q[2j] | q[2j+1] << 4 (Utils.h:1838-1852)."""
    q = np.asarray(q, np.uint8)
    return (q[..., 0::2] | (q[..., 1::2] << 4)).astype(np.uint8)


def unpack_nibbles(b: np.ndarray) -> np.ndarray:
    b = np.asarray(b, np.uint8)
    out = np.empty(b.shape[:-1] + (b.shape[-1] * 2,), np.uint8)
    out[..., 0::2] = b & 0xF
    out[..., 1::2] = b >> 4
    return out


def kv4_dequant_f16(q: np.ndarray, scale_f16, zero_f16) -> np.ndarray:
    """Reference dequant: ``__hfma2(half(q), half(scale), half(-scale*zero))`` (Utils.h:2196-2211)."""
    s = np.asarray(scale_f16, f16)
    z = np.asarray(zero_f16, f16)
    hz = (-(s.astype(f32)) * z.astype(f32)).astype(f16)
    v = np.asarray(q).astype(f64) * s.astype(f64)[..., None] + hz.astype(f64)[..., None]
    return v.astype(f16)


def kv4_dequant_exact(q: np.ndarray, scale_f16, zero_f16) -> np.ndarray:
    """Mathematically exact dequant (q - zero) * scale in float64 (for the pure-precision oracle)."""
    s = np.asarray(scale_f16, f16).astype(f64)[..., None]
    z = np.asarray(zero_f16, f16).astype(f64)[..., None]
    return (np.asarray(q).astype(f64) - z) * s


# ----------------------------------------------------------------------------------------------
# RoPE
# ----------------------------------------------------------------------------------------------
def rope_neox(x_f16: np.ndarray, pos, rotary_dim: int, base: float, scale: float = 1.0) -> np.ndarray:
    """NeoX-style rotary on the last axis of x fp16 [..., Dh]; ``pos`` broadcastable to x.shape[:-1].

    inv_freq = (pos*scale) / base^(2i/rot_dim) in fp32, cos/sin fp32, rotate in fp32, round to fp16
    (Utils.h:1147-1167).  ``scale`` is the *already inverted* linear factor (1/factor).
    """
    x = np.asarray(x_f16, f16).astype(f32)
    out = x.copy()
    half = rotary_dim // 2
    i = np.arange(half, dtype=f32)
    denom = np.power(f32(base), (2.0 * i / f32(rotary_dim)).astype(f32)).astype(f32)
    p = np.asarray(pos, f32)[..., None]
    ang = ((p * f32(scale)) / denom).astype(f32)
    c, s = np.cos(ang.astype(f64)).astype(f32), np.sin(ang.astype(f64)).astype(f32)
    a, b = x[..., :half], x[..., half:rotary_dim]
    out[..., :half] = c * a - s * b
    out[..., half:rotary_dim] = c * b + s * a
    return out.astype(f16)


# ----------------------------------------------------------------------------------------------
# writers
# ----------------------------------------------------------------------------------------------
def write_token(cache: PagedKV4, which: str, page: int, slot: int, head_rank: int, x_f16: np.ndarray):
    q, s, z = kv4_quant(x_f16)
    cache.data(which, page)[head_rank, slot, :] = pack_nibbles(q)
    cache.scales(which, page)[head_rank, slot] = s
    cache.zeros(which, page)[head_rank, slot] = z


def prefill_write(cache: PagedKV4, block_table: np.ndarray, qkv_f16: np.ndarray, seq_lens, n_q_heads: int,
                  rotary_dim: int, base: float, scale: float = 1.0):
    """apply_bias_rope_update_kv_cache for dense (all-retrieval) heads, no bias
    (fine_grained_common/applyBiasRopeUpdateKVCache.h:99-556): in-place NeoX RoPE of q and k inside
    the packed qkv buffer [T, (Hq+2Hkv)*Dh]; quantise post-RoPE k and raw v into the pages.
    Sequences are packed back to back (unpadded).  Mutates ``qkv_f16`` and ``cache``.
    """
    H, Dh = cache.H, cache.Dh
    t0 = 0
    for b, L in enumerate(seq_lens):
        blk = qkv_f16[t0 : t0 + L]
        pos = np.arange(L)
        qv = blk[:, : n_q_heads * Dh].reshape(L, n_q_heads, Dh)
        kv = blk[:, n_q_heads * Dh : (n_q_heads + H) * Dh].reshape(L, H, Dh)
        vv = blk[:, (n_q_heads + H) * Dh :].reshape(L, H, Dh)
        qv[:] = rope_neox(qv, pos[:, None], rotary_dim, base, scale)
        kv[:] = rope_neox(kv, pos[:, None], rotary_dim, base, scale)
        for t in range(L):
            page, slot = int(block_table[b, t // TOKENS_PER_BLOCK]), t % TOKENS_PER_BLOCK
            for h in range(H):
                write_token(cache, "k", page, slot, h, kv[t, h])
                write_token(cache, "v", page, slot, h, vv[t, h])
        t0 += L


def fill_random(cache: PagedKV4, block_table: np.ndarray, seq_lens, rng: np.random.Generator):
    """Fill pages 0..len-1 of each sequence with N(0,1) K/V through the oracle quantiser.
    Returns the fp16 (k, v) that were quantised, as lists of [L, H, Dh] arrays."""
    ks, vs = [], []
    for b, L in enumerate(seq_lens):
        k = rng.standard_normal((L, cache.H, cache.Dh)).astype(f16)
        v = rng.standard_normal((L, cache.H, cache.Dh)).astype(f16)
        qk, sk, zk = kv4_quant(k)
        qv, sv, zv = kv4_quant(v)
        for t in range(L):
            page, slot = int(block_table[b, t // TOKENS_PER_BLOCK]), t % TOKENS_PER_BLOCK
            cache.data("k", page)[:, slot, :] = pack_nibbles(qk[t])
            cache.scales("k", page)[:, slot] = sk[t]
            cache.zeros("k", page)[:, slot] = zk[t]
            cache.data("v", page)[:, slot, :] = pack_nibbles(qv[t])
            cache.scales("v", page)[:, slot] = sv[t]
            cache.zeros("v", page)[:, slot] = zv[t]
        ks.append(k)
        vs.append(v)
    return ks, vs


def gather_seq(cache: PagedKV4, which: str, block_row: np.ndarray, positions, head_rank: int):
    """Read (q u4 [n, Dh], scale f16 [n], zero f16 [n]) of the given token positions of one head."""
    positions = np.asarray(positions)
    pages = np.asarray(block_row)[positions // TOKENS_PER_BLOCK].astype(np.int64)
    slots = positions % TOKENS_PER_BLOCK
    pool = cache._pool(which)
    H, Dh = cache.H, cache.Dh
    d = pool[:, : cache.data_bytes].reshape(cache.P, H, TOKENS_PER_BLOCK, Dh // 2)[pages, head_rank, slots]
    o = cache.data_bytes
    sc = pool[:, o : o + cache.sz_bytes // 2].view(f16).reshape(cache.P, H, TOKENS_PER_BLOCK)[pages, head_rank, slots]
    o2 = o + cache.sz_bytes // 2
    ze = pool[:, o2 : o2 + cache.sz_bytes // 2].view(f16).reshape(cache.P, H, TOKENS_PER_BLOCK)[pages, head_rank, slots]
    return unpack_nibbles(d), sc, ze


# ----------------------------------------------------------------------------------------------
# decode attention
# ----------------------------------------------------------------------------------------------
def _qk_cached_mimic(q_r: np.ndarray, kd: np.ndarray) -> np.ndarray:
    """Reference cached-key dot (Template.hpp:450-467): 16 lanes x 8 elements; per lane two fp16 FMA
    chains (elements 0-3 and 4-7), ``__hadd`` of the two, fp32 butterfly (xor 8,4,2,1) over lanes."""
    n = kd.shape[0]
    qq = q_r.astype(f64).reshape(16, 8)
    kk = kd.astype(f64).reshape(n, 16, 8)
    lo = (qq[None, :, 0] * kk[:, :, 0]).astype(f16)
    hi = (qq[None, :, 4] * kk[:, :, 4]).astype(f16)
    for i in (1, 2, 3):
        lo = (qq[None, :, i] * kk[:, :, i] + lo.astype(f64)).astype(f16)
        hi = (qq[None, :, 4 + i] * kk[:, :, 4 + i] + hi.astype(f64)).astype(f16)
    v = (lo.astype(f32) + hi.astype(f32)).astype(f16).astype(f32)  # [n, 16]
    lanes = np.arange(16)
    for m in (8, 4, 2, 1):
        v = (v + v[:, lanes ^ m]).astype(f32)
    return v[:, 0]


def decode_attention(q, k, v, cache: PagedKV4, block_table, lengths, rotary_dim: int, base: float,
                     scale: float = 1.0, mimic: bool = True, positions_fn=None, head_rank=None,
                     append: bool = True, update_stats_sub_chunk: int = 0):
    """``single_query_attention`` of fused_attention_pure_dense (and, through ``positions_fn`` /
    ``head_rank``, the LServe masked variants).

    q fp16 [B,Hq,Dh]; k, v fp16 [B,Hkv,Dh] (new token, pre-RoPE); lengths int [B] = ctx INCLUDING the
    new token.  Appends the new token's quantised K (post-RoPE) / V into ``cache`` (if ``append``)
    and returns fp16 [B,Hq,Dh].

    mimic=True reproduces the reference's rounding points (fp16 dequant, fp16 q.k partials, fp16
    probabilities, fp32 P.V per 16-token-stride group, fp16 tree reduce over the 16 groups);
    mimic=False is the exact-arithmetic oracle (float64 everywhere, fp16 only at the output).

    positions_fn(b, hq, tl) -> int array of *cached* token positions to attend (default: all 0..tl-1);
    head_rank[hkv] -> row of that kv head inside the page (default: identity).
    update_stats_sub_chunk > 0 (fused_attention_fine_grained_sparse, retrieval heads): the appended post-RoPE key
    is folded into the page's kmax / kmin of its sub-chunk, element-wise against the stored values
    (sparse_attention/decoderMaskedMultiheadAttentionTemplate.hpp:1414-1429).
    """
    q = np.asarray(q, f16)
    k = np.asarray(k, f16)
    v = np.asarray(v, f16)
    B, Hq, Dh = q.shape
    Hkv = k.shape[1]
    g = Hq // Hkv
    inv_sqrt = f32(1.0 / np.sqrt(f32(Dh)))
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
            if n:
                qk_q, qk_s, qk_z = gather_seq(cache, "k", block_table[b], pos, rank)
                qv_q, qv_s, qv_z = gather_seq(cache, "v", block_table[b], pos, rank)
            if mimic:
                logits = np.empty(n + 1, f32)
                if n:
                    kd = kv4_dequant_f16(qk_q, qk_s, qk_z)
                    logits[:n] = _qk_cached_mimic(q_r[hq], kd) * inv_sqrt
                logits[n] = f32((q_r[hq].astype(f64) * k_r[hk].astype(f64)).sum()) * inv_sqrt
                mxl = logits.max()
                e = np.exp((logits - mxl).astype(f64)).astype(f32)
                inv_sum = f32(1.0) / (e.sum(dtype=f64).astype(f32) + f32(1e-6))
                p = (e * inv_sum).astype(f16)
                acc = np.zeros((16, Dh), f32)  # 16 token groups (V_PER_ITER), fp32 accumulators
                if n:
                    vd = kv4_dequant_f16(qv_q, qv_s, qv_z).astype(f32)
                    for j in range(n):  # ordinal j within the attended list plays the role of ti
                        acc[j % 16] = (p[j].astype(f32) * vd[j] + acc[j % 16]).astype(f32)
                acc[n % 16] = (p[n].astype(f32) * v[b, hk].astype(f32) + acc[n % 16]).astype(f32)
                active = 16
                while active >= 2:
                    mid = active // 2
                    upper = acc[mid:active].astype(f16).astype(f32)
                    acc[:mid] = (upper + acc[:mid]).astype(f32)
                    active = mid
                out[b, hq] = acc[0].astype(f16)
            else:
                logits = np.empty(n + 1, f64)
                if n:
                    kd = kv4_dequant_exact(qk_q, qk_s, qk_z)
                    logits[:n] = kd @ q_r[hq].astype(f64)
                logits[n] = (q_r[hq].astype(f64) * k_r[hk].astype(f64)).sum()
                logits *= 1.0 / np.sqrt(Dh)
                e = np.exp(logits - logits.max())
                p = e / e.sum()
                o = p[n] * v[b, hk].astype(f64)
                if n:
                    o = o + p[:n] @ kv4_dequant_exact(qv_q, qv_s, qv_z)
                out[b, hq] = o.astype(f16)
        if append:
            for hk in range(Hkv):
                rank = hk if head_rank is None else int(head_rank[hk])
                write_token(cache, "k", page, slot, rank, k_r[hk])
                write_token(cache, "v", page, slot, rank, v[b, hk])
                if update_stats_sub_chunk:
                    kmax, kmin = cache.kstats(page)
                    sub = slot // update_stats_sub_chunk
                    sl = slice(rank * Dh, (rank + 1) * Dh)
                    kmax[sub, sl] = np.maximum(kmax[sub, sl], k_r[hk])
                    kmin[sub, sl] = np.minimum(kmin[sub, sl], k_r[hk])
    return out


# ----------------------------------------------------------------------------------------------
# LServe pieces (page statistics, selector score, top-k page choice)
# ----------------------------------------------------------------------------------------------
def paged_min_max_pool(cache: PagedKV4, block_table, keys_post_rope_f16, seq_lens, sub_chunk: int = 16,
                       pooling_heads_idx=None):
    """paged_min_max_pool (sparse_utils/ContextPool/context_pool_kernel.cu:16-69): per ``sub_chunk``-token
    sub-chunk channel-wise max/min of post-RoPE K written after the scales/zeros of each K page.
    keys: fp16 [T, H_in, Dh] (sequences packed back to back); pooled head j reads input head
    ``pooling_heads_idx[j]`` (default identity) and writes row j of the statistics (:31,:45).  A partial last
    sub-chunk pools its valid tokens only (the clamp of :39 repeats the last token)."""
    t0 = 0
    n_sub = TOKENS_PER_BLOCK // sub_chunk
    assert cache.n_sub == n_sub
    idx = np.arange(cache.H) if pooling_heads_idx is None else np.asarray(pooling_heads_idx)
    assert len(idx) == cache.H
    keys = np.asarray(keys_post_rope_f16, f16)
    keys = keys.reshape(keys.shape[0], -1, cache.Dh)
    for b, L in enumerate(seq_lens):
        kk = keys[t0 : t0 + L][:, idx, :].reshape(L, cache.H * cache.Dh)
        for c0 in range(0, L, sub_chunk):
            page = int(block_table[b, c0 // TOKENS_PER_BLOCK])
            sub = (c0 % TOKENS_PER_BLOCK) // sub_chunk
            kmax, kmin = cache.kstats(page)
            chunk = kk[c0 : min(c0 + sub_chunk, L)]
            kmax[sub] = chunk.max(axis=0)
            kmin[sub] = chunk.min(axis=0)
        t0 += L


def _score_fp16(q_r: np.ndarray, kmax: np.ndarray, kmin: np.ndarray) -> np.ndarray:
    """qk_hmma_dot_min_max<16> (KVPageSelectorTemplate.hpp:478-503) for rows kmax/kmin [n, Dh]:
    16 lanes x 8 channels; per lane four half2 products rounded to fp16, ``__hmax2``, sequential fp16 ``__hadd2``
    over the four pairs, ``__hadd`` of the two halves, then an fp32 butterfly (xor 8,4,2,1) across the 16 lanes."""
    n = kmax.shape[0]
    qq = q_r.astype(f32).reshape(1, 16, 4, 2)
    a = (qq * kmax.astype(f32).reshape(n, 16, 4, 2)).astype(f16)    # fp16 x fp16 product is exact in fp32 -> one rounding
    b = (qq * kmin.astype(f32).reshape(n, 16, 4, 2)).astype(f16)
    m = np.maximum(a, b)
    acc = m[:, :, 0, :]
    for i in (1, 2, 3):
        acc = (acc.astype(f32) + m[:, :, i, :].astype(f32)).astype(f16)
    v = (acc[:, :, 0].astype(f32) + acc[:, :, 1].astype(f32)).astype(f16).astype(f32)   # [n, 16]
    lanes = np.arange(16)
    for msk in (8, 4, 2, 1):
        v = (v + v[:, lanes ^ msk]).astype(f32)
    return v[:, 0].astype(f16)


def page_selector(q, cache: PagedKV4, block_table, lengths, timestep: int, rotary_dim: int, base: float,
                  scale: float = 1.0, retrieval_flags=None, head_rank=None, n_kv_heads: int | None = None,
                  sub_chunk: int = 16):
    """single_query_page_selector (fused_kv_page_selector.cpp:171-334 + KVPageSelectorTemplate.hpp:786-1290).

    q fp16 [B,Hq,Dh] pre-RoPE; lengths [B] incl. the new token (None -> timestep + 1).  Returns fp16
    [B, Hq, padded(timestep)] zero-initialised (:274-277); retrieval heads get score[sub] for sub < ceil(tl/sub_chunk)
    at row pitch ``padded(tl_b)`` -- the kernel derives the pitch from the sample's own length (:1130-1133), which
    equals the host's only when lengths[b] == timestep + 1 (kept as is; writes past the tensor are dropped here).
    Streaming heads are skipped (:791-794)."""
    q = np.asarray(q, f16)
    B, Hq, Dh = q.shape
    Hkv = n_kv_heads if n_kv_heads is not None else (len(retrieval_flags) if retrieval_flags is not None else cache.H)
    g = Hq // Hkv
    grp = TOKENS_PER_BLOCK // sub_chunk
    n_sub_host = (timestep + sub_chunk - 1) // sub_chunk
    padded_host = (n_sub_host + grp - 1) // grp * grp
    out = np.zeros(B * Hq * padded_host, f16)
    for b in range(B):
        tl = timestep if lengths is None else int(lengths[b]) - 1
        n_sub = (tl + sub_chunk - 1) // sub_chunk
        padded = (n_sub + grp - 1) // grp * grp
        q_r = rope_neox(q[b], tl, rotary_dim, base, scale)
        for hq in range(Hq):
            hk = hq // g
            if retrieval_flags is not None and not retrieval_flags[hk]:
                continue
            rank = hk if head_rank is None else int(head_rank[hk])
            sl = slice(rank * Dh, (rank + 1) * Dh)
            rows_max = np.empty((n_sub, Dh), f16)
            rows_min = np.empty((n_sub, Dh), f16)
            for sc in range(n_sub):
                kmax, kmin = cache.kstats(int(block_table[b, sc // grp]))
                rows_max[sc] = kmax[sc % grp, sl]
                rows_min[sc] = kmin[sc % grp, sl]
            sc_vals = _score_fp16(q_r[hq], rows_max, rows_min) if n_sub else np.zeros(0, f16)
            o = (b * Hq + hq) * padded
            hi = min(o + n_sub, out.size)
            if hi > o:
                out[o:hi] = sc_vals[: hi - o]
    return out.reshape(B, Hq, padded_host)


def select_topk_pages(stats, timestep: int, token_budget: int, sub_chunk: int = 16):
    """The Python half of the selector (omniserve/modeling/layers/decoding_attention.py:88-143):
    stats fp16 [B,Hq,padded] -> int32 [B,Hq,P]: max over the sub-chunks of a page, top-(k-1) over all pages but the
    last (k = min(max(3, budget/64), pages)), the newest page appended last.  Ties are broken towards the lower
    page index here (torch.topk's order among equal values is unspecified)."""
    stats = np.asarray(stats)
    B, Hq, padded = stats.shape
    grp = TOKENS_PER_BLOCK // sub_chunk
    if timestep <= token_budget:
        n = timestep // TOKENS_PER_BLOCK + 1
        return np.broadcast_to(np.arange(n, dtype=np.int32), (B, Hq, n)).copy()
    budget = min(token_budget, timestep)
    page_stats = stats.reshape(B, Hq, padded // grp, grp).astype(f32).max(axis=-1)
    total = page_stats.shape[-1]
    kk = min(max(3, budget // TOKENS_PER_BLOCK), total) - 1
    order = np.argsort(-page_stats[:, :, :-1], axis=-1, kind="stable")[:, :, :kk]
    last = np.full((B, Hq, 1), total - 1, np.int64)
    return np.concatenate([order, last], axis=-1).astype(np.int32)
