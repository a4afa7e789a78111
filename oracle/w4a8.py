"""CPU restatement of the reference's W4A8 quantiser, packer and GEMM semantics.

TEST INFRASTRUCTURE -- see oracle/__init__.py.  numpy only.

Reference files restated (all under /root/reference):
  * fake-quant:            scripts/ckpt_converter/quant_utils.py:96-138
  * zero-point shift:      scripts/ckpt_converter/checkpoint_converter.py:105-107
  * packer (per-channel):  omniserve/modeling/layers/quantized_linear/w4a8_linear.py:284-335
  * packer (per-group):    .../w4a8_linear.py:170-282
  * GEMM per-channel:      kernels/csrc/qgemm/w4a8_per_chn/gemm_cuda.cu:281-306 (unpack), :569-598 (epilogue)
  * GEMM per-group:        kernels/csrc/qgemm/w4a8_per_group/gemm_cuda.cu:276-331 (dequant), :605-632 (epilogue)
"""
from __future__ import annotations

import numpy as np

G = 128  # group size of the per-group variant (gemm_cuda.cu:657 `constexpr int G = 128`)


# ----------------------------------------------------------------------------------------------
# fake quantisation (quant_utils.py:96-138), zero_point=True branch
# ----------------------------------------------------------------------------------------------
def pseudo_quantize_tensor(w: np.ndarray, n_bit: int = 4, q_group_size: int = -1):
    """Returns (w_fake, scales, zeros) exactly like ``pseudo_quantize_tensor(get_scale_zp=True)``.

    ``w`` float32 [N, K].  scales/zeros come back as [N, K/group] float32.
    """
    w = np.asarray(w, dtype=np.float32)
    org_shape = w.shape
    if q_group_size > 0:
        assert org_shape[-1] % q_group_size == 0
        w2 = w.reshape(-1, q_group_size)
    else:
        w2 = w.reshape(org_shape[0], -1)
    max_val = w2.max(axis=1, keepdims=True)
    min_val = w2.min(axis=1, keepdims=True)
    max_int = 2**n_bit - 1
    scales = np.maximum(max_val - min_val, np.float32(1e-5)) / np.float32(max_int)
    zeros = np.clip(-np.round(min_val / scales), 0, max_int)  # torch.round == np.round (half-to-even)
    q = np.clip(np.round(w2 / scales) + zeros, 0, max_int)
    w_fake = ((q - zeros) * scales).astype(np.float32).reshape(org_shape)
    return w_fake, scales.reshape(org_shape[0], -1), zeros.reshape(org_shape[0], -1)


def quantize_per_channel(w: np.ndarray, s1: np.ndarray, zeros: np.ndarray) -> np.ndarray:
    """w4a8_linear.py:286-294: ``q = round(w / s1).to(int8) + zeros``; returns uint8 in [0, 15]."""
    q = np.round(np.asarray(w, np.float32) / np.asarray(s1, np.float32).reshape(-1, 1)).astype(np.int8)
    q = q + np.asarray(zeros).reshape(-1, 1).astype(np.int8)
    assert q.min() >= 0 and q.max() <= 15, "Quantized weight out of range"
    return q.astype(np.uint8)


# ----------------------------------------------------------------------------------------------
# the tile packer (w4a8_linear.py:297-327; identical in both branches)
# ----------------------------------------------------------------------------------------------
def pack_w4(q: np.ndarray) -> np.ndarray:
    """uint4 values [N, K] (uint8 storage) -> int8 [N, K/2] in the reference layout.

    Contiguous as [N/32][K/32][32 lanes][16 B]; lane = c*4+e, byte = d*8+b*4+f holds
    ``(W[n+16,k] << 4) | W[n,k]`` with n = n32*32 + b*8 + c, k = k32*32 + d*16 + e*4 + f.
    """
    q = np.asarray(q)
    N, K = q.shape
    assert N % 32 == 0 and K % 32 == 0
    # reshape(N/32, 2[hi], 2[b], 8[c], K/32, 2[d], 4[e], 4[f]).permute(0,4,3,6,1,5,2,7)
    r = q.reshape(N // 32, 2, 2, 8, K // 32, 2, 4, 4).transpose(0, 4, 3, 6, 1, 5, 2, 7)
    # -> [n32, k32, c, e, hi, d, b, f]; .permute(0,1,2,3,5,6,7,4) -> [n32, k32, c, e, d, b, f, hi]
    r = r.transpose(0, 1, 2, 3, 5, 6, 7, 4).astype(np.int8)
    packed = ((r[..., 1].astype(np.int16) << 4) + r[..., 0]).astype(np.uint8).view(np.int8)
    return np.ascontiguousarray(packed.reshape(N // 32, K // 32, 32, 16).reshape(N, K // 2))


def unpack_w4(packed: np.ndarray) -> np.ndarray:
    """Inverse of :func:`pack_w4`: int8 [N, K/2] -> uint8 [N, K] values in [0, 15].

    Follows the consumer's view (gemm_cuda.cu:291-303): per 16-byte lane, words (x,y,z,w) =
    (d0b0, d0b1, d1b0, d1b1); ``& 0x0F0F0F0F`` are rows n, ``>> 4`` rows n+16.
    """
    p = np.asarray(packed).view(np.uint8)
    N, K2 = p.shape
    K = K2 * 2
    r = p.reshape(N // 32, K // 32, 8, 4, 2, 2, 4)  # [n32, k32, c, e, d, b, f]
    lo = r & 0xF
    hi = r >> 4
    both = np.stack([lo, hi], axis=0)  # [hi, n32, k32, c, e, d, b, f]
    # target [n32, hi, b, c, k32, d, e, f]
    out = both.transpose(1, 0, 6, 3, 2, 5, 4, 7)
    return np.ascontiguousarray(out.reshape(N, K))


def pack_s2(x: np.ndarray) -> np.ndarray:
    """[N, K/G] -> [K/G, N] with the N axis permuted inside each 32-block: pos = c*4 + j <-> n = j*8 + c.

    w4a8_linear.py:236-253 (``reshape(.., 4, 8).transpose(-2, -1)``).
    """
    x = np.asarray(x)
    N, ng = x.shape
    t = x.T.reshape(ng, N // 32, 4, 8).transpose(0, 1, 3, 2)
    return np.ascontiguousarray(t.reshape(ng, N))


def unpack_s2(p: np.ndarray) -> np.ndarray:
    p = np.asarray(p)
    ng, N = p.shape
    t = p.reshape(ng, N // 32, 8, 4).transpose(0, 1, 3, 2).reshape(ng, N)
    return np.ascontiguousarray(t.T)


def quantize_per_group(w, s1, s2, zeros, group_size: int = G):
    """Two-level quantiser of w4a8_linear.py:170-282.

    w [N,K] f32; s1 [N] (level-1 per-channel scale); s2 [N, K/G] (u8-valued level-2 scale);
    zeros [N, K/G] (u4 zero point, *unscaled*).  Returns (qweight int8 [N,K/2],
    s2_scales int8 [K/G,N], s2_zeros int8 [K/G,N] (= -z*s2, two's complement), q u4 [N,K]).
    """
    w = np.asarray(w, np.float32)
    N, K = w.shape
    ng = K // group_size
    lw = np.round(w / np.asarray(s1, np.float32).reshape(N, 1))
    assert lw.min() >= -128 and lw.max() <= 127, "Stage 1: Quantized weight out of range"
    lw = lw.reshape(N, ng, group_size)
    s2h = np.asarray(s2, np.float32).astype(np.float16).astype(np.float32).reshape(N, ng, 1)
    zh = np.asarray(zeros, np.float32).astype(np.float16).astype(np.float32).reshape(N, ng, 1)
    q = lw / s2h + zh
    assert q.min() >= 0 and q.max() <= 15, "Stage 2: Quantized weight out of range"
    q = q.reshape(N, K).astype(np.int8).astype(np.uint8)  # .to(torch.int8) truncates toward zero
    qweight = pack_w4(q)
    s2_p = pack_s2(np.asarray(s2).reshape(N, ng).astype(np.int64))
    z_p = pack_s2((-np.asarray(zeros).reshape(N, ng)).astype(np.int32).astype(np.int64))
    s2_zeros = (z_p * s2_p).astype(np.int8)  # assignment into an int8 buffer wraps
    return qweight, s2_p.astype(np.int8), s2_zeros, q


# ----------------------------------------------------------------------------------------------
# GEMM semantics
# ----------------------------------------------------------------------------------------------
def _exact_int_gemm(a_i8: np.ndarray, w_i: np.ndarray) -> np.ndarray:
    """sum_k a[m,k]*w[n,k] exactly.  float64 BLAS is exact here (|sum| < 2^53)."""
    acc = a_i8.astype(np.float64) @ w_i.astype(np.float64).T
    return np.rint(acc).astype(np.int64).astype(np.int32)  # s32 accumulate wraps like the MMA


def gemm_per_chn(in_feats, qweight, wscales, ascales, w_szs, a_ssums):
    """``C = (A . Wu4^T) * s1[n] * sa[m] - sz[n] * ssum[m]`` (per_chn/gemm_cuda.cu:583-590).

    Returns (acc int32 [M,N], out float16 [M,N]).  The unsigned 4-bit weights enter the MMA as
    s8 values 0..15 with *no* zero-point subtraction in the main loop (:291-303).
    Epilogue evaluated in fp32 in the reference's association order and rounded once to fp16.
    """
    a = np.asarray(in_feats, np.int8)
    w = unpack_w4(qweight).astype(np.int32)
    acc = _exact_int_gemm(a, w)
    ws = np.asarray(wscales, np.float16).astype(np.float32)[None, :]
    sa = np.asarray(ascales, np.float16).astype(np.float32)[:, None]
    sz = np.asarray(w_szs, np.float16).astype(np.float32)[None, :]
    ss = np.asarray(a_ssums, np.float16).astype(np.float32)[:, None]
    ps = acc.astype(np.float32)  # __int2float_rn
    out = (ps * ws) * sa - sz * ss
    return acc, out.astype(np.float16)


def dequant_per_group_w8(qweight, zeros_i8, scales_i8, group_size: int = G) -> np.ndarray:
    """Level-2 dequant exactly as per_group/gemm_cuda.cu:289-329.

    Four packed u4 (one per byte of a 32-bit word) are multiplied by the u8 scale with ONE 32-bit
    multiply (so a byte product > 255 carries into its neighbour), then ``__vadd4`` adds the
    replicated zero byte modulo 256.  The four bytes of a word are 4 consecutive k of one row.
    Returns int8 [N, K].
    """
    q = unpack_w4(qweight).astype(np.uint64)  # [N, K]
    N, K = q.shape
    s = unpack_s2(np.asarray(scales_i8).view(np.uint8)).astype(np.uint64)  # [N, K/G], unsigned byte
    z = unpack_s2(np.asarray(zeros_i8).view(np.uint8)).astype(np.uint64)
    s = np.repeat(s, group_size, axis=1)
    z = np.repeat(z, group_size, axis=1)
    words = q.reshape(N, K // 4, 4)
    word32 = words[..., 0] | (words[..., 1] << 8) | (words[..., 2] << 16) | (words[..., 3] << 24)
    prod = (word32 * s.reshape(N, K // 4, 4)[..., 0]) & 0xFFFFFFFF
    pb = np.stack([(prod >> (8 * i)) & 0xFF for i in range(4)], axis=-1)
    w8 = (pb + z.reshape(N, K // 4, 4)) & 0xFF
    return w8.reshape(N, K).astype(np.uint8).view(np.int8)


def gemm_per_group(in_feats, qweight, zeros_i8, scales_i8, wscales, ascales, group_size: int = G):
    """``C = (A . w8^T) * (s1[n] * sa[m])`` (per_group/gemm_cuda.cu:622-627)."""
    a = np.asarray(in_feats, np.int8)
    w8 = dequant_per_group_w8(qweight, zeros_i8, scales_i8, group_size).astype(np.int32)
    acc = _exact_int_gemm(a, w8)
    ws = np.asarray(wscales, np.float16).astype(np.float32)[None, :]
    sa = np.asarray(ascales, np.float16).astype(np.float32)[:, None]
    out = acc.astype(np.float32) * (ws * sa)
    return acc, out.astype(np.float16)


# ----------------------------------------------------------------------------------------------
# BASELINE config 1: per-channel quant -> pack -> unpack -> dequant round trip
# ----------------------------------------------------------------------------------------------
def gemm_w8a8(in_feats, weight, wscales, ascales):
    """W8A8 GEMM of kernels/csrc/qgemm/w8a8/w8a8_gemm_cuda.cu:515-530: ``C = rn_f16( f32(A . W^T) * (ws[n] * as[m]) )`` with
    the scale product formed first in fp32 (``psum *= wscale * ascale``).  in_feats int8 [M,K], weight int8 [N,K] row-major.
    Returns (acc int32 [M,N], out float16 [M,N])."""
    a = np.asarray(in_feats, np.int8)
    w = np.asarray(weight, np.int8)
    acc = _exact_int_gemm(a, w.astype(np.int32))
    ws = np.asarray(wscales, np.float16).astype(np.float32)[None, :]
    sa = np.asarray(ascales, np.float16).astype(np.float32)[:, None]
    out = (acc.astype(np.float32) * (ws * sa).astype(np.float32)).astype(np.float32)
    return acc, out.astype(np.float16)


def moe_gemm_per_chn(x, qweights, wscales, ascales, w_szs, a_ssums, problem_sizes):
    """Grouped per-channel GEMM of a mixture-of-experts layer (interface of w4a8_moe_linear.py:83-94; the reference never
    released the kernel, so this restates the only possible semantics: rows of ``x`` are sorted by expert, expert ``e`` owns
    ``problem_sizes[e]`` consecutive rows and applies :func:`gemm_per_chn` with its own packed weight / scales).
    x int8 [T,K]; qweights int8 [E,N,K/2]; wscales, w_szs f16 [E,N]; ascales, a_ssums f16 [T].  Returns f16 [T,N]."""
    x = np.asarray(x, np.int8)
    E, N = np.asarray(wscales).shape
    out = np.zeros((x.shape[0], N), np.float16)
    r = 0
    for e in range(E):
        m = int(problem_sizes[e])
        if m:
            _, o = gemm_per_chn(x[r:r + m], qweights[e], wscales[e], ascales[r:r + m], w_szs[e], a_ssums[r:r + m])
            out[r:r + m] = o
        r += m
    assert r == x.shape[0]
    return out


def roundtrip_per_channel(w: np.ndarray):
    """Returns (w_fake, w_roundtrip, packed).  ``w_roundtrip`` must equal ``w_fake`` exactly."""
    w_fake, scales, zeros = pseudo_quantize_tensor(w, n_bit=4, q_group_size=-1)
    q = quantize_per_channel(w_fake, scales[:, 0], zeros[:, 0])
    packed = pack_w4(q)
    q2 = unpack_w4(packed).astype(np.float32)
    w_rt = ((q2 - zeros) * scales).astype(np.float32)
    return w_fake, w_rt, packed
