"""Pin the W4A8 oracle against golden vectors produced by the reference's own Python packer
(tests/golden/make_golden.py) and against structural properties."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import w4a8

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_per_channel_golden_small():
    g = np.load(f"{GOLD}/w4a8_per_chn_64x128.npz")
    w_fake, scales, zeros = w4a8.pseudo_quantize_tensor(g["w"], 4, -1)
    np.testing.assert_array_equal(w_fake, g["w_fake"])
    np.testing.assert_array_equal(scales, g["scales"])
    np.testing.assert_array_equal(zeros, g["zeros"])
    q = w4a8.quantize_per_channel(w_fake, scales[:, 0], zeros[:, 0])
    np.testing.assert_array_equal(w4a8.pack_w4(q), g["qweight"])
    np.testing.assert_array_equal(w4a8.unpack_w4(g["qweight"]), q)
    # s1_szeros = zeros * s1 in fp16 (w4a8_linear.py:333-335)
    szs = (zeros[:, 0] * scales[:, 0]).astype(np.float16)
    np.testing.assert_array_equal(szs, g["s1_szeros"])


def test_per_channel_golden_4096_checksum():
    """BASELINE config 1 at full size: the packed bytes must hash to what the reference produced."""
    meta = json.load(open(f"{GOLD}/golden.json"))["per_chn_4096"]
    rng = np.random.default_rng(meta["seed"])
    w = (rng.standard_normal((4096, 4096)) * 0.02).astype(np.float32)
    w_fake, w_rt, packed = w4a8.roundtrip_per_channel(w)
    assert hashlib.sha256(w_fake.tobytes()).hexdigest() == meta["w_fake_sha256"]
    assert hashlib.sha256(packed.tobytes()).hexdigest() == meta["qweight_sha256"]
    np.testing.assert_array_equal(w_rt, w_fake)  # quant -> pack -> unpack -> dequant is exact


def test_per_group_golden_small():
    g = np.load(f"{GOLD}/w4a8_per_group_64x256.npz")
    qweight, s2s, s2z, q = w4a8.quantize_per_group(g["w"], g["s1"], g["s2"], g["zeros"])
    np.testing.assert_array_equal(qweight, g["qweight"])
    np.testing.assert_array_equal(s2s, g["s2_scales"])
    np.testing.assert_array_equal(s2z, g["s2_zeros"])
    # the in-register dequant must give back the level-1 int8 weight the fixture was built from
    w8 = w4a8.dequant_per_group_w8(g["qweight"], g["s2_zeros"], g["s2_scales"])
    np.testing.assert_array_equal(w8.astype(np.float32), g["w8"])


@pytest.mark.parametrize("N,K", [(32, 32), (64, 96), (128, 256)])
def test_pack_bijection_and_index_formula(N, K):
    rng = np.random.default_rng(N * 1000 + K)
    q = rng.integers(0, 16, (N, K), dtype=np.uint8)
    p = w4a8.pack_w4(q)
    np.testing.assert_array_equal(w4a8.unpack_w4(p), q)
    # explicit (lane, byte, nibble) <-> (n, k) formula of SURVEY 8(a1)
    blk = p.view(np.uint8).reshape(N // 32, K // 32, 32, 16)
    for _ in range(200):
        n32, k32 = rng.integers(N // 32), rng.integers(K // 32)
        c, e, d, b, f, hi = (rng.integers(x) for x in (8, 4, 2, 2, 4, 2))
        byte = blk[n32, k32, c * 4 + e, d * 8 + b * 4 + f]
        val = (byte >> 4) if hi else (byte & 0xF)
        assert val == q[n32 * 32 + hi * 16 + b * 8 + c, k32 * 32 + d * 16 + e * 4 + f]


def test_s2_permutation():
    x = np.arange(64 * 3).reshape(64, 3)
    p = w4a8.pack_s2(x)
    np.testing.assert_array_equal(w4a8.unpack_s2(p), x)
    for n32 in range(2):
        for c in range(8):
            for j in range(4):
                assert p[1, n32 * 32 + c * 4 + j] == x[n32 * 32 + j * 8 + c, 1]


def test_gemm_per_chn_matches_dequantised_matmul():
    rng = np.random.default_rng(7)
    M, N, K = 5, 64, 128
    a = rng.integers(-127, 128, (M, K), dtype=np.int8)
    q = rng.integers(0, 16, (N, K), dtype=np.uint8)
    z = rng.integers(0, 16, N).astype(np.float32)
    s1 = rng.uniform(0.005, 0.02, N).astype(np.float16)
    sa = rng.uniform(0.01, 0.05, M).astype(np.float16)
    x = a.astype(np.float32) * sa.astype(np.float32)[:, None]
    ssum = x.sum(1).astype(np.float16)
    szs = (z * s1.astype(np.float32)).astype(np.float16)
    acc, out = w4a8.gemm_per_chn(a, w4a8.pack_w4(q), s1, sa, szs, ssum)
    np.testing.assert_array_equal(acc, a.astype(np.int64) @ q.astype(np.int64).T)
    ref = x @ ((q.astype(np.float32) - z[:, None]) * s1.astype(np.float32)[:, None]).T
    assert np.abs(out.astype(np.float32) - ref).max() <= 2e-2 * np.abs(ref).max()


def test_per_group_wraparound_is_modelled():
    """__vadd4 wraps mod 256 and the packed 32-bit multiply carries across bytes (gemm_cuda.cu:289-329)."""
    N, K = 32, 128
    q = np.full((N, K), 15, np.uint8)
    s2 = np.full((N, 1), 20, np.int64)      # 15*20 = 300 > 255 -> carry into the next byte
    z = np.full((N, 1), 3, np.int64)
    s2p = w4a8.pack_s2(s2).astype(np.int8)
    zp = (w4a8.pack_s2(-z) * w4a8.pack_s2(s2)).astype(np.int8)
    w8 = w4a8.dequant_per_group_w8(w4a8.pack_w4(q), zp, s2p).view(np.uint8)
    word = (0x0F0F0F0F * 20) & 0xFFFFFFFF
    expect = [(((word >> (8 * i)) & 0xFF) + ((-3 * 20) & 0xFF)) & 0xFF for i in range(4)]
    np.testing.assert_array_equal(w8[0, :4], expect)


def test_w8a8_oracle_matches_plain_definition():
    """oracle.w4a8.gemm_w8a8: exact INT32 accumulate, fp32 scale product formed first (w8a8_gemm_cuda.cu:527-529)."""
    from oracle import w4a8 as ow
    rng = np.random.default_rng(0)
    a = rng.integers(-127, 128, (5, 256), dtype=np.int8)
    w = rng.integers(-127, 128, (24, 256), dtype=np.int8)
    ws = rng.uniform(0.002, 0.01, 24).astype(np.float16)
    sa = rng.uniform(0.01, 0.05, 5).astype(np.float16)
    acc, out = ow.gemm_w8a8(a, w, ws, sa)
    np.testing.assert_array_equal(acc, a.astype(np.int64) @ w.astype(np.int64).T)
    want = (acc.astype(np.float32) * (ws.astype(np.float32)[None, :] * sa.astype(np.float32)[:, None])).astype(np.float16)
    np.testing.assert_array_equal(out, want)


def test_moe_oracle_applies_each_experts_gemm_to_its_rows():
    from oracle import w4a8 as ow
    rng = np.random.default_rng(4)
    E, N, K, sizes = 3, 64, 256, [2, 0, 5]
    q = rng.integers(0, 16, (E, N, K), dtype=np.uint8)
    qw = np.stack([ow.pack_w4(q[e]) for e in range(E)])
    x = rng.integers(-127, 128, (7, K), dtype=np.int8)
    s1 = rng.uniform(0.005, 0.02, (E, N)).astype(np.float16)
    szs = (8 * s1.astype(np.float32)).astype(np.float16)
    sa = rng.uniform(0.01, 0.05, 7).astype(np.float16)
    ss = rng.standard_normal(7).astype(np.float16)
    out = ow.moe_gemm_per_chn(x, qw, s1, sa, szs, ss, sizes)
    _, o0 = ow.gemm_per_chn(x[:2], qw[0], s1[0], sa[:2], szs[0], ss[:2])
    _, o2 = ow.gemm_per_chn(x[2:], qw[2], s1[2], sa[2:], szs[2], ss[2:])
    np.testing.assert_array_equal(out[:2], o0)
    np.testing.assert_array_equal(out[2:], o2)
