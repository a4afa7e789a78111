"""Oracle self-checks for the small ops and the KV4 format (CPU)."""
import numpy as np

from oracle import act, kv4


def test_quant_fuse_sum_properties():
    rng = np.random.default_rng(0)
    x = (rng.standard_normal((6, 512)) * 3).astype(np.float16)
    q, s, sm = act.quant_fuse_sum(x)
    assert q.dtype == np.int8 and np.abs(q).max() == 127  # the amax element maps to +-127
    deq = q.astype(np.float32) * s.astype(np.float32)[:, None]
    assert np.abs(deq - x.astype(np.float32)).max() <= 0.51 * s.astype(np.float32).max() + 1e-3
    np.testing.assert_allclose(sm.astype(np.float32), x.astype(np.float32).sum(1), rtol=2e-3, atol=2e-2)


def test_rmsnorm_quirks():
    rng = np.random.default_rng(1)
    x = (rng.standard_normal((3, 1024)) + 0.5).astype(np.float16)  # non-zero mean exposes quirk (i)
    g = np.ones(1024, np.float16)
    q, s, sm, nh = act.rms_norm_general_fuse_sum(x, g, 1e-5)
    xf = x.astype(np.float32)
    expect = (xf - xf.mean(1, keepdims=True)) / np.sqrt((xf ** 2).mean(1, keepdims=True) + 1e-5)
    np.testing.assert_allclose(nh.astype(np.float32), expect, atol=2e-3)
    # the sum is of the mean-subtracted values -> ~0, NOT sum of rmsnorm(x)
    assert np.abs(sm.astype(np.float32)).max() < 0.5


def test_silu_and_mul():
    x = np.array([[0.0, 1.0, -2.0, 8.0, 1.0, 2.0, 3.0, 0.5]], np.float16)
    out = act.silu_and_mul(x).astype(np.float32)
    g = x[0, :4].astype(np.float32)
    ref = g / (1 + np.exp(-g)) * x[0, 4:].astype(np.float32)
    np.testing.assert_allclose(out[0], ref, rtol=2e-3, atol=1e-3)


def test_kv4_page_sizes_match_cache_engine():
    # cache_engine.py:73-88 @ Llama-3-8B: 8*64*128/2 + 64*8*4 = 34816 bytes per K (and V) page
    c = kv4.PagedKV4(2, 8, 128)
    assert c.k_page_bytes == 34816 and c.v_page_bytes == 34816
    c2 = kv4.PagedKV4(2, 4, 128, k_stats_subchunks=4)
    assert c2.k_page_bytes == 4 * 64 * 64 + 64 * 4 * 4 + 2 * 4 * 4 * 128 * 2


def test_kv4_quant_roundtrip_and_wrap():
    rng = np.random.default_rng(2)
    x = rng.standard_normal((5, 128)).astype(np.float16)
    q, s, z = kv4.kv4_quant(x)
    assert q.max() <= 15
    d = kv4.kv4_dequant_f16(q, s, z).astype(np.float32)
    assert np.abs(d - x.astype(np.float32)).max() <= 0.6 * s.astype(np.float32).max()
    np.testing.assert_array_equal(kv4.unpack_nibbles(kv4.pack_nibbles(q)), q)
    de = kv4.kv4_dequant_exact(q, s, z)
    assert np.abs(de - d).max() < 5e-3


def test_decode_attention_mimic_close_to_exact_and_appends():
    rng = np.random.default_rng(3)
    B, Hq, Hkv, Dh = 2, 4, 2, 128
    lens = [37, 70]
    cache = kv4.PagedKV4(4, Hkv, Dh)
    bt = np.array([[0, 1], [2, 3]])
    kv4.fill_random(cache, bt, [l - 1 for l in lens], rng)
    before = cache.k_pool.copy()
    q = rng.standard_normal((B, Hq, Dh)).astype(np.float16)
    k = rng.standard_normal((B, Hkv, Dh)).astype(np.float16)
    v = rng.standard_normal((B, Hkv, Dh)).astype(np.float16)
    c2 = kv4.PagedKV4(4, Hkv, Dh)
    c2.k_pool[:], c2.v_pool[:] = cache.k_pool, cache.v_pool
    om = kv4.decode_attention(q, k, v, cache, bt, lens, 128, 5e5, mimic=True).astype(np.float32)
    oe = kv4.decode_attention(q, k, v, c2, bt, lens, 128, 5e5, mimic=False).astype(np.float32)
    assert np.abs(om - oe).max() <= 3e-3 * np.abs(oe).max()
    assert (cache.k_pool != before).sum() > 0  # the new token was appended
    np.testing.assert_array_equal(cache.k_pool, c2.k_pool)


def test_rope_is_a_rotation():
    rng = np.random.default_rng(4)
    x = rng.standard_normal((3, 128)).astype(np.float16)
    y = kv4.rope_neox(x, np.array([0, 5, 1000]), 128, 5e5).astype(np.float32)
    np.testing.assert_allclose(y[0], x[0].astype(np.float32), atol=1e-3)  # position 0 = identity
    np.testing.assert_allclose(np.linalg.norm(y, axis=1), np.linalg.norm(x.astype(np.float32), axis=1), rtol=2e-3)
