"""CPU: the per-tensor INT8 KV oracle (oracle/kv8.py) -- quantiser known answers, page layout, and that the reference-rounding
and exact-arithmetic variants of the decode agree to the tolerance the GPU tests use."""
import numpy as np

from oracle import kv8


def test_kv8_quant_known_answers_round_half_even_and_saturate():
    x = np.asarray([0.0, 0.5, 1.5, 2.5, -0.5, -1.5, 200.0, -200.0, 126.5, 127.5], np.float16)
    np.testing.assert_array_equal(kv8.kv8_quant(x, 1.0), [0, 0, 2, 2, 0, -2, 127, -128, 126, 127])
    np.testing.assert_array_equal(kv8.kv8_quant(np.asarray([1.0, -1.0], np.float16), 28.0), [28, -28])
    np.testing.assert_array_equal(kv8.kv8_dequant_f16(np.asarray([127, -128], np.int8), 0.03125), np.asarray([3.96875, -4.0], np.float16))


def test_kv8_page_layout_matches_cache_engine_sizes():
    c = kv8.PagedKV8(3, 8, 128, k_stats_subchunks=4)
    assert c.data_bytes == 8 * 64 * 128 and c.sz_bytes == 64 * 8 * 4         # cache_engine.py:73-88 with 1-byte elements
    assert c.k_page_bytes == c.data_bytes + c.sz_bytes + 2 * 4 * 8 * 128 * 2 and c.v_page_bytes == c.data_bytes + c.sz_bytes
    c.data("k", 1)[2, 5, :] = 7
    assert c.k_pool[1, (2 * 64 + 5) * 128: (2 * 64 + 5) * 128 + 128].tolist() == [7] * 128


def test_kv8_prefill_then_decode_roundtrip_and_mimic_close_to_exact():
    rng = np.random.default_rng(0)
    Hq, Hkv, Dh, lens = 4, 2, 128, [70, 9]
    T = sum(lens)
    cache = kv8.PagedKV8(4, Hkv, Dh)
    bt = np.asarray([[0, 1], [2, 3]], np.int64)
    sqo = np.asarray([4.5 / 127, 4.0 / 127], np.float32)
    soq = (1 / sqo).astype(np.float32)
    qkv = rng.standard_normal((T, (Hq + 2 * Hkv) * Dh)).astype(np.float16)
    v_raw = qkv[:, (Hq + Hkv) * Dh:].copy().reshape(T, Hkv, Dh)
    kv8.prefill_write(cache, bt, qkv, lens, Hq, 128, 10000.0, soq)
    # round trip: dequantised V within half a code step of the input (|x| < 4 here except rare tails -> saturation allowed)
    vd = kv8.kv8_dequant_exact(cache.data("v", 0)[:, :64, :], sqo[1]).transpose(1, 0, 2)
    ok = np.abs(v_raw[:64]) < 3.9
    assert np.abs(vd - v_raw[:64].astype(np.float64))[ok].max() <= 0.5 * sqo[1] + 1e-6
    q = rng.standard_normal((2, Hq, Dh)).astype(np.float16)
    k = rng.standard_normal((2, Hkv, Dh)).astype(np.float16)
    v = rng.standard_normal((2, Hkv, Dh)).astype(np.float16)
    lens1 = [l + 1 for l in lens]
    c2 = kv8.PagedKV8(4, Hkv, Dh); c2.k_pool[:], c2.v_pool[:] = cache.k_pool, cache.v_pool
    oe = kv8.decode_attention(q, k, v, cache, bt, lens1, 128, 10000.0, sqo, soq, mimic=False).astype(np.float32)
    om = kv8.decode_attention(q, k, v, c2, bt, lens1, 128, 10000.0, sqo, soq, mimic=True).astype(np.float32)
    assert np.abs(oe - om).max() <= 2e-3 * np.abs(oe).max()
    np.testing.assert_array_equal(cache.k_pool, c2.k_pool)                    # the append does not depend on the variant
    assert (cache.data("k", 1)[:, 6, :] != 0).any()                           # token 70 of sequence 0 -> page 1 slot 6
