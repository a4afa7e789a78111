"""CPU: the oracle's LServe pieces -- page statistics, selector score, top-k page
choice -- checked against independent restatements of the reference semantics
(sparse_utils/ContextPool/context_pool_kernel.cu:16-69, sparse_utils/KVPageSelector/KVPageSelectorTemplate.hpp:478-503,
1130-1283, omniserve/modeling/layers/decoding_attention.py:88-143)."""
import numpy as np
import torch

from oracle import kv4


def _case(lens, H=2, seed=0):
    rng = np.random.default_rng(seed)
    n_pages = sum((l + 63) // 64 for l in lens) + 1
    cache = kv4.PagedKV4(n_pages, H, 128, k_stats_subchunks=4)
    perm = rng.permutation(n_pages)
    bt = np.zeros((len(lens), max((l + 63) // 64 for l in lens)), np.int64)
    c = 0
    for b, l in enumerate(lens):
        for j in range((l + 63) // 64):
            bt[b, j] = perm[c]
            c += 1
    return cache, bt, rng


def test_pool_partial_subchunk_and_head_mapping():
    lens = [70, 33]
    cache, bt, rng = _case(lens)
    keys = rng.standard_normal((sum(lens), 5, 128)).astype(np.float16)   # 5 input heads, pool heads 3 and 1
    idx = np.array([3, 1])
    kv4.paged_min_max_pool(cache, bt, keys, lens, 16, pooling_heads_idx=idx)
    # sequence 0: tokens 64..69 are sub-chunk 0 of its second page (6 valid tokens)
    kmax, kmin = cache.kstats(int(bt[0, 1]))
    np.testing.assert_array_equal(kmax[0, :128], keys[64:70, 3].max(0))
    np.testing.assert_array_equal(kmin[0, 128:], keys[64:70, 1].min(0))
    assert not kmax[1].any()   # untouched sub-chunks stay as they were
    # sequence 1 starts at packed token 70; its token 32 alone forms sub-chunk 2 of page 0
    kmax, kmin = cache.kstats(int(bt[1, 0]))
    np.testing.assert_array_equal(kmax[2, :128], keys[70 + 32, 3])
    np.testing.assert_array_equal(kmin[2, :128], keys[70 + 32, 3])
    np.testing.assert_array_equal(kmax[1, 128:], keys[70 + 16:70 + 32, 1].max(0))


def test_score_is_fp16_rounded_upper_bound_of_qk():
    rng = np.random.default_rng(1)
    q = rng.standard_normal(128).astype(np.float16)
    ks = rng.standard_normal((40, 16, 128)).astype(np.float16)
    kmax, kmin = ks.max(1), ks.min(1)
    got = kv4._score_fp16(q, kmax, kmin).astype(np.float64)
    exact = np.maximum(q.astype(np.float64) * kmax, q.astype(np.float64) * kmin).sum(-1)
    assert np.abs(got - exact).max() <= 3e-3 * np.abs(exact).max() + 0.05   # fp16 accumulation of 8 terms per lane
    true_max = (ks.astype(np.float64) @ q.astype(np.float64)).max(1)
    assert (exact >= true_max - 1e-9).all()                                 # the statistic bounds every key of the sub-chunk


def test_selector_layout_streaming_rows_and_pitch_quirk():
    lens = [200, 137]     # cached tokens tl = 199, 136
    cache, bt, rng = _case(lens, H=1)
    keys = rng.standard_normal((sum(l - 1 for l in lens), 1, 128)).astype(np.float16)
    kv4.paged_min_max_pool(cache, bt, keys, [l - 1 for l in lens], 16)
    q = rng.standard_normal((2, 4, 128)).astype(np.float16)
    flags = np.array([1, 0])   # kv head 0 retrieval (rank 0), kv head 1 streaming
    rank = np.array([0, 0])
    timestep = 199
    out = kv4.page_selector(q, cache, bt, lens, timestep, 128, 500000.0, retrieval_flags=flags, head_rank=rank, n_kv_heads=2)
    assert out.shape == (2, 4, 16)          # ceil(199/16) = 13 -> padded to 16
    flat = out.reshape(-1)
    # sample 0 (lengths == timestep + 1): plain layout, 13 valid scores per retrieval q-head, padding zero
    assert out[0, 0, :13].astype(np.float32).any() and not out[0, 0, 13:].any()
    assert not out[0, 2].any()              # streaming head (row 3 is overwritten by sample 1, see below)
    # sample 1: the kernel derives the row pitch from the sample's own length, padded(tl=136) = 12, so its q-head h
    # starts at flat element (1*4 + h) * 12 -- inside sample 0's block -- not at (1*4 + h) * 16
    # (KVPageSelectorTemplate.hpp:1130-1133 vs fused_kv_page_selector.cpp:274-277)
    q_r = kv4.rope_neox(q[1], 136, 128, 500000.0)
    rows_max = np.stack([cache.kstats(int(bt[1, s // 4]))[0][s % 4, :128] for s in range(9)])
    rows_min = np.stack([cache.kstats(int(bt[1, s // 4]))[1][s % 4, :128] for s in range(9)])
    for h in (0, 1):
        o = (1 * 4 + h) * 12
        np.testing.assert_array_equal(flat[o:o + 9], kv4._score_fp16(q_r[h], rows_max, rows_min))
    assert not flat[(1 * 4 + 2) * 12:].any()   # streaming heads of sample 1 and everything after stay zero


def test_topk_choice_matches_the_reference_python():
    rng = np.random.default_rng(3)
    B, Hq, pages = 2, 4, 40
    stats = rng.standard_normal((B, Hq, pages * 4)).astype(np.float16)
    timestep, budget = pages * 64 - 10, 8 * 64
    got = kv4.select_topk_pages(stats, timestep, budget)
    # decoding_attention.py:132-141 restated with torch
    s = torch.from_numpy(stats).view(B, Hq, -1, 4)
    s = torch.max(s, dim=-1).values
    total = s.size(-1)
    _, idx = s[:, :, :-1].float().topk(k=(min(max(3, budget // 64), total) - 1), dim=-1)
    idx = torch.cat([idx, torch.ones_like(idx[..., :1]) * (total - 1)], dim=-1).to(torch.int32).numpy()
    assert got.shape == idx.shape == (B, Hq, 8)
    assert (got[..., -1] == pages - 1).all()
    for b in range(B):
        for h in range(Hq):
            assert set(got[b, h, :-1]) == set(idx[b, h, :-1])
    # short contexts: every page, in order
    short = kv4.select_topk_pages(stats, 130, 4096)
    np.testing.assert_array_equal(short[0, 0], [0, 1, 2])


def test_decode_append_folds_key_into_statistics():
    lens = [70]
    cache, bt, rng = _case(lens, H=2)
    kv4.fill_random(cache, bt, [69], rng)
    page = int(bt[0, 1])
    kmax, kmin = cache.kstats(page)
    kmax[:] = 0.25
    kmin[:] = -0.25
    q = rng.standard_normal((1, 4, 128)).astype(np.float16)
    k = rng.standard_normal((1, 2, 128)).astype(np.float16)
    v = rng.standard_normal((1, 2, 128)).astype(np.float16)
    kv4.decode_attention(q, k, v, cache, bt, lens, 128, 500000.0, mimic=False, update_stats_sub_chunk=16)
    k_r = kv4.rope_neox(k[0], 69, 128, 500000.0).reshape(-1)
    np.testing.assert_array_equal(kmax[0], np.maximum(np.float16(0.25), k_r))
    np.testing.assert_array_equal(kmin[0], np.minimum(np.float16(-0.25), k_r))
    assert (kmax[1:] == np.float16(0.25)).all()
