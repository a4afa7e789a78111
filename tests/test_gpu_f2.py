"""-m gpu: SURVEY.md section 8 row f2 -- (a) the fused prefill pass (RoPE + KV4 page write + kmax / kmin statistics in one
kernel) against the two-op chain it replaces (apply_bias_rope_update_kv_cache then paged_min_max_pool, ctx_update_kv.py:104-178),
bit for bit; (b) the device-side page choice against the reference's torch chain (decoding_attention.py:132-141)."""
import numpy as np
import pytest
import torch

from tests.gpu_util import t

pytestmark = pytest.mark.gpu


def _pools(B, lens, Hr, Hs, sink_blk, local_blk, seed):
    from oracle import kv4
    rng = np.random.default_rng(seed)
    n_r = sum((l + 63) // 64 for l in lens)
    rc = kv4.PagedKV4(max(n_r, 1), max(Hr, 1), 128, k_stats_subchunks=4)
    rc.k_pool[:] = rng.integers(0, 256, rc.k_pool.shape, dtype=np.uint8)     # stale bytes: everything written must be overwritten
    rbt = np.zeros((B, max((l + 63) // 64 for l in lens)), np.int64)
    perm = rng.permutation(n_r)
    c = 0
    for b, l in enumerate(lens):
        for j in range((l + 63) // 64):
            rbt[b, j] = perm[c]; c += 1
    sc = kv4.PagedKV4(B * (sink_blk + local_blk), max(Hs, 1), 128)
    stab = rng.permutation(B * (sink_blk + local_blk)).reshape(B, -1)
    return rc, rbt, sc, stab


def _ptrs(cache, tables):
    kpool, vpool = t(cache.k_pool), t(cache.v_pool)
    B, P = tables.shape
    p = np.zeros((B, 2, P), np.int64)
    p[:, 0] = kpool.data_ptr() + tables * cache.k_page_bytes
    p[:, 1] = vpool.data_ptr() + tables * cache.v_page_bytes
    return kpool, vpool, t(p)


@pytest.mark.parametrize("lens,flags", [((200, 64, 333), (1, 1, 1, 1)), ((17, 1, 1000), (1, 0, 0, 1)), ((129,), (0, 1, 1, 1)),
                                        ((48, 31), (1, 1, 0, 0))])
def test_fused_prefill_write_and_pool_equals_two_op_chain(lens, flags):
    from omniserve_b200.backend import fused_attention_ctx_pool as pool
    from omniserve_b200.backend import fused_attention_fine_grained_dense as wr
    Hq, Hkv = 8, 4
    flags_np = np.asarray(flags, np.int32)
    rank_np = np.zeros(Hkv, np.int32)
    rank_np[flags_np == 1] = np.arange((flags_np == 1).sum())
    rank_np[flags_np == 0] = np.arange((flags_np == 0).sum())
    Hr, Hs = int(flags_np.sum()), int(Hkv - flags_np.sum())
    sink, local, sink_blk, local_blk = 64, 128, 1, 3
    B, T = len(lens), sum(lens)
    rng = np.random.default_rng(sum(lens))
    qkv = rng.standard_normal((T, (Hq + 2 * Hkv) * 128)).astype(np.float16)
    cu = t(np.concatenate([[0], np.cumsum(lens)]).astype(np.int32))
    sl = t(np.asarray(lens, np.int32))
    res = []
    for fused in (False, True):
        rc, rbt, sc, stab = _pools(B, lens, Hr, Hs, sink_blk, local_blk, seed=7)
        rk, rv, rptrs = _ptrs(rc, rbt)
        sk, sv, sptrs = _ptrs(sc, stab)
        tq = t(qkv)
        pad = wr.compute_padding_offsets(cu, max(lens), T)
        args = (tq, sl, sl if Hs else None, pad, rptrs if Hr else None, sptrs if Hs else None, t(flags_np), t(rank_np), Hq, Hkv,
                max(lens), 64, Hr * 64, Hs * 64, sink, local, sink_blk, local_blk, Hr, Hs, 128, 500000.0, 1.0, 8192, True, True, True)
        if fused:
            wr.apply_bias_rope_update_kv_cache_pool(*args, tokens_per_sub_chunk=16)
        else:
            wr.apply_bias_rope_update_kv_cache(*args)
            if Hr:
                keys = tq[:, Hq * 128:(Hq + Hkv) * 128].view(T, Hkv, 128).contiguous()     # post-RoPE keys, as llama:316-325
                pool.paged_min_max_pool(keys, rptrs, cu, t(np.nonzero(flags_np)[0].astype(np.int32)), max(lens), 16, 64, Hr * 64, True)
        torch.cuda.synchronize()
        res.append((tq.cpu(), rk.cpu(), rv.cpu(), sk.cpu(), sv.cpu()))
    names = ("qkv (RoPE in place)", "retrieval K pages + statistics", "retrieval V pages", "streaming K pages", "streaming V pages")
    for a, b, n in zip(res[0], res[1], names):
        assert torch.equal(a, b), n


@pytest.mark.parametrize("B,Hq,total,k_out", [(1, 8, 100, 17), (2, 4, 4097, 64), (1, 32, 65, 64), (3, 2, 5, 3), (1, 1, 16385, 64)])
def test_device_page_topk_matches_torch_chain(B, Hq, total, k_out):
    from omniserve_b200.backend import fused_attention_selector as sel
    g = torch.Generator(device="cuda").manual_seed(total + k_out)
    group = 4
    stats = (torch.randn((B, Hq, total * group), generator=g, device="cuda") * 20).half()
    stats[:, 0] = 0                                            # a streaming head's row: all zero (every page ties)
    idx = sel.page_topk(stats, group, k_out)
    torch.cuda.synchronize()
    assert idx.dtype == torch.int32 and idx.shape == (B, Hq, k_out)
    page = stats.view(B, Hq, total, group).max(-1).values.float()
    _, ref = page[:, :, :-1].topk(k_out - 1, dim=-1)          # the reference's choice (decoding_attention.py:137)
    got = idx.cpu().numpy()
    page_np, ref_np = page.cpu().numpy(), ref.cpu().numpy()
    for b in range(B):
        for h in range(Hq):
            row = got[b, h]
            assert row[-1] == total - 1                        # newest page last
            chosen = row[:-1]
            assert len(set(chosen.tolist())) == k_out - 1 and chosen.min() >= 0 and chosen.max() < total - 1
            sc = page_np[b, h]
            rest = np.setdiff1d(np.arange(total - 1), chosen)
            if len(rest) and len(chosen):
                assert sc[chosen].min() >= sc[rest].max()      # a valid top-(k-1) set
            # same multiset of scores as torch's choice (sets may differ only among exact ties)
            np.testing.assert_array_equal(np.sort(sc[chosen]), np.sort(sc[ref_np[b, h]]))
