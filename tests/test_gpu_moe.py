"""-m gpu: grouped W4A8 GEMM for mixture-of-experts layers (BASELINE config 5) against the
CPU oracle -- the reference ships only the interface (w4a8_moe_linear.py:83-94), so parity is pinned by the oracle's
per-expert application of the per-channel GEMM, which itself is pinned against the reference's dense kernel."""
import numpy as np
import pytest
import torch

from tests.gpu_util import t

pytestmark = pytest.mark.gpu


def _case(sizes, N, K, seed):
    from oracle import w4a8 as ow
    rng = np.random.default_rng(seed)
    E, T = len(sizes), int(sum(sizes))
    q = rng.integers(0, 16, (E, N, K), dtype=np.uint8)
    qw = np.stack([ow.pack_w4(q[e]) for e in range(E)])
    x = rng.integers(-127, 128, (T, K), dtype=np.int8)
    s1 = rng.uniform(0.005, 0.02, (E, N)).astype(np.float16)
    z = rng.integers(0, 16, (E, N)).astype(np.float32)
    szs = (z * s1.astype(np.float32)).astype(np.float16)
    sa = rng.uniform(0.01, 0.05, T).astype(np.float16)
    ssum = (x.astype(np.float32) * sa.astype(np.float32)[:, None]).sum(1).astype(np.float16)
    return x, qw, s1, sa, szs, ssum


@pytest.mark.parametrize("sizes,N,K", [((5, 0, 17, 1), 256, 512), ((64, 64), 128, 128), ((100, 3, 0, 70, 9, 0, 1, 30), 512, 1024),
                                        ((8,) * 8, 4096, 4096)])
def test_moe_grouped_gemm_vs_oracle(sizes, N, K):
    from omniserve_b200.backend import qgemm_w4a8_moe as op
    from oracle import w4a8 as ow
    x, qw, s1, sa, szs, ssum = _case(sizes, N, K, seed=sum(sizes) + N)
    out = op.moe_gemm_forward_cuda_api(t(x), t(qw), t(s1), t(sa), t(szs), t(ssum), list(sizes))
    torch.cuda.synchronize()
    ref = ow.moe_gemm_per_chn(x, qw, s1, sa, szs, ssum, sizes)
    got = out.cpu().numpy()
    g32, r32 = got.astype(np.float32), ref.astype(np.float32)
    assert np.abs(g32 - r32).max() <= 1e-3 * np.abs(r32).max()
    assert (got == ref).mean() > 0.999


def test_moe_more_chunks_than_one_launch_holds():
    """> 64 chunks of <= 64 rows: the op issues several launches; problem_sizes given as a tensor."""
    from omniserve_b200.backend import qgemm_w4a8_moe as op
    from oracle import w4a8 as ow
    sizes = [70] * 40 + [1] * 10          # 40 experts x 2 chunks + 10 = 90 chunks
    x, qw, s1, sa, szs, ssum = _case(sizes, 128, 256, seed=1)
    out = op.moe_gemm_forward_cuda_api(t(x), t(qw), t(s1), t(sa), t(szs), t(ssum), torch.tensor(sizes, dtype=torch.int32))
    torch.cuda.synchronize()
    ref = ow.moe_gemm_per_chn(x, qw, s1, sa, szs, ssum, sizes)
    assert (out.cpu().numpy() == ref).mean() > 0.999


def test_moe_equals_dense_op_per_expert_at_mixtral_decode_shape():
    """BASELINE config 5 shape (Mixtral-8x7B gate_up: 8 experts x [28672, 4096], bs=32 tokens x top-2 = 64 rows): the grouped
    launch must equal the dense per-channel op applied expert by expert (size-independent property, bit for bit)."""
    from omniserve_b200.backend import qgemm_w4a8_moe as op
    from omniserve_b200.backend import qgemm_w4a8_per_chn as dense
    E, N, K = 8, 28672, 4096
    g = torch.Generator(device="cuda").manual_seed(0)
    sizes = [11, 3, 0, 14, 9, 8, 12, 7]
    T = sum(sizes)
    qw = torch.randint(-128, 128, (E, N, K // 2), generator=g, device="cuda", dtype=torch.int8)
    x = torch.randint(-127, 128, (T, K), generator=g, device="cuda", dtype=torch.int8)
    s1 = (torch.rand((E, N), generator=g, device="cuda") * 0.01 + 0.005).half()
    szs = (s1.float() * 8).half()
    sa = (torch.rand(T, generator=g, device="cuda") * 0.02 + 0.01).half()
    ss = torch.randn(T, generator=g, device="cuda").half()
    out = op.moe_gemm_forward_cuda_api(x, qw, s1, sa, szs, ss, sizes)
    ref = torch.empty_like(out)
    r = 0
    for e, m in enumerate(sizes):
        if m:
            dense.gemm_forward_cuda(x[r:r + m], qw[e], s1[e], sa[r:r + m], szs[e], ss[r:r + m], ref[r:r + m])
        r += m
    torch.cuda.synchronize()
    assert torch.equal(out, ref)
