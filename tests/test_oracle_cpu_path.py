"""Pin the vectorised torch-CPU path (bench.py's cpu_baseline) against the loop oracles."""
import numpy as np
import torch

from oracle import act, cpu_path, kv4, w4a8


def test_linear_matches_loop_oracle():
    rng = np.random.default_rng(0)
    M, N, K = 5, 64, 256
    qw = rng.integers(0, 16, (N, K), dtype=np.uint8)
    x = rng.standard_normal((M, K)).astype(np.float16)
    qa, sa, ss = act.quant_fuse_sum(x)
    s1 = rng.uniform(0.005, 0.02, N).astype(np.float16)
    szs = (8 * s1.astype(np.float32)).astype(np.float16)
    _, ref = w4a8.gemm_per_chn(qa, w4a8.pack_w4(qw), s1, sa, szs, ss)
    q2, sa2, ss2 = cpu_path.quant_per_token(torch.from_numpy(x))
    np.testing.assert_array_equal(q2.numpy().astype(np.int8), qa)
    got = cpu_path.w4a8_linear(q2, sa2, ss2, torch.from_numpy(qw.astype(np.float32)), torch.from_numpy(s1),
                               torch.from_numpy(szs))
    np.testing.assert_allclose(got.float().numpy(), ref.astype(np.float32), rtol=2e-3, atol=1e-3)


def test_kv_fake_quant_and_rope_match_loop_oracle():
    rng = np.random.default_rng(1)
    x = rng.standard_normal((3, 2, 128)).astype(np.float16)
    q, s, z = kv4.kv4_quant(x)
    ref = kv4.kv4_dequant_f16(q, s, z)
    got = cpu_path.kv4_fake_quant(torch.from_numpy(x)).numpy()
    assert (got != ref).mean() < 5e-3
    pos = np.array([3, 77, 1200])
    r = kv4.rope_neox(x, pos[:, None], 128, 5e5)
    g = cpu_path.rope_neox(torch.from_numpy(x), torch.from_numpy(pos), 5e5).numpy()
    assert np.abs(r.astype(np.float32) - g.astype(np.float32)).max() < 4e-3


def test_decode_layer_runs_and_is_finite():
    cfg = dict(hidden=256, inter=512, hq=4, hkv=2, dh=128, eps=1e-5, base=5e5)
    g = torch.Generator().manual_seed(0)
    p = cpu_path.random_layer(cfg, g)
    B, ctx = 3, 40
    kc = torch.randn(B, 2, 64, 128, generator=g).half()
    vc = torch.randn(B, 2, 64, 128, generator=g).half()
    x = torch.randn(B, 256, generator=g).half()
    lens = torch.tensor([ctx, 5, 0])
    y = cpu_path.decode_layer(x, p, kc, vc, lens, cfg)
    assert y.shape == (B, 256) and torch.isfinite(y.float()).all()
