"""-m gpu: fused small ops vs the oracle.  INT8 codes: bit-exact expected; because the product uses the same
fast-math intrinsics as the reference build while the oracle divides exactly, up to 1e-3 of the codes may
differ by 1 LSB (tolerance stated here)."""
import numpy as np
import pytest
import torch

from tests.gpu_util import t

pytestmark = pytest.mark.gpu


def _cmp_q(got, ref):
    d = np.abs(got.astype(np.int32) - ref.astype(np.int32))
    assert d.max() <= 1
    assert (d > 0).mean() <= 1e-3


@pytest.mark.parametrize("T,H", [(1, 64), (3, 256), (64, 4096), (5, 14336), (4, 8192), (2, 28672)])
def test_quant_fuse_sum(T, H):
    from omniserve_b200.backend import fused_kernels
    from oracle import act
    rng = np.random.default_rng(T * 7 + H)
    x = (rng.standard_normal((T, H)) * 2).astype(np.float16)
    q, s, sm = act.quant_fuse_sum(x)
    out = torch.empty((T, H), dtype=torch.int8, device="cuda")
    sc = torch.empty(T, dtype=torch.float16, device="cuda")
    su = torch.empty(T, dtype=torch.float16, device="cuda")
    fused_kernels.invoke_quant_fuse_sum(out, t(x), su, sc)
    out2 = torch.empty_like(out)
    sc2 = torch.empty_like(sc)
    fused_kernels.invoke_quant(out2, t(x), sc2)
    torch.cuda.synchronize()
    _cmp_q(out.cpu().numpy(), q)
    assert torch.equal(out, out2) and torch.equal(sc, sc2)
    np.testing.assert_array_equal(sc.cpu().numpy(), s)
    np.testing.assert_allclose(su.cpu().numpy().astype(np.float32), sm.astype(np.float32), rtol=2e-3, atol=2e-2)


@pytest.mark.parametrize("T,H", [(3, 256), (64, 4096), (4, 8192), (2, 1024)])
def test_rms_norm_general_fuse_sum(T, H):
    from omniserve_b200.backend import layernorm_ops
    from oracle import act
    rng = np.random.default_rng(T + H)
    x = (rng.standard_normal((T, H)) * 1.5 + 0.2).astype(np.float16)
    g = (rng.standard_normal(H) * 0.2 + 1.0).astype(np.float16)
    q, s, sm, _ = act.rms_norm_general_fuse_sum(x, g, 1e-5)
    out = torch.empty((T, H), dtype=torch.int8, device="cuda")
    sc = torch.empty(T, dtype=torch.float16, device="cuda")
    su = torch.empty(T, dtype=torch.float16, device="cuda")
    layernorm_ops.rms_norm_general_fuse_sum(out, t(x), t(g), su, sc, 1e-5, True)
    out2 = torch.empty_like(out)
    sc2 = torch.empty_like(sc)
    layernorm_ops.rms_norm_general(out2, t(x), t(g), sc2, 1e-5, True)
    torch.cuda.synchronize()
    _cmp_q(out.cpu().numpy(), q)
    assert torch.equal(out, out2)
    assert np.abs(sc.cpu().numpy().astype(np.float32) - s.astype(np.float32)).max() <= 1e-3 * s.astype(np.float32).max()
    # the fp16 partial-sum quirk makes this sum tiny and noisy; compare absolutely
    assert np.abs(su.cpu().numpy().astype(np.float32) - sm.astype(np.float32)).max() <= 0.05


def test_rms_norm_and_silu():
    from omniserve_b200.backend import activation_ops, layernorm_ops
    from oracle import act
    rng = np.random.default_rng(5)
    x = (rng.standard_normal((7, 4096))).astype(np.float16)
    g = (rng.standard_normal(4096) * 0.1 + 1).astype(np.float16)
    o = torch.empty((7, 4096), dtype=torch.float16, device="cuda")
    layernorm_ops.rms_norm(o, t(x), t(g), 1e-5)
    ref = act.rms_norm(x, g, 1e-5).astype(np.float32)
    assert np.abs(o.cpu().numpy().astype(np.float32) - ref).max() <= 2e-3 * np.abs(ref).max()
    y = (rng.standard_normal((5, 2 * 14336)) * 2).astype(np.float16)
    so = torch.empty((5, 14336), dtype=torch.float16, device="cuda")
    activation_ops.silu_and_mul(so, t(y))
    rs = act.silu_and_mul(y).astype(np.float32)
    assert np.abs(so.cpu().numpy().astype(np.float32) - rs).max() <= 2e-3 * np.abs(rs).max()
    # fused silu*mul -> quant == the two-kernel chain of the reference (activation.py:54-64), bit-for-bit
    from omniserve_b200.backend import fused_kernels
    q1 = torch.empty((5, 14336), dtype=torch.int8, device="cuda")
    q2 = torch.empty_like(q1)
    s1 = torch.empty(5, dtype=torch.float16, device="cuda"); s2 = torch.empty_like(s1)
    m1 = torch.empty_like(s1); m2 = torch.empty_like(s1)
    activation_ops.silu_and_mul_quant(q1, t(y), m1, s1)
    fused_kernels.invoke_quant_fuse_sum(q2, so, m2, s2)
    torch.cuda.synchronize()
    assert torch.equal(q1, q2) and torch.equal(s1, s2)
    assert (m1.float() - m2.float()).abs().max() <= 2e-3 * m2.float().abs().max() + 1e-2


def test_empty_input_is_a_noop():
    from omniserve_b200.backend import fused_kernels
    fused_kernels.invoke_quant(torch.empty((0, 64), dtype=torch.int8, device="cuda"),
                               torch.empty((0, 64), dtype=torch.float16, device="cuda"),
                               torch.empty((0,), dtype=torch.float16, device="cuda"))
    torch.cuda.synchronize()
