"""-m gpu: INTEGRATION.md section B -- the reference's pybind11 binding with the entry-function bodies replaced by calls into
the C ABI (integration/pybind_stub.cpp, compiled by __graft_entry__.build()) gives the same results as the ctypes mirror."""
import importlib.machinery
import importlib.util
import os

import numpy as np
import pytest
import torch

from tests.gpu_util import ROOT, make_gemm_inputs, t

pytestmark = pytest.mark.gpu


def _stub():
    path = os.path.join(ROOT, "integration", "_build", "ob_pybind_stub.so")
    if not os.path.exists(path):
        pytest.skip("integration/_build/ob_pybind_stub.so not built (python integration/build_stub.py)")
    loader = importlib.machinery.ExtensionFileLoader("ob_pybind_stub", path)
    spec = importlib.util.spec_from_loader("ob_pybind_stub", loader)
    m = importlib.util.module_from_spec(spec)
    loader.exec_module(m)
    return m


def test_pybind_binding_over_the_c_abi_matches_the_ctypes_mirror_and_the_oracle():
    from omniserve_b200.backend import fused_kernels, qgemm_w4a8_per_chn
    from oracle import w4a8 as ow
    stub = _stub()
    M, N, K = 64, 1024, 2048
    d = make_gemm_inputs(M, N, K, seed=5)
    args = [t(d[k]) for k in ("a", "qw", "s1", "sa", "szs", "ssum")]
    o1 = torch.empty((M, N), dtype=torch.float16, device="cuda")
    o2 = torch.empty_like(o1)
    stub.gemm_forward_cuda(*args, o1)
    qgemm_w4a8_per_chn.gemm_forward_cuda(*args, o2)
    torch.cuda.synchronize()
    assert torch.equal(o1, o2)
    _, ref = ow.gemm_per_chn(d["a"], d["qw"], d["s1"], d["sa"], d["szs"], d["ssum"])
    assert (o1.cpu().numpy() == ref).mean() > 0.999
    x = (torch.randn((7, 4096), device="cuda") * 2).half()
    q1 = torch.empty((7, 4096), dtype=torch.int8, device="cuda"); q2 = torch.empty_like(q1)
    s1 = torch.empty(7, dtype=torch.float16, device="cuda"); s2 = torch.empty_like(s1)
    m1 = torch.empty(7, dtype=torch.float16, device="cuda"); m2 = torch.empty_like(m1)
    stub.invoke_quant_fuse_sum(q1, x, m1, s1)
    fused_kernels.invoke_quant_fuse_sum(q2, x, m2, s2)
    torch.cuda.synchronize()
    assert torch.equal(q1, q2) and torch.equal(s1, s2) and torch.equal(m1, m2)
    with pytest.raises(RuntimeError):     # C-ABI error codes surface as the reference's TORCH_CHECK failures do
        stub.gemm_forward_cuda(args[0], args[1], args[2], args[3], args[4], args[5], torch.empty((M, 48), dtype=torch.float16, device="cuda"))
