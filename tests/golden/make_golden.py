#!/usr/bin/env python
"""Generate golden vectors by running the REFERENCE's own Python on CPU.

Run in the authoring container only (needs /root/reference):  python tests/golden/make_golden.py

What is imported from the reference, unmodified, by file path:
  * scripts/ckpt_converter/quant_utils.py   -> pseudo_quantize_tensor            (:96-138)
  * omniserve/modeling/layers/quantized_linear/w4a8_linear.py
                                            -> W4A8OF16LinearDynamicInputScale.from_linear (:141-337)
The latter imports ``omniserve_backend.*`` and calls ``torch.cuda.current_device()`` /
``Tensor.cuda()`` at import/use; those three are stubbed here (no CUDA in this container) -- the
packing arithmetic itself is the reference's.  Outputs (committed): tests/golden/*.npz + golden.json.
"""
import hashlib
import importlib.util
import json
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def main():
    # --- stubs so the reference file imports on a CPU-only box
    be = types.ModuleType("omniserve_backend")
    for sub in ("qgemm_w4a8_per_chn", "qgemm_w4a8_per_group"):
        sm = types.ModuleType(f"omniserve_backend.{sub}")
        setattr(be, sub, sm)
        sys.modules[f"omniserve_backend.{sub}"] = sm
    sys.modules["omniserve_backend"] = be
    torch.cuda.current_device = lambda: "cpu"
    torch.Tensor.cuda = lambda self, *a, **k: self

    qu = _load("ref_quant_utils", f"{REF}/scripts/ckpt_converter/quant_utils.py")
    wl = _load("ref_w4a8_linear", f"{REF}/omniserve/modeling/layers/quantized_linear/w4a8_linear.py")
    Lin = wl.W4A8OF16LinearDynamicInputScale
    meta = {}

    # ---------------- per-channel (BASELINE config 1 semantics), small fixture
    g = torch.Generator().manual_seed(1234)
    N, K = 64, 128
    w = torch.randn(N, K, generator=g) * 0.02
    w_fake, scales, zeros = qu.pseudo_quantize_tensor(w, n_bit=4, zero_point=True, q_group_size=-1, get_scale_zp=True)
    lin = torch.nn.Linear(K, N, bias=False)
    lin.weight.data = w_fake.clone()
    ql = Lin.from_linear(lin, 4, -1, init_only=False, s1_scale=scales.reshape(-1), zeros=zeros.reshape(-1).to(torch.int8))
    np.savez_compressed(
        f"{HERE}/w4a8_per_chn_64x128.npz",
        w=w.numpy(), w_fake=w_fake.numpy(), scales=scales.numpy(), zeros=zeros.numpy(),
        qweight=ql.qweight.numpy(), s1_scales=ql.s1_scales.numpy(), s1_szeros=ql.s1_szeros.numpy(),
    )

    # ---------------- per-channel 4096x4096: checksum only (config 1 full size)
    rng = np.random.default_rng(20240522)
    w_big = torch.from_numpy((rng.standard_normal((4096, 4096)) * 0.02).astype(np.float32))
    wf, sc, zr = qu.pseudo_quantize_tensor(w_big, n_bit=4, zero_point=True, q_group_size=-1, get_scale_zp=True)
    lin = torch.nn.Linear(4096, 4096, bias=False)
    lin.weight.data = wf.clone()
    qb = Lin.from_linear(lin, 4, -1, init_only=False, s1_scale=sc.reshape(-1), zeros=zr.reshape(-1).to(torch.int8))
    meta["per_chn_4096"] = {
        "seed": 20240522,
        "qweight_sha256": hashlib.sha256(qb.qweight.numpy().tobytes()).hexdigest(),
        "w_fake_sha256": hashlib.sha256(wf.numpy().tobytes()).hexdigest(),
        "s1_scales_sha256": hashlib.sha256(qb.s1_scales.numpy().tobytes()).hexdigest(),
        "s1_szeros_sha256": hashlib.sha256(qb.s1_szeros.numpy().tobytes()).hexdigest(),
    }

    # ---------------- per-group (g128) two-level, small fixture
    N, K, G = 64, 256, 128
    g = torch.Generator().manual_seed(99)
    ng = K // G
    s1 = (torch.rand(N, generator=g) * 0.01 + 0.005).half().float()
    s2 = torch.randint(1, 9, (N, ng), generator=g).float()
    z = torch.randint(0, 16, (N, ng), generator=g).float()
    q4 = torch.randint(0, 16, (N, K), generator=g).float()
    w8 = (q4.reshape(N, ng, G) - z[..., None]) * s2[..., None]          # level-1 integer in [-120, 120]
    w8 = w8.clamp(-127, 127).reshape(N, K)
    wfp = w8 * s1[:, None]
    lin = torch.nn.Linear(K, N, bias=False)
    lin.weight.data = wfp.clone()
    ql = Lin.from_linear(lin, 4, G, init_only=False, s1_scale=s1, s2_scale=s2, zeros=z)
    np.savez_compressed(
        f"{HERE}/w4a8_per_group_64x256.npz",
        w=wfp.numpy(), s1=s1.numpy(), s2=s2.numpy(), zeros=z.numpy(), w8=w8.numpy(),
        qweight=ql.qweight.numpy(), s1_scales=ql.s1_scales.numpy(),
        s2_scales=ql.s2_scales.numpy(), s2_zeros=ql.s2_zeros.numpy(),
    )

    with open(f"{HERE}/golden.json", "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
