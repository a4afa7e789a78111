#!/usr/bin/env python
"""One warm-up + one measured launch of a single hot kernel on realistic (HBM-cold) inputs, for Nsight Compute:

    ncu --set full --clock-control none -k regex:<kernel> --launch-skip 1 --launch-count 1 -o out \\
        python tools/ncu_targets.py <target>

targets: gemm_qkv gemm_o gemm_gate_up gemm_down (decode, M=64) | prefill_qkv prefill_gate_up (M=8192) |
         prefill_qkv_pair (CTA-pair kernel) | attention (bs=64, ctx=1280) | selector (256K context, bs=1)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from omniserve_b200 import _lib as L  # noqa: E402

dev = "cuda"
torch.cuda.set_device(0)
target = sys.argv[1]
SHAPES = {"qkv": (6144, 4096), "o": (4096, 4096), "gate_up": (28672, 4096), "down": (4096, 14336)}


def gemm(M, N, K):
    ws = [torch.randint(-128, 128, (N, K // 2), dtype=torch.int8, device=dev) for _ in range(2)]   # cold weights per launch
    x = torch.randint(-127, 128, (M, K), dtype=torch.int8, device=dev)
    s1 = torch.full((N,), 0.01, dtype=torch.float16, device=dev)
    sz = torch.full((N,), 0.08, dtype=torch.float16, device=dev)
    sa = torch.full((M,), 0.02, dtype=torch.float16, device=dev)
    ss = torch.full((M,), 0.1, dtype=torch.float16, device=dev)
    out = torch.empty((M, N), dtype=torch.float16, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for w in ws:
        flush.zero_()   # evict L2 (126 MB)
        assert L.lib().ob_w4a8_gemm_per_chn(L.ptr(x), L.ptr(w), L.ptr(s1), L.ptr(sa), L.ptr(sz), L.ptr(ss), L.ptr(out), M, N, K, N,
                                            L.stream()) == 0
        torch.cuda.synchronize()


if target.startswith("gemm_"):
    gemm(64, *SHAPES[target[5:]])
elif target == "prefill_qkv_pair":
    os.environ["OB_GEMM_2CTA"] = "1"
    gemm(8192, *SHAPES["qkv"])
elif target.startswith("prefill_"):
    gemm(8192, *SHAPES[target[8:]])
elif target == "attention":
    from omniserve_b200.backend import _attn_common as A
    B, Hq, Hkv, ctx = 64, 32, 8, 1280
    pages = ctx // 64 + 2
    pb = Hkv * 64 * 64 + Hkv * 64 * 4
    for _ in range(2):
        kp = torch.randint(0, 256, (B * pages, pb), dtype=torch.uint8, device=dev)
        vp = torch.randint(0, 256, (B * pages, pb), dtype=torch.uint8, device=dev)
        for p_ in (kp, vp):
            sz = p_[:, Hkv * 4096:].view(torch.float16)
            sz[:, :Hkv * 64] = 0.25
            sz[:, Hkv * 64:] = 7.5
        perm = torch.randperm(B * pages, device=dev).view(B, pages)
        tab = torch.empty((B, 2, pages), dtype=torch.int64, device=dev)
        tab[:, 0] = kp.data_ptr() + perm * pb
        tab[:, 1] = vp.data_ptr() + perm * pb
        qkv = torch.randn((B, 6144), dtype=torch.float16, device=dev)
        q3, k3, v3 = qkv[:, :4096].view(B, Hq, 128), qkv[:, 4096:5120].view(B, Hkv, 128), qkv[:, 5120:].view(B, Hkv, 128)
        lens = torch.full((B,), ctx + 1, dtype=torch.int32, device=dev)
        A.single_query(q3, k3, v3, tab, None, None, None, None, lens, 64, Hkv, 0, 0, 0, 0, 0, ctx, 128, 5e5, 1.0)
        torch.cuda.synchronize()
elif target == "selector":
    from omniserve_b200.backend import fused_attention_selector as op
    ctx, Hq, Hkv, Hr = 262144, 32, 8, 4
    n_pages = ctx // 64
    pb = Hr * 64 * 64 + Hr * 64 * 4 + 2 * 4 * Hr * 128 * 2
    flags = torch.tensor([1, 0, 1, 0, 1, 0, 1, 0], dtype=torch.int32, device=dev)
    rank = torch.tensor([0, 0, 1, 1, 2, 2, 3, 3], dtype=torch.int32, device=dev)
    for _ in range(2):
        pool = torch.randint(0, 256, (n_pages, pb), dtype=torch.uint8, device=dev)
        pool[:, Hr * 64 * 64 + Hr * 64 * 4:].view(torch.float16).normal_()
        ptrs = torch.zeros((1, 2, n_pages), dtype=torch.int64, device=dev)
        ptrs[0, 0] = pool.data_ptr() + torch.randperm(n_pages, device=dev) * pb
        q = torch.randn((1, Hq, 128), dtype=torch.float16, device=dev)
        k = torch.randn((1, Hkv, 128), dtype=torch.float16, device=dev)
        lens = torch.tensor([ctx + 1], dtype=torch.int32, device=dev)
        op.single_query_page_selector(q, k, k, ptrs, None, flags, rank, None, lens, None, 1 << 20, 64, Hr * 64, 0, 0, 0, 0, 0, Hr, 0,
                                      ctx, 128, 5e5, 1.0, True, True, True, 16, Hr * 128, 1000000)
        torch.cuda.synchronize()
else:
    raise SystemExit(f"unknown target {target}")
print("done", target)
