#!/usr/bin/env python
"""Graph-timed KV4 decode-attention micro-benchmark: context sweep at bs=64 (Llama-3-8B heads) to separate the
per-CTA fixed cost (prologue / merge) from the per-page streaming cost, plus split sweeps."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from omniserve_b200.backend import _attn_common as A  # noqa: E402

dev = "cuda"
torch.cuda.set_device(0)
B, Hq, Hkv, Dh = 64, 32, 8, 128
NSETS = 8


def run(ctx, split=0, B=B, tag=""):
    pages = (ctx + 64) // 64 + 1
    page_bytes = Hkv * 64 * 64 + Hkv * 64 * 4
    pools = []
    for _ in range(NSETS):
        kp = torch.randint(0, 256, (B * pages, page_bytes), dtype=torch.uint8, device=dev)
        vp = torch.randint(0, 256, (B * pages, page_bytes), dtype=torch.uint8, device=dev)
        for p_ in (kp, vp):
            sz = p_[:, Hkv * 4096:].view(torch.float16)
            sz[:, :Hkv * 64] = 0.25
            sz[:, Hkv * 64:] = 7.5
        perm = torch.randperm(B * pages, device=dev).view(B, pages)
        tab = torch.empty((B, 2, pages), dtype=torch.int64, device=dev)
        tab[:, 0] = kp.data_ptr() + perm * page_bytes
        tab[:, 1] = vp.data_ptr() + perm * page_bytes
        pools.append((kp, vp, tab))
    qkv = torch.randn((B, (Hq + 2 * Hkv) * Dh), dtype=torch.float16, device=dev)
    q3 = qkv[:, :Hq * Dh].view(B, Hq, Dh)
    k3 = qkv[:, Hq * Dh:(Hq + Hkv) * Dh].view(B, Hkv, Dh)
    v3 = qkv[:, (Hq + Hkv) * Dh:].view(B, Hkv, Dh)
    lens = torch.full((B,), ctx + 1, dtype=torch.int32, device=dev)

    def launch_all():
        for _, _, tab in pools:
            A.single_query(q3, k3, v3, tab, None, None, None, None, lens, 64, Hkv, 0, 0, 0, 0, 0, ctx, 128, 5e5, 1.0,
                           force_split=split)
    launch_all()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        launch_all()
    best = 1e9
    for _ in range(5):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    us = best / NSETS * 1e3
    by = B * ctx * Hkv * 136
    print(f"{tag:20s} bs={B:3d} ctx={ctx:6d} split={split}: {us:8.2f} us  {by / us / 1e3:8.1f} GB/s", flush=True)


if __name__ == "__main__":
    for ctx in (63, 320, 640, 1280, 2560, 5120):
        run(ctx)
    for s in (2, 3):
        run(1280, split=s, tag="forced split")
    run(1280, B=37, tag="bs=37")
    run(8192, B=8, tag="bs=8 long")
    run(32768, B=1, tag="bs=1 32K")
