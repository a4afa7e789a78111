#!/usr/bin/env python
"""Ablations of the decode GEMM's per-K-block pipeline (instrumented build, see tools/dec_waits.py): graph-timed us per
K-block with parts of the pipeline switched off (results are then wrong; timing only)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from omniserve_b200 import _lib as L  # noqa: E402

dev = "cuda"
torch.cuda.set_device(0)


def run(M, N, K, ctas, reps=8):
    ws = [torch.randint(-128, 128, (N, K // 2), dtype=torch.int8, device=dev) for _ in range(reps)]
    x = torch.randint(-127, 128, (M, K), dtype=torch.int8, device=dev)
    s1 = torch.full((N,), 0.01, dtype=torch.float16, device=dev)
    sa = torch.full((M,), 0.02, dtype=torch.float16, device=dev)
    out = torch.empty((M, N), dtype=torch.float16, device=dev)

    def go():
        for w in ws:
            assert L.lib().ob_w4a8_gemm_ex(0, L.ptr(x), L.ptr(w), 0, 0, L.ptr(s1), L.ptr(sa), L.ptr(s1), L.ptr(sa), L.ptr(out), M, N, K, N,
                                           0, 3, ctas, L.stream()) == 0
    go(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        go()
    best = 1e9
    for _ in range(4):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best / reps * 1e3


for name, mask in (("full", 0), ("skeleton (no MMA/st/actTMA/lds)", 15), ("skeleton, plain arrive instead of commit", 15 | 16),
                   ("skeleton+arrive, unpack ignores ba_empty", 15 | 16 | 32), ("skeleton+arrive, MMA ignores b_full", 15 | 16 | 64),
                   ("skeleton+arrive, MMA ignores a_full", 15 | 16 | 128), ("skeleton+arrive, MMA ignores both", 15 | 16 | 64 | 128),
                   ("skeleton+arrive, nobody waits", 15 | 16 | 32 | 64 | 128)):
    os.environ["OB_DEC_DBG"] = str(mask)
    a = run(64, 18944, 14336, 148)
    b = run(64, 4096, 14336, 32)
    c = run(16, 4096, 14336, 32)
    print(f"{name:44s}: 148x112kb {a:7.2f} us ({a / 112:.3f} us/kb)   32x112kb {b:7.2f} us ({b / 112:.3f})   M=16 32x112kb {c:7.2f} ({c / 112:.3f})",
          flush=True)
