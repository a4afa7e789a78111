#!/usr/bin/env python
"""Globaltimer timeline of PDL-chained decode GEMM launches (w4a8_gemm_decode.cu).  Needs the instrumented build:

    OB_DEC_TIMING=1 python -m omniserve_b200.build --force && python tools/dec_waits.py ; python -m omniserve_b200.build --force
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
dev = "cuda"
torch.cuda.set_device(0)
dbg = torch.zeros((64, 1024, 32), dtype=torch.int64, device=dev)   # one block per launch (round robin)
os.environ["OB_DEC_DBGT"] = str(dbg.data_ptr())
from omniserve_b200 import _lib as L  # noqa: E402

U = ["wait w_full", "wait ba_empty", "wait::st", "wait acc_full", "epilogue", "lds + arrive w_empty", "convert + tcgen05.st issue",
     "fence + arrive a_full"]
launches = 0


def timeline(shapes, M=64, layers=3, tag=""):
    """Globaltimer stamps of a PDL chain of GEMMs (graph replay): when does each launch enter, finish its prologue, get its
    dependency resolved, complete its accumulators, finish its epilogue and exit -- relative to the first launch's entry."""
    global launches
    ws = [[torch.randint(-128, 128, (N, K // 2), dtype=torch.int8, device=dev) for (N, K) in shapes] for _ in range(layers)]
    Kmax, Nmax = max(k for _, k in shapes), max(n for n, _ in shapes)
    x = torch.randint(-127, 128, (M, Kmax), dtype=torch.int8, device=dev)
    xs = [x[:, :K].contiguous() for (N, K) in shapes]
    s1 = torch.full((Nmax,), 0.01, dtype=torch.float16, device=dev)
    sz = torch.full((Nmax,), 0.08, dtype=torch.float16, device=dev)
    sa = torch.full((M,), 0.02, dtype=torch.float16, device=dev)
    ss = torch.full((M,), 0.1, dtype=torch.float16, device=dev)
    outs = [torch.empty((M, N), dtype=torch.float16, device=dev) for (N, K) in shapes]

    def go():
        for l in range(layers):
            for i, (N, K) in enumerate(shapes):
                assert L.lib().ob_w4a8_gemm_ex(0, L.ptr(xs[i]), L.ptr(ws[l][i]), 0, 0, L.ptr(s1), L.ptr(sa), L.ptr(sz), L.ptr(ss),
                                               L.ptr(outs[i]), M, N, K, N, 0, 3, 0, L.stream()) == 0
    go()
    torch.cuda.synchronize()
    first = launches + layers * len(shapes)      # the launches made during capture do not run; replay reuses their slots
    launches += layers * len(shapes)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        go()
    launches += layers * len(shapes)
    dbg.zero_()
    g.replay()
    torch.cuda.synchronize()
    d = dbg.cpu()
    print(f"== timeline {tag}: us relative to the first launch's first CTA entry; min..max over CTAs")
    t0 = None
    for j in range(layers * len(shapes)):
        blk = d[(first + j) % 64]
        used = blk[:, 24] > 0
        if not used.any():
            continue
        b = blk[used]
        if t0 is None:
            t0 = int(b[:, 24].min())
        N, K = shapes[j % len(shapes)]
        f = lambda c: f"{(int(b[:, c].min()) - t0) / 1e3:7.2f}..{(int(b[:, c].max()) - t0) / 1e3:7.2f}"  # noqa: E731
        print(f"   launch {j:2d} N={N:6d} K={K:6d} grid={int(used.sum()):3d}: entry {f(24)} | prologue {f(25)} | unpacked {f(26)} | "
              f"dep {f(27)} | acc {f(28)} | epi {f(29)} | exit {f(30)}")


LL = [(6144, 4096), (4096, 4096), (28672, 4096), (4096, 14336)]
timeline(LL, tag="Llama-3-8B layer GEMMs, M=64")
timeline([(768, 4096), (4096, 512), (3584, 4096), (4096, 1792)], tag="TP=8 shards, M=64")
