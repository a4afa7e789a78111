#!/usr/bin/env python
"""Graph-timed W4A8 GEMM micro-benchmarks: separates fixed per-launch cost, per-K-block cost and split-K cost."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from omniserve_b200 import _lib as L  # noqa: E402

dev = "cuda"
torch.cuda.set_device(0)
NSETS = 16


def run(M, N, K, mode=-1, bn=0, ctas=0, tag=""):
    ws = [torch.randint(-128, 128, (N, K // 2), dtype=torch.int8, device=dev) for _ in range(NSETS)]
    x = torch.randint(-127, 128, (M, K), dtype=torch.int8, device=dev)
    s1 = torch.full((N,), 0.01, dtype=torch.float16, device=dev)
    sz = torch.full((N,), 0.08, dtype=torch.float16, device=dev)
    sa = torch.full((M,), 0.02, dtype=torch.float16, device=dev)
    ss = torch.full((M,), 0.1, dtype=torch.float16, device=dev)
    out = torch.empty((M, N), dtype=torch.float16, device=dev)

    def launch_all():
        for w in ws:
            c = L.lib().ob_w4a8_gemm_ex(0, L.ptr(x), L.ptr(w), 0, 0, L.ptr(s1), L.ptr(sa), L.ptr(sz), L.ptr(ss), L.ptr(out),
                                        M, N, K, N, bn, mode, ctas, L.stream())
            assert c == 0
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
    by = N * K // 2
    print(f"{tag:28s} M={M:5d} N={N:6d} K={K:6d} mode={mode:2d} bn={bn:3d} ctas={ctas:3d}: {us:8.2f} us  "
          f"{by / us / 1e3:8.1f} GB/s  {2.0 * M * N * K / us / 1e6:8.1f} TOPS", flush=True)


def chain(M, shapes, mode, ctas, layers=8, tag=""):
    """The four GEMMs of `layers` decoder layers back to back in ONE graph (PDL-chained, distinct weights per layer):
    what the decode step's GEMMs cost when each launch can overlap its predecessor's tail."""
    ws = [[torch.randint(-128, 128, (N, K // 2), dtype=torch.int8, device=dev) for (N, K) in shapes] for _ in range(layers)]
    Kmax, Nmax = max(k for _, k in shapes), max(n for n, _ in shapes)
    x = torch.randint(-127, 128, (M, Kmax), dtype=torch.int8, device=dev)
    s1 = torch.full((Nmax,), 0.01, dtype=torch.float16, device=dev)
    sz = torch.full((Nmax,), 0.08, dtype=torch.float16, device=dev)
    sa = torch.full((M,), 0.02, dtype=torch.float16, device=dev)
    ss = torch.full((M,), 0.1, dtype=torch.float16, device=dev)
    outs = [torch.empty((M, N), dtype=torch.float16, device=dev) for (N, K) in shapes]
    xs = [x[:, :K].contiguous() for (N, K) in shapes]

    def launch_all():
        for l in range(layers):
            for i, (N, K) in enumerate(shapes):
                c = L.lib().ob_w4a8_gemm_ex(0, L.ptr(xs[i]), L.ptr(ws[l][i]), 0, 0, L.ptr(s1), L.ptr(sa), L.ptr(sz), L.ptr(ss),
                                            L.ptr(outs[i]), M, N, K, N, 0, mode, ctas, L.stream())
                assert c == 0
    launch_all()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        launch_all()
    best = 1e9
    for _ in range(5):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    us = best / layers * 1e3
    by = sum(N * K // 2 for N, K in shapes)
    print(f"{tag:34s} M={M:3d} mode={mode:2d} ctas={ctas:3d}: {us:8.2f} us per layer (4 GEMMs)  {by / us / 1e3:8.1f} GB/s", flush=True)


if __name__ == "__main__" and os.environ.get("OB_UPC_SWEEP"):
    # decode kernel: time every "units per CTA" choice (through force_ctas) for the shapes that matter
    SH = {"qkv": (6144, 4096), "o": (4096, 4096), "gate_up": (28672, 4096), "down": (4096, 14336),
          "tp8 qkv": (768, 4096), "tp8 o": (4096, 512), "tp8 gate_up": (3584, 4096), "tp8 down": (4096, 1792),
          "tp2 qkv": (3072, 4096), "tp2 o": (4096, 2048), "tp2 gate_up": (14336, 4096), "tp2 down": (4096, 7168)}
    for nm, (N, K) in SH.items():
        tiles, KB = N // 128, K // 128
        units = tiles * KB
        seen = set()
        for upc in sorted(set([max(1, (units + 295) // 296), 1, 2, 4, 7, 8, 14, 16, 25, 28, 32, 56, KB])):
            if upc > KB or upc < (units + 295) // 296:
                continue
            ctas = (units + upc - 1) // upc
            if ctas in seen or ctas > 296:
                continue
            seen.add(ctas)
            run(64, N, K, mode=3, ctas=ctas, tag=f"{nm} upc={upc}")
    sys.exit(0)

if __name__ == "__main__" and os.environ.get("OB_DEC_EXP"):
    LL = [(6144, 4096), (4096, 4096), (28672, 4096), (4096, 14336)]
    names = ["qkv", "o_proj", "gate_up", "down"]
    for M in (64, 16):
        for (N, K), nm in zip(LL, names):
            run(M, N, K, mode=1, tag=f"{nm} v1 stream-K")
            run(M, N, K, mode=3, ctas=0, tag=f"{nm} v2 auto")
            run(M, N, K, mode=3, ctas=148, tag=f"{nm} v2 148 CTAs")
            run(M, N, K, mode=3, ctas=296, tag=f"{nm} v2 296 CTAs")
    run(64, 18944, 14336, mode=0, tag="v1 148 tiles x 112 kb")
    run(64, 18944, 14336, mode=3, ctas=148, tag="v2 148 x 112 kb")
    run(64, 18944, 14336, mode=3, ctas=296, tag="v2 296 x 56 kb")
    run(64, 18944, 128, mode=3, ctas=148, tag="v2 fixed cost (148 x 1 kb)")
    run(64, 4096, 14336, mode=3, ctas=32, tag="v2 32 CTAs x 112 kb")
    for M in (64,):
        chain(M, LL, 1, 0, tag="layer chain v1 stream-K")
        chain(M, LL, 3, 0, tag="layer chain v2 auto")
        chain(M, LL, 3, 148, tag="layer chain v2 148")
        chain(M, LL, 3, 296, tag="layer chain v2 296")
    # TP=8 shards of the same layer (strong scaling: per-GPU shapes)
    TP8 = [(768, 4096), (4096, 512), (3584, 4096), (4096, 1792)]
    chain(64, TP8, 1, 0, tag="tp8 shard chain v1")
    chain(64, TP8, 3, 0, tag="tp8 shard chain v2 auto")
    sys.exit(0)

if __name__ == "__main__" and os.environ.get("OB_CL_EXP"):
    # split-K reduction path A/B: run once with OB_GEMM_DEC_CLUSTER=0 (L2 reds + last-CTA finalise) and once with =1
    # (cluster, distributed shared memory); -> profiles/r2_cluster_splitk.log
    print("OB_GEMM_DEC_CLUSTER =", os.environ.get("OB_GEMM_DEC_CLUSTER"))
    LL = [(6144, 4096), (4096, 4096), (28672, 4096), (4096, 14336)]
    for M in (64, 16):
        for (N, K), nm in zip(LL, ["qkv", "o_proj", "gate_up", "down"]):
            run(M, N, K, mode=3, ctas=0, tag=f"{nm} v2 auto")
    chain(64, LL, 3, 0, tag="layer chain v2 auto")
    chain(16, LL, 3, 0, tag="layer chain v2 auto")
    chain(64, [(768, 4096), (4096, 512), (3584, 4096), (4096, 1792)], 3, 0, tag="tp8 shard chain v2 auto")
    chain(64, [(3072, 4096), (4096, 2048), (14336, 4096), (4096, 7168)], 3, 0, tag="tp2 shard chain v2 auto")
    chain(16, [(10240, 8192), (8192, 8192), (57344, 8192), (8192, 28672)], 3, 0, layers=2, tag="70B layer chain (bs 16)")
    sys.exit(0)

if __name__ == "__main__" and os.environ.get("OB_2CTA_SWEEP"):
    for two in ("0", "1"):
        os.environ["OB_GEMM_2CTA"] = two
        for M in (256, 512, 1024, 2048, 4096, 8192):
            run(M, 6144, 4096, mode=0, bn=128, tag=f"qkv 2cta={two}")
    sys.exit(0)

if __name__ == "__main__" and os.environ.get("OB_2CTA_EXP"):
    for two in ("0", "1"):
        os.environ["OB_GEMM_2CTA"] = two
        run(8192, 6144, 4096, tag=f"prefill qkv 2cta={two}")
        run(8192, 28672, 4096, tag=f"prefill gate_up 2cta={two}")
        run(8192, 4096, 14336, tag=f"prefill down 2cta={two}")
        run(8192, 4096, 4096, tag=f"prefill o 2cta={two}")
        run(2048, 28672, 4096, tag=f"M=2048 gate_up 2cta={two}")
        run(512, 28672, 4096, tag=f"M=512 gate_up 2cta={two}")
    sys.exit(0)

if __name__ == "__main__" and os.environ.get("OB_MC_EXP"):
    for mc in ("1", "2", "4"):
        os.environ["OB_GEMM_MC"] = mc
        run(8192, 6144, 4096, tag=f"prefill qkv mc={mc}")
        run(8192, 28672, 4096, tag=f"prefill gate_up mc={mc}")
        run(8192, 4096, 14336, tag=f"prefill down mc={mc}")
        run(8192, 4096, 4096, tag=f"prefill o mc={mc}")
    del os.environ["OB_GEMM_MC"]
    run(2048, 28672, 4096, tag="M=2048 gate_up auto")
    run(512, 28672, 4096, tag="M=512 gate_up auto")
    sys.exit(0)

if __name__ == "__main__" and os.environ.get("OB_STREAM_EXP"):
    # what bounds the per-K-block time of the load pipeline?  (results invalid for dbg != 0: timing only)
    for w2k in ("0", "1"):
        os.environ["OB_GEMM_W2K"] = w2k
        for dbg, what in ((0, "full"), (6, "no unpack, no MMA"), (6 + 32, "W only (no act TMA)"), (6 + 64, "act only (no W TMA)")):
            os.environ["OB_GEMM_DBG"] = str(dbg)
            for M in (64, 16):
                run(M, 18944, 14336, mode=0, tag=f"w2k={w2k} {what} M={M}")
            run(64, 4096, 14336, mode=0, tag=f"w2k={w2k} {what} 32 CTAs")
    os.environ["OB_GEMM_DBG"] = "0"
    for w2k in ("0", "1"):
        os.environ["OB_GEMM_W2K"] = w2k
        run(64, 28672, 4096, tag=f"gate_up auto w2k={w2k}")
        run(64, 4096, 14336, tag=f"down auto w2k={w2k}")
        run(8192, 6144, 4096, tag=f"prefill qkv w2k={w2k}")
    sys.exit(0)

if __name__ == "__main__" and os.environ.get("OB_CLUSTER_EXP"):
    print("cluster experiment: OB_GEMM_DBG=%s OB_NO_PDL=%s" % (os.environ.get("OB_GEMM_DBG"), os.environ.get("OB_NO_PDL")))
    for k in (2, 4):
        run(64, 4096, 4096, mode=2, ctas=k, tag=f"o_proj cluster k={k}")
    sys.exit(0)

if __name__ == "__main__" and os.environ.get("OB_GEMM_DBG"):
    print("OB_GEMM_DBG =", os.environ["OB_GEMM_DBG"], "(results invalid, timing only)")
    run(64, 18944, 4096, mode=0, tag="148 tiles x 32 kb, no split")
    run(64, 18944, 14336, mode=0, tag="148 tiles x 112 kb, no split")
    run(8192, 6144, 4096, tag="prefill qkv")
    sys.exit(0)

if __name__ == "__main__":
    run(64, 4096, 4096, tag="o_proj auto (cluster)")
    run(64, 6144, 4096, tag="qkv auto (cluster)")
    run(64, 28672, 4096, tag="gate_up auto")
    run(64, 4096, 14336, tag="down auto (cluster)")
    run(64, 4096, 4096, mode=2, ctas=2, tag="o_proj cluster k=2")
    run(64, 4096, 4096, mode=2, ctas=8, tag="o_proj cluster k=8")
    run(64, 4096, 14336, mode=2, ctas=8, tag="down cluster k=8")
    run(16, 4096, 4096, tag="o_proj M=16 auto")
    run(64, 18944, 128, mode=0, tag="fixed cost (148 tiles,1kb)")
    run(64, 18944, 1024, mode=0, tag="148 tiles x 8 kb, no split")
    run(64, 18944, 4096, mode=0, tag="148 tiles x 32 kb, no split")
    run(64, 18944, 14336, mode=0, tag="148 tiles x 112 kb, no split")
    run(64, 4096, 4096, mode=0, tag="o_proj, 32 CTAs no split")
    run(64, 4096, 4096, mode=1, tag="o_proj stream-K")
    run(64, 6144, 4096, mode=1, tag="qkv stream-K")
    run(64, 28672, 4096, mode=1, tag="gate_up stream-K")
    run(64, 28672, 4096, mode=0, tag="gate_up tiles (224)")
    run(64, 4096, 14336, mode=1, tag="down stream-K")
    run(16, 4096, 4096, mode=1, tag="o_proj M=16")
    run(128, 28672, 4096, tag="gate_up M=128")
    run(8192, 6144, 4096, tag="prefill qkv")
    run(8192, 28672, 4096, tag="prefill gate_up")
    run(8192, 4096, 14336, tag="prefill down")
