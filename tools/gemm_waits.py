#!/usr/bin/env python
"""Where do the warp roles of the W4A8 GEMM wait?  Runs one prefill-shaped launch per variant with OB_GEMM_DBGT set and
prints, per role, the cycles spent inside mbarrier waits (mean over CTAs; leader / peer separately for the CTA pair).
Needs the instrumented build:  OB_GEMM_TIMING=1 python -m omniserve_b200.build --force  (the counters are compiled out of
the product build: they cost ~20 % of the kernel's throughput even when disabled at run time)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from omniserve_b200 import _lib as L  # noqa: E402

dev = "cuda"
torch.cuda.set_device(0)
SLOTS = ["Wprod:w_empty", "Bprod:ba_empty", "MMA:acc_empty", "MMA:b_full", "MMA:a_full", "MMA:peer_ready", "unpack:w_full",
         "unpack:ba_empty", "epi:acc_full", "kernel cycles"]


def run(M, N, K, two):
    os.environ["OB_GEMM_2CTA"] = two
    w = torch.randint(-128, 128, (N, K // 2), dtype=torch.int8, device=dev)
    x = torch.randint(-127, 128, (M, K), dtype=torch.int8, device=dev)
    s1 = torch.full((N,), 0.01, dtype=torch.float16, device=dev)
    sz = torch.full((N,), 0.08, dtype=torch.float16, device=dev)
    sa = torch.full((M,), 0.02, dtype=torch.float16, device=dev)
    ss = torch.full((M,), 0.1, dtype=torch.float16, device=dev)
    out = torch.empty((M, N), dtype=torch.float16, device=dev)
    dbg = torch.zeros((148, 16), dtype=torch.int64, device=dev)
    os.environ["OB_GEMM_DBGT"] = "0"
    for _ in range(2):
        os.environ.pop("OB_GEMM_DBGT", None)
        L.lib().ob_w4a8_gemm_per_chn(L.ptr(x), L.ptr(w), L.ptr(s1), L.ptr(sa), L.ptr(sz), L.ptr(ss), L.ptr(out), M, N, K, N, L.stream())
    torch.cuda.synchronize()
    os.environ["OB_GEMM_DBGT"] = str(dbg.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    L.lib().ob_w4a8_gemm_per_chn(L.ptr(x), L.ptr(w), L.ptr(s1), L.ptr(sa), L.ptr(sz), L.ptr(ss), L.ptr(out), M, N, K, N, L.stream())
    e1.record()
    torch.cuda.synchronize()
    os.environ.pop("OB_GEMM_DBGT", None)
    d = dbg.cpu().double()
    print(f"--- M={M} N={N} K={K} 2cta={two}: {e0.elapsed_time(e1) * 1e3:.1f} us")
    groups = [("all CTAs", slice(None))] if two == "0" else [("leaders", slice(0, None, 2)), ("peers", slice(1, None, 2))]
    for name, sl in groups:
        m = d[sl].mean(0)
        tot = float(m[9]) or 1.0
        print("   " + name + ": " + "  ".join(f"{SLOTS[i]}={m[i] / tot * 100:5.1f}%" for i in range(9)) + f"  ({tot:.0f} cycles)")


if __name__ == "__main__":
    for two in ("0", "1"):
        run(8192, 6144, 4096, two)
    run(64, 18944, 14336, "0")
