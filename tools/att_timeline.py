#!/usr/bin/env python
"""Globaltimer timeline of the decode attention inside its PDL chain (qkv GEMM -> attention + fused output quant -> o_proj GEMM),
bs=64, Llama-3-8B heads, ctx 1280, CUDA-graph replay.  Needs the instrumented library variant:

    OB_ATT_TIMING=1 OB_DEC_TIMING=1 OB_BUILD_LIBDIR=lib_timing python -m omniserve_b200.build --force
    OB_LIB_PATH=omniserve_b200/lib_timing/libomniserve_b200.so python tools/att_timeline.py     # -> profiles/r2_att_timeline.log
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
dev = "cuda"
torch.cuda.set_device(0)
gdbg = torch.zeros((64, 1024, 32), dtype=torch.int64, device=dev)
adbg = torch.zeros((4096, 16), dtype=torch.int64, device=dev)
os.environ["OB_DEC_DBGT"] = str(gdbg.data_ptr())
os.environ["OB_ATT_DBGT"] = str(adbg.data_ptr())
from omniserve_b200 import _lib as L  # noqa: E402
from omniserve_b200.backend import _attn_common as A  # noqa: E402

B, Hq, Hkv, Dh, H = 64, 32, 8, 128, 4096
LAYERS = 3


def main(ctx=1280, quant=True, stable=True, dbg=0):
    os.environ["OB_ATT_DBG"] = str(dbg)      # instrumented library only; 1 = no math, 2 = no scale-row copies (results invalid)
    pages = (ctx + 64) // 64 + 1
    page_bytes = Hkv * 64 * 64 + Hkv * 64 * 4
    tabs, keep = [], []
    for _ in range(LAYERS):
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
        tabs.append(tab); keep.append((kp, vp))
    NQ = (Hq + 2 * Hkv) * Dh
    wq = [torch.randint(-128, 128, (NQ, H // 2), dtype=torch.int8, device=dev) for _ in range(LAYERS)]
    wo = [torch.randint(-128, 128, (H, H // 2), dtype=torch.int8, device=dev) for _ in range(LAYERS)]
    x = torch.randint(-127, 128, (B, H), dtype=torch.int8, device=dev)
    s1 = torch.full((NQ,), 0.002, dtype=torch.float16, device=dev)
    sz = torch.full((NQ,), 0.01, dtype=torch.float16, device=dev)
    sa = torch.full((B,), 0.02, dtype=torch.float16, device=dev)
    ss = torch.full((B,), 0.1, dtype=torch.float16, device=dev)
    qkv = torch.empty((B, NQ), dtype=torch.float16, device=dev)
    od = torch.empty((B, H), dtype=torch.float16, device=dev)
    q3 = qkv[:, :Hq * Dh].view(B, Hq, Dh)
    k3 = qkv[:, Hq * Dh:(Hq + Hkv) * Dh].view(B, Hkv, Dh)
    v3 = qkv[:, (Hq + Hkv) * Dh:].view(B, Hkv, Dh)
    lens = torch.full((B,), ctx + 1, dtype=torch.int32, device=dev)
    qa = torch.empty((B, H), dtype=torch.int8, device=dev)
    qs = torch.empty((B,), dtype=torch.float16, device=dev)
    qsum = torch.empty((B,), dtype=torch.float16, device=dev)

    def go():
        for l in range(LAYERS):
            assert L.lib().ob_w4a8_gemm_ex(0, L.ptr(x), L.ptr(wq[l]), 0, 0, L.ptr(s1), L.ptr(sa), L.ptr(sz), L.ptr(ss), L.ptr(qkv),
                                           B, NQ, H, NQ, 0, 3, 0, L.stream()) == 0
            A.single_query(q3, k3, v3, tabs[l], None, None, None, None, lens, 64, Hkv, 0, 0, 0, 0, 0, ctx, 128, 5e5, 1.0,
                           quant=(qa, qs, qsum) if quant else None, history_is_stable=stable)
            assert L.lib().ob_w4a8_gemm_ex(0, L.ptr(qa), L.ptr(wo[l]), 0, 0, L.ptr(s1), L.ptr(qs), L.ptr(sz), L.ptr(qsum), L.ptr(od),
                                           B, H, H, H, 0, 3, 0, L.stream()) == 0
    go()
    torch.cuda.synchronize()
    n_gemm = 2 * LAYERS
    first = n_gemm * (main.calls * 2 + 1)      # eager launches, then capture-time launches own the slots the replay writes
    main.calls += 1
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        go()
    gdbg.zero_(); adbg.zero_()
    g.replay()
    torch.cuda.synchronize()
    gd, ad = gdbg.cpu(), adbg.cpu()
    # the last layer's launches: qkv GEMM, attention (the attention buffer holds the last launch), o_proj GEMM
    bq = gd[(first + n_gemm - 2) % 64]; bo = gd[(first + n_gemm - 1) % 64]
    bq, bo = bq[bq[:, 24] > 0], bo[bo[:, 24] > 0]
    a = ad[ad[:, 0] > 0]
    t0 = int(bq[:, 24].min())
    us = lambda t: (int(t) - t0) / 1e3  # noqa: E731
    print(f"== ctx {ctx}, fused quant {quant}, history_is_stable {stable}, debug mode {dbg}: us relative to the qkv GEMM's first CTA entry")
    print(f"   qkv GEMM  ({len(bq)} CTAs): entry {us(bq[:, 24].min()):7.2f}  acc done {us(bq[:, 28].max()):7.2f}  last exit {us(bq[:, 30].max()):7.2f}")
    names = {0: "entry", 1: "pre-dependency prologue done", 2: "dependency resolved", 3: "page loop starts", 4: "page loop done",
             5: "outputs stored", 6: "exit", 8: "producer: first page issued", 9: "producer: last page issued"}
    print(f"   attention ({len(a)} CTAs): slot: min / median / max")
    for k in (0, 8, 1, 2, 3, 9, 4, 5, 6):
        col = a[:, k]
        col = col[col > 0]
        if len(col):
            print(f"      {names[k]:32s} {us(col.min()):7.2f} / {us(col.median()):7.2f} / {us(col.max()):7.2f}")
    d = (a[:, 4] - a[:, 3]).float() / 1e3
    print(f"      page loop duration per CTA: min {float(d.min()):.2f} median {float(d.median()):.2f} max {float(d.max()):.2f} us; "
          f"loop start -> dep: {float(((a[:, 3] - a[:, 2]).float() / 1e3).median()):.2f} us; loop end -> stored: "
          f"{float(((a[:, 5] - a[:, 4]).float() / 1e3).median()):.2f} us; stored -> exit: median "
          f"{float(((a[:, 6] - a[:, 5]).float() / 1e3).median()):.2f} max {float(((a[:, 6] - a[:, 5]).float() / 1e3).max()):.2f} us")
    print(f"   o_proj GEMM ({len(bo)} CTAs): entry {us(bo[:, 24].min()):7.2f}  dependency resolved {us(bo[:, 27].min()):7.2f}..{us(bo[:, 27].max()):7.2f}"
          f"  last exit {us(bo[:, 30].max()):7.2f}")


main.calls = 0
if __name__ == "__main__":
    main(1280, True, True)
    main(1280, False, True)
    main(1280, True, False)
    main(320, True, True)
    main(2560, True, True)
    print("# ---- experiments (outputs invalid, timing only): what bounds the page loop")
    main(1280, False, True, dbg=1)      # loads only
    main(1280, False, True, dbg=3)      # loads only, without the four 128-byte scale / zero copies per page
    main(1280, False, True, dbg=2)      # math, without the scale / zero copies
