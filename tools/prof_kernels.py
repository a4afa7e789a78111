#!/usr/bin/env python
"""Launch each hot kernel a few times on realistic shapes (for ncu / quick CUDA-event timing).
usage: python tools/prof_kernels.py [gemm_decode] [gemm_prefill] [attention] [small]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from omniserve_b200.model import LlamaConfig, LlamaW4A8, Ops, W4A8Linear  # noqa: E402

what = set(sys.argv[1:]) or {"gemm_decode", "gemm_prefill", "attention", "small"}
dev = "cuda"
torch.cuda.set_device(0)
gen = torch.Generator().manual_seed(0)
ops = Ops()
SHAPES = {"qkv": (6144, 4096), "o": (4096, 4096), "gate_up": (28672, 4096), "down": (4096, 14336)}
REPS = int(os.environ.get("REPS", "3"))
GROUP = int(os.environ.get("GROUP", "-1"))


def ev():
    return torch.cuda.Event(enable_timing=True)


def timeit(fn, n):
    torch.cuda.synchronize()
    a, b = ev(), ev()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def gemm(M, tag):
    for name, (N, K) in SHAPES.items():
        # several distinct weight sets so that successive launches are HBM-cold
        lins = [W4A8Linear(K, N, GROUP, dev, ops).random_init_(gen) for _ in range(6 if M <= 256 else 1)]
        x = torch.randint(-127, 128, (M, K), dtype=torch.int8, device=dev)
        sc = torch.full((M,), 0.02, dtype=torch.float16, device=dev)
        sm = torch.full((M,), 0.1, dtype=torch.float16, device=dev)
        out = torch.empty((M, N), dtype=torch.float16, device=dev)
        state = {"i": 0}

        def run():
            lins[state["i"] % len(lins)](x, sc, sm, out)
            state["i"] += 1
        for _ in range(2):
            run()
        ms = timeit(run, REPS * len(lins))
        by = N * K // 2
        print(f"[{tag}] {name:8s} M={M} N={N} K={K}: {ms * 1e3:8.1f} us  {2.0 * M * N * K / ms / 1e9:8.1f} TOPS  "
              f"weights {by / ms / 1e6:7.1f} GB/s", flush=True)
        del lins, out


if "gemm_decode" in what:
    gemm(64, "gemm_decode")
if "gemm_prefill" in what:
    gemm(8192, "gemm_prefill")

if "attention" in what or "small" in what:
    cfg = LlamaConfig.llama3_8b()
    cfg.num_hidden_layers = 4
    m = LlamaW4A8(cfg, dev)
    B, ctx = 64, 1280
    m.alloc(B, 1536, 64)
    for pool in m.kv.k_pools + m.kv.v_pools:
        pool.view(torch.uint8).random_(0, 256)
        v = pool.view(torch.uint8)[:, 32768:].view(torch.float16)  # plausible fp16 scales / zeros
        v[:, :512] = 0.25
        v[:, 512:] = 7.5
    if "attention" in what:
        qkv = torch.randn((B, 6144), dtype=torch.float16, device=dev)
        q3, k3, v3 = qkv[:, :4096].view(B, 32, 128), qkv[:, 4096:5120].view(B, 8, 128), qkv[:, 5120:].view(B, 8, 128)
        lens = torch.full((B,), ctx + 1, dtype=torch.int32, device=dev)
        st = {"i": 0}

        def attn():
            li = st["i"] % 4
            st["i"] += 1
            ops.fused_attention_pure_dense.single_query_attention(q3, k3, v3, m.kv.tables[li], lens, None, 1536, 64, 512,
                                                                  ctx, 128, 5e5, True, True, True)
        for _ in range(2):
            attn()
        ms = timeit(attn, 4 * REPS)
        by = B * ctx * 8 * 136
        print(f"[attention] bs={B} ctx={ctx}: {ms * 1e3:8.1f} us  {by / ms / 1e6:7.1f} GB/s", flush=True)
    if "small" in what:
        # prefill-sized rows (T = 8192): these ops are pure HBM streams there
        T = 8192
        x = torch.randn((T, 4096), dtype=torch.float16, device=dev)
        d = torch.randn((T, 4096), dtype=torch.float16, device=dev)
        ho = torch.empty_like(x)
        g = torch.ones(4096, dtype=torch.float16, device=dev)
        q = torch.empty((T, 4096), dtype=torch.int8, device=dev)
        sc = torch.empty(T, dtype=torch.float16, device=dev)
        sm = torch.empty(T, dtype=torch.float16, device=dev)
        gu = torch.randn((T, 28672), dtype=torch.float16, device=dev)
        qm = torch.empty((T, 14336), dtype=torch.int8, device=dev)
        qkvp = torch.randn((T, 6144), dtype=torch.float16, device=dev)
        rows = [("add_rms_norm_general", lambda: ops.layernorm_ops.add_rms_norm_general(q, x, d, ho, g, sm, sc, 1e-5), T * 4096 * 7),
                ("rms_norm_general_fuse_sum", lambda: ops.layernorm_ops.rms_norm_general_fuse_sum(q, x, g, sm, sc, 1e-5, True), T * 4096 * 3),
                ("invoke_quant_fuse_sum", lambda: ops.fused_kernels.invoke_quant_fuse_sum(q, x, sm, sc), T * 4096 * 3),
                ("silu_and_mul_quant", lambda: ops.activation_ops.silu_and_mul_quant(qm, gu, sm, sc), T * 14336 * 5)]
        for name, fn, by in rows:
            fn()
            t = timeit(fn, 10)
            print(f"[small T={T}] {name:28s} {t * 1e3:8.1f} us  {by / t / 1e6:7.1f} GB/s", flush=True)
        sl = torch.full((8,), 1024, dtype=torch.int32, device=dev)
        cu = torch.arange(0, 8193, 1024, dtype=torch.int32, device=dev)
        pad = ops.fused_attention_fine_grained_dense.compute_padding_offsets(cu, 1024, T)
        m8 = LlamaW4A8(cfg, dev)
        m8.alloc(8, 1536, 64)
        fl, rk = torch.ones(8, dtype=torch.int32, device=dev), torch.arange(8, dtype=torch.int32, device=dev)

        def kvw():
            ops.fused_attention_fine_grained_dense.apply_bias_rope_update_kv_cache(
                qkvp, sl, None, pad, m8.kv.tables[0], None, fl, rk, 32, 8, 1024, 64, 512, 0, 0, 0, 0, 0, 8, 0, 128, 5e5, 1.0,
                8192, True, True, True)
        kvw()
        t = timeit(kvw, 10)
        by = T * 6144 * 2 + T * 5120 * 2 + T * 2 * 8 * 68
        print(f"[small T={T}] {'apply_bias_rope_update_kv_cache':28s} {t * 1e3:8.1f} us  {by / t / 1e6:7.1f} GB/s", flush=True)
        del x, d, ho, q, gu, qm, qkvp
        x = torch.randn((B, 4096), dtype=torch.float16, device=dev)
        g = torch.ones(4096, dtype=torch.float16, device=dev)
        q = torch.empty((B, 4096), dtype=torch.int8, device=dev)
        sc = torch.empty(B, dtype=torch.float16, device=dev)
        sm = torch.empty(B, dtype=torch.float16, device=dev)
        t1 = timeit(lambda: ops.layernorm_ops.rms_norm_general_fuse_sum(q, x, g, sm, sc, 1e-5, True), 20)
        t2 = timeit(lambda: ops.fused_kernels.invoke_quant_fuse_sum(q, x, sm, sc), 20)
        gu = torch.randn((B, 28672), dtype=torch.float16, device=dev)
        qm = torch.empty((B, 14336), dtype=torch.int8, device=dev)
        t3 = timeit(lambda: ops.activation_ops.silu_and_mul_quant(qm, gu, sm, sc), 20)
        print(f"[small] rmsnorm_quant {t1 * 1e3:.1f} us, quant {t2 * 1e3:.1f} us, silu_mul_quant {t3 * 1e3:.1f} us")
