#!/usr/bin/env python
"""Where does prefill time go?  Times one 8 x 1024-token chunk of the Llama-3-8B W4A8KV4 prefill per op class
(CUDA events, eager) and compares the candidate fp16 prefill-attention libraries (torch SDPA backends, flash_attn)."""
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

dev = "cuda"
torch.cuda.set_device(0)


def ev(fn, n=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def attention_candidates():
    B, L, Hq, Hkv, Dh = 8, 1024, 32, 8, 128
    qkv = torch.randn(B * L, (Hq + 2 * Hkv) * Dh, device=dev, dtype=torch.float16)
    q = qkv[:, :Hq * Dh].view(B * L, Hq, Dh)
    k = qkv[:, Hq * Dh:(Hq + Hkv) * Dh].view(B * L, Hkv, Dh)
    v = qkv[:, (Hq + Hkv) * Dh:].view(B * L, Hkv, Dh)
    flops = 4 * B * Hq * L * L * Dh / 2

    def sdpa(gqa=True):
        qq = q.reshape(B, L, Hq, Dh).transpose(1, 2)
        kk = k.reshape(B, L, Hkv, Dh).transpose(1, 2)
        vv = v.reshape(B, L, Hkv, Dh).transpose(1, 2)
        if not gqa:
            kk = kk.repeat_interleave(Hq // Hkv, dim=1)
            vv = vv.repeat_interleave(Hq // Hkv, dim=1)
            return F.scaled_dot_product_attention(qq, kk, vv, is_causal=True)
        return F.scaled_dot_product_attention(qq, kk, vv, is_causal=True, enable_gqa=True)

    from torch.nn.attention import SDPBackend, sdpa_kernel
    for name, be in (("default", None), ("flash", SDPBackend.FLASH_ATTENTION), ("cudnn", SDPBackend.CUDNN_ATTENTION),
                     ("efficient", SDPBackend.EFFICIENT_ATTENTION)):
        for gqa in (True, False):
            try:
                if be is None:
                    ms = ev(lambda: sdpa(gqa))
                else:
                    with sdpa_kernel(be):
                        ms = ev(lambda: sdpa(gqa))
                print(f"sdpa[{name:9s}] gqa={gqa}: {ms:8.3f} ms  {flops / ms / 1e9:7.1f} TFLOP/s", flush=True)
            except Exception as e:  # noqa: BLE001
                print(f"sdpa[{name:9s}] gqa={gqa}: FAILED {str(e)[:100]}", flush=True)
    try:
        from flash_attn import flash_attn_varlen_func
        cu = torch.arange(0, (B + 1) * L, L, device=dev, dtype=torch.int32)
        ms = ev(lambda: flash_attn_varlen_func(q, k, v, cu, cu, L, L, causal=True))
        print(f"flash_attn_varlen_func      : {ms:8.3f} ms  {flops / ms / 1e9:7.1f} TFLOP/s", flush=True)
    except Exception as e:  # noqa: BLE001
        print("flash_attn FAILED", str(e)[:200], flush=True)


def chunk_breakdown():
    from omniserve_b200.model import LlamaConfig, LlamaW4A8
    cfg = LlamaConfig.llama3_8b()
    cfg.num_hidden_layers = 4
    m = LlamaW4A8(cfg, dev)
    m.alloc(8, 1536, 8192)
    toks = torch.randint(0, cfg.vocab_size, (8192,), device=dev)
    for i in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        m.prefill(toks, [1024] * 8)
        torch.cuda.synchronize()
        print(f"prefill chunk (4 layers) call {i}: {(time.perf_counter() - t0) * 1e3:.1f} ms wall", flush=True)
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        m.prefill(toks, [1024] * 8)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=70))


if __name__ == "__main__":
    attention_candidates()
    chunk_breakdown()
