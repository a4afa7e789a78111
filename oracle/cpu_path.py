"""torch-CPU restatement of one W4A8KV4 decoder layer (decode step), vectorised: the "torch-fp16 CPU path on
the host cores" that BASELINE.json's north star asks to be timed next to the GPU numbers.

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Used only by bench.py's cpu_baseline leg and tests.
Same math as oracle/{w4a8,act,kv4}.py (which follow the reference file:line cited there) with fp16 storage and
fp32 arithmetic; `tests/test_oracle_cpu_path.py` pins it against those loop oracles on small cases.
"""
from __future__ import annotations

import math

import torch


def quant_per_token(x: torch.Tensor):
    """fp16 [T,H] -> (int8-valued fp32 [T,H], scale fp16 [T], sum fp16 [T])  (fused_kernels.cu:97-142)."""
    xf = x.float()
    amax = xf.abs().amax(dim=1)
    q = torch.clamp(torch.round(xf * (127.0 / amax)[:, None]), -128, 127)
    return q, (amax / 127.0).half(), xf.sum(dim=1).half()


def norm_quant(x: torch.Tensor, gamma: torch.Tensor, eps: float):
    """layernorm_kernels.cu:194-331 semantics without the fp16 partial-sum quirk on the sum."""
    xf = x.float()
    mean = xf.mean(dim=1, keepdim=True)
    rstd = torch.rsqrt((xf * xf).mean(dim=1, keepdim=True) + eps)
    v = (xf - mean) * rstd * gamma.float()
    vh = v.half().float()
    amax = vh.abs().amax(dim=1).clamp_min(1e-6)
    q = torch.clamp(torch.round(v * (127.0 / amax)[:, None]), -128, 127)
    return q, (amax / 127.0).half(), vh.sum(dim=1).half()


def w4a8_linear(q_act, sa, ssum, w_u4_f32, s1, szs):
    """(A.Wu4^T)*s1*sa - sz*ssum  (per_chn/gemm_cuda.cu:583-590); operands held in fp32 (exact for K<=4096)."""
    acc = q_act @ w_u4_f32.t()
    return ((acc * s1.float()[None, :]) * sa.float()[:, None] - szs.float()[None, :] * ssum.float()[:, None]).half()


def rope_neox(x: torch.Tensor, pos: torch.Tensor, base: float):
    """x fp16 [B,H,Dh], pos [B]."""
    Dh = x.shape[-1]
    half = Dh // 2
    inv = 1.0 / (base ** (torch.arange(0, half, dtype=torch.float32) * 2.0 / Dh))
    ang = pos.float()[:, None] * inv[None, :]
    c, s = torch.cos(ang)[:, None, :], torch.sin(ang)[:, None, :]
    xf = x.float()
    a, b = xf[..., :half], xf[..., half:]
    return torch.cat([c * a - s * b, c * b + s * a], dim=-1).half()


def kv4_fake_quant(x: torch.Tensor):
    """per-(token,head) asymmetric 4-bit quant -> dequantised fp16 (Template.hpp:1063-1081, Utils.h:2196-2211)."""
    xf = x.float()
    mx, mn = xf.amax(-1, keepdim=True), xf.amin(-1, keepdim=True)
    scale = ((mx - mn) / 15.0).half().float()
    zero = (-15.0 * mn / (mx - mn)).half().float()
    q = (torch.clamp(torch.round(xf / scale + zero), 0, 255).to(torch.int32) & 0xF).float()
    return (q * scale - (scale * zero).half().float()).half()


def decode_layer(x, p, k_cache, v_cache, lens, cfg):
    """One decoder layer, decode step.  x fp16 [B,H]; k_cache/v_cache fp16 [B,Hkv,maxctx,Dh] (dequantised KV4),
    lens [B] cached tokens.  Returns fp16 [B,H]; appends the new token to the caches."""
    B = x.shape[0]
    Hq, Hkv, Dh = cfg["hq"], cfg["hkv"], cfg["dh"]
    q8, sa, ss = norm_quant(x, p["ln1"], cfg["eps"])
    qkv = w4a8_linear(q8, sa, ss, *p["qkv"])
    q = qkv[:, :Hq * Dh].view(B, Hq, Dh)
    k = qkv[:, Hq * Dh:(Hq + Hkv) * Dh].view(B, Hkv, Dh)
    v = qkv[:, (Hq + Hkv) * Dh:].view(B, Hkv, Dh)
    q, k = rope_neox(q, lens, cfg["base"]), rope_neox(k, lens, cfg["base"])
    idx = lens.long()
    bi = torch.arange(B)
    k_cache[bi, :, idx] = kv4_fake_quant(k)
    v_cache[bi, :, idx] = kv4_fake_quant(v)
    g = Hq // Hkv
    maxc = int(lens.max()) + 1
    kk = k_cache[:, :, :maxc].float()
    vv = v_cache[:, :, :maxc].float()
    qq = q.float().view(B, Hkv, g, Dh)
    sc = torch.einsum("bhgd,bhtd->bhgt", qq, kk) / math.sqrt(Dh)
    mask = torch.arange(maxc)[None, :] > idx[:, None]
    sc = sc.masked_fill(mask[:, None, None, :], float("-inf"))
    pr = torch.softmax(sc, dim=-1).half().float()
    o = torch.einsum("bhgt,bhtd->bhgd", pr, vv).reshape(B, Hq * Dh).half()
    q8, sa, ss = quant_per_token(o)
    h = (x.float() + w4a8_linear(q8, sa, ss, *p["o"]).float()).half()
    q8, sa, ss = norm_quant(h, p["ln2"], cfg["eps"])
    gu = w4a8_linear(q8, sa, ss, *p["gate_up"])
    d = gu.shape[1] // 2
    gf = gu[:, :d].float()
    act = ((gf / (1.0 + torch.exp(-gf))).half().float() * gu[:, d:].float()).half()
    q8, sa, ss = quant_per_token(act)
    return (h.float() + w4a8_linear(q8, sa, ss, *p["down"]).float()).half()


def random_layer(cfg, gen):
    H, I, Hq, Hkv, Dh = cfg["hidden"], cfg["inter"], cfg["hq"], cfg["hkv"], cfg["dh"]

    def lin(n, k):
        w = torch.randint(0, 16, (n, k), generator=gen, dtype=torch.int8).float()
        s1 = ((torch.rand(n, generator=gen) * 0.5 + 0.75) * (0.02 / 4.6)).half()
        return w, s1, (s1.float() * 8).half()
    return {"qkv": lin((Hq + 2 * Hkv) * Dh, H), "o": lin(H, Hq * Dh), "gate_up": lin(2 * I, H), "down": lin(H, I),
            "ln1": torch.ones(H).half(), "ln2": torch.ones(H).half()}
