#!/usr/bin/env python
"""First-contact diagnostics on a B200 (run under gpurun): exercises every kernel on small structured
inputs and prints what differs from the oracle, so one GPU round trip yields maximum information.
Test infrastructure (imports oracle/)."""
import os
import sys
import time
import traceback

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import act as oact  # noqa: E402
from oracle import kv4 as okv  # noqa: E402
from oracle import w4a8 as ow  # noqa: E402
from omniserve_b200 import _lib as L  # noqa: E402
from omniserve_b200.backend import (activation_ops, fused_attention_pure_dense, fused_kernels, layernorm_ops,  # noqa: E402
                                    qgemm_w4a8_per_chn, qgemm_w4a8_per_group)

dev = "cuda"


def t(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev)


def section(name):
    print(f"\n===== {name} =====", flush=True)


def gemm_case(M, N, K, per_group=False, seed=0, force_bn=0, force_mode=-1, force_ctas=0, onehot=False, verbose=True):
    rng = np.random.default_rng(seed)
    q = rng.integers(0, 16, (N, K), dtype=np.uint8)
    a = rng.integers(-127, 128, (M, K), dtype=np.int8)
    if onehot:
        a[:] = 0
        for m in range(M):
            a[m, (m * 37) % K] = 1
    s1 = rng.uniform(0.005, 0.02, N).astype(np.float16)
    sa = rng.uniform(0.01, 0.05, M).astype(np.float16)
    if onehot:
        s1[:] = 1.0
        sa[:] = 1.0
    qw = ow.pack_w4(q)
    out = torch.full((M, N), float("nan"), dtype=torch.float16, device=dev)
    if per_group:
        ng = K // 128
        s2 = rng.integers(1, 9, (N, ng)).astype(np.int64)
        z = rng.integers(0, 16, (N, ng)).astype(np.int64)
        s2p = ow.pack_s2(s2).astype(np.int8)
        zp = (ow.pack_s2(-z) * ow.pack_s2(s2)).astype(np.int8)
        acc, ref = ow.gemm_per_group(a, qw, zp, s2p, s1, sa)
        ta, tq, tz, ts2, ts1, tsa = t(a), t(qw), t(zp), t(s2p), t(s1), t(sa)
        code = L.lib().ob_w4a8_gemm_ex(1, L.ptr(ta), L.ptr(tq), L.ptr(tz), L.ptr(ts2), L.ptr(ts1),
                                       L.ptr(tsa), 0, 0, L.ptr(out), M, N, K, N, force_bn, force_mode, force_ctas,
                                       L.stream())
    else:
        z = rng.integers(0, 16, N).astype(np.float32)
        szs = (z * s1.astype(np.float32)).astype(np.float16)
        ssum = (a.astype(np.float32) * sa.astype(np.float32)[:, None]).sum(1).astype(np.float16)
        if onehot:
            szs[:] = 0
        acc, ref = ow.gemm_per_chn(a, qw, s1, sa, szs, ssum)
        ta, tq, ts1, tsa, tsz, tss = t(a), t(qw), t(s1), t(sa), t(szs), t(ssum)
        code = L.lib().ob_w4a8_gemm_ex(0, L.ptr(ta), L.ptr(tq), 0, 0, L.ptr(ts1), L.ptr(tsa), L.ptr(tsz), L.ptr(tss),
                                       L.ptr(out), M, N, K, N, force_bn, force_mode, force_ctas, L.stream())
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    bad = ~(got == ref) & ~((np.abs(got.astype(np.float32) - ref.astype(np.float32))
                             <= 1e-3 * np.abs(ref.astype(np.float32)) + 1e-3))
    nbad = int(bad.sum())
    tag = f"M={M} N={N} K={K} pg={int(per_group)} bn={force_bn} mode={force_mode} ctas={force_ctas} onehot={int(onehot)}"
    exact = int((got == ref).sum())
    print(f"[gemm] {tag}: code={code} mismatches={nbad}/{got.size} bit-exact={exact}/{got.size} nan={int(np.isnan(got).sum())}",
          flush=True)
    if nbad and verbose:
        idx = np.argwhere(bad)[:8]
        for (m, n) in idx:
            print(f"    out[{m},{n}] got={got[m, n]} ref={ref[m, n]} acc={acc[m, n]}")
        if onehot:
            print("    onehot view got[0:4, 0:40]:\n", got[:4, :40].astype(np.float32))
            print("    onehot view ref[0:4, 0:40]:\n", ref[:4, :40].astype(np.float32))
    return nbad == 0 and code == 0


def small_ops():
    section("small ops")
    rng = np.random.default_rng(1)
    for (T, H) in [(3, 256), (5, 4096), (2, 14336), (4, 8192)]:
        x = (rng.standard_normal((T, H)) * 1.5).astype(np.float16)
        q, s, sm = oact.quant_fuse_sum(x)
        out = torch.empty((T, H), dtype=torch.int8, device=dev)
        sc = torch.empty(T, dtype=torch.float16, device=dev)
        su = torch.empty(T, dtype=torch.float16, device=dev)
        fused_kernels.invoke_quant_fuse_sum(out, t(x), su, sc)
        torch.cuda.synchronize()
        dq = np.abs(out.cpu().numpy().astype(np.int32) - q.astype(np.int32))
        print(f"[quant_fuse_sum] T={T} H={H}: q maxdiff={dq.max()} n_diff={int((dq > 0).sum())} "
              f"scale_eq={np.array_equal(sc.cpu().numpy(), s)} sum_maxrel="
              f"{np.max(np.abs(su.cpu().numpy().astype(np.float32) - sm.astype(np.float32)) / (np.abs(sm.astype(np.float32)) + 1e-3)):.2e}")
    for (T, H) in [(3, 256), (5, 4096), (4, 8192)]:
        x = (rng.standard_normal((T, H)) * 1.5 + 0.1).astype(np.float16)
        g = (rng.standard_normal(H) * 0.2 + 1.0).astype(np.float16)
        q, s, sm, _ = oact.rms_norm_general_fuse_sum(x, g, 1e-5)
        out = torch.empty((T, H), dtype=torch.int8, device=dev)
        sc = torch.empty(T, dtype=torch.float16, device=dev)
        su = torch.empty(T, dtype=torch.float16, device=dev)
        layernorm_ops.rms_norm_general_fuse_sum(out, t(x), t(g), su, sc, 1e-5, True)
        torch.cuda.synchronize()
        dq = np.abs(out.cpu().numpy().astype(np.int32) - q.astype(np.int32))
        print(f"[rms_norm_general_fuse_sum] T={T} H={H}: q maxdiff={dq.max()} n_diff={int((dq > 0).sum())} "
              f"scale_eq={np.array_equal(sc.cpu().numpy(), s)} sum got={su.cpu().numpy()[:3]} ref={sm[:3]}")
    T, d = 4, 14336
    x = (rng.standard_normal((T, 2 * d)) * 2).astype(np.float16)
    ref = oact.silu_and_mul(x)
    out = torch.empty((T, d), dtype=torch.float16, device=dev)
    activation_ops.silu_and_mul(out, t(x))
    torch.cuda.synchronize()
    diff = np.abs(out.cpu().numpy().astype(np.float32) - ref.astype(np.float32))
    print(f"[silu_and_mul] maxabs={diff.max():.3e} exact={int((out.cpu().numpy() == ref).sum())}/{ref.size}")
    h = t(x[:, :4096].copy())
    o2 = torch.empty_like(h)
    layernorm_ops.rms_norm(o2, h, t(np.ones(4096, np.float16)), 1e-5)
    torch.cuda.synchronize()
    r2 = oact.rms_norm(x[:, :4096], np.ones(4096, np.float16), 1e-5)
    print(f"[rms_norm] maxabs={np.abs(o2.cpu().numpy().astype(np.float32) - r2.astype(np.float32)).max():.3e}")


def attention(B=3, Hq=8, Hkv=2, lens=(70, 200, 129), seed=3, force_split=0):
    rng = np.random.default_rng(seed)
    Dh = 128
    n_pages = sum((l + 63) // 64 for l in lens) + 2
    cache = okv.PagedKV4(n_pages, Hkv, Dh)
    perm = rng.permutation(n_pages)
    max_pages = max((l + 63) // 64 for l in lens)
    bt = np.zeros((B, max_pages), np.int64)
    c = 0
    for b, l in enumerate(lens):
        for j in range((l + 63) // 64):
            bt[b, j] = perm[c]
            c += 1
    okv.fill_random(cache, bt, [l - 1 for l in lens], rng)
    q = rng.standard_normal((B, Hq, Dh)).astype(np.float16)
    k = rng.standard_normal((B, Hkv, Dh)).astype(np.float16)
    v = rng.standard_normal((B, Hkv, Dh)).astype(np.float16)
    kpool, vpool = t(cache.k_pool), t(cache.v_pool)
    ptrs = np.zeros((B, 2, max_pages), np.int64)
    ptrs[:, 0] = kpool.data_ptr() + bt * cache.k_page_bytes
    ptrs[:, 1] = vpool.data_ptr() + bt * cache.v_page_bytes
    qkv = torch.cat([t(q).reshape(B, -1), t(k).reshape(B, -1), t(v).reshape(B, -1)], dim=1).contiguous()
    tq = qkv[:, : Hq * Dh].view(B, Hq, Dh)
    tk = qkv[:, Hq * Dh: (Hq + Hkv) * Dh].view(B, Hkv, Dh)
    tv = qkv[:, (Hq + Hkv) * Dh:].view(B, Hkv, Dh)
    lens_t = t(np.asarray(lens, np.int32))
    from omniserve_b200.backend import _attn_common as AC
    out = AC.single_query(tq, tk, tv, t(ptrs), None, None, None, None, lens_t, 64, Hkv, 0, 0, 0, 0, 0,
                          max(lens) - 1, 128, 500000.0, 1.0, force_split=force_split)
    torch.cuda.synchronize()
    got = out.cpu().numpy().astype(np.float32)
    c1 = okv.PagedKV4(n_pages, Hkv, Dh)
    c1.k_pool[:] = cache.k_pool
    c1.v_pool[:] = cache.v_pool
    ref_m = okv.decode_attention(q, k, v, c1, bt, lens, 128, 500000.0, mimic=True).astype(np.float32)
    c2 = okv.PagedKV4(n_pages, Hkv, Dh)
    c2.k_pool[:] = cache.k_pool
    c2.v_pool[:] = cache.v_pool
    ref_e = okv.decode_attention(q, k, v, c2, bt, lens, 128, 500000.0, mimic=False).astype(np.float32)
    sc = np.abs(ref_e).max()
    print(f"[attention] B={B} Hq={Hq} Hkv={Hkv} lens={lens} split={force_split}: "
          f"max|got-exact|/max|ref|={np.abs(got - ref_e).max() / sc:.3e} "
          f"max|got-mimic|/max={np.abs(got - ref_m).max() / sc:.3e} "
          f"max|mimic-exact|/max={np.abs(ref_m - ref_e).max() / sc:.3e} nan={int(np.isnan(got).sum())}")
    # appended page bytes
    kb = kpool.cpu().numpy()
    vb = vpool.cpu().numpy()
    print(f"    K pool bytes equal to oracle after append: {np.array_equal(kb, c1.k_pool)} "
          f"(diff bytes {int((kb != c1.k_pool).sum())}); V: {np.array_equal(vb, c1.v_pool)} (diff {int((vb != c1.v_pool).sum())})")
    return np.abs(got - ref_e).max() / sc


def main():
    print("device:", torch.cuda.get_device_name(0), "lib version", L.lib().ob_version(), flush=True)
    try:
        small_ops()
    except Exception:
        traceback.print_exc()
    try:
        section("attention")
        attention()
        attention(B=2, Hq=32, Hkv=8, lens=(300, 65), seed=5)
        attention(B=1, Hq=8, Hkv=2, lens=(1000,), seed=6, force_split=3)
        attention(B=2, Hq=4, Hkv=4, lens=(1, 64), seed=7)
    except Exception:
        traceback.print_exc()
    section("gemm")
    ok = True
    try:
        ok &= gemm_case(16, 128, 128, onehot=True)
        ok &= gemm_case(16, 128, 128)
        ok &= gemm_case(16, 128, 512)
        ok &= gemm_case(64, 256, 1024)
        ok &= gemm_case(64, 256, 1024, per_group=True)
        ok &= gemm_case(7, 4096, 4096)
        ok &= gemm_case(64, 6144, 4096, verbose=False)
        ok &= gemm_case(64, 4096, 14336, verbose=False)
        ok &= gemm_case(200, 4096, 4096, verbose=False)
        ok &= gemm_case(1000, 1024, 1024, force_mode=0, verbose=False)
        ok &= gemm_case(1000, 1024, 1024, per_group=True, force_mode=0, verbose=False)
        ok &= gemm_case(64, 1024, 2048, force_mode=1, force_ctas=24, verbose=False)
        ok &= gemm_case(4096, 6144, 4096, verbose=False)
    except Exception:
        traceback.print_exc()
        ok = False
    print("GEMM ALL OK" if ok else "GEMM HAS FAILURES", flush=True)


if __name__ == "__main__":
    main()
