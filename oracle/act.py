"""CPU restatement of the small fused ops around the W4A8 GEMMs.

TEST INFRASTRUCTURE -- see oracle/__init__.py.  numpy only.

Reference files restated:
  * per-token INT8 activation quant (+sum): kernels/csrc/fused_kernels.cu:57-142
  * "RMSNorm"+quant(+sum) (TRT-LLM generalLayerNorm): kernels/csrc/layernorm_kernels.cu:26-34,194-331
  * plain rms_norm: kernels/csrc/layernorm_kernels.cu:335-364
  * silu_and_mul: kernels/csrc/activation_kernels.cu:10-30
  * int8 rounding everywhere = ``cvt.rni.sat.s8.f32`` (kernels/csrc/utils.cuh:79-84)

The reference is compiled with ``--use_fast_math`` (kernels/setup.py:33), so ``127.f/amax``,
``rsqrtf`` and ``expf`` are approximate on the GPU.  This oracle evaluates them exactly in fp32;
GPU-vs-oracle tests therefore allow |dq| <= 1 LSB on a tiny fraction of int8 elements, while the
GPU-vs-rebuilt-reference tests (tests/test_gpu_vs_ref.py) compare the same intrinsics.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32
f16 = np.float16


def rni_sat_s8(x: np.ndarray) -> np.ndarray:
    """cvt.rni.sat.s8.f32: round-half-even, saturate to [-128, 127] (NaN -> 0)."""
    x = np.nan_to_num(np.asarray(x, f32), nan=0.0, posinf=127.0, neginf=-128.0)
    return np.clip(np.rint(x), -128, 127).astype(np.int8)


def quant_fuse_sum(x_f16: np.ndarray):
    """invoke_quant_fuse_sum (fused_kernels.cu:97-142).  x fp16 [T, H].

    Returns (q int8 [T,H], scale fp16 [T], sum fp16 [T]).  ``scale = f16(amax/127)``;
    ``q = rni(x * (127/amax))`` with the UNROUNDED fp32 ``127/amax`` (:137-140);
    ``sum`` = fp32 sum of the fp activations rounded to fp16 (:115,:127).
    """
    x = np.asarray(x_f16, f16).astype(f32)
    amax = np.abs(x).max(axis=1)
    scale = (amax / f32(127.0)).astype(f16)
    with np.errstate(divide="ignore", invalid="ignore"):
        tmp = f32(127.0) / amax
        q = rni_sat_s8(x * tmp[:, None])
    s = x.astype(np.float64).sum(axis=1).astype(f32).astype(f16)
    return q, scale, s


def quant(x_f16: np.ndarray):
    """invoke_quant, per-token branch (fused_kernels.cu:57-85)."""
    q, scale, _ = quant_fuse_sum(x_f16)
    return q, scale


def _f16_seq_sum(cols: np.ndarray) -> np.ndarray:
    """Sequential fp16 accumulation along axis 0 (``T_scalar sum; sum += half`` -> __hadd)."""
    acc = np.zeros(cols.shape[1:], dtype=f16)
    for r in cols:
        acc = (acc.astype(f32) + r.astype(f32)).astype(f16)
    return acc


def rms_norm_general_fuse_sum(x_f16, gamma_f16, eps: float, block: int | None = None):
    """rms_norm_general_fuse_sum, per-token branch (layernorm_kernels.cu:194-331, 471-513).

    Quirks reproduced:
      (i)   output = (x - mean(x)) * rsqrt(mean(x^2) + eps) * gamma      (:28 with :263 uncentred)
      (ii)  the normalised value is rounded to fp16 before amax / sum      (:286-291)
      (iii) each CUDA thread accumulates its strided elements' sum in a *half* register (:280,:291),
            threads = min(H,1024) rounded up to 32 (:479-480); block-level reduce is fp32
      (iv)  amax floor 1e-6 (as half: 1e-6 -> 1.013e-6)                   (:279)
      (v)   int8 = rni(val_f32 * (127/amax)) recomputed from fp32          (:310-321)
    Returns (q int8 [T,H], scale fp16 [T], sum fp16 [T], normed fp16 [T,H]).
    """
    x = np.asarray(x_f16, f16).astype(f32)
    g = np.asarray(gamma_f16, f16).astype(f32)
    T, H = x.shape
    if block is None:
        block = min(H, 1024)
        block = 32 * ((block + 31) // 32)
    mean = (x.astype(np.float64).sum(axis=1) / H).astype(f32)
    var = (x.astype(np.float64) ** 2).sum(axis=1).astype(f32)
    s_var = (1.0 / np.sqrt((var / f32(H) + f32(eps)).astype(np.float64))).astype(f32)
    val_f = ((x - mean[:, None]) * s_var[:, None]) * g[None, :]  # compute_layernorm (:28)
    val_h = val_f.astype(f16)
    amax = np.maximum(np.abs(val_h).max(axis=1), f16(1e-6)).astype(f32)
    # (iii): thread t owns elements t, t+block, ...; half accumulation per thread, fp32 across threads
    n_iter = (H + block - 1) // block
    pad = n_iter * block - H
    vh = np.pad(val_h, ((0, 0), (0, pad))) if pad else val_h
    per_thread = _f16_seq_sum(vh.reshape(T, n_iter, block).transpose(1, 0, 2))  # [T, block] fp16
    sum_f = per_thread.astype(np.float64).sum(axis=1).astype(f32)
    scale = (amax / f32(127.0)).astype(f16)
    dyn = f32(127.0) / amax
    q = rni_sat_s8(val_f * dyn[:, None])
    return q, scale, sum_f.astype(f16), val_h


def rms_norm_general(x_f16, gamma_f16, eps: float):
    """rms_norm_general per-token (no sum): same kernel minus input_sum (layernorm_kernels.cu:60-190)."""
    q, scale, _, _ = rms_norm_general_fuse_sum(x_f16, gamma_f16, eps)
    return q, scale


def rms_norm(x_f16, gamma_f16, eps: float) -> np.ndarray:
    """vllm rms_norm_kernel, fp16 out (layernorm_kernels.cu:335-364): ``((half)(x*rsqrt)) * w`` in half."""
    x = np.asarray(x_f16, f16).astype(f32)
    H = x.shape[1]
    var = (x.astype(np.float64) ** 2).sum(axis=1).astype(f32)
    s = (1.0 / np.sqrt((var / f32(H) + f32(eps)).astype(np.float64))).astype(f32)
    t = (x * s[:, None]).astype(f16)
    return (t.astype(f32) * np.asarray(gamma_f16, f16).astype(f32)[None, :]).astype(f16)


def silu_and_mul(x_f16: np.ndarray) -> np.ndarray:
    """silu_and_mul (activation_kernels.cu:10-30): out = half(silu_fp32(g)) *half u.  x [T, 2d]."""
    x = np.asarray(x_f16, f16)
    d = x.shape[1] // 2
    g = x[:, :d].astype(f32)
    u = x[:, d:]
    with np.errstate(over="ignore"):
        s = (g / (f32(1.0) + np.exp(-g, dtype=f32))).astype(f16)
    return (s.astype(f32) * u.astype(f32)).astype(f16)  # __hmul: exact product rounded once
