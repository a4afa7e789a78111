"""Weight quantiser / packer: host-side mirror of the reference's checkpoint path, in torch (CPU or CUDA).

  * pseudo_quantize_tensor      /root/reference/scripts/ckpt_converter/quant_utils.py:96-138
  * from_linear (per-channel)   /root/reference/omniserve/modeling/layers/quantized_linear/w4a8_linear.py:284-335
  * from_linear (per-group)     .../w4a8_linear.py:170-282
The byte layout produced here is the reference's ([N/32][K/32][32 lanes][16 B], nibble = rows n / n+16), so
released QServe checkpoints load unchanged and our GEMM consumes them unchanged.
"""
from __future__ import annotations

import torch


def pseudo_quantize_tensor(w: torch.Tensor, n_bit: int = 4, q_group_size: int = -1):
    """Asymmetric fake quantisation; returns (w_fake, scales [N, K/g], zeros [N, K/g])."""
    org_shape = w.shape
    w2 = w.reshape(-1, q_group_size) if q_group_size > 0 else w.reshape(org_shape[0], -1)
    max_val = w2.amax(dim=1, keepdim=True)
    min_val = w2.amin(dim=1, keepdim=True)
    max_int = 2 ** n_bit - 1
    scales = (max_val - min_val).clamp(min=1e-5) / max_int
    zeros = (-torch.round(min_val / scales)).clamp_(0, max_int)
    wq = (torch.clamp(torch.round(w2 / scales) + zeros, 0, max_int) - zeros) * scales
    return wq.reshape(org_shape), scales.view(org_shape[0], -1), zeros.view(org_shape[0], -1)


def pack_w4(q: torch.Tensor) -> torch.Tensor:
    """uint4 values [N, K] (any integer dtype) -> int8 [N, K/2] in the reference tile layout."""
    N, K = q.shape
    assert N % 32 == 0 and K % 32 == 0
    r = q.reshape(N // 32, 2, 2, 8, K // 32, 2, 4, 4).permute(0, 4, 3, 6, 1, 5, 2, 7)
    r = r.permute(0, 1, 2, 3, 5, 6, 7, 4).contiguous().to(torch.int8)
    packed = (r[..., 1] << 4) + r[..., 0]
    return packed.reshape(N // 32, K // 32, 32, 16).reshape(N, K // 2).contiguous()


def unpack_w4(packed: torch.Tensor) -> torch.Tensor:
    N, K2 = packed.shape
    K = 2 * K2
    p = packed.view(torch.uint8).reshape(N // 32, K // 32, 8, 4, 2, 2, 4)
    both = torch.stack([p & 0xF, p >> 4], dim=0)  # [hi, n32, k32, c, e, d, b, f]
    return both.permute(1, 0, 6, 3, 2, 5, 4, 7).reshape(N, K).contiguous()


def pack_s2(x: torch.Tensor) -> torch.Tensor:
    """[N, K/G] -> [K/G, N], N permuted inside each 32-block (pos = c*4+j <-> n = j*8+c)."""
    N, ng = x.shape
    return x.t().reshape(ng, N // 32, 4, 8).transpose(-2, -1).reshape(ng, N).contiguous()


def quantize_per_channel(weight: torch.Tensor, s1_scale: torch.Tensor, zeros: torch.Tensor):
    """-> dict(qweight, s1_scales, s1_szeros) like from_linear(group_size=-1)."""
    N = weight.shape[0]
    q = torch.round(weight.float() / s1_scale.reshape(N, 1).float()).to(torch.int8) + zeros.reshape(N, 1).to(torch.int8)
    if q.min() < 0 or q.max() > 15:
        raise ValueError("Quantized weight out of range")
    return {
        "qweight": pack_w4(q),
        "s1_scales": s1_scale.reshape(N).half(),
        "s1_szeros": (zeros.reshape(N).float() * s1_scale.reshape(N).float()).half(),
    }


def quantize_per_group(weight, s1_scale, s2_scale, zeros, group_size: int = 128):
    """-> dict(qweight, s1_scales, s2_scales, s2_zeros) like from_linear(group_size=128)."""
    N, K = weight.shape
    ng = K // group_size
    lw = torch.round(weight.float() / s1_scale.reshape(N, 1).float())
    if lw.min() < -128 or lw.max() > 127:
        raise ValueError("Stage 1: Quantized weight out of range")
    lw = lw.reshape(N, ng, group_size)
    q = lw / s2_scale.reshape(N, ng, 1).half().float() + zeros.reshape(N, ng, 1).half().float()
    if q.min() < 0 or q.max() > 15:
        raise ValueError("Stage 2: Quantized weight out of range")
    q = q.reshape(N, K).to(torch.int8)
    s2p = pack_s2(s2_scale.reshape(N, ng).to(torch.int64))
    zp = pack_s2((-zeros.reshape(N, ng)).to(torch.int32).to(torch.int64))
    return {
        "qweight": pack_w4(q),
        "s1_scales": s1_scale.reshape(N).half(),
        "s2_scales": s2p.to(torch.int8),
        "s2_zeros": (zp * s2p).to(torch.int8),
    }
