"""Tensor-parallel sharding of packed W4A8 parameters (the reference has no TP: SURVEY.md F1, section 5).

Column-parallel (qkv_proj, gate_up_proj): shard output rows N -- contiguous row ranges that are multiples of 32.
Row-parallel (o_proj, down_proj): shard the reduction dim K on the K/32 *tile axis* of the
[N/32][K/32][32][16] packing (slicing `[:, start:end]` of the [N, K/2] view, as
omniserve/utils/weight_utils.py:207-212 would, is wrong for this layout).  s1_scales / s1_szeros are
replicated for row-parallel layers: each rank dequantises its own partial product
(acc_r*s1*sa_r - sz*sum_r) and the all-reduce sums them, which is exact because the per-token scale and
sum are computed on the rank's own activation slice.
"""
from __future__ import annotations

from typing import Dict, Sequence

import torch


def _rows(t: torch.Tensor, ranges: Sequence[range]) -> torch.Tensor:
    return torch.cat([t[r.start:r.stop] for r in ranges], dim=0).contiguous()


def shard_column(p: Dict[str, torch.Tensor], ranges: Sequence[range]) -> Dict[str, torch.Tensor]:
    """Keep the output rows in `ranges` (each a multiple-of-32 aligned range), in order."""
    for r in ranges:
        assert r.start % 32 == 0 and r.stop % 32 == 0
    out = {"qweight": _rows(p["qweight"], ranges), "s1_scales": _rows(p["s1_scales"], ranges)}
    if "s1_szeros" in p:
        out["s1_szeros"] = _rows(p["s1_szeros"], ranges)
    for k in ("s2_scales", "s2_zeros"):
        if k in p:
            out[k] = torch.cat([p[k][:, r.start:r.stop] for r in ranges], dim=1).contiguous()
    return out


def shard_row(p: Dict[str, torch.Tensor], k_range: range, group_size: int = 128) -> Dict[str, torch.Tensor]:
    """Keep reduction indices k_range (aligned to 128) of a row-parallel layer."""
    assert k_range.start % 128 == 0 and k_range.stop % 128 == 0
    N, K2 = p["qweight"].shape
    K = 2 * K2
    tiles = p["qweight"].reshape(N // 32, K // 32, 512)
    q = tiles[:, k_range.start // 32:k_range.stop // 32].reshape(N, (k_range.stop - k_range.start) // 2).contiguous()
    out = {"qweight": q, "s1_scales": p["s1_scales"].clone()}
    if "s1_szeros" in p:
        out["s1_szeros"] = p["s1_szeros"].clone()
    for k in ("s2_scales", "s2_zeros"):
        if k in p:
            out[k] = p[k][k_range.start // group_size:k_range.stop // group_size].contiguous()
    return out


def qkv_ranges(num_heads: int, num_kv_heads: int, head_dim: int, rank: int, size: int):
    """Row ranges of a fused [q | k | v] projection owned by `rank` (heads sharded contiguously)."""
    hq, hkv = num_heads // size, num_kv_heads // size
    q0 = rank * hq * head_dim
    k0 = num_heads * head_dim + rank * hkv * head_dim
    v0 = (num_heads + num_kv_heads) * head_dim + rank * hkv * head_dim
    return [range(q0, q0 + hq * head_dim), range(k0, k0 + hkv * head_dim), range(v0, v0 + hkv * head_dim)]


def gate_up_ranges(intermediate: int, rank: int, size: int):
    part = intermediate // size
    return [range(rank * part, (rank + 1) * part), range(intermediate + rank * part, intermediate + (rank + 1) * part)]
