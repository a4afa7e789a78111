"""Host-side driver of the W4A8KV4 hot path: a Llama decoder stack sequenced exactly like the reference
model code, on top of the `omniserve_backend`-compatible ops, plus what the reference lacks: tensor
parallelism (column/row sharding + one all-reduce after o_proj and after down_proj) and CUDA-graph replay
of the decode step.

Mirrors (call order, buffer aliasing rules, shapes):
  * LlamaDecoderLayer.forward   /root/reference/omniserve/modeling/models/llama_w4a8_unpad.py:406-438
  * LlamaAttention.forward      .../llama_w4a8_unpad.py:265-361
  * LlamaMLP.forward            .../llama_w4a8_unpad.py:83-112
  * W4A8OF16LinearDynamicInputScale  omniserve/modeling/layers/quantized_linear/w4a8_linear.py:16-139
  * ActivationBuffer            omniserve/utils/input_metadata.py:60-104
  * CacheEngine page sizing     omniserve/worker/cache_engine.py:73-88
PyTorch is used for device memory, streams, embedding / lm_head / argmax (as in the reference) and
torch.distributed; every quantised op goes through omniserve_b200.backend (the C ABI).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional

import torch

from . import backend as _backend_pkg  # noqa: F401  (the ctypes mirror of omniserve_backend)


class Ops:
    """The seven `omniserve_backend` modules the W4A8KV4 model code imports (llama_w4a8_unpad.py:32-37,
    w4a8_linear.py:12-13, layernorm.py:16, activation.py:16, decoding_attention.py:5-10, ctx_update_kv.py:3-5).
    Default: ours.  bench.py --impl reference passes the reference's own rebuilt modules instead."""

    def __init__(self, loader=None):
        import importlib
        names = ("activation_ops", "fused_attention_fine_grained_dense", "fused_attention_pure_dense", "fused_kernels",
                 "layernorm_ops", "qgemm_w4a8_per_chn", "qgemm_w4a8_per_group")
        for n in names:
            setattr(self, n, loader(n) if loader else importlib.import_module(f"omniserve_b200.backend.{n}"))

TOKENS_PER_BLOCK = 64


@dataclass
class LlamaConfig:
    hidden_size: int = 4096
    intermediate_size: int = 14336
    num_hidden_layers: int = 32
    num_attention_heads: int = 32
    num_key_value_heads: int = 8
    head_dim: int = 128
    vocab_size: int = 128256
    rope_theta: float = 500000.0
    rms_norm_eps: float = 1e-5
    group_size: int = -1  # -1 = per-channel (QServe default for the A100 numbers), 128 = per-group

    @staticmethod
    def llama3_8b(**kw):
        return LlamaConfig(**kw)

    @staticmethod
    def llama3_70b(**kw):
        return LlamaConfig(hidden_size=8192, intermediate_size=28672, num_hidden_layers=80,
                           num_attention_heads=64, num_key_value_heads=8, **kw)

    @staticmethod
    def tiny(**kw):
        d = dict(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4,
                 num_key_value_heads=2, vocab_size=1024)
        d.update(kw)
        return LlamaConfig(**d)


class W4A8Linear:
    """Buffers and forward of W4A8OF16LinearDynamicInputScale (w4a8_linear.py:16-139)."""

    def __init__(self, in_features: int, out_features: int, group_size: int, device, ops: "Ops" = None):
        assert in_features % 128 == 0 and out_features % 32 == 0
        self.ops = ops
        self.in_features, self.out_features, self.group_size = in_features, out_features, group_size
        self.per_channel = group_size == -1
        self.qweight = torch.zeros((out_features, in_features // 2), dtype=torch.int8, device=device)
        self.s1_scales = torch.zeros((out_features,), dtype=torch.float16, device=device)
        if self.per_channel:
            self.s1_szeros = torch.zeros((out_features,), dtype=torch.float16, device=device)
        else:
            ng = in_features // group_size
            self.s2_scales = torch.zeros((ng, out_features), dtype=torch.int8, device=device)
            self.s2_zeros = torch.zeros((ng, out_features), dtype=torch.int8, device=device)

    def random_init_(self, gen: torch.Generator, std: float = 0.02):
        """Synthetic non-zero weights (the reference benchmark runs all-zero weights, which
        makes KV scales 0/15 -> NaN).  Any byte pattern is a valid packed-u4 weight."""
        dev = self.qweight.device
        n = self.qweight.numel()
        self.qweight.view(torch.uint8).copy_(
            torch.randint(0, 256, (n,), generator=gen, dtype=torch.uint8).view(self.qweight.shape).to(dev))
        if self.per_channel:
            s1 = (torch.rand(self.out_features, generator=gen) * 0.5 + 0.75) * (std / 4.6)
            self.s1_scales.copy_(s1.half().to(dev))
            self.s1_szeros.copy_((s1.half().float() * 8.0).half().to(dev))
        else:
            ng = self.in_features // self.group_size
            s1 = (torch.rand(self.out_features, generator=gen) * 0.5 + 0.75) * (std / 4.6 / 4.0)
            self.s1_scales.copy_(s1.half().to(dev))
            s2 = torch.randint(2, 8, (ng, self.out_features), generator=gen)
            z = torch.randint(6, 10, (ng, self.out_features), generator=gen)
            self.s2_scales.copy_(s2.to(torch.int8).to(dev))
            self.s2_zeros.copy_((-z * s2).to(torch.int8).to(dev))
        return self

    def fused_add_norm_quant(self, x_i8, input_scales, input_sum, output_buffer, hidden_in, hidden_out, norm_weight, norm_out,
                             norm_sum, norm_scale, eps) -> bool:
        """GEMM + residual add + norm + quant in one launch (our ops only).  False -> not fusable, nothing launched."""
        if self.per_channel:
            f = getattr(self.ops.qgemm_w4a8_per_chn, "gemm_forward_cuda_add_norm_quant", None)
            return bool(f) and f(x_i8, self.qweight, self.s1_scales, input_scales, self.s1_szeros, input_sum, output_buffer,
                                 hidden_in, hidden_out, norm_weight, norm_out, norm_sum, norm_scale, eps)
        f = getattr(self.ops.qgemm_w4a8_per_group, "gemm_forward_cuda_add_norm_quant", None)
        return bool(f) and f(x_i8, self.qweight, self.s2_zeros, self.s2_scales, self.s1_scales, input_scales, output_buffer,
                             hidden_in, hidden_out, norm_weight, norm_out, norm_scale, eps)

    def __call__(self, x_i8, input_scales, input_sum, output_buffer):
        if self.per_channel:  # forward_per_chn (w4a8_linear.py:109-123)
            self.ops.qgemm_w4a8_per_chn.gemm_forward_cuda(x_i8, self.qweight, self.s1_scales, input_scales, self.s1_szeros,
                                                 input_sum, output_buffer)
        else:  # forward_per_group (:125-139)
            self.ops.qgemm_w4a8_per_group.gemm_forward_cuda(x_i8, self.qweight, self.s2_zeros, self.s2_scales, self.s1_scales,
                                                   input_scales, output_buffer)


class ActivationBuffer:
    """Pre-allocated activation arena (omniserve/utils/input_metadata.py:60-104), sized for T tokens."""

    def __init__(self, T: int, hidden: int, inter_local: int, qkv_local: int, q_local: int, device):
        self.T = T
        f16, i8 = torch.float16, torch.int8
        self.quantized_hidden_states_buffer = torch.empty((T, hidden), dtype=i8, device=device)
        self.quantized_attn_buffer = torch.empty((T, q_local), dtype=i8, device=device)
        self.quantized_mlp_act_buffer = torch.empty((T, inter_local), dtype=i8, device=device)
        self.quantized_scale_buffer = torch.empty((T,), dtype=f16, device=device)
        self.quantized_sum_buffer = torch.empty((T,), dtype=f16, device=device)
        self.qkv_proj_act_buffer = torch.empty((T, qkv_local), dtype=f16, device=device)
        self.gate_up_proj_act_buffer = torch.empty((T, 2 * inter_local), dtype=f16, device=device)
        self.out_down_proj_act_buffer = torch.empty((T, hidden), dtype=f16, device=device)
        self.hidden_a = torch.empty((T, hidden), dtype=f16, device=device)
        self.hidden_b = torch.empty((T, hidden), dtype=f16, device=device)


class PagedKVCache:
    """Per-layer K and V page pools + int64 pointer tables (cache_engine.py:73-136, block_table_utils.py:64-121)."""

    def __init__(self, cfg: LlamaConfig, n_kv_local: int, batch: int, max_ctx: int, device):
        self.pages_per_seq = (max_ctx + TOKENS_PER_BLOCK - 1) // TOKENS_PER_BLOCK
        self.num_pages = batch * self.pages_per_seq
        self.page_bytes = n_kv_local * TOKENS_PER_BLOCK * cfg.head_dim // 2 + TOKENS_PER_BLOCK * n_kv_local * 4
        L = cfg.num_hidden_layers
        self.k_pools = [torch.zeros((self.num_pages, self.page_bytes), dtype=torch.int8, device=device) for _ in range(L)]
        self.v_pools = [torch.zeros((self.num_pages, self.page_bytes), dtype=torch.int8, device=device) for _ in range(L)]
        # deterministic shuffled page assignment
        g = torch.Generator().manual_seed(1234)
        perm = torch.randperm(self.num_pages, generator=g).view(batch, self.pages_per_seq)
        self.block_ids = perm
        self.tables: List[torch.Tensor] = []
        for l in range(L):
            tab = torch.empty((batch, 2, self.pages_per_seq), dtype=torch.int64)
            tab[:, 0] = self.k_pools[l].data_ptr() + perm * self.page_bytes
            tab[:, 1] = self.v_pools[l].data_ptr() + perm * self.page_bytes
            self.tables.append(tab.to(device))

    def bytes(self) -> int:
        return 2 * len(self.k_pools) * self.num_pages * self.page_bytes


class LlamaW4A8:
    """Llama decoder stack over the W4A8KV4 ops.  tp_rank / tp_size shard heads and MLP columns."""

    def __init__(self, cfg: LlamaConfig, device="cuda", tp_rank: int = 0, tp_size: int = 1, seed: int = 0,
                 fuse_silu_quant: bool = True, process_group=None, ops: Ops = None):
        self.ops = ops or Ops()
        self.cfg, self.device, self.tp_rank, self.tp_size, self.pg = cfg, device, tp_rank, tp_size, process_group
        assert cfg.num_attention_heads % tp_size == 0 and cfg.num_key_value_heads % tp_size == 0
        assert cfg.intermediate_size % (tp_size * 128) == 0
        self.hq = cfg.num_attention_heads // tp_size
        self.hkv = cfg.num_key_value_heads // tp_size
        self.q_size = self.hq * cfg.head_dim
        self.kv_size = self.hkv * cfg.head_dim
        self.inter = cfg.intermediate_size // tp_size
        self.fuse_silu_quant = fuse_silu_quant
        self.fuse_add_norm = hasattr(self.ops.layernorm_ops, "add_rms_norm_general")
        self.fuse_attn_quant = hasattr(self.ops.fused_attention_pure_dense, "single_query_attention_quant")
        # decode, tp_size == 1: `residual + o_proj/down_proj -> layernorm -> int8` as the tail of the GEMM launch.  Opt-in
        # (OB_FUSE_GEMM_NORM=1): bit-identical, but measured SLOWER than the PDL-chained separate norm kernel on B200
        # (4.07 vs 3.99 ms per decode step, same box back to back) -- the grid-wide barrier costs more than the launch.
        import os as _os
        self.fuse_gemm_norm = (hasattr(self.ops.qgemm_w4a8_per_chn, "gemm_forward_cuda_add_norm_quant") and tp_size == 1
                               and _os.environ.get("OB_FUSE_GEMM_NORM", "0") == "1")
        self.act_sum = cfg.group_size == -1
        gen = torch.Generator().manual_seed(seed * 1000 + tp_rank)
        gen_rep = torch.Generator().manual_seed(seed * 1000 + 999)  # replicated parameters: same on every rank
        H, gs = cfg.hidden_size, cfg.group_size
        self.layers = []
        for _ in range(cfg.num_hidden_layers):
            ly = {
                "qkv_proj": W4A8Linear(H, self.q_size + 2 * self.kv_size, gs, device, self.ops).random_init_(gen),
                "o_proj": W4A8Linear(self.q_size, H, gs, device, self.ops).random_init_(gen),
                "gate_up_proj": W4A8Linear(H, 2 * self.inter, gs, device, self.ops).random_init_(gen),
                "down_proj": W4A8Linear(self.inter, H, gs, device, self.ops).random_init_(gen),
                "input_layernorm": (1.0 + 0.05 * torch.randn(H, generator=gen_rep)).half().to(device),
                "post_attention_layernorm": (1.0 + 0.05 * torch.randn(H, generator=gen_rep)).half().to(device),
            }
            self.layers.append(ly)
        self.norm_weight = (1.0 + 0.05 * torch.randn(H, generator=gen_rep)).half().to(device)
        self.embed_tokens = (torch.randn(cfg.vocab_size, H, generator=gen_rep) * 0.5).half().to(device)
        # vocab-parallel lm_head (each rank owns vocab_size / tp rows); fp16 GEMM via torch like the reference
        assert cfg.vocab_size % tp_size == 0
        self.vocab_local = cfg.vocab_size // tp_size
        gen_v = torch.Generator().manual_seed(seed * 1000 + 500 + tp_rank)
        self.lm_head = (torch.randn(self.vocab_local, H, generator=gen_v) * 0.02).half().to(device)
        self.kv: Optional[PagedKVCache] = None
        self.buf: Optional[ActivationBuffer] = None
        self.peer = False
        # all heads are retrieval heads on the QServe dense path (ctx_attn_init.py:11-85)
        self._flags = torch.ones(self.hkv, dtype=torch.int32, device=device)
        self._rank = torch.arange(self.hkv, dtype=torch.int32, device=device)

    # ------------------------------------------------------------------ tensor-parallel sharding of a full model
    def load_shard_of(self, full: "LlamaW4A8"):
        """Overwrite this rank's parameters with its shard of `full` (a tp_size == 1 model on the same device):
        qkv_proj / gate_up_proj column-parallel, o_proj / down_proj row-parallel on the K/32 tile axis (tp.py)."""
        from . import tp
        cfg, r, n = self.cfg, self.tp_rank, self.tp_size

        def params(lin):
            d = {"qweight": lin.qweight, "s1_scales": lin.s1_scales}
            for k in ("s1_szeros", "s2_scales", "s2_zeros"):
                if hasattr(lin, k):
                    d[k] = getattr(lin, k)
            return d

        def assign(lin, d):
            for k, v in d.items():
                getattr(lin, k).copy_(v)

        qkv_r = tp.qkv_ranges(cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim, r, n)
        gu_r = tp.gate_up_ranges(cfg.intermediate_size, r, n)
        qs, it = full.q_size // n, cfg.intermediate_size // n
        for mine, src in zip(self.layers, full.layers):
            assign(mine["qkv_proj"], tp.shard_column(params(src["qkv_proj"]), qkv_r))
            assign(mine["gate_up_proj"], tp.shard_column(params(src["gate_up_proj"]), gu_r))
            assign(mine["o_proj"], tp.shard_row(params(src["o_proj"]), range(r * qs, (r + 1) * qs)))
            assign(mine["down_proj"], tp.shard_row(params(src["down_proj"]), range(r * it, (r + 1) * it)))
            mine["input_layernorm"].copy_(src["input_layernorm"])
            mine["post_attention_layernorm"].copy_(src["post_attention_layernorm"])
        self.norm_weight.copy_(full.norm_weight)
        self.embed_tokens.copy_(full.embed_tokens)
        self.lm_head.copy_(full.lm_head[r * self.vocab_local:(r + 1) * self.vocab_local])

    # ------------------------------------------------------------------ memory
    def weight_bytes(self) -> int:
        n = 0
        for ly in self.layers:
            for k in ("qkv_proj", "o_proj", "gate_up_proj", "down_proj"):
                n += ly[k].qweight.numel()
        return n

    def alloc(self, batch: int, max_ctx: int, max_tokens: int):
        self.kv = PagedKVCache(self.cfg, self.hkv, batch, max_ctx, self.device)
        self.buf = ActivationBuffer(max_tokens, self.cfg.hidden_size, self.inter, self.q_size + 2 * self.kv_size,
                                    self.q_size, self.device)
        self.batch, self.max_ctx = batch, max_ctx
        self.context_lens = torch.zeros((batch,), dtype=torch.int32, device=self.device)

    def _all_reduce(self, t):
        if self.tp_size > 1:
            torch.distributed.all_reduce(t, group=self.pg)

    def enable_peer_allreduce(self):
        """Decode path, tp_size > 1: replace `NCCL all-reduce + add + norm + quant` after o_proj / down_proj by ONE kernel
        that sums the ranks' partial results straight out of NVLink peer memory (csrc/small_ops.cu: PeerCtx).  Two
        symmetric buffers alternate (o_proj -> A, down_proj -> B): a rank can only overwrite A after it has passed the
        entry barrier of the next fused all-reduce on B, i.e. after every rank has finished reading A."""
        from .peer import PeerGroup
        assert self.tp_size > 1 and self.buf is not None
        self.peer_group = PeerGroup(self.pg, self.device)
        self.peer_a = self.peer_group.buffer(self.batch, self.cfg.hidden_size)
        self.peer_b = self.peer_group.buffer(self.batch, self.cfg.hidden_size)
        self.peer = True

    # ------------------------------------------------------------------ one decoder layer
    def _norm_quant(self, out_i8, hidden, delta, hidden_out, weight, sm, sc):
        """rms_norm_general(_fuse_sum) of (hidden [+ delta]).  With our ops the residual add of
        llama_w4a8_unpad.py:425,437 is fused into the norm (bit-identical to torch.add + norm); with the
        reference's ops it is a separate torch.add like in the reference.  Returns the tensor holding hidden+delta."""
        lo, eps = self.ops.layernorm_ops, self.cfg.rms_norm_eps
        if delta is not None and not torch.is_tensor(delta):   # PeerBuffer: all-reduce fused into the norm
            lo.peer_add_rms_norm_general(out_i8, hidden, delta, hidden_out, weight, sm if self.act_sum else None, sc, eps)
            return hidden_out
        if delta is not None:
            if self.fuse_add_norm:
                lo.add_rms_norm_general(out_i8, hidden, delta, hidden_out, weight, sm if self.act_sum else None, sc, eps)
                return hidden_out
            torch.add(hidden, delta, out=hidden_out)
            hidden = hidden_out
        if self.act_sum:
            lo.rms_norm_general_fuse_sum(out_i8, hidden, weight, sm, sc, eps, True)
        else:
            lo.rms_norm_general(out_i8, hidden, weight, sc, eps, True)
        return hidden

    def _layer(self, li: int, hidden, delta, T: int, is_prompt: bool, meta):
        """hidden (+ delta, the previous layer's not-yet-added MLP output) -> returns (hidden', delta')."""
        cfg, ly, b = self.cfg, self.layers[li], self.buf
        fused_kernels, activation_ops = self.ops.fused_kernels, self.ops.activation_ops
        fused_attention_fine_grained_dense = self.ops.fused_attention_fine_grained_dense
        fused_attention_pure_dense = self.ops.fused_attention_pure_dense
        qh = b.quantized_hidden_states_buffer[:T]
        sc, sm = b.quantized_scale_buffer[:T], b.quantized_sum_buffer[:T]
        qkv = b.qkv_proj_act_buffer[:T]
        od = b.out_down_proj_act_buffer[:T]
        ha, hb = b.hidden_a[:T], b.hidden_b[:T]
        # 1. (residual add of the previous MLP +) input_layernorm -> int8 (+sum)  (llama:416-421, layernorm.py:86-101)
        if isinstance(delta, str):   # "done": the previous layer's down_proj launch already ran this norm as its tail
            h1 = hidden
        else:
            h1 = self._norm_quant(qh, hidden, delta, ha, ly["input_layernorm"], sm, sc)
        fuse_tail = self.fuse_gemm_norm and not is_prompt
        eps = cfg.rms_norm_eps
        # 2. qkv_proj
        ly["qkv_proj"](qh, sc, sm, qkv)
        q3 = qkv[:, : self.q_size].view(T, self.hq, cfg.head_dim)
        k3 = qkv[:, self.q_size: self.q_size + self.kv_size].view(T, self.hkv, cfg.head_dim)
        v3 = qkv[:, self.q_size + self.kv_size:].view(T, self.hkv, cfg.head_dim)
        attn_quant_done = False
        if is_prompt:
            # 3a. RoPE in place + KV4 page write, then fp16 flash attention (third party, llama:309-325)
            fused_attention_fine_grained_dense.apply_bias_rope_update_kv_cache(
                qkv, meta["seq_lens"], None, meta["padding_offset"], self.kv.tables[li], None, meta["flags"],
                meta["rank"], self.hq, self.hkv, meta["max_seq_len"], TOKENS_PER_BLOCK, self.kv_size // 2, 0, 0, 0, 0,
                0, self.hkv, 0, cfg.head_dim, cfg.rope_theta, 1.0, 8192, True, True, True)
            attn = meta["prefill_attn"](q3, k3, v3).reshape(T, self.q_size)
        else:
            # 3b. KV4 decode attention (decoding_attention.py:146-182); with our ops the quant of step 4 is fused in
            if self.fuse_attn_quant:
                attn_quant_done = True
                attn = fused_attention_pure_dense.single_query_attention_quant(
                    q3, k3, v3, self.kv.tables[li], meta["context_lens"], None, self.max_ctx, TOKENS_PER_BLOCK,
                    self.kv_size // 2, meta["timestep"], cfg.head_dim, cfg.rope_theta, True, True, True,
                    b.quantized_attn_buffer[:T], b.quantized_sum_buffer[:T] if self.act_sum else None,
                    b.quantized_scale_buffer[:T], history_is_stable=True).reshape(T, self.q_size)
            else:
                attn = fused_attention_pure_dense.single_query_attention(
                    q3, k3, v3, self.kv.tables[li], meta["context_lens"], None, self.max_ctx, TOKENS_PER_BLOCK,
                    self.kv_size // 2, meta["timestep"], cfg.head_dim, cfg.rope_theta, True, True, True).reshape(T, self.q_size)
        # 4. quant(+sum) of the attention output (llama:257-263,354)
        qa = b.quantized_attn_buffer[:T]
        if attn_quant_done:
            pass
        elif self.act_sum:
            fused_kernels.invoke_quant_fuse_sum(qa, attn, sm, sc)
        else:
            fused_kernels.invoke_quant(qa, attn, sc)
        # 5. o_proj (row-parallel) -> all-reduce (NCCL, or fused into the norm below over peer memory)
        use_peer = self.peer and not is_prompt
        if fuse_tail and ly["o_proj"].fused_add_norm_quant(qa, sc, sm, od, h1, hb, ly["post_attention_layernorm"], qh,
                                                            sm if self.act_sum else None, sc, eps):
            h2 = hb   # 5-7 in one launch: o_proj, residual add, post_attention_layernorm, int8 quant
        else:
            if use_peer:
                ly["o_proj"](qa, sc, sm, self.peer_a.tensor[:T])
                od_attn = self.peer_a
            else:
                ly["o_proj"](qa, sc, sm, od)
                self._all_reduce(od)
                od_attn = od
            # 6-7. residual add + post_attention_layernorm
            h2 = self._norm_quant(qh, h1, od_attn, hb, ly["post_attention_layernorm"], sm, sc)
        # 8-10. MLP (llama:83-112): gate_up -> silu*mul -> quant -> down (row-parallel) -> all-reduce
        gu = b.gate_up_proj_act_buffer[:T]
        ly["gate_up_proj"](qh, sc, sm, gu)
        qm = b.quantized_mlp_act_buffer[:T]
        if self.fuse_silu_quant:
            activation_ops.silu_and_mul_quant(qm, gu, sm if self.act_sum else None, sc)
        else:
            tmp = meta["silu_tmp"][:T]
            activation_ops.silu_and_mul(tmp, gu)
            if self.act_sum:
                fused_kernels.invoke_quant_fuse_sum(qm, tmp, sm, sc)
            else:
                fused_kernels.invoke_quant(qm, tmp, sc)
        if fuse_tail and li + 1 < cfg.num_hidden_layers and ly["down_proj"].fused_add_norm_quant(
                qm, sc, sm, od, h2, ha, self.layers[li + 1]["input_layernorm"], qh, sm if self.act_sum else None, sc, eps):
            return ha, "done"  # down_proj + residual add + the NEXT layer's input_layernorm + quant in one launch
        if use_peer:
            ly["down_proj"](qm, sc, sm, self.peer_b.tensor[:T])
            return h2, self.peer_b  # the all-reduce and the add are folded into the next norm
        ly["down_proj"](qm, sc, sm, od)
        self._all_reduce(od)
        return h2, od  # the add of `od` is folded into the next norm

    # ------------------------------------------------------------------ whole model
    def _run_layers(self, hidden0, T, is_prompt, meta):
        cur, delta = hidden0, None
        for li in range(self.cfg.num_hidden_layers):
            cur, delta = self._layer(li, cur, delta, T, is_prompt, meta)
        return cur, delta

    def _final_hidden(self, hidden, delta):
        """hidden + delta (the last layer's MLP output), materialised (prefill gathers rows from it)."""
        out = torch.empty_like(hidden)
        torch.add(hidden, delta, out=out)
        return out

    def _sample(self, hidden_last, delta=None):
        """final rms_norm (+ last residual add) + vocab-parallel lm_head + argmax (torch, as in the reference)."""
        x = torch.empty_like(hidden_last)
        if delta is not None and not torch.is_tensor(delta):
            self.ops.layernorm_ops.peer_add_rms_norm(x, hidden_last, delta, self.norm_weight, self.cfg.rms_norm_eps)
        elif delta is not None and self.fuse_add_norm:
            self.ops.layernorm_ops.add_rms_norm(x, hidden_last, delta, self.norm_weight, self.cfg.rms_norm_eps)
        else:
            if delta is not None:
                hidden_last = self._final_hidden(hidden_last, delta)
            self.ops.layernorm_ops.rms_norm(x, hidden_last, self.norm_weight, self.cfg.rms_norm_eps, False)
        logits = torch.matmul(x, self.lm_head.t())
        val, idx = logits.max(dim=-1)
        if self.tp_size == 1:
            return idx
        idx = idx + self.tp_rank * self.vocab_local
        vals = [torch.empty_like(val) for _ in range(self.tp_size)]
        idxs = [torch.empty_like(idx) for _ in range(self.tp_size)]
        torch.distributed.all_gather(vals, val, group=self.pg)
        torch.distributed.all_gather(idxs, idx, group=self.pg)
        vals, idxs = torch.stack(vals), torch.stack(idxs)
        best = vals.argmax(dim=0, keepdim=True)
        return idxs.gather(0, best).squeeze(0)

    @torch.no_grad()
    def prefill(self, tokens: torch.Tensor, seq_lens: List[int], seq_offset: int = 0):
        """tokens: int64 [sum(seq_lens)] for sequences seq_offset .. seq_offset+len(seq_lens)-1 of the batch."""
        from . import prefill_attention

        T = tokens.numel()
        B = len(seq_lens)
        dev = self.device
        assert self.buf is not None and T <= self.buf.T, f"prefill of {T} tokens exceeds the activation arena ({self.buf.T})"
        assert T == sum(seq_lens) and seq_offset + B <= self.batch
        assert max(seq_lens) <= self.kv.pages_per_seq * TOKENS_PER_BLOCK, "prompt longer than the per-sequence page budget"
        sl = torch.tensor(seq_lens, dtype=torch.int32, device=dev)
        cu = torch.zeros(B + 1, dtype=torch.int32, device=dev)
        cu[1:] = torch.cumsum(sl, 0)
        max_len = max(seq_lens)
        pad = self.ops.fused_attention_fine_grained_dense.compute_padding_offsets(cu, max_len, T)
        meta = {
            "seq_lens": sl, "padding_offset": pad, "max_seq_len": max_len, "flags": self._flags, "rank": self._rank,
            "prefill_attn": prefill_attention.make(cu, max_len, self.hq, self.hkv, self.cfg.head_dim),
        }
        if not self.fuse_silu_quant:
            meta["silu_tmp"] = torch.empty((T, self.inter), dtype=torch.float16, device=dev)
        # the page tables of this sub-batch
        saved = self.kv.tables
        self.kv.tables = [t[seq_offset: seq_offset + B] for t in saved]
        try:
            h0 = self.embed_tokens[tokens]
            h, delta = self._run_layers(h0, T, True, meta)
            self.last_hidden = self._final_hidden(h, delta)  # kept for tests
            last = self.last_hidden[(cu[1:] - 1).long()]
            nxt = self._sample(last)
        finally:
            self.kv.tables = saved
        self.context_lens[seq_offset: seq_offset + B] = sl
        return nxt

    @torch.no_grad()
    def decode_step(self, tokens: torch.Tensor, timestep: int):
        """One decode iteration for the whole batch.  tokens int64 [B] (device).  `timestep` = upper bound of
        the cached context over the batch (host int, baked into launch configs / CUDA graphs).  context_lens
        (device) is advanced by one inside the step."""
        B = tokens.numel()
        # `timestep` bounds the cached context of every sequence; the token appended by this step lands at position
        # <= timestep, which must exist in the page tables (the kernels trust the pointers)
        assert B <= self.buf.T and timestep < self.kv.pages_per_seq * TOKENS_PER_BLOCK + 1 and timestep <= self.max_ctx
        self.context_lens.add_(1)  # length incl. the new token (decoding_attention.py:151-153)
        meta = {"context_lens": self.context_lens, "timestep": timestep}
        if not self.fuse_silu_quant:
            meta["silu_tmp"] = self._silu_tmp
        h0 = self.embed_tokens[tokens]
        h, delta = self._run_layers(h0, B, False, meta)
        self.last_decode_state = (h, delta)
        return self._sample(h, delta)

    def prepare_decode(self):
        if not self.fuse_silu_quant:
            self._silu_tmp = torch.empty((self.batch, self.inter), dtype=torch.float16, device=self.device)


class DecodeGraph:
    """CUDA-graph replay of `decode_step` (the reference has no graphs: ~390 launches per step)."""

    def __init__(self, model: LlamaW4A8, max_timestep: int, warmup: int = 2):
        self.m = model
        B = model.batch
        self.tokens = torch.zeros((B,), dtype=torch.int64, device=model.device)
        self.out = torch.zeros((B,), dtype=torch.int64, device=model.device)
        model.prepare_decode()
        saved = model.context_lens.clone()
        # the warm-up steps append `warmup` tokens per sequence before the lengths are restored
        assert int(saved.max()) + warmup <= model.kv.pages_per_seq * TOKENS_PER_BLOCK, "no page room for the warm-up steps"
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):  # allocates lazy workspaces, tensor maps, NCCL channels
                self.out.copy_(model.decode_step(self.tokens.clone(), max_timestep))
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        model.context_lens.copy_(saved)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            nxt = model.decode_step(self.tokens, max_timestep)
            self.out.copy_(nxt)
            self.tokens.copy_(nxt)  # the next replay consumes this step's samples without touching the host
        model.context_lens.copy_(saved)
        torch.cuda.synchronize()

    def step(self):
        """tokens <- previous output is chained on device by the caller when desired."""
        self.graph.replay()
        return self.out


def kernel_launches_per_decode_step(cfg: LlamaConfig, fuse_silu_quant: bool = True) -> int:
    """Count of OUR kernels launched per decode step (torch's embedding/add/matmul/argmax not included)."""
    per_layer = 2 + 4 + 1 + (1 if fuse_silu_quant else 2)  # (add+)norms, gemms, attention(+quant), silu(+quant)
    return per_layer * cfg.num_hidden_layers + 1  # + final (add+)rms_norm
