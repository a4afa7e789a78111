#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/tp_profile.py : per-kernel device time of ONE tensor-parallel decode layer at the
bench's shapes (Llama-3-8B, bs=64, ctx 1280), CUDA-graph timed on every rank (max over ranks printed by rank 0), plus the
exchange step in its three forms: NCCL all-reduce + add+norm+quant, the fused peer-memory kernel, and (if the fabric offers
multicast) the in-switch reduction.  Output -> profiles/r2_tp_step.md (copy of stdout)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from omniserve_b200.model import LlamaConfig, LlamaW4A8  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = LlamaConfig.llama3_8b()
cfg.num_hidden_layers = 8
B, ctx = 64, 1280
m = LlamaW4A8(cfg, dev, rank, world)
m.alloc(B, 1536, B)
for pool in m.kv.k_pools + m.kv.v_pools:      # valid random pages (profiling only)
    u8 = pool.view(torch.uint8)
    u8.random_(0, 256)
    sz = u8[:, m.hkv * 4096:].view(torch.float16)
    sz[:, :m.hkv * 64] = 0.25
    sz[:, m.hkv * 64:] = 7.5
m.enable_peer_allreduce()
b, ops = m.buf, m.ops
L = cfg.num_hidden_layers
qh, sc, sm = b.quantized_hidden_states_buffer[:B], b.quantized_scale_buffer[:B], b.quantized_sum_buffer[:B]
qh.random_(-127, 127); sc.fill_(0.02); sm.fill_(0.1)
b.quantized_attn_buffer[:B].random_(-127, 127); b.quantized_mlp_act_buffer[:B].random_(-127, 127)
qkv = b.qkv_proj_act_buffer[:B]; qkv.normal_()
q3 = qkv[:, :m.q_size].view(B, m.hq, 128)
k3 = qkv[:, m.q_size:m.q_size + m.kv_size].view(B, m.hkv, 128)
v3 = qkv[:, m.q_size + m.kv_size:].view(B, m.hkv, 128)
lens = torch.full((B,), ctx + 1, dtype=torch.int32, device=dev)
hid = torch.randn((B, cfg.hidden_size), device=dev).half()
hout = torch.empty_like(hid)
od = b.out_down_proj_act_buffer[:B]; od.normal_()
gu = b.gate_up_proj_act_buffer[:B]; gu.normal_()


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    best = 1e9
    for _ in range(reps):
        dist.barrier(); torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); e.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(e))
    t = torch.tensor([best], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) / L * 1e3   # us per launch


rows = []
for name, x, out in (("qkv_proj", qh, qkv), ("o_proj", b.quantized_attn_buffer[:B], od), ("gate_up_proj", qh, gu),
                     ("down_proj", b.quantized_mlp_act_buffer[:B], od)):
    rows.append((f"W4A8 GEMM {name} [{m.layers[0][name].out_features} x {m.layers[0][name].in_features}]",
                 timed(lambda name=name, x=x, out=out: [ly[name](x, sc, sm, out) for ly in m.layers])))
rows.append((f"KV4 decode attention (Hq={m.hq}, Hkv={m.hkv} local, ctx {ctx})",
             timed(lambda: [ops.fused_attention_pure_dense.single_query_attention_quant(
                 q3, k3, v3, m.kv.tables[li], lens, None, 1536, 64, m.kv_size // 2, ctx, 128, cfg.rope_theta, True, True, True,
                 b.quantized_attn_buffer[:B], sm, sc, history_is_stable=True) for li in range(L)])))
rows.append(("silu*mul + quant", timed(lambda: [ops.activation_ops.silu_and_mul_quant(b.quantized_mlp_act_buffer[:B], gu, sm, sc)
                                              for _ in range(L)])))
rows.append(("add + norm + quant (no exchange)", timed(lambda: [ops.layernorm_ops.add_rms_norm_general(
    qh, hid, od, hout, m.layers[0]["input_layernorm"], sm, sc, 1e-5) for _ in range(L)])))
rows.append(("NCCL all-reduce [64, 4096] fp16", timed(lambda: [dist.all_reduce(od) for _ in range(L)])))
m.peer_a.tensor[:B].normal_()
rows.append(("fused exchange: peer-memory all-reduce + add + norm + quant",
             timed(lambda: [ops.layernorm_ops.peer_add_rms_norm_general(qh, hid, m.peer_a if i % 2 == 0 else m.peer_b, hout,
                                                                         m.layers[0]["input_layernorm"], sm, sc, 1e-5) for i in range(L)])))
if rank == 0:
    print(f"## tensor-parallel decode layer, tp{world}, Llama-3-8B bs=64 ctx={ctx}: us per launch (graph-timed, max over ranks)")
    tot = 0.0
    for n, t in rows:
        print(f"| {n} | {t:7.2f} |")
    per_layer = sum(t for n, t in rows[:6]) + rows[6][1] * 0 + 2 * rows[8][1]
    print(f"layer estimate: 4 GEMMs + attention + silu + 2 fused exchanges = {per_layer:.1f} us -> {per_layer * 32 / 1e3:.2f} ms per 32-layer step")
    by = (world - 1) * B * cfg.hidden_size * 2
    print(f"NVLink bytes read per rank per fused exchange: (W-1) x 64 x 4096 x 2 B = {by / 1e6:.2f} MB; at {rows[8][1]:.1f} us -> "
          f"{by / rows[8][1] / 1e3:.0f} GB/s per GPU")
sys.stdout.flush()
os._exit(0)
