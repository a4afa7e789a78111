#!/usr/bin/env python
"""torchrun --nproc-per-node 2 tools/tp_check.py : tensor-parallel (NCCL all-reduce after o_proj / down_proj) vs the
same model on one GPU.  Row-parallel shards quantise their own activation slices, so hidden states agree to int8
rounding noise, not bit-exactly; the check bounds the relative difference and exercises graph capture with NCCL."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from omniserve_b200.model import DecodeGraph, LlamaConfig, LlamaW4A8  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = LlamaConfig(hidden_size=1024, intermediate_size=2048, num_hidden_layers=3, num_attention_heads=8,
                  num_key_value_heads=4, vocab_size=2048)
full = LlamaW4A8(cfg, dev, 0, 1, seed=3)
full.alloc(batch=4, max_ctx=256, max_tokens=4 * 64)
shard = LlamaW4A8(cfg, dev, rank, world, seed=3)
shard.load_shard_of(full)
shard.alloc(batch=4, max_ctx=256, max_tokens=4 * 64)
g = torch.Generator().manual_seed(0)
toks = torch.randint(0, cfg.vocab_size, (4 * 50,), generator=g).to(dev)
lens = [50] * 4
a = full.prefill(toks, lens)
b = shard.prefill(toks, lens)
ha, hb = full.last_hidden.float(), shard.last_hidden.float()
rel = float((ha - hb).abs().max() / ha.abs().max())
full.prepare_decode(); shard.prepare_decode()
gr = DecodeGraph(shard, 256)       # NCCL all-reduce captured in the CUDA graph
gr.tokens.copy_(a)
t1 = full.decode_step(a.clone(), 256)
gr.step()
torch.cuda.synchronize()
h1 = full.last_decode_state[0].float() + full.last_decode_state[1].float()
h2 = shard.last_decode_state[0].float() + shard.last_decode_state[1].float()
rel2 = float((h1 - h2).abs().max() / h1.abs().max())
agree = float((t1 == gr.out).float().mean())
ok = rel < 5e-2 and rel2 < 5e-2
if rank == 0:
    print(f"tp{world}: prefill hidden rel diff {rel:.3e}, decode hidden rel diff {rel2:.3e}, argmax agreement {agree:.2f}, "
          f"{'OK' if ok else 'FAIL'}")
sys.stdout.flush()
os._exit(0 if ok else 1)   # destroy_process_group() can hang after NCCL work was captured in a CUDA graph
