#!/usr/bin/env python
"""Run THE REFERENCE'S OWN model code -- `LlamaDecoderLayer` of omniserve/modeling/models/llama_w4a8_unpad.py:365-438 with
its LlamaAttention / LlamaMLP / RMSNormGeneral / SiluAndMulQuant / W4A8OF16LinearDynamicInputScale /
DecodingAttentionWrapper / ApplyBiasRopeUpdateKVCacheWrapper, all UNCHANGED, imported from /root/reference -- on top of
this repository's `omniserve_backend` shim, with the C library replaced by a recorder, and write the sequence of C-ABI
calls it makes (entry point, every scalar argument, which pointers are null) to tests/golden/ref_layer_trace.json.

This container has no GPU, so nothing is computed: what is proven is that the reference's callers import and run
unchanged over the boundary (every positional argument binds to our mirror), and WHICH ops they invoke in WHICH order
with WHICH shapes.  tests/test_ref_callers_cpu.py re-derives the trace when /root/reference is present and compares it
with (a) the committed fixture and (b) the trace of omniserve_b200/model.py driving the same layer through the same
(unfused) ops; tests/test_gpu_model.py replays model.py's unfused path on the GPU and requires the same trace plus
bit-identical results against the fused production path.

    python tests/golden/make_ref_trace.py            # writes the fixture
Test infrastructure; authoring container only."""
from __future__ import annotations

import ctypes
import json
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("OMNISERVE_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "ref_layer_trace.json")

DIMS = dict(hidden=4096, inter=14336, heads=32, kv_heads=8, head_dim=128, eps=1e-5, rope=500000.0, vocab=1024)


class Recorder:
    """Stands in for the ctypes CDLL: every `ob_*` call is recorded and returns 0 (OB_OK)."""

    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        if not name.startswith("ob_"):
            raise AttributeError(name)

        def f(*args):
            self.calls.append([name, [self._norm(a) for a in args]])
            return 0
        return f

    @staticmethod
    def _norm(a):
        if a is None:
            return "null"
        if isinstance(a, bool):
            return int(a)
        if isinstance(a, int):
            return "ptr" if abs(a) >= (1 << 24) else a     # device / host addresses vs sizes and flags
        if isinstance(a, float):
            return round(a, 9)
        obj = getattr(a, "_obj", None)                        # ctypes.byref(struct)
        if isinstance(obj, ctypes.Structure):
            d = {}
            for fname, ftype in obj._fields_:
                v = getattr(obj, fname)
                if ftype is ctypes.c_void_p:
                    d[fname] = "null" if not v else "ptr"
                elif isinstance(v, float):
                    d[fname] = round(v, 9)
                elif isinstance(v, int):
                    d[fname] = v
                else:
                    d[fname] = "opaque"
            return d
        return "opaque"


def install_recorder():
    sys.path.insert(0, ROOT)
    from omniserve_b200 import _lib as L
    rec = Recorder()
    L.lib = lambda: rec
    L.require_cuda = lambda *a, **k: None
    L.stream = lambda: 0
    return rec


def import_reference():
    """The reference's Python package, unchanged, over our shim.  Three things of the environment are stubbed: the
    un-vendored `block_sparse_attn` dependency, `torch.cuda.current_device()` (evaluated at class-definition
    time in w4a8_linear.py:24; there is no GPU here) and flash-attn's prefill kernel (third party, CUDA only)."""
    if REF not in sys.path:
        sys.path.insert(1, REF)
    bs = types.ModuleType("block_sparse_attn")
    for n in ("block_streaming_attn_func", "block_sparse_attn_func", "token_streaming_attn_func", "flash_attn_varlen_func"):
        setattr(bs, n, lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("block_sparse_attn is not installed")))
    sys.modules.setdefault("block_sparse_attn", bs)
    torch.cuda.current_device = lambda: "cpu"
    import importlib
    import omniserve_backend  # noqa: F401  (this repository's shim package)
    m = importlib.import_module("omniserve.modeling.models.llama_w4a8_unpad")
    return m


class _SpAttn:
    def sparse_kv_cache_enabled(self): return False
    def sparse_context_enabled(self): return False
    def get_dec_sub_chunk_per_block(self): return 4
    def get_sparse_decode_mode(self): return 0
    def get_dec_dynamic_sparse_token_budget(self): return 4096
    def get_dec_selector_update_interval(self): return 4
    def get_static_sparsity(self): return 0.0


def reference_layer_trace(rec, T_decode=64, ctx=1280, prompt_lens=(96, 160)):
    from transformers import LlamaConfig
    m = import_reference()
    d = DIMS
    hf = LlamaConfig(hidden_size=d["hidden"], intermediate_size=d["inter"], num_attention_heads=d["heads"],
                     num_key_value_heads=d["kv_heads"], rms_norm_eps=d["eps"], rope_theta=d["rope"], vocab_size=d["vocab"],
                     max_position_embeddings=8192, num_hidden_layers=1)
    hf.rope_scaling = None     # recent transformers fill in a default dict; the reference (transformers 4.x) sees None
    hf.rope_theta = d["rope"]  # ... and keep rope_theta inside rope_parameters; the reference reads the attribute
    model_config = types.SimpleNamespace(chunk_prefill_size=8192, kv_quant_granularity="fine_grained", multiblock_switch=2048,
                                         sp_attn_config=_SpAttn())
    layer = m.LlamaDecoderLayer(hf, model_config, group_size=-1, layer_idx=0,
                                kv_cache_config={"INT4_ENABLED": True, "ZEROS_ENABLED": True})
    att = layer.self_attn
    # what ctx_attn_init.py:11-85 registers on every attention module (dense: all heads are retrieval heads)
    att.retrieval_head_flags = torch.ones(d["kv_heads"], dtype=torch.int32)
    att.head_rank_table = torch.arange(d["kv_heads"], dtype=torch.int32)
    att.pooling_heads_idx = torch.arange(d["kv_heads"], dtype=torch.int32)
    att.num_retrieval_kv_heads, att.num_streaming_kv_heads = d["kv_heads"], 0
    att.sink_size = att.local_size = att.sink_blocks = att.local_blocks = 0
    att.head_mask_type, att.streaming_info = None, None

    class LlamaForCausalLM:   # what ActivationBuffer reads from the model (input_metadata.py:27-47)
        pass
    fake = LlamaForCausalLM()
    fake.model = types.SimpleNamespace(embed_tokens=types.SimpleNamespace(weight=torch.zeros(1, dtype=torch.float16)))
    fake.model_config = model_config
    fake.q_size, fake.kv_size, fake.config = att.q_size, att.kv_size, hf
    from omniserve.utils.input_metadata import ActivationBuffer

    traces = {}   # the prompt pass comes first, as in the engine (it also creates cached_dynamic_sparse_page_idx, :309-313)
    # ---------------------------------------------------------------- prefill chunk (:309-325)
    T = sum(prompt_lens)
    ab = ActivationBuffer(fake, T)
    ab.allocate_activation_buffer()
    cu = torch.tensor([0] + list(torch.tensor(prompt_lens).cumsum(0)), dtype=torch.int32)
    pages = (max(prompt_lens) + 63) // 64
    meta = types.SimpleNamespace(
        activation_buffer=ab, is_prompt=True, max_seq_len=max(prompt_lens), cu_seqlens=cu,
        padding_offsets=torch.zeros(T, dtype=torch.int32),
        retrieval_context_lens=torch.tensor(prompt_lens, dtype=torch.int32), streaming_context_lens=None,
        retrieval_block_tables=[torch.zeros((len(prompt_lens), 2, pages), dtype=torch.int64)], streaming_block_tables=[None])
    m.attention_wrapper = lambda q, k, v, **kw: torch.zeros_like(q)   # flash-attn varlen prefill: third party, CUDA only
    rec.calls.clear()
    layer(torch.zeros((T, d["hidden"]), dtype=torch.float16), meta)
    traces["prefill"] = {"T": T, "prompt_lens": list(prompt_lens), "calls": list(rec.calls)}
    # ---------------------------------------------------------------- decode step (llama_w4a8_unpad.py:326-361,406-438)
    ab = ActivationBuffer(fake, T_decode)
    ab.allocate_activation_buffer()
    pages = (ctx + 64) // 64
    meta = types.SimpleNamespace(
        activation_buffer=ab, is_prompt=False, max_seq_len=ctx,
        retrieval_context_lens=torch.full((T_decode,), ctx + 1, dtype=torch.int32),
        retrieval_block_tables=[torch.zeros((T_decode, 2, pages), dtype=torch.int64)], streaming_block_tables=[None])
    rec.calls.clear()
    hidden = torch.zeros((T_decode, d["hidden"]), dtype=torch.float16)
    layer(hidden, meta)
    traces["decode"] = {"T": T_decode, "ctx": ctx, "calls": list(rec.calls)}
    return traces


def main():
    if not os.path.isdir(os.path.join(REF, "omniserve")):
        print("reference tree absent: nothing to do")
        return 0
    rec = install_recorder()
    tr = reference_layer_trace(rec)
    out = {"what": "C-ABI calls made by the reference's unchanged LlamaDecoderLayer over this repository's omniserve_backend shim",
           "dims": DIMS, "traces": tr}
    txt = json.dumps(out, indent=1, sort_keys=True) + "\n"
    if "--check" in sys.argv:
        return 0 if os.path.exists(OUT) and open(OUT).read() == txt else 1
    open(OUT, "w").write(txt)
    for k, v in tr.items():
        print(k, [c[0] for c in v["calls"]])
    return 0


if __name__ == "__main__":
    sys.exit(main())
