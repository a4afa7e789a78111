"""Subprocess body of tests/test_ref_callers_cpu.py: the C-ABI call trace of omniserve_b200/model.py driving ONE decoder
layer (prefill chunk, then a decode step) through the reference's op set only -- no fused extension ops -- with the C
library replaced by the recorder of tests/golden/make_ref_trace.py.  Prints the trace as JSON.  CPU only."""
import json
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_ref_trace as T  # noqa: E402


class ReferenceOpsOnly:
    """omniserve_b200.model.Ops restricted to the functions the reference's extension modules export
    (tests/golden/ref_api.json): hides this repository's fused extensions so that model.py takes the reference's chain."""

    def __init__(self):
        import importlib
        api = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_api.json")))
        for n in ("activation_ops", "fused_attention_fine_grained_dense", "fused_attention_pure_dense", "fused_kernels",
                  "layernorm_ops", "qgemm_w4a8_per_chn", "qgemm_w4a8_per_group"):
            real = importlib.import_module(f"omniserve_b200.backend.{n}")
            ns = types.SimpleNamespace(**{f: getattr(real, f) for f in api[n]})
            setattr(self, n, ns)


def main():
    rec = T.install_recorder()
    from omniserve_b200.model import LlamaConfig, LlamaW4A8
    d = T.DIMS
    cfg = LlamaConfig(hidden_size=d["hidden"], intermediate_size=d["inter"], num_hidden_layers=1, num_attention_heads=d["heads"],
                      num_key_value_heads=d["kv_heads"], vocab_size=d["vocab"], rope_theta=d["rope"], rms_norm_eps=d["eps"])
    m = LlamaW4A8(cfg, "cpu", fuse_silu_quant=False, ops=ReferenceOpsOnly())
    assert not m.fuse_add_norm and not m.fuse_attn_quant and not m.fuse_gemm_norm
    fx = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_layer_trace.json")))["traces"]
    out = {}
    # ---- prefill chunk
    lens = fx["prefill"]["prompt_lens"]
    Tn = sum(lens)
    from omniserve_b200.model import ActivationBuffer
    m.buf = ActivationBuffer(Tn, cfg.hidden_size, m.inter, m.q_size + 2 * m.kv_size, m.q_size, "cpu")
    pages = (max(lens) + 63) // 64
    m.kv = types.SimpleNamespace(tables=[torch.zeros((len(lens), 2, pages), dtype=torch.int64)], pages_per_seq=pages)
    meta = {"seq_lens": torch.tensor(lens, dtype=torch.int32), "padding_offset": torch.zeros(Tn, dtype=torch.int32),
            "max_seq_len": max(lens), "flags": m._flags, "rank": m._rank,
            "prefill_attn": lambda q, k, v: torch.zeros_like(q), "silu_tmp": torch.empty((Tn, m.inter), dtype=torch.float16)}
    rec.calls.clear()
    m._layer(0, torch.zeros((Tn, cfg.hidden_size), dtype=torch.float16), None, Tn, True, meta)
    out["prefill"] = list(rec.calls)
    # ---- decode step
    Td, ctx = fx["decode"]["T"], fx["decode"]["ctx"]
    m.buf = ActivationBuffer(Td, cfg.hidden_size, m.inter, m.q_size + 2 * m.kv_size, m.q_size, "cpu")
    pages = (ctx + 64) // 64
    m.kv = types.SimpleNamespace(tables=[torch.zeros((Td, 2, pages), dtype=torch.int64)], pages_per_seq=pages)
    m.max_ctx = 8192     # kv_max_seq_len of the reference layer: min(max_seq_len, max_position_embeddings)
    meta = {"context_lens": torch.full((Td,), ctx + 1, dtype=torch.int32), "timestep": ctx,
            "silu_tmp": torch.empty((Td, m.inter), dtype=torch.float16)}
    rec.calls.clear()
    m._layer(0, torch.zeros((Td, cfg.hidden_size), dtype=torch.float16), None, Td, False, meta)
    out["decode"] = list(rec.calls)
    print("TRACE_JSON " + json.dumps(out))


if __name__ == "__main__":
    main()
