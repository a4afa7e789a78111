#!/usr/bin/env python
"""Derive the retrieval / streaming KV-head split LServe uses for Llama-3-8B-Instruct-Gradient-1048k at
`--static-sparsity 0.5` from the reference's data fixture attn_patterns/<model>/full_attention_heads.tsv, with the
reference's own rule (omniserve/attn_config.py:113-150: clip to [0,1], threshold = quantile(static_sparsity) over the
whole [layers, kv_heads] matrix, head is a retrieval head iff score >= threshold; the reference adds U(0,1e-6) random
noise to break ties (many scores clip to exactly 0) -- replaced here by a deterministic ramp of the same magnitude so
that the fixture is reproducible) plus the head_rank_table of
omniserve/modeling/layers/ctx_attn/ctx_attn_init.py:52-76.  Output: tests/golden/head_split_llama3_8b_1048k_s50.json.
Run in the authoring container only (needs /root/reference)."""
import json
import os

import numpy as np

REF = "/root/reference/attn_patterns/Llama-3-8B-Instruct-Gradient-1048k"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    m = np.clip(np.loadtxt(os.path.join(REF, "full_attention_heads.tsv"), dtype=float, delimiter="\t"), 0, 1)
    m = m + np.linspace(0, 1e-6, m.size).reshape(m.shape)
    thr = np.quantile(m, 0.5)
    flags = (m >= thr).astype(int)
    ranks = np.zeros_like(flags)
    for l in range(flags.shape[0]):
        ranks[l][flags[l] == 0] = np.arange((flags[l] == 0).sum())
        ranks[l][flags[l] == 1] = np.arange((flags[l] == 1).sum())
    cfg = json.load(open(os.path.join(REF, "config.json")))
    out = {"model": "Llama-3-8B-Instruct-Gradient-1048k", "static_sparsity": 0.5, "threshold": float(thr),
           "actual_sparsity": float(1 - flags.mean()), "retrieval_head_flags": flags.tolist(),
           "head_rank_table": ranks.tolist(), "pattern_config": cfg}
    json.dump(out, open(os.path.join(HERE, "head_split_llama3_8b_1048k_s50.json"), "w"), indent=1)
    print("retrieval heads per layer:", flags.sum(1).tolist())


if __name__ == "__main__":
    main()
