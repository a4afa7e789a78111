"""Subprocess body: run THE REFERENCE'S page-selector kernel (oracle/_ref, unmodified sources) on the case stored in
argv[1] (.npz) and write its scores to argv[2].  Runs in its own process because (a) the kernel needs the launch
interposer of oracle/ref_launch_shim.c preloaded (LD_PRELOAD must be set before the process starts) and (b) a faulting
reference kernel must not take the test session's CUDA context with it.  Test infrastructure."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.gpu_util import ref_module, t  # noqa: E402


def main(inp, outp):
    d = np.load(inp)
    ref = ref_module("fused_attention_selector")
    if ref is None:
        print("reference selector module not shipped")
        return 3
    kpool = t(d["k_pool"])
    bt = d["bt"]
    B, P = bt.shape
    ptrs = np.zeros((B, 2, P), np.int64)
    ptrs[:, 0] = kpool.data_ptr() + bt * int(d["k_page_bytes"])
    ptrs[:, 1] = ptrs[:, 0]
    q, k, v = d["q"], d["k"], d["v"]
    Hq, Hkv = q.shape[1], k.shape[1]
    qkv = torch.cat([t(q).reshape(B, -1), t(k).reshape(B, -1), t(v).reshape(B, -1)], dim=1).contiguous()
    tq = qkv[:, :Hq * 128].view(B, Hq, 128)
    tk = qkv[:, Hq * 128:(Hq + Hkv) * 128].view(B, Hkv, 128)
    tv = qkv[:, (Hq + Hkv) * 128:].view(B, Hkv, 128)
    Hr = int(d["Hr"])
    out = ref.single_query_page_selector(tq, tk, tv, t(ptrs), None, t(d["flags"]), t(d["rank"]), None, t(d["lens"]), None,
                                         1 << 20, 64, Hr * 64, 0, 0, 0, 0, 0, Hr, 0, int(d["timestep"]), 128, 500000.0, 1.0,
                                         True, True, True, 16, Hr * 128, 1000000)
    torch.cuda.synchronize()
    np.savez(outp, out=out.cpu().numpy())
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1], sys.argv[2]))
