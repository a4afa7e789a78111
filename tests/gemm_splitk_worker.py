"""Run by tests/test_gpu_gemm.py::test_decode_split_k_reduction_paths in a fresh process per value of OB_GEMM_DEC_CLUSTER (the
library reads it once): decode-kernel GEMMs whose tiles are split 2 / 4 / 8 ways along K, against the oracle, with unit scales so
that the fp16 outputs ARE the INT32 accumulators (|acc| <= 2048) -- integer equality, no tolerance.  Prints OK."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from omniserve_b200 import _lib as L  # noqa: E402
from oracle import w4a8 as ow  # noqa: E402


def t(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def case(M, N, K, ctas, per_group, seed):
    rng = np.random.default_rng(seed)
    q = rng.integers(0, 16, (N, K), dtype=np.uint8)
    a = rng.integers(-1, 2, (M, K), dtype=np.int8)               # |acc| <= 15 K / ... stays far below 2048 for K <= 1024
    a[rng.random((M, K)) < 0.7] = 0
    qw = ow.pack_w4(q)
    s1 = np.ones(N, np.float16)
    sa = np.ones(M, np.float16)
    out = torch.full((M, N), float("nan"), dtype=torch.float16, device="cuda")
    ta, tq, ts1, tsa = t(a), t(qw), t(s1), t(sa)     # keep every device tensor alive until the launch has finished
    if per_group:
        ng = K // 128
        ts2 = t(ow.pack_s2(np.ones((N, ng), np.int64)).astype(np.int8))
        tz2 = t(ow.pack_s2(np.zeros((N, ng), np.int64)).astype(np.int8))
        code = L.lib().ob_w4a8_gemm_ex(1, L.ptr(ta), L.ptr(tq), L.ptr(tz2), L.ptr(ts2), L.ptr(ts1), L.ptr(tsa), 0, 0,
                                       L.ptr(out), M, N, K, N, 0, 3, ctas, L.stream())
    else:
        tszs, tssum = t(np.zeros(N, np.float16)), t(np.zeros(M, np.float16))
        code = L.lib().ob_w4a8_gemm_ex(0, L.ptr(ta), L.ptr(tq), 0, 0, L.ptr(ts1), L.ptr(tsa), L.ptr(tszs), L.ptr(tssum),
                                       L.ptr(out), M, N, K, N, 0, 3, ctas, L.stream())
    torch.cuda.synchronize()
    assert code == 0, code
    want = a.astype(np.int64) @ q.astype(np.int64).T
    assert np.abs(want).max() <= 2048
    got = out.cpu().numpy().astype(np.float64)
    assert np.array_equal(got, want.astype(np.float64)), (M, N, K, ctas, per_group, np.abs(got - want).max())


def main():
    n = 0
    for per_group in (False, True):
        for M in (7, 16, 32, 64):
            for N, K, tiles in ((256, 1024, 2), (384, 2048, 3), (256, 1792, 2)):
                kb = K // 128
                for s in (2, 4, 8):
                    if kb % s:
                        continue                              # K = 1792: 14 K-blocks -> only the 2-way split (7 per CTA: odd tail step)
                    for rep in range(2):                      # twice: the L2 path must leave its workspace clean
                        case(M, N, K, tiles * s, per_group, seed=1000 * M + s + rep)
                        n += 1
                case(M, N, K, tiles, per_group, seed=M)        # whole tiles
                case(M, N, K, tiles * kb, per_group, seed=M)   # one K-block per CTA (kb-way split: never a cluster for kb = 16)
    print(f"OK {n} split cases, OB_GEMM_DEC_CLUSTER={os.environ.get('OB_GEMM_DEC_CLUSTER')}")


if __name__ == "__main__":
    main()
