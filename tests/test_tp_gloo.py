"""world_size-2 gloo test of the tensor-parallel algebra on CPU: row-parallel partial GEMMs (each rank with its
own per-token activation scale/sum) all-reduced == the unsharded per-channel GEMM, up to fp16 rounding."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from omniserve_b200 import tp
    from oracle import act, w4a8
    rng = np.random.default_rng(5)  # same data on both ranks
    M, N, K = 6, 64, 512
    qw = rng.integers(0, 16, (N, K), dtype=np.uint8)
    z = rng.integers(0, 16, N).astype(np.float32)
    s1 = rng.uniform(0.005, 0.02, N).astype(np.float16)
    x = (rng.standard_normal((M, K))).astype(np.float16)
    p = {"qweight": torch.from_numpy(w4a8.pack_w4(qw)), "s1_scales": torch.from_numpy(s1),
         "s1_szeros": torch.from_numpy((z * s1.astype(np.float32)).astype(np.float16))}
    kr = range(rank * K // world, (rank + 1) * K // world)
    sh = tp.shard_row(p, kr)
    xq, sa, ss = act.quant_fuse_sum(x[:, kr.start:kr.stop])  # local per-token quant of the local slice
    _, part = w4a8.gemm_per_chn(xq, sh["qweight"].numpy(), sh["s1_scales"].numpy(), sa, sh["s1_szeros"].numpy(), ss)
    t = torch.from_numpy(part.astype(np.float32))
    dist.all_reduce(t)
    if rank == 0:
        w_real = (qw.astype(np.float32) - z[:, None]) * s1.astype(np.float32)[:, None]
        ref = x.astype(np.float32) @ w_real.T
        err = float(np.abs(t.numpy() - ref).max() / np.abs(ref).max())
        q.put(err)
    dist.destroy_process_group()


def test_row_parallel_allreduce_matches_full_gemm():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert err < 2e-2  # int8 activation quantisation noise only
