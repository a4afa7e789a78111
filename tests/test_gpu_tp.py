"""-m gpu: tensor parallelism on 2 / 4 / 8 GPUs of one box, one process per GPU under torchrun,
NCCL + the fused peer-memory exchange.  Skipped when the box has fewer GPUs (the round-end GPU test box has one; run
`gpurun --gpus 8 -- python -m pytest tests/test_gpu_tp.py -m gpu -q -s` for the full matrix; log: profiles/r2_tp_tests.log).
What is asserted is in tests/tp_worker.py."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 4, 8])
def test_tensor_parallel_exact_partials_exchange_and_end_to_end(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, this box has {torch.cuda.device_count()}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "tp_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    tail = r.stdout[-3000:] + "\n--- stderr (tail) ---\n" + r.stderr[-2500:]
    print(tail)
    assert r.returncode == 0, tail
    assert f"tp{world}:" in r.stdout and "OK" in r.stdout
