"""-m gpu: the multi-GPU data plane (user-sharded iterations) against one engine.

Two ranks are launched with torch.distributed.run.  On a box with >= 2 GPUs they use one GPU each over NCCL;
on the driver's single-GPU test box both ranks share cuda:0 (see tests/multi_rank_worker.py).  Every exchange
mode must reproduce the single-engine fp64 result to <= 1e-10 with bit-identical item replicas."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_sharded_iterations_equal_single_engine():
    import torch
    shared = torch.cuda.device_count() < 2
    env = dict(os.environ, HPF_TEST_SHARED_GPU="1" if shared else "0")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multi_rank_worker.py")]
    res = subprocess.run(cmd, env=env, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    lines = [json.loads(l[len("PARITY "):]) for l in res.stdout.splitlines() if l.startswith("PARITY ")]
    assert res.returncode == 0, res.stdout[-3000:]
    assert len(lines) >= 3, res.stdout[-3000:]
    for r in lines:
        assert r["ok"] and r["max_rel_err"] < 1e-10 and r["item_replicas_bit_identical"], r
