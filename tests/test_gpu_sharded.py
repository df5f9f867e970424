"""Hypothesis sharding on real GPUs (BASELINE cfg 3): the N-rank result equals the 1-GPU result.
Two entries: a test that runs in-process when pytest itself is launched under torchrun (WORLD_SIZE > 1), and a
self-spawning 2-rank test (the two ranks share cuda:0 when the box has one GPU -- a functional check of the
peer-mailbox exchange; on a box with >= 2 GPUs they take one each)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _spawn(world):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "shard_worker.py")]
    env = dict(os.environ)
    env.pop("WORLD_SIZE", None)
    env.pop("RANK", None)
    env.setdefault("NCCL_DEBUG", "WARN")
    return subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=420)


def test_two_ranks_equal_single_gpu():
    r = _spawn(2)
    out = r.stdout + r.stderr
    if "SHARD_SKIP" in out:
        pytest.skip("peer-memory communicator unavailable on this box: " + out[-300:])
    if torch.cuda.device_count() < 2 and r.returncode != 0 and ("timeout" in out or "did not answer" in out):
        pytest.skip("two contexts on one GPU did not time-slice the exchange: " + out[-300:])
    assert r.returncode == 0 and "SHARD_OK" in out, out[-3000:]


@pytest.mark.skipif(int(os.environ.get("WORLD_SIZE", "1")) < 2, reason="run pytest under torchrun for this one")
def test_sharded_under_torchrun():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import shard_worker
    shard_worker.main()
