"""CPU suite, N > 1 path: the sharded-map protocol (two exchanges, owner-side verdicts, apply order) run as a
numpy model over torch.distributed/gloo with 2 and 3 processes, checked against the oracle after every scan."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world,port,exchange", [(2, 29711, "gather"), (3, 29712, "gather"), (2, 29713, "mailbox"), (3, 29714, "mailbox")])
def test_shard_protocol_model_over_gloo(world, port, exchange):
    """gather: the exchanges as collectives; mailbox: one-sided stores into the owners' (shared-memory) inboxes + arrival
    stamps + the end-of-scan flag exchange, i.e. the peer-memory protocol of the CUDA path"""
    env = dict(os.environ, OMP_NUM_THREADS="1", BNX_MODEL_EXCHANGE=exchange)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tests", "shard_model.py")], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and f"SHARD_MODEL_OK {world} {exchange}" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]


def test_split_points_covers_every_index_once():
    from bonxai_b200.sharded import split_points
    for n in (0, 1, 7, 131072, 1_024_000):
        for world in (1, 2, 3, 8):
            parts = split_points(n, world)
            assert parts[0][0] == 0 and parts[-1][1] == n and all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            assert max(h - l for l, h in parts) - min(h - l for l, h in parts) <= 1


@pytest.mark.parametrize("world,port", [(2, 29721), (3, 29722)])
def test_host_bootstrap_callback_over_gloo(world, port):
    """the caller-supplied all-gather of bnx_map_shard_host_init (mailbox handles without NCCL), called through its C
    function pointer by 2 and 3 gloo processes"""
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tests", "bootstrap_worker.py")], capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert r.returncode == 0 and f"BOOTSTRAP_OK {world}" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]
