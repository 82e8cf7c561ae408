"""torchrun worker for tests/test_gpu_sharded.py::test_nccl_two_ranks (and manual N-GPU checks):
every rank inserts its slice of each scan into a root-key-sharded map over NCCL; rank 0 gathers the shard dumps
and compares their union with the CPU oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bonxai_b200 import capi, synth  # noqa: E402
from bonxai_b200.sharded import ShardedMap, split_points  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    # BNX_SHARD_BOOTSTRAP=host: the processes find each other over gloo and hand the mailbox handles to the library
    # through a callback, so the ranks may share GPUs (CUDA IPC between processes on one device): the same mailbox
    # kernels, ld.acquire.sys / st.release.sys stamps and freeze + replay logic as on an NVLink box
    host = os.environ.get("BNX_SHARD_BOOTSTRAP", "nccl") == "host"
    local = local % torch.cuda.device_count()
    torch.cuda.set_device(local)
    if host:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mode = os.environ.get("BNX_SHARD_TEST_MODE", "sync")
    sm = ShardedMap(0.1, bootstrap="host" if host else "nccl")
    if rank == 0:
        import oracle
        om = oracle.load("port").map(0.1)
    if mode == "fleet":
        # 80 pipelined FLEET steps without a sync in between: the queue fills (64), so the library drains inside an insert
        # whose fleet origins are already armed; with tiny pools and 1 MB growth steps that drain finds a frozen pipeline and
        # replays queued steps (each with its own origins) before the armed step runs
        keep, steps = [], 80
        for step in range(steps):
            scans = []
            for v in range(world):
                pts, origin = synth.lidar_scan(step, beams=16, azimuths=256, seed=7 + v)
                shift = np.float32([0.0, 150.0 * v, 0.0])
                scans.append((np.ascontiguousarray(pts[:, :3] + shift), origin + shift))
            mine = torch.from_numpy(scans[rank][0]).cuda()
            keep.append(mine)
            sm.insert_fleet(capi.DevPtr(mine.data_ptr()), len(scans[rank][0]), 12, len(scans[rank][0]), np.array([o for _, o in scans]), 40.0,
                            use_async=True)
            if rank == 0:
                for pts, origin in scans:
                    om.insert(pts, origin, 40.0)
        sm.sync()
        dig = sm.digest()
        st = sm.stats()
        if rank == 0:
            assert dig == capi.digest_of_dump(*om.dump()), "fleet pipeline with replays: sharded map differs from the oracle"
            assert st["replays"] > 0, st
        dist.barrier()
        sm.close()
        if rank == 0:
            print("SHARDED_OK", world, "fleet", st)
        dist.destroy_process_group()
        return
    keep = []
    for scan in range(6 if mode == "async" else 3):
        pts, origin = synth.lidar_scan(scan * 2, beams=32, azimuths=1024)
        parts = split_points(len(pts), world)
        lo, hi = parts[rank]
        n_max = max(h - l for l, h in parts)
        if scan % 2 == 0:
            dev = torch.from_numpy(pts[lo:hi]).cuda()
            keep.append(dev)
            sm.insert(capi.DevPtr(dev.data_ptr()), hi - lo, 16, lo, n_max, origin, 40.0, use_async=(mode == "async"))
        else:  # host input: staged through the copy-stream ring when pipelined
            host = np.ascontiguousarray(pts[lo:hi])
            keep.append(host)
            sm.insert(host, hi - lo, 16, lo, n_max, origin, 40.0, use_async=(mode == "async"))
        if mode == "async" and scan % 3 != 2:
            if rank == 0:
                om.insert(pts, origin, 40.0)
            continue  # compare only every third scan: the others stay in the pipeline
        xyz, w = sm.map.dump(sort=False)
        gathered = [None] * world
        dist.all_gather_object(gathered, (xyz, w))
        if rank == 0:
            om.insert(pts, origin, 40.0)
            gx = np.concatenate([g[0] for g in gathered])
            gw = np.concatenate([g[1] for g in gathered])
            order = np.lexsort((gx[:, 2], gx[:, 1], gx[:, 0]))
            ox, ow = om.dump()
            assert np.array_equal(gx[order], ox) and np.array_equal(gw[order], ow), f"scan {scan}: sharded map differs from the oracle"
    want = os.environ.get("BNX_SHARD_EXCHANGE", "p2p")
    assert sm.exchange_kind() == want, (sm.exchange_kind(), want)
    # the digest of the sharded map (combined over the ranks on the device side) equals the digest of the oracle's dump
    dig = sm.digest()
    if rank == 0:
        assert dig == capi.digest_of_dump(*om.dump()), "digest of the sharded map differs from the oracle's"
    st = sm.stats()
    if os.environ.get("BNX_EXPECT_REPLAY") == "1":
        assert st["replays"] + st["sync_retries"] > 0, st
    kind = sm.exchange_kind()
    dist.barrier()
    sm.close()
    if rank == 0:
        print("SHARDED_OK", world, kind, st)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
