"""CPU model of the sharded insertPointCloud protocol, one process per rank over gloo.

It restates, in numpy, what the CUDA stages do on each rank — classify + local dedupe, exchange 1 (endpoint
records to the owner of their root), owner-side lowest-global-index / stale verdict, ray casting by the endpoint's
owner, exchange 2 (ray cells to the owner of *their* root), endpoint apply before ray apply — so that the
protocol itself (who decides what, in which order, with which two exchanges) is checked against the CPU oracle
without a GPU. Run by tests/test_shard_protocol_cpu.py with world_size 2 and 3.

gloo has no all_to_all, so an exchange is an all_gather_object of the per-destination buckets.
"""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bonxai_b200 import synth  # noqa: E402
from bonxai_b200.sharded import split_points  # noqa: E402

MISS, HIT, CMIN, CMAX = -405465, 847297, -1992430, 3476099


def owner_of(cells: np.ndarray, world: int) -> np.ndarray:
    """any deterministic function of the root key (coord >> 5) works for the protocol"""
    r = (cells >> 5).astype(np.int64)
    h = (r[:, 0] * 73856093) ^ (r[:, 1] * 19349663) ^ (r[:, 2] * 83492791)
    return ((h ^ (h >> 13)) % world).astype(np.int64)


def classify(pts, origin, max_range, res):
    """probabilistic_map.hpp:146-158 + bonxai.hpp:404-410 in fp64, one rounding per operation"""
    p = pts[:, :3].astype(np.float64)
    o = origin.astype(np.float64)
    v = p - o
    sq = (v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1]) + v[:, 2] * v[:, 2]
    miss = sq >= max_range * max_range
    with np.errstate(invalid="ignore", divide="ignore"):
        trunc = o + (v / np.sqrt(sq)[:, None]) * max_range
    e = np.where(miss[:, None], trunc, p)
    return np.floor(e * (1.0 / res)).astype(np.int32), miss.astype(np.int64)


def ray_cells(O, E):
    d = E.astype(np.int64) - O
    m = int(np.abs(d).max())
    if m == 0:
        return np.zeros((0, 3), np.int32)
    k = np.arange(m)[:, None]
    return (O + np.sign(d) * ((2 * k * np.abs(d) + m) // (2 * m))).astype(np.int32)


def exchange(buckets):
    """all-to-all through all_gather_object: buckets[o] goes to rank o; returns what every rank sent to me"""
    world, rank = dist.get_world_size(), dist.get_rank()
    gathered = [None] * world
    dist.all_gather_object(gathered, buckets)
    return [gathered[src][rank] for src in range(world)]


def sharded_insert(shard: dict, c: int, pts_local, index_base, origin, max_range, res):
    world = dist.get_world_size()
    # ---- stage A: classify my slice, keep the lowest LOCAL index per voxel, send to the root's owner
    cells, miss = classify(pts_local, origin, max_range, res)
    _, first = np.unique(cells, axis=0, return_index=True)
    first.sort()
    own = owner_of(cells[first], world)
    rec = [(cells[first][own == o], (index_base + first[own == o]) * 2 + miss[first][own == o]) for o in range(world)]
    got = exchange(rec)
    # ---- stage B (owner): lowest GLOBAL index wins, stale test against MY shard, rays only for fresh endpoints
    ecells = np.concatenate([g[0] for g in got])
    eprio = np.concatenate([g[1] for g in got])
    order = np.argsort(eprio, kind="stable")
    ecells, eprio = ecells[order], eprio[order]
    _, keep = np.unique(ecells, axis=0, return_index=True)
    O = np.floor(origin.astype(np.float64) * (1.0 / res)).astype(np.int64)
    endpoints, out_cells = [], [[] for _ in range(world)]
    for i in sorted(keep):
        key = tuple(int(v) for v in ecells[i])
        if (shard.get(key, 0) & 0xF) == c:
            continue  # stale update_id: skipped AND no ray is cast
        endpoints.append((key, int(eprio[i]) & 1))
        rc = ray_cells(O, ecells[i])
        if len(rc):
            ro = owner_of(rc, world)
            for o in range(world):
                out_cells[o].append(rc[ro == o])
    send = [np.unique(np.concatenate(b), axis=0) if b else np.zeros((0, 3), np.int32) for b in out_cells]
    got = exchange(send)
    # ---- stage C (owner): endpoints first, then every ray cell whose id is not the current one
    for key, is_miss in endpoints:
        w = shard.get(key, 0)
        p = (w >> 4) if w < 2**31 else ((w - 2**32) >> 4)
        p = max(p + MISS, CMIN) if is_miss else min(p + HIT, CMAX)
        shard[key] = ((p << 4) | c) & 0xFFFFFFFF
    for block in got:
        for cell in block:
            key = (int(cell[0]), int(cell[1]), int(cell[2]))
            w = shard.get(key, 0)
            if (w & 0xF) != c:
                p = (w >> 4) if w < 2**31 else ((w - 2**32) >> 4)
                shard[key] = ((max(p + MISS, CMIN) << 4) | c) & 0xFFFFFFFF
    return 1 if c == 3 else c + 1


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    res, max_range = 0.25, 12.0
    shard, c = {}, 1
    scans = [synth.lidar_scan(s, beams=8, azimuths=96) for s in (0, 1, 1, 2, 0, 0)]  # repeats exercise the stale rule
    if rank == 0:
        import oracle
        om = oracle.load("port").map(res)
    for k, (pts, origin) in enumerate(scans):
        lo, hi = split_points(len(pts), world)[rank]
        c = sharded_insert(shard, c, pts[lo:hi], lo, origin, max_range, res)
        gathered = [None] * world
        dist.all_gather_object(gathered, shard)
        if rank == 0:
            om.insert(pts, origin, max_range)
            union = {}
            for g in gathered:
                assert not (set(g) & set(union)), "shards overlap"
                union.update(g)
            xyz, words = om.dump()
            want = {tuple(int(v) for v in x): int(w) for x, w in zip(xyz, words)}
            assert union == want, f"scan {k}: sharded model differs from the oracle ({len(union)} vs {len(want)} cells)"
    dist.barrier()
    if rank == 0:
        print("SHARD_MODEL_OK", world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
