"""CPU model of the sharded insertPointCloud protocol, one process per rank over gloo.

It restates, in numpy, what the CUDA stages do on each rank — classify + local dedupe, exchange 1 (endpoint
records to the owner of their root), owner-side lowest-global-index / stale verdict, ray casting by the endpoint's
owner, exchange 2 (ray cells to the owner of *their* root), endpoint apply before ray apply — so that the
protocol itself (who decides what, in which order, with which two exchanges) is checked against the CPU oracle
without a GPU; then FLEET steps (every rank inserts the scan of its own sensor): the sender of an endpoint
record names its sensor, rays start at that sensor's origin, every touched cell is stamped with the update id the
sensor's scan would have had in a sequence of consecutive inserts. Run by tests/test_shard_protocol_cpu.py with
world_size 2 and 3.

Two flavours of the exchanges (BNX_MODEL_EXCHANGE):
  gather    an all_gather_object of the per-destination buckets (gloo has no all_to_all)
  mailbox   the peer-memory protocol of the CUDA path, with POSIX shared memory standing in for NVLink-mapped device
            memory: every rank owns a mailbox [stamps | inbox [world][cap]], a sender stores its records straight into
            block [rank] of the OWNER's inbox, then the count into the block header, then its arrival stamp; the
            receiver spins on the stamps of all senders and reads the blocks as one dense range. The inboxes are single
            buffered: what makes that safe is the flag exchange at the end of every scan (nobody starts scan k+1 before
            every rank has consumed scan k), exactly as in DESIGN.md §7. gloo is only used to hand the segment names
            around and to collect the result.
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


class Mailboxes:
    """one mailbox per rank in shared memory; peers store into it one-sidedly"""
    KINDS, COLS = 3, 4  # exchange 1, exchange 2, flags; a record = 4 x int64

    def __init__(self, cap: int):
        from multiprocessing import shared_memory
        self.world, self.rank, self.cap = dist.get_world_size(), dist.get_rank(), cap
        self.words = self.KINDS * self.world + 2 * self.world * (1 + cap * self.COLS)  # stamps + 2 inboxes with headers
        name = f"bnx_mbox_{os.environ.get('MASTER_PORT', '0')}_{os.getpid()}"
        self.mine = shared_memory.SharedMemory(create=True, size=self.words * 8, name=name)
        np.ndarray((self.words,), np.int64, self.mine.buf)[:] = 0
        names = [None] * self.world
        dist.all_gather_object(names, name)  # the "IPC handles"
        self.segs = [self.mine if r == self.rank else shared_memory.SharedMemory(name=names[r]) for r in range(self.world)]
        self.box = [np.ndarray((self.words,), np.int64, seg.buf) for seg in self.segs]
        self.seq = [0] * self.KINDS
        dist.barrier()

    def _block(self, owner: int, inbox: int, src: int) -> slice:
        base = self.KINDS * self.world + (inbox * self.world + src) * (1 + self.cap * self.COLS)
        return slice(base, base + 1 + self.cap * self.COLS)

    def exchange(self, kind: int, buckets):
        """buckets[o]: int64 (n, 4) records for owner o. Returns the records of all senders as one dense array."""
        self.seq[kind] += 1
        for o, rec in enumerate(buckets):
            rec = np.asarray(rec, np.int64).reshape(-1, self.COLS)
            assert len(rec) <= self.cap, "inbox overflow"
            blk = self.box[o][self._block(o, kind, self.rank)]
            blk[1:1 + rec.size] = rec.ravel()          # records first
            blk[0] = len(rec)                           # then the header
            self.box[o][kind * self.world + self.rank] = self.seq[kind]  # then the arrival stamp
        mine = self.box[self.rank]
        import time
        t0 = time.time()
        while any(mine[kind * self.world + src] < self.seq[kind] for src in range(self.world)):
            assert time.time() - t0 < 60, "a peer never arrived"
            time.sleep(0)
        out = []
        for src in range(self.world):  # dense range over the per-sender blocks
            blk = mine[self._block(self.rank, kind, src)]
            out.append(blk[1:1 + int(blk[0]) * self.COLS].reshape(-1, self.COLS).copy())
        return out

    def flags(self, mine_flag: int) -> int:
        """the end-of-scan flag exchange: also the barrier that makes the single-buffered inboxes safe"""
        kind = 2
        self.seq[kind] += 1
        for o in range(self.world):  # value and stamp share one word here (the CUDA path stores {flags, stamp} 16 bytes)
            self.box[o][kind * self.world + self.rank] = self.seq[kind] * 4 + mine_flag
        me = self.box[self.rank]
        import time
        t0 = time.time()
        while any(me[kind * self.world + src] // 4 < self.seq[kind] for src in range(self.world)):
            assert time.time() - t0 < 60, "a peer never arrived"
            time.sleep(0)
        return max(int(me[kind * self.world + src]) % 4 for src in range(self.world))

    def close(self):
        dist.barrier()
        self.box = None
        for r, seg in enumerate(self.segs):
            seg.close()
        self.mine.unlink()


MBOX = None


def exchange(buckets, kind=0):
    """all-to-all: buckets[o] goes to rank o; returns what every rank sent to me"""
    world, rank = dist.get_world_size(), dist.get_rank()
    if MBOX is not None:
        packed = []
        for b in buckets:
            if isinstance(b, tuple):  # endpoint records: (cells (n,3), prio (n,))
                packed.append(np.concatenate([b[0].astype(np.int64), b[1].astype(np.int64)[:, None]], axis=1))
            else:                     # ray cells (n,3) or, in a fleet step, (n,4): cell + sensor
                b = np.asarray(b).reshape(len(b), -1)
                pad = np.zeros((len(b), 4 - b.shape[1]), np.int64)
                packed.append(np.concatenate([b.astype(np.int64), pad], axis=1))
        got = MBOX.exchange(kind, packed)
        if isinstance(buckets[0], tuple):
            return [(g[:, :3].astype(np.int32), g[:, 3]) for g in got]
        width = max([np.asarray(b).reshape(len(b), -1).shape[1] for b in buckets if len(b)] + [3])
        return [g[:, :width].astype(np.int32) for g in got]
    gathered = [None] * world
    dist.all_gather_object(gathered, buckets)
    return [gathered[src][rank] for src in range(world)]


def scan_c(c: int, sensor: int) -> int:
    """update id of the sensor-th consecutive insert starting at id c (probabilistic_map.cpp:103-105)"""
    return (c - 1 + sensor) % 3 + 1


def sharded_insert(shard: dict, c: int, pts_local, index_base, origin, max_range, res, fleet=None):
    """fleet = None: one scan, this rank holds the points [index_base, ...) of it. fleet = origins of ALL sensors: a fleet
    step, this rank holds the whole scan of sensor `rank`; the sender of an endpoint record (= the inbox block it arrives
    in) names its sensor, whose origin the ray starts at and whose update id stamps everything the scan touches."""
    world, rank = dist.get_world_size(), dist.get_rank()
    if fleet is not None:
        origin = fleet[rank]
    # ---- stage A: classify my slice, keep the lowest LOCAL index per voxel, send to the root's owner
    cells, miss = classify(pts_local, origin, max_range, res)
    _, first = np.unique(cells, axis=0, return_index=True)
    first.sort()
    own = owner_of(cells[first], world)
    rec = [(cells[first][own == o], (index_base + first[own == o]) * 2 + miss[first][own == o]) for o in range(world)]
    got = exchange(rec, 0)
    # ---- stage B (owner): lowest GLOBAL index wins, stale test against MY shard, rays only for fresh endpoints
    ecells = np.concatenate([g[0] for g in got])
    eprio = np.concatenate([g[1] for g in got])
    esrc = np.concatenate([np.full(len(g[1]), src, np.int64) for src, g in enumerate(got)])  # which rank sent it = its sensor
    order = np.argsort(eprio, kind="stable")
    ecells, eprio, esrc = ecells[order], eprio[order], esrc[order]
    _, keep = np.unique(ecells, axis=0, return_index=True)
    origins = fleet if fleet is not None else [origin] * world
    Os = [np.floor(np.asarray(o, np.float64) * (1.0 / res)).astype(np.int64) for o in origins]
    endpoints, out_cells = [], [[] for _ in range(world)]
    for i in sorted(keep):
        key = tuple(int(v) for v in ecells[i])
        sensor = int(esrc[i]) if fleet is not None else 0
        ci = scan_c(c, sensor)
        if (shard.get(key, 0) & 0xF) == ci:
            continue  # stale update_id: skipped AND no ray is cast
        endpoints.append((key, int(eprio[i]) & 1, ci))
        rc = ray_cells(Os[sensor], ecells[i])
        if len(rc):
            ro = owner_of(rc, world)
            tagged = np.concatenate([rc, np.full((len(rc), 1), sensor, np.int32)], axis=1)  # the sensor travels with the cell
            for o in range(world):
                out_cells[o].append(tagged[ro == o])
    send = [np.unique(np.concatenate(b), axis=0) if b else np.zeros((0, 4), np.int32) for b in out_cells]
    got = exchange(send, 1)
    # ---- stage C (owner): endpoints first, then every ray cell whose id is not its scan's
    for key, is_miss, ci in endpoints:
        w = shard.get(key, 0)
        p = (w >> 4) if w < 2**31 else ((w - 2**32) >> 4)
        p = max(p + MISS, CMIN) if is_miss else min(p + HIT, CMAX)
        shard[key] = ((p << 4) | ci) & 0xFFFFFFFF
    for block in got:
        for cell in block:
            key = (int(cell[0]), int(cell[1]), int(cell[2]))
            ci = scan_c(c, int(cell[3]))
            w = shard.get(key, 0)
            if (w & 0xF) != ci:
                p = (w >> 4) if w < 2**31 else ((w - 2**32) >> 4)
                shard[key] = ((max(p + MISS, CMIN) << 4) | ci) & 0xFFFFFFFF
    if MBOX is not None:
        assert MBOX.flags(0) == 0  # every rank has consumed both inboxes of this scan: they may be overwritten now
    return scan_c(c, world if fleet is not None else 1)


def main():
    global MBOX
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    if os.environ.get("BNX_MODEL_EXCHANGE", "gather") == "mailbox":
        MBOX = Mailboxes(cap=1 << 16)
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
    # ---- fleet steps: every rank holds the scan of its OWN sensor (parallel streets, reach not overlapping); the result
    # must equal the sensors' scans inserted one after the other, each with its own update id. Sensor 1 stands still,
    # so its endpoints turn stale when the id sequence comes round.
    for step in range(4):
        scans = []
        for v in range(world):
            pts, origin = synth.lidar_scan(step if v != 1 else 0, beams=8, azimuths=96, seed=7 + v)
            shift = np.float32([0.0, 40.0 * v, 0.0])  # max_range 12 m: streets 40 m apart never share a cell
            scans.append((np.ascontiguousarray(pts[:, :3] + shift), origin + shift))
        c = sharded_insert(shard, c, scans[rank][0], rank * len(scans[0][0]), None, max_range, res, fleet=[o for _, o in scans])
        gathered = [None] * world
        dist.all_gather_object(gathered, shard)
        if rank == 0:
            for pts, origin in scans:
                om.insert(pts, origin, max_range)
            union = {}
            for g in gathered:
                assert not (set(g) & set(union)), "shards overlap"
                union.update(g)
            xyz, words = om.dump()
            want = {tuple(int(v) for v in x): int(w) for x, w in zip(xyz, words)}
            assert union == want, f"fleet step {step}: sharded model differs from the oracle ({len(union)} vs {len(want)} cells)"
    dist.barrier()
    if MBOX is not None:
        MBOX.close()
    if rank == 0:
        print("SHARD_MODEL_OK", world, "mailbox" if MBOX is not None else "gather")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
