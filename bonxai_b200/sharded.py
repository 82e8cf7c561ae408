"""One ProbabilisticMap sharded over several GPUs by root key.

Each rank owns the roots with shard_owner(root) == rank and holds 1/world of every scan's points. The CUDA
stages live behind the C ABI (bnx_map_shard_*); this module only moves the staged device buffers between ranks:

    ShardedMap       one process per GPU; the library drives the scan itself (bnx_map_shard_insert): records go
                     straight into the owners' mailboxes over NVLink (peer memory, CUDA IPC; NCCL only bootstraps),
                     or through NCCL send/recv collectives with BNX_SHARD_EXCHANGE=nccl
    LocalShardGroup  all shards in ONE process on one GPU; exchange="transpose": the staged buffers are moved by
                     block transposes on the device, exchange="p2p": the mailbox kernels, with the raw pointers of
                     the other shards as "peers". Used by the single-GPU tests to run the whole protocol

Bit-exactness: the union of the shards' forEachCell dumps equals the unsharded map's dump after every scan.
"""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

from . import capi

REC_WORDS = 4     # endpoint record: int4 {x, y, z, global index << 1 | type}
LEAF_WORDS = 20   # leaf record: int4 {leaf origin xyz, 0} + u64 mask[8]


def split_points(n: int, world: int):
    """contiguous index ranges [lo, hi) per rank"""
    bounds = [(n * r) // world for r in range(world + 1)]
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


class _Shard:
    """a ProbabilisticMap shard + its staging buffers (torch CUDA tensors)"""

    def __init__(self, resolution: float, rank: int, world: int, device, cap_records: int = 1 << 15, cap_leaves: int = 1 << 13, p2p: bool = False):
        import torch
        self.torch = torch
        self.rank, self.world, self.device = rank, world, device
        self.map = capi.ProbabilisticMap(resolution)
        self.lib = self.map.lib
        capi._check(self.lib.bnx_map_shard_config(self.map.h, rank, world))
        self.cap_records, self.cap_leaves = 0, 0
        self.p2p = p2p
        self.mailbox = None
        if p2p:
            self.cap_records, self.cap_leaves = int(cap_records), int(cap_leaves)
            self.alloc_mailbox()
        else:
            self._alloc_records(cap_records)
            self._alloc_leaves(cap_leaves)
            self.flags = torch.zeros(4, dtype=torch.int32, device=device)
        self.n_local = 0

    # ---- peer-memory exchange: this shard's mailbox (bnx_map_shard_p2p_*)
    def alloc_mailbox(self):
        ptr = C.c_void_p(0)
        capi._check(self.lib.bnx_map_shard_p2p_alloc(self.map.h, C.c_int64(self.cap_records), C.c_int64(self.cap_leaves), None, C.byref(ptr)))
        self.mailbox = ptr.value

    def attach(self, mailboxes):
        arr = (C.c_void_p * self.world)(*mailboxes)
        capi._check(self.lib.bnx_map_shard_p2p_attach(self.map.h, None, arr))

    def _alloc_records(self, cap):
        t = self.torch
        self.cap_records = int(cap)
        self.send1 = t.empty((self.world, self.cap_records, REC_WORDS), dtype=t.int32, device=self.device)
        self.recv1 = t.empty_like(self.send1)

    def _alloc_leaves(self, cap):
        t = self.torch
        self.cap_leaves = int(cap)
        self.send2 = t.empty((self.world, self.cap_leaves, LEAF_WORDS), dtype=t.int32, device=self.device)
        self.recv2 = t.empty_like(self.send2)

    def set_stream(self, stream: int):
        self.map.set_stream(stream)

    # ---- the four stages (bnx_map_shard_*)
    def begin(self, pts, n, stride_bytes, f64, index_base, origin, max_range, fleet=None):
        if n + 2 > self.cap_records and not self.p2p:
            self._alloc_records(max(n + 2, self.cap_records * 2))
        o = np.ascontiguousarray(origin, dtype=np.float64)
        if fleet is not None:  # fleet step: this rank holds the scan of sensor `rank`, origins of all sensors given
            fo = np.ascontiguousarray(fleet, dtype=np.float64).reshape(self.world, 3)
            capi._check(self.lib.bnx_map_shard_set_fleet(self.map.h, C.c_void_p(fo.ctypes.data)))
        if isinstance(pts, capi.DevPtr):
            p, where = C.c_void_p(pts.address), capi.BNX_DEVICE
        else:
            arr = np.ascontiguousarray(pts)
            p, where = C.c_void_p(arr.ctypes.data), capi.BNX_HOST
        self.n_local = n
        send1 = None if self.p2p else C.c_void_p(self.send1.data_ptr())
        capi._check(self.lib.bnx_map_shard_begin(self.map.h, p, C.c_int64(stride_bytes), C.c_int64(n), int(bool(f64)), C.c_uint32(index_base),
                                                 C.c_void_p(o.ctypes.data), C.c_double(max_range), send1, C.c_int64(self.cap_records), where))

    def resolve_mark(self):
        if self.p2p:
            capi._check(self.lib.bnx_map_shard_resolve_mark(self.map.h, None, None, C.c_int64(0)))
        else:
            capi._check(self.lib.bnx_map_shard_resolve_mark(self.map.h, C.c_void_p(self.recv1.data_ptr()), C.c_void_p(self.send2.data_ptr()),
                                                            C.c_int64(self.cap_leaves)))

    def merge(self):
        if self.p2p:
            capi._check(self.lib.bnx_map_shard_merge(self.map.h, None, None))
        else:
            capi._check(self.lib.bnx_map_shard_merge(self.map.h, C.c_void_p(self.recv2.data_ptr()), C.c_void_p(self.flags.data_ptr())))

    def finish(self) -> int:
        retry = C.c_int(0)
        capi._check(self.lib.bnx_map_shard_finish(self.map.h, None if self.p2p else C.c_void_p(self.flags.data_ptr()), C.byref(retry)))
        return retry.value

    def grow_after(self, retry: int):
        if retry & (4 << 8):   # OVF_RECORDS cannot happen: cap_records >= n_local + 2
            raise RuntimeError("endpoint record exchange overflowed")
        if retry & (8 << 8):
            self._alloc_leaves(self.cap_leaves * 4)


def nccl_library_path():
    """torch's bundled NCCL (already loaded by the process once torch.distributed uses it), else the system one"""
    try:
        import torch
        cand = os.path.join(os.path.dirname(os.path.dirname(torch.__file__)), "nvidia", "nccl", "lib", "libnccl.so.2")
        if os.path.exists(cand):
            return cand
    except Exception:
        pass
    return "libnccl.so.2"


GATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64)  # bnx_allgather_fn (include/bonxai_b200.h)


def host_allgather_callback(group=None):
    """the bnx_allgather_fn a host program hands to bnx_map_shard_host_init, here over torch.distributed (any backend that
    moves CPU tensors, e.g. gloo): gathers `nbytes` bytes of host memory from every rank into recv[world][nbytes]"""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)

    def gather(ctx, send, recv, nbytes):
        try:
            mine = torch.frombuffer(bytearray(C.string_at(send, nbytes)), dtype=torch.uint8)
            parts = [torch.empty(nbytes, dtype=torch.uint8) for _ in range(world)]
            dist.all_gather(parts, mine, group=group)
            C.memmove(recv, b"".join(bytes(p.tolist()) for p in parts), nbytes * world)
            return 0
        except Exception as e:  # noqa: BLE001
            print("bootstrap all-gather failed:", repr(e), file=sys.stderr)
            return 1

    return GATHER_FN(gather)


class ShardedMap:
    """one rank of a map sharded over torch.distributed ranks (one process per GPU). torch.distributed is only used to
    hand the NCCL unique id around; the per-scan exchanges are issued by the library itself (bnx_map_shard_insert)."""

    def __init__(self, resolution: float, group=None, bootstrap: str = "nccl"):
        """bootstrap="nccl": the mailbox handles travel through one NCCL all-gather issued by the library (needs one GPU
        per rank); bootstrap="host": through torch.distributed.all_gather on `group` (any backend, e.g. gloo) handed to
        the library as a callback (bnx_map_shard_host_init) — ranks may then share a GPU."""
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        assert self.world > 1, "use capi.ProbabilisticMap for a single GPU"
        self.map = capi.ProbabilisticMap(resolution)
        self.lib = self.map.lib
        self.map.set_stream(torch.cuda.current_stream().cuda_stream)
        if bootstrap == "host":
            self._gather_cb = host_allgather_callback(group)  # keep alive as long as the map
            capi._check(self.lib.bnx_map_shard_host_init(self.map.h, self.rank, self.world, self._gather_cb, None))
            return
        path = nccl_library_path().encode()
        uid = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            buf = (C.c_uint8 * 128)()
            capi._check(self.lib.bnx_nccl_unique_id(path, buf))
            uid = torch.tensor(list(buf), dtype=torch.uint8)
        uid = uid.cuda()
        dist.broadcast(uid, src=0, group=group)
        raw = bytes(uid.cpu().tolist())
        capi._check(self.lib.bnx_map_shard_comm_init(self.map.h, path, raw, self.rank, self.world))

    def insert(self, pts_local, n_local, stride_bytes, index_base, n_max, origin, max_range, f64=False, use_async=False):
        """this rank's slice of the scan: points [index_base, index_base + n_local) of the global cloud; n_max = the
        largest slice over all ranks (every rank must pass the same value)"""
        o = np.ascontiguousarray(origin, dtype=np.float64)
        if isinstance(pts_local, capi.DevPtr):
            p, where = C.c_void_p(pts_local.address), capi.BNX_DEVICE
        else:
            assert pts_local.flags["C_CONTIGUOUS"]
            p, where = C.c_void_p(pts_local.ctypes.data), capi.BNX_HOST
        capi._check(self.lib.bnx_map_shard_insert(self.map.h, p, C.c_int64(stride_bytes), C.c_int64(n_local), int(bool(f64)), C.c_uint32(index_base),
                                                  C.c_int64(n_max), C.c_void_p(o.ctypes.data), C.c_double(max_range), where, int(use_async)))

    def insert_fleet(self, pts_local, n_local, stride_bytes, n_max, origins, max_range, f64=False, use_async=False):
        """fleet step: this rank holds the WHOLE scan of sensor `rank`; origins = (world, 3) origins of all sensors (the
        same array on every rank); n_max = the largest scan of the step. Result = the scans of sensors 0..world-1
        inserted one after the other; needs sensors whose reach does not overlap (else BonxaiError UNSUPPORTED)."""
        fo = np.ascontiguousarray(origins, dtype=np.float64).reshape(self.world, 3)
        capi._check(self.lib.bnx_map_shard_set_fleet(self.map.h, C.c_void_p(fo.ctypes.data)))
        self.insert(pts_local, n_local, stride_bytes, self.rank * n_max, n_max, fo[self.rank], max_range, f64=f64, use_async=use_async)

    def close(self):
        """releases the shard now, at a point of the program every rank reaches (not whenever the garbage collector runs)"""
        self.map.close()

    def sync(self):
        """completes the pipelined scans; collective: every rank must call it at the same point"""
        self.map.sync()

    def exchange_kind(self) -> str:
        """how the per-scan records travel: "p2p" (mailboxes in peer memory; after the first insert) or "nccl" """
        kind = C.c_int(0)
        capi._check(self.lib.bnx_map_shard_exchange(self.map.h, C.byref(kind)))
        return {0: "caller", 1: "nccl", 2: "p2p"}[kind.value]

    def counters(self):
        return self.map.counters()

    def totals(self):
        return self.map.totals()

    def stats(self):
        """what the sharded pipeline really did (bnx_map_shard_stats)"""
        return self.map.shard_stats()

    def digest(self):
        """digest of the WHOLE map (all shards): collective — (sum, xor, count) combined over the ranks"""
        s, x, c = self.map.digest()
        # 64-bit values travel as pairs of 32-bit halves (no unsigned 64-bit reductions in torch.distributed)
        t = self.torch.tensor([s & 0xFFFFFFFF, s >> 32, x & 0xFFFFFFFF, x >> 32, c], dtype=self.torch.int64)
        if self.dist.get_backend(self.group) == "nccl":
            t = t.cuda()
        parts = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(parts, t, group=self.group)
        return capi.combine_digests([(int(p[0]) | (int(p[1]) << 32), int(p[2]) | (int(p[3]) << 32), int(p[4])) for p in parts])


class LocalShardGroup:
    """all `world` shards of one map inside ONE process on one GPU: the exchanges are device-side block
    transposes. Same stages, kernels and record formats as ShardedMap; no NCCL needed."""

    def __init__(self, resolution: float, world: int, device="cuda:0", cap_leaves: int = 1 << 13, exchange: str = "transpose",
                 cap_records: int = 1 << 15):
        import torch
        assert exchange in ("transpose", "p2p")
        self.torch = torch
        self.world = world
        self.p2p = exchange == "p2p"
        self.shards = [_Shard(resolution, r, world, torch.device(device), cap_records=cap_records, cap_leaves=cap_leaves, p2p=self.p2p)
                       for r in range(world)]
        # one stream for all shards: every producer of a stage is enqueued before any consumer of the next stage
        stream = torch.cuda.current_stream().cuda_stream
        for s in self.shards:
            s.set_stream(stream)
        if self.p2p:
            self._attach_all()
        self.attempts = 0

    def _attach_all(self):
        boxes = [s.mailbox for s in self.shards]
        for s in self.shards:
            s.attach(boxes)

    def _regrow_mailboxes(self, cap_records=None, cap_leaves=None):
        for s in self.shards:
            s.cap_records = int(cap_records or s.cap_records)
            s.cap_leaves = int(cap_leaves or s.cap_leaves)
            s.alloc_mailbox()
        self._attach_all()

    def insert(self, pts: np.ndarray, origin, max_range):
        pts = np.ascontiguousarray(pts)
        f64 = pts.dtype == np.float64
        stride = pts.shape[1] * pts.dtype.itemsize
        parts = split_points(len(pts), self.world)
        need = max(hi - lo for lo, hi in parts) + 2
        if self.p2p:
            if need > self.shards[0].cap_records:
                self._regrow_mailboxes(cap_records=max(need, self.shards[0].cap_records * 2))
        else:
            for s in self.shards:
                if need > s.cap_records:
                    s._alloc_records(max(need, s.cap_records * 2))
        while True:
            for s, (lo, hi) in zip(self.shards, parts):
                s.begin(pts[lo:hi], hi - lo, stride, f64, lo, origin, max_range)
            if not self.p2p:
                self._exchange("send1", "recv1")
            restart = False
            while not restart:
                self.attempts += 1
                for s in self.shards:
                    s.resolve_mark()
                if not self.p2p:
                    self._exchange("send2", "recv2")
                for s in self.shards:
                    s.merge()
                if not self.p2p:
                    flags = self.torch.stack([s.flags for s in self.shards]).max(dim=0).values
                    for s in self.shards:
                        s.flags.copy_(flags)
                retries = [s.finish() for s in self.shards]
                assert len(set(retries)) == 1
                if not retries[0]:
                    return
                if self.p2p:
                    if retries[0] & (4 << 8):
                        raise RuntimeError("endpoint record exchange overflowed")
                    if retries[0] & (8 << 8):
                        # larger mailboxes replace the old ones, and with them the received endpoint records:
                        # the scan starts over (a failed attempt has changed nothing)
                        self._regrow_mailboxes(cap_leaves=self.shards[0].cap_leaves * 4)
                        restart = True
                else:
                    for s in self.shards:
                        s.grow_after(retries[0])

    def insert_fleet(self, scans, max_range):
        """fleet step: scans = [(points of sensor s, origin of sensor s)] * world, shard s holds scan s. Equal to
        inserting the scans one after the other (sensor 0 first) when the sensors' reach does not overlap."""
        assert len(scans) == self.world
        pts = [np.ascontiguousarray(p) for p, _ in scans]
        origins = np.array([np.asarray(o, np.float64) for _, o in scans])
        f64 = pts[0].dtype == np.float64
        stride = pts[0].shape[1] * pts[0].dtype.itemsize
        n_max = max(len(p) for p in pts)
        need = n_max + 2
        if self.p2p:
            if need > self.shards[0].cap_records:
                self._regrow_mailboxes(cap_records=max(need, self.shards[0].cap_records * 2))
        else:
            for s in self.shards:
                if need > s.cap_records:
                    s._alloc_records(max(need, s.cap_records * 2))
        while True:
            for r, s in enumerate(self.shards):
                s.begin(pts[r], len(pts[r]), stride, f64, r * n_max, origins[r], max_range, fleet=origins)
            if not self.p2p:
                self._exchange("send1", "recv1")
            restart = False
            while not restart:
                self.attempts += 1
                for s in self.shards:
                    s.resolve_mark()
                if not self.p2p:
                    self._exchange("send2", "recv2")
                for s in self.shards:
                    s.merge()
                if not self.p2p:
                    flags = self.torch.stack([s.flags for s in self.shards]).max(dim=0).values
                    for s in self.shards:
                        s.flags.copy_(flags)
                retries = [s.finish() for s in self.shards]
                assert len(set(retries)) == 1
                if not retries[0]:
                    return
                if self.p2p:
                    if retries[0] & (8 << 8):
                        self._regrow_mailboxes(cap_leaves=self.shards[0].cap_leaves * 4)
                        restart = True
                else:
                    for s in self.shards:
                        s.grow_after(retries[0])

    def _exchange(self, send, recv):
        # all-to-all: block o of rank r's send buffer becomes block r of rank o's receive buffer
        for r, s in enumerate(self.shards):
            for o, d in enumerate(self.shards):
                getattr(d, recv)[r].copy_(getattr(s, send)[o])

    def dump(self, sort=True):
        xs, ws = zip(*[s.map.dump(sort=False) for s in self.shards])
        xyz, w = np.concatenate(xs), np.concatenate(ws)
        if sort and len(xyz):
            order = np.lexsort((xyz[:, 2], xyz[:, 1], xyz[:, 0]))
            xyz, w = xyz[order], w[order]
        return xyz, w

    def counters(self):
        cs = [s.map.counters() for s in self.shards]
        n = sum(c["N"] for c in cs)
        return dict(N=n, E=sum(c["E"] for c in cs), V=sum(c["V"] for c in cs) + n, U=sum(c["U"] for c in cs),
                    retries=max(c["retries"] for c in cs))
