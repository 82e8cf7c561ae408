"""The golden workloads: small, named, deterministic insertPointCloud sequences. Shared by make_golden.py (which
runs them through the compiled reference) and by the tests (which run them through the port / the GPU path)."""
import os

import numpy as np

from bonxai_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))


def digest(xyz: np.ndarray, words: np.ndarray) -> dict:
    """order-independent 2x64-bit digest of a (coord, value) dump + its size"""
    if len(xyz) == 0:
        return {"cells": 0, "sum": "0" * 16, "xor": "0" * 16}
    with np.errstate(over="ignore"):
        h = np.full(len(xyz), 0xCBF29CE484222325, np.uint64)
        for col in (xyz[:, 0], xyz[:, 1], xyz[:, 2], np.asarray(words).view(np.uint32)):
            h = (h ^ col.astype(np.uint32).astype(np.uint64)) * np.uint64(0x100000001B3)
            h ^= h >> np.uint64(29)
            h *= np.uint64(0x9E3779B97F4A7C15)
        s = int(np.add.reduce(h, dtype=np.uint64))
        x = int(np.bitwise_xor.reduce(h))
    return {"cells": int(len(xyz)), "sum": f"{s:016x}", "xor": f"{x:016x}"}


def _rand_scans(seed, res, n_scans, n_pts, spread, max_range, dtype):
    rng = np.random.default_rng(seed)
    scans = []
    for _ in range(n_scans):
        origin = rng.uniform(-2, 2, 3).astype(dtype)
        pts = (origin + rng.normal(0, spread, (n_pts, 3))).astype(dtype)
        pts[: n_pts // 8] = pts[0]
        scans.append((pts, origin, max_range))
    return res, scans


def workloads():
    """name -> (resolution, [(points, origin, max_range), ...])"""
    w = {}
    quirk = np.array([[1, 0, 0], [1, .05, 0], [0, 3, 0], [-.55, -.72, .33]], np.float32)
    w["quirk"] = (0.1, [(quirk, np.zeros(3, np.float32), 2.0)])
    a = np.array([[1.0, 0.02, 0.01]], np.float32)
    b = np.array([[0.0, 1.0, 0.0]], np.float32)
    w["stale_id"] = (0.1, [(p, np.zeros(3, np.float32), 10.0) for p in (a, b, b, a, a)])
    w["clamp"] = (0.1, [(np.array([[0.75, 0.31, -0.2]], np.float32), np.zeros(3, np.float32), 5.0)] * 12)
    apple = np.load(os.path.join(HERE, "apple_xyz_f32.npy"))
    w["apple_inf"] = (0.02, [(apple, np.zeros(3, np.float32), float("inf"))] * 3)
    w["apple_074"] = (0.02, [(apple, np.zeros(3, np.float32), 0.74)] * 3)
    room, o = synth.room_synth()
    w["room_inf"] = (0.02, [(room, o, float("inf"))])
    w["room_25"] = (0.02, [(room, o, 2.5)])
    w["rand_f32"] = _rand_scans(1, 0.1, 5, 3000, 4.0, 6.0, np.float32)
    w["rand_f64"] = _rand_scans(2, 0.037, 4, 2000, 1.0, float("inf"), np.float64)
    w["lidar_0_2"] = (0.1, [(*synth.lidar_scan(s), 50.0) for s in range(3)])
    w["depth_small"] = (0.01, [(*synth.depth_scan(s, width=320, height=200), 5.0) for s in range(2)])
    w["fleet_3x4"] = (0.1, [(pts, origin, 40.0) for step in fleet_steps() for pts, origin in step])
    return w


def fleet_steps(world: int = 3, steps: int = 4):
    """[[(points, origin) of sensor 0..world-1] for every step]: three vehicles on parallel streets 150 m apart (reach 40 m:
    no cell is shared), sensor 1 standing still. Inserted one after the other — sensor 0, 1, 2, then the next step — this
    is the sequence a FLEET step of a sharded map must reproduce (bnx_map_shard_set_fleet); the golden digests after
    every insert come from the unmodified reference."""
    out = []
    for step in range(steps):
        scans = []
        for s in range(world):
            pts, origin = synth.lidar_scan(step if s != 1 else 0, beams=16, azimuths=512, seed=7 + s)
            shift = np.float32([0.0, 150.0 * s, 0.0])
            scans.append((np.ascontiguousarray(pts[:, :3] + shift), origin + shift))
        out.append(scans)
    return out


def run(make_map, name_filter=None):
    """-> {name: [digest after each scan]} using make_map(resolution) -> object with insert()/dump()"""
    out = {}
    for name, (res, scans) in workloads().items():
        if name_filter and name not in name_filter:
            continue
        m = make_map(res)
        out[name] = []
        for pts, origin, max_range in scans:
            m.insert(pts, origin, max_range)
            out[name].append(digest(*m.dump(sort=False)))
    return out
