"""Synthetic sensor workloads: deterministic numpy generators shared by tests and bench.

All generators return float32 point clouds already transformed to the world frame in float
(as the reference's ROS caller does, bonxai_ros/src/bonxai_server.cpp:167-171) plus the float32
sensor origin. Randomness is counter based (splitmix64 of (seed, scan, index)), so any scan can be
generated independently.
"""
from __future__ import annotations

import numpy as np

_U64 = np.uint64


def splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = (x.astype(np.uint64) + _U64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> _U64(30))) * _U64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> _U64(27))) * _U64(0x94D049BB133111EB)
        return z ^ (z >> _U64(31))


def uniform01(seed: int, stream: int, idx: np.ndarray) -> np.ndarray:
    """U[0,1) doubles from (seed, stream, idx), 53 bits."""
    with np.errstate(over="ignore"):
        key = splitmix64(np.asarray(idx, dtype=np.uint64)
                         + _U64(seed) * _U64(0xD1342543DE82EF95)
                         + _U64(stream) * _U64(0xA24BAED4963EE407))
    return (key >> _U64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def _ray_exit_box(o, d, lo, hi):
    """distance along d at which a ray starting INSIDE the box [lo,hi] leaves it."""
    with np.errstate(divide="ignore", invalid="ignore"):
        t1 = (lo - o) / d
        t2 = (hi - o) / d
    t = np.where(d > 0, t2, np.where(d < 0, t1, np.inf))
    return t.min(axis=1)


def _ray_enter_boxes(o, d, boxes, t_best):
    """nearest entry distance into any of the axis-aligned boxes [(lo3, hi3), ...] (rays start outside)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / d
    for lo, hi in boxes:
        ta = (lo - o) * inv
        tb = (hi - o) * inv
        tmin = np.nanmax(np.minimum(ta, tb), axis=1)
        tmax = np.nanmin(np.maximum(ta, tb), axis=1)
        hit = (tmax >= np.maximum(tmin, 0.0)) & (tmin > 1e-6)
        t_best = np.where(hit & (tmin < t_best), tmin, t_best)
    return t_best


# --------------------------------------------------------------------------------------------
# config #1 stand-in: room_synth (the reference's data/room_scan.pcd is a Git-LFS stub)
# --------------------------------------------------------------------------------------------
def room_synth(n: int = 50_000, seed: int = 1234):
    i = np.arange(n, dtype=np.float64)
    golden = np.pi * (3.0 - np.sqrt(5.0))
    zc = 1.0 - 2.0 * (i + 0.5) / n
    r = np.sqrt(np.maximum(0.0, 1.0 - zc * zc))
    d = np.stack([r * np.cos(golden * i), r * np.sin(golden * i), zc], axis=1)
    o = np.zeros(3)
    t = _ray_exit_box(o, d, np.array([-3.0, -2.5, -1.4]), np.array([3.0, 2.5, 1.4]))
    t = t + (uniform01(seed, 0, np.arange(n)) * 0.02 - 0.01)
    pts = (d * t[:, None]).astype(np.float32)
    return pts, np.zeros(3, np.float32)


# --------------------------------------------------------------------------------------------
# config #3: 64-beam LiDAR driving down a street
# --------------------------------------------------------------------------------------------
def _street_boxes(x_center: float, seed: int):
    """box 'buildings' hashed from the 40 m block index, on both sides of the street."""
    boxes = []
    b0 = int(np.floor((x_center - 140.0) / 40.0))
    b1 = int(np.floor((x_center + 140.0) / 40.0))
    for b in range(b0, b1 + 1):
        u = uniform01(seed, 101, np.arange(8) + 8 * (b + 100000))
        for side in (0, 1):
            if u[4 * side] < 0.25:
                continue  # empty lot
            x0 = 40.0 * b + 4.0 + 10.0 * u[4 * side + 1]
            w = 8.0 + 16.0 * u[4 * side + 2]
            h = 4.0 + 18.0 * u[4 * side + 3]
            if side == 0:
                y0, y1 = 2.5 + 3.0 * u[4 * side], 7.5
            else:
                y0, y1 = -7.5, -2.5 - 3.0 * u[4 * side]
            boxes.append((np.array([x0, y0, 0.0]), np.array([x0 + w, y1, h])))
    return boxes


def lidar_scan(scan: int, beams: int = 64, azimuths: int = 2048, seed: int = 7, speed: float = 1.0,
               stride4: bool = True, path: str = "line", index_range=None):
    """One beams x azimuths scan (131,072 points by default), beam-major point order.
    Returns (points float32 (n,4|3), origin float32 (3,)). index_range=(lo, hi) generates only the points
    [lo, hi) of the scan (every point depends on its own index only), e.g. one rank's slice."""
    if path == "line":
        ox, oy, yaw = speed * scan, 0.0, 0.0
    else:  # serpentine city path for the multi-GPU config: long parallel streets 60 m apart
        leg = 2000.0
        s = speed * scan
        k = int(s // leg)
        f = s - k * leg
        ox = f if k % 2 == 0 else leg - f
        oy = 60.0 * k
        yaw = 0.0 if k % 2 == 0 else np.pi
    o = np.array([ox, oy, 1.8])
    lo, hi = (0, beams * azimuths) if index_range is None else index_range
    idx = np.arange(lo, hi, dtype=np.int64)
    el = np.deg2rad(np.linspace(-24.8, 2.0, beams))[idx // azimuths]
    az = yaw + (idx % azimuths) * (2.0 * np.pi / azimuths)
    ce, se = np.cos(el), np.sin(el)
    d = np.stack([ce * np.cos(az), ce * np.sin(az), se], axis=1)
    n = d.shape[0]
    t = np.full(n, 120.0)  # no return -> beyond max_range -> truncated miss ray
    with np.errstate(divide="ignore", invalid="ignore"):
        tg = np.where(d[:, 2] < 0, (0.0 - o[2]) / d[:, 2], np.inf)          # ground z = 0
        yw = oy + 8.0 * np.sign(d[:, 1])
        tw = np.where(d[:, 1] != 0, (yw - o[1]) / d[:, 1], np.inf)           # street walls y = oy +- 8
    wall_z = o[2] + tw * d[:, 2]
    tw = np.where((tw > 0) & (wall_z >= 0.0) & (wall_z <= 12.0), tw, np.inf)
    t = np.minimum(t, np.minimum(tg, tw))
    boxes = [(b0 + np.array([0.0, oy, 0.0]), b1 + np.array([0.0, oy, 0.0])) for b0, b1 in _street_boxes(ox, seed)]
    t = _ray_enter_boxes(o, d, boxes, t)
    noise = uniform01(seed, 1 + scan, idx) * 0.04 - 0.02
    t = np.where(t < 120.0, t + noise, t)
    p = (o[None, :] + d * t[:, None]).astype(np.float32)
    if stride4:
        p = np.concatenate([p, np.zeros((n, 1), np.float32)], axis=1)
    return np.ascontiguousarray(p), o.astype(np.float32)


# --------------------------------------------------------------------------------------------
# config #4: dense depth camera inside a room
# --------------------------------------------------------------------------------------------
def _room_boxes(seed: int, count: int = 20):
    u = uniform01(seed, 202, np.arange(count * 6)).reshape(count, 6)
    boxes = []
    for k in range(count):
        c = np.array([-3.6 + 7.2 * u[k, 0], -2.6 + 5.2 * u[k, 1], 0.0])
        s = np.array([0.2 + 0.8 * u[k, 2], 0.2 + 0.8 * u[k, 3], 0.3 + 1.5 * u[k, 4]])
        lo = np.array([c[0] - s[0] / 2, c[1] - s[1] / 2, 0.0])
        hi = np.array([c[0] + s[0] / 2, c[1] + s[1] / 2, s[2]])
        if np.hypot(c[0], c[1]) < 1.9:  # keep the camera circle (radius 1.2) clear
            continue
        boxes.append((lo, hi))
    return boxes


def depth_scan(scan: int, width: int = 1280, height: int = 800, seed: int = 11, poses: int = 40,
               stride4: bool = True):
    """One 1280x800 pinhole depth frame (1,024,000 points), 87x58 deg FOV, room 8x6x3 m."""
    ang = 2.0 * np.pi * (scan % poses) / poses
    o = np.array([1.2 * np.cos(ang), 1.2 * np.sin(ang), 1.4])
    yaw = ang + 0.5 * np.pi + 0.35  # looks roughly along the circle tangent, slightly outward
    fwd = np.array([np.cos(yaw), np.sin(yaw), -0.12])
    fwd /= np.linalg.norm(fwd)
    right = np.cross(fwd, np.array([0.0, 0.0, 1.0]))
    right /= np.linalg.norm(right)
    up = np.cross(right, fwd)
    tx = np.tan(np.deg2rad(87.0) / 2) * (2.0 * (np.arange(width) + 0.5) / width - 1.0)
    ty = np.tan(np.deg2rad(58.0) / 2) * (1.0 - 2.0 * (np.arange(height) + 0.5) / height)
    d = (fwd[None, None, :] + tx[None, :, None] * right[None, None, :] + ty[:, None, None] * up[None, None, :])
    d = d.reshape(-1, 3)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    n = d.shape[0]
    t = _ray_exit_box(o, d, np.array([-4.0, -3.0, 0.0]), np.array([4.0, 3.0, 3.0]))
    t = _ray_enter_boxes(o, d, _room_boxes(seed), t)
    t = t + (uniform01(seed, 1 + scan, np.arange(n)) * 0.004 - 0.002)
    p = (o[None, :] + d * t[:, None]).astype(np.float32)
    if stride4:
        p = np.concatenate([p, np.zeros((n, 1), np.float32)], axis=1)
    return np.ascontiguousarray(p), o.astype(np.float32)


# --------------------------------------------------------------------------------------------
# config #2: bulk VoxelGrid sweeps
# --------------------------------------------------------------------------------------------
def coherent_coords(n: int, order: str = "x") -> np.ndarray:
    """dense cube of side ceil(n^(1/3)) enumerated x-fastest ('x') or z-fastest ('z'), first n coords."""
    side = int(np.ceil(round(n ** (1.0 / 3.0), 9)))
    i = np.arange(n, dtype=np.int64)
    a, b, c = i % side, (i // side) % side, i // (side * side)
    half = side // 2
    if order == "x":
        xyz = np.stack([a, b, c], axis=1)
    else:
        xyz = np.stack([c, b, a], axis=1)
    return (xyz - half).astype(np.int32)


def random_coords(n: int, seed: int = 42) -> np.ndarray:
    """uniform int32 coords in a bounded cube of side ceil((2n)^(1/3)) centred at 0."""
    side = int(np.ceil((2.0 * n) ** (1.0 / 3.0)))
    idx = np.arange(n)
    xyz = np.stack([(uniform01(seed, k, idx) * side).astype(np.int64) - side // 2 for k in range(3)], axis=1)
    return xyz.astype(np.int32)


def sweep_values(n: int) -> np.ndarray:
    return (np.arange(n, dtype=np.int64) & 0xFFFF).astype(np.float32)
