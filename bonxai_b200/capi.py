"""ctypes binding of include/bonxai_b200.h (the C-ABI shared library built from bonxai_b200/csrc).

This is the host-side mirror used by the tests and bench.py; the drop-in for C++ callers is
include/bonxai/bonxai.hpp + include/bonxai_map/probabilistic_map.hpp. There is NO fallback: if the library
is missing, or there is no CUDA device, the calls raise.

Array arguments are numpy arrays (host memory, BNX_HOST) or `DevPtr(address)` wrappers around raw device
addresses, e.g. `DevPtr(tensor.data_ptr())` for a torch CUDA tensor (BNX_DEVICE).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
# BNX_LIB selects another build of the same library (tuning variants made by bonxai_b200.build.build_variant)
LIB_PATH = os.environ.get("BNX_LIB") or os.path.join(PKG, "libbonxai_b200.so")

BNX_HOST, BNX_DEVICE = 0, 1
BNX_CLEAR_MEMORY, BNX_SET_ALL_CELLS_OFF = 0, 1
BNX_OCCUPIED, BNX_UNKNOWN, BNX_FREE = 0, 1, 2
STATUS_NAMES = {0: "BNX_OK", 1: "BNX_ERR_INVALID", 2: "BNX_ERR_CUDA", 3: "BNX_ERR_NOMEM", 4: "BNX_ERR_CAPACITY",
                5: "BNX_ERR_UNSUPPORTED"}

# every symbol include/bonxai_b200.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "bnx_version", "bnx_last_error", "bnx_launch_count", "bnx_device_count", "bnx_host_alloc", "bnx_host_free",
    "bnx_grid_create", "bnx_grid_destroy", "bnx_grid_set_stream", "bnx_grid_sync", "bnx_grid_info",
    "bnx_grid_pos_to_coord", "bnx_grid_coord_to_pos", "bnx_grid_set_values", "bnx_grid_get_values",
    "bnx_grid_get_or_create", "bnx_grid_update_values", "bnx_grid_set_on", "bnx_grid_set_off", "bnx_grid_is_on",
    "bnx_grid_active_count", "bnx_grid_digest", "bnx_grid_dump", "bnx_grid_clear", "bnx_grid_release_unused", "bnx_grid_mem_usage",
    "bnx_grid_stats", "bnx_grid_serialize", "bnx_grid_deserialize",
    "bnx_map_create", "bnx_map_destroy", "bnx_map_set_stream", "bnx_map_sync", "bnx_map_grid", "bnx_map_set_options",
    "bnx_map_get_options", "bnx_map_insert_f32", "bnx_map_insert_f64", "bnx_map_add_hit", "bnx_map_add_miss",
    "bnx_map_query", "bnx_map_get_voxels", "bnx_map_get_voxel_points", "bnx_map_counters", "bnx_map_update_count",
    "bnx_map_set_profiling", "bnx_map_phase_times", "bnx_map_set_marking",
    "bnx_map_publish_occupied_f32", "bnx_map_insert_transformed_f32", "bnx_map_insert_async_f32", "bnx_map_insert_async_f64", "bnx_map_totals",
    "bnx_map_shard_config", "bnx_map_shard_begin", "bnx_map_shard_resolve_mark", "bnx_map_shard_merge", "bnx_map_shard_finish",
    "bnx_nccl_unique_id", "bnx_map_shard_comm_init", "bnx_map_shard_insert",
    "bnx_map_shard_p2p_alloc", "bnx_map_shard_p2p_attach", "bnx_map_shard_exchange", "bnx_map_shard_host_init", "bnx_map_shard_stats", "bnx_map_shard_set_fleet",
]


class BonxaiError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {message}")
        self.status = status


class DevPtr:
    """A raw device address (int) handed to the C ABI as BNX_DEVICE memory."""

    def __init__(self, address: int):
        self.address = int(address)


_lib = None


def load_library(path: str | None = None) -> C.CDLL:
    """Load the CUDA library. Raises if it has not been built: there is no CPU implementation to fall back to."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: build it with `python -m bonxai_b200.build` "
                          "(bonxai_b200 has no CPU fallback)")
    lib = C.CDLL(path)
    lib.bnx_last_error.restype = C.c_char_p
    for name in SYMBOLS:
        fn = getattr(lib, name)
        if name == "bnx_launch_count":
            fn.restype = C.c_int64
        elif name != "bnx_last_error":
            fn.restype = C.c_int
    _lib = lib
    return lib


def _check(status: int):
    if status != 0:
        raise BonxaiError(status, load_library().bnx_last_error().decode(errors="replace"))


def launch_count() -> int:
    return int(load_library().bnx_launch_count())


def device_count() -> int:
    n = C.c_int(0)
    _check(load_library().bnx_device_count(C.byref(n)))
    return n.value


def _arg(a, dtype=None, cols=None):
    """-> (c_void_p, where, keepalive)"""
    if a is None:
        return C.c_void_p(None), None, None
    if isinstance(a, DevPtr):
        return C.c_void_p(a.address), BNX_DEVICE, a
    arr = np.ascontiguousarray(a, dtype=dtype)
    if cols is not None:
        arr = arr.reshape(-1, cols)
    return C.c_void_p(arr.ctypes.data), BNX_HOST, arr


def _where(*ws):
    ws = [w for w in ws if w is not None]
    if not ws:
        return BNX_HOST
    if any(w != ws[0] for w in ws):
        raise ValueError("host and device buffers cannot be mixed in one call")
    return ws[0]


def _mix64(h):
    h = h ^ (h >> np.uint64(33))
    h = h * np.uint64(0xFF51AFD7ED558CCD)
    h = h ^ (h >> np.uint64(33))
    h = h * np.uint64(0xC4CEB9FE1A85EC53)
    return h ^ (h >> np.uint64(33))


def digest_of_dump(xyz, values):
    """bnx_grid_digest restated in numpy for a host dump (xyz int32 (n,3), values any fixed-size dtype):
    (sum, xor, count) over mix64(hash3(x,y,z) + FNV1a64(value bytes) * 0x9E3779B97F4A7C15), all mod 2^64"""
    xyz = np.ascontiguousarray(xyz, np.int32).reshape(-1, 3)
    vals = np.ascontiguousarray(values)
    n = len(xyz)
    if n == 0:
        return 0, 0, 0
    with np.errstate(over="ignore"):
        x, y, z = (xyz[:, k].view(np.uint32).astype(np.uint64) for k in range(3))
        h = x * np.uint64(0x9E3779B97F4A7C15)
        h = h ^ (y * np.uint64(0xC2B2AE3D27D4EB4F) + (h >> np.uint64(29)))
        h = h ^ (z * np.uint64(0x165667B19E3779F9) + (h << np.uint64(7)))
        h = _mix64(h)
        raw = vals.view(np.uint8).reshape(n, -1)
        f = np.full(n, 0xCBF29CE484222325, np.uint64)
        for k in range(raw.shape[1]):
            f = (f ^ raw[:, k].astype(np.uint64)) * np.uint64(0x100000001B3)
        d = _mix64(h + f * np.uint64(0x9E3779B97F4A7C15))
        return int(d.sum(dtype=np.uint64)), int(np.bitwise_xor.reduce(d)), n


def combine_digests(digests):
    """digest of the union of disjoint shards"""
    s = x = c = 0
    for a, b, n in digests:
        s = (s + a) & 0xFFFFFFFFFFFFFFFF
        x ^= b
        c += n
    return s, x, c


class VoxelGrid:
    """Bonxai::VoxelGrid<DataT> (bonxai_core/include/bonxai/bonxai.hpp:114-333) with batched accessor calls."""

    def __init__(self, voxel_size: float, inner_bits: int = 2, leaf_bits: int = 3, dtype=np.uint32, _handle=None, _owner=None):
        self.lib = load_library()
        self.dtype = np.dtype(dtype)
        self._owner = _owner
        if _handle is not None:
            self.h = _handle
        else:
            h = C.c_void_p()
            _check(self.lib.bnx_grid_create(C.c_double(voxel_size), int(inner_bits), int(leaf_bits), self.dtype.itemsize, C.byref(h)))
            self.h = h

    def __del__(self):
        if getattr(self, "h", None) and self._owner is None:
            self.lib.bnx_grid_destroy(self.h)
            self.h = None

    # ---- helpers
    def _vals(self, values, n):
        if isinstance(values, DevPtr):
            return values
        v = np.ascontiguousarray(values)
        if v.dtype.itemsize != self.dtype.itemsize:
            v = v.astype(self.dtype)
        assert v.size == n, "one value per coordinate"
        return v

    def info(self):
        vs, ib, lb, cb = C.c_double(), C.c_int(), C.c_int(), C.c_int()
        _check(self.lib.bnx_grid_info(self.h, C.byref(vs), C.byref(ib), C.byref(lb), C.byref(cb)))
        return dict(voxel_size=vs.value, inner_bits=ib.value, leaf_bits=lb.value, cell_bytes=cb.value)

    def set_stream(self, stream: int):
        _check(self.lib.bnx_grid_set_stream(self.h, C.c_void_p(stream)))

    def sync(self):
        _check(self.lib.bnx_grid_sync(self.h))

    def pos_to_coord(self, xyz):
        p, w, keep = _arg(xyz, np.float64, 3)
        out = np.empty((len(keep), 3), np.int32)
        _check(self.lib.bnx_grid_pos_to_coord(self.h, p, C.c_int64(len(keep)), C.c_void_p(out.ctypes.data), BNX_HOST))
        return out

    def coord_to_pos(self, xyz):
        p, w, keep = _arg(xyz, np.int32, 3)
        out = np.empty((len(keep), 3), np.float64)
        _check(self.lib.bnx_grid_coord_to_pos(self.h, p, C.c_int64(len(keep)), C.c_void_p(out.ctypes.data), BNX_HOST))
        return out

    # ---- batched accessor operations (host numpy in, numpy out)
    def set_values(self, xyz, values, n=None, was_on=None):
        px, wx, kx = _arg(xyz, np.int32, 3)
        n = len(kx) if n is None else n
        pv, wv, kv = _arg(self._vals(values, n))
        if wx == BNX_HOST:
            out = np.empty(n, np.uint8)
            _check(self.lib.bnx_grid_set_values(self.h, px, pv, C.c_int64(n), C.c_void_p(out.ctypes.data), _where(wx, wv)))
            return out.astype(bool)
        pw, ww, kw = _arg(was_on)
        _check(self.lib.bnx_grid_set_values(self.h, px, pv, C.c_int64(n), pw, _where(wx, wv, ww)))
        return None

    def get_values(self, xyz, n=None, values=None, found=None):
        px, wx, kx = _arg(xyz, np.int32, 3)
        n = len(kx) if n is None else n
        if wx == BNX_HOST:
            vals = np.zeros(n, self.dtype)
            fnd = np.empty(n, np.uint8)
            _check(self.lib.bnx_grid_get_values(self.h, px, C.c_int64(n), C.c_void_p(vals.ctypes.data), C.c_void_p(fnd.ctypes.data), BNX_HOST))
            return vals, fnd.astype(bool)
        pv, wv, kv = _arg(values)
        pf, wf, kf = _arg(found)
        _check(self.lib.bnx_grid_get_values(self.h, px, C.c_int64(n), pv, pf, BNX_DEVICE))
        return None

    def get_or_create(self, xyz):
        px, wx, kx = _arg(xyz, np.int32, 3)
        vals = np.zeros(len(kx), self.dtype)
        _check(self.lib.bnx_grid_get_or_create(self.h, px, C.c_int64(len(kx)), C.c_void_p(vals.ctypes.data), BNX_HOST))
        return vals

    def update_values(self, xyz, values):
        px, wx, kx = _arg(xyz, np.int32, 3)
        pv, wv, kv = _arg(self._vals(values, len(kx)))
        _check(self.lib.bnx_grid_update_values(self.h, px, pv, C.c_int64(len(kx)), BNX_HOST))

    def set_on(self, xyz, default_value=0):
        px, wx, kx = _arg(xyz, np.int32, 3)
        d = np.array([default_value]).astype(self.dtype) if not isinstance(default_value, np.ndarray) else default_value
        out = np.empty(len(kx), np.uint8)
        _check(self.lib.bnx_grid_set_on(self.h, px, C.c_int64(len(kx)), C.c_void_p(d.ctypes.data), C.c_void_p(out.ctypes.data), BNX_HOST))
        return out.astype(bool)

    def set_off(self, xyz):
        px, wx, kx = _arg(xyz, np.int32, 3)
        out = np.empty(len(kx), np.uint8)
        _check(self.lib.bnx_grid_set_off(self.h, px, C.c_int64(len(kx)), C.c_void_p(out.ctypes.data), BNX_HOST))
        return out.astype(bool)

    def is_on(self, xyz):
        px, wx, kx = _arg(xyz, np.int32, 3)
        out = np.empty(len(kx), np.uint8)
        _check(self.lib.bnx_grid_is_on(self.h, px, C.c_int64(len(kx)), C.c_void_p(out.ctypes.data), BNX_HOST))
        return out.astype(bool)

    # ---- whole grid
    def active_count(self) -> int:
        n = C.c_int64()
        _check(self.lib.bnx_grid_active_count(self.h, C.byref(n)))
        return n.value

    def digest(self):
        """order-independent digest (sum, xor, count) of all (coord, value) pairs, computed on the device
        (bnx_grid_digest); `digest_of_dump` below is the same function of a host dump"""
        out = (C.c_uint64 * 3)()
        _check(self.lib.bnx_grid_digest(self.h, out))
        return int(out[0]), int(out[1]), int(out[2])

    def dump(self, sort=True):
        """forEachCell as arrays: (xyz int32 (n,3), values). Sorted by (x,y,z) for comparison by default."""
        n = self.active_count()
        xyz = np.empty((n, 3), np.int32)
        vals = np.empty(n, self.dtype)
        cnt = C.c_int64()
        _check(self.lib.bnx_grid_dump(self.h, C.c_void_p(xyz.ctypes.data), C.c_void_p(vals.ctypes.data), C.c_int64(n), C.byref(cnt), BNX_HOST))
        assert cnt.value == n
        if sort and n:
            order = np.lexsort((xyz[:, 2], xyz[:, 1], xyz[:, 0]))
            xyz, vals = xyz[order], vals[order]
        return xyz, vals

    def dump_device(self, xyz: DevPtr, values: DevPtr | None, cap: int) -> int:
        cnt = C.c_int64()
        pv = C.c_void_p(values.address) if values is not None else C.c_void_p(None)
        _check(self.lib.bnx_grid_dump(self.h, C.c_void_p(xyz.address), pv, C.c_int64(cap), C.byref(cnt), BNX_DEVICE))
        return cnt.value

    def clear(self, option: int):
        _check(self.lib.bnx_grid_clear(self.h, int(option)))

    def release_unused(self):
        _check(self.lib.bnx_grid_release_unused(self.h))

    def mem_usage(self) -> int:
        n = C.c_int64()
        _check(self.lib.bnx_grid_mem_usage(self.h, C.byref(n)))
        return n.value

    def stats(self):
        a = (C.c_int64 * 8)()
        _check(self.lib.bnx_grid_stats(self.h, a))
        return dict(roots=a[0], inner=a[1], leaves=a[2], free_leaves=a[3], root_slots=a[4], leaf_capacity=a[5], mapped_bytes=a[6])

    def serialize(self, type_name: str) -> bytes:
        size = C.c_int64()
        _check(self.lib.bnx_grid_serialize(self.h, type_name.encode(), None, C.c_int64(0), C.byref(size)))
        buf = np.empty(size.value, np.uint8)
        _check(self.lib.bnx_grid_serialize(self.h, type_name.encode(), C.c_void_p(buf.ctypes.data), C.c_int64(size.value), C.byref(size)))
        return buf.tobytes()

    @classmethod
    def deserialize(cls, data: bytes, dtype, type_name: str):
        lib = load_library()
        buf = np.frombuffer(data, np.uint8)
        h = C.c_void_p()
        _check(lib.bnx_grid_deserialize(C.c_void_p(buf.ctypes.data), C.c_int64(len(buf)), np.dtype(dtype).itemsize, type_name.encode(), C.byref(h)))
        return cls(0.0, dtype=dtype, _handle=h)


class ProbabilisticMap:
    """Bonxai::ProbabilisticMap (bonxai_map/include/bonxai_map/probabilistic_map.hpp:27-139)."""

    def __init__(self, resolution: float):
        self.lib = load_library()
        h = C.c_void_p()
        _check(self.lib.bnx_map_create(C.c_double(resolution), C.byref(h)))
        self.h = h
        gh = C.c_void_p()
        _check(self.lib.bnx_map_grid(self.h, C.byref(gh)))
        self._grid = VoxelGrid(resolution, dtype=np.uint32, _handle=gh, _owner=self)

    def close(self):
        """destroys the map now (the object and its grid() wrapper reference each other, so without this the device
        memory is only released when Python's cycle collector gets to it)"""
        if getattr(self, "h", None):
            self.lib.bnx_map_destroy(self.h)
            self.h = None
            self._grid.h = None

    def __del__(self):
        self.close()

    def grid(self) -> VoxelGrid:
        return self._grid

    def set_stream(self, stream: int):
        _check(self.lib.bnx_map_set_stream(self.h, C.c_void_p(stream)))

    def sync(self):
        _check(self.lib.bnx_map_sync(self.h))

    def set_options(self, opts):
        a = np.ascontiguousarray(opts, dtype=np.int32)
        assert a.shape == (5,)
        _check(self.lib.bnx_map_set_options(self.h, C.c_void_p(a.ctypes.data)))

    def options(self):
        a = np.empty(5, np.int32)
        _check(self.lib.bnx_map_get_options(self.h, C.c_void_p(a.ctypes.data)))
        return a

    def insert(self, pts, origin, max_range, n=None, stride_bytes=None, f64=False):
        """insertPointCloud. pts: float32 (n,3)/(n,4) or float64 (n,3) numpy array, or DevPtr (+ n, stride_bytes, f64)."""
        if isinstance(pts, DevPtr):
            p, where = C.c_void_p(pts.address), BNX_DEVICE
            assert n is not None and stride_bytes is not None
        else:
            arr = np.ascontiguousarray(pts)
            assert arr.ndim == 2
            f64 = arr.dtype == np.float64
            if not f64 and arr.dtype != np.float32:
                raise TypeError(arr.dtype)
            p, where, n, stride_bytes = C.c_void_p(arr.ctypes.data), BNX_HOST, len(arr), arr.shape[1] * arr.dtype.itemsize
        if f64:
            o = np.ascontiguousarray(origin, dtype=np.float64)
            _check(self.lib.bnx_map_insert_f64(self.h, p, C.c_int64(stride_bytes), C.c_int64(n), C.c_void_p(o.ctypes.data), C.c_double(max_range), where))
        else:
            o = np.ascontiguousarray(origin, dtype=np.float32)
            _check(self.lib.bnx_map_insert_f32(self.h, p, C.c_int64(stride_bytes), C.c_int64(n), C.c_void_p(o.ctypes.data), C.c_double(max_range), where))

    def insert_async(self, pts, origin, max_range, n=None, stride_bytes=None, f64=False):
        """pipelined insertPointCloud: enqueue and return. The buffer behind `pts` must stay alive until sync()."""
        if isinstance(pts, DevPtr):
            p, where = C.c_void_p(pts.address), BNX_DEVICE
        else:
            assert pts.flags["C_CONTIGUOUS"] and pts.ndim == 2, "pass the array itself (no temporary copies): it must outlive the call"
            f64 = pts.dtype == np.float64
            p, where, n, stride_bytes = C.c_void_p(pts.ctypes.data), BNX_HOST, len(pts), pts.shape[1] * pts.dtype.itemsize
        if f64:
            o = np.ascontiguousarray(origin, dtype=np.float64)
            _check(self.lib.bnx_map_insert_async_f64(self.h, p, C.c_int64(stride_bytes), C.c_int64(n), C.c_void_p(o.ctypes.data), C.c_double(max_range), where))
        else:
            o = np.ascontiguousarray(origin, dtype=np.float32)
            _check(self.lib.bnx_map_insert_async_f32(self.h, p, C.c_int64(stride_bytes), C.c_int64(n), C.c_void_p(o.ctypes.data), C.c_double(max_range), where))

    def insert_transformed(self, pts, sensor_to_world, origin, max_range, use_async=False):
        """fused ROS pre-step: drop non-finite points, apply the 4x4 float transform, insert (float32 points only)"""
        arr = pts if use_async else np.ascontiguousarray(pts)
        assert arr.dtype == np.float32 and arr.ndim == 2 and arr.flags["C_CONTIGUOUS"]
        T = np.ascontiguousarray(sensor_to_world, dtype=np.float32).reshape(16)
        o = np.ascontiguousarray(origin, dtype=np.float32)
        _check(self.lib.bnx_map_insert_transformed_f32(self.h, C.c_void_p(arr.ctypes.data), C.c_int64(arr.shape[1] * 4), C.c_int64(len(arr)),
                                                       C.c_void_p(T.ctypes.data), C.c_void_p(o.ctypes.data), C.c_double(max_range), BNX_HOST, int(use_async)))

    def totals(self):
        a = (C.c_int64 * 4)()
        _check(self.lib.bnx_map_totals(self.h, a))
        return dict(N=a[0], E=a[1], V=a[2], U=a[3])

    def add_hit(self, p):
        a = np.ascontiguousarray(p, dtype=np.float64)
        _check(self.lib.bnx_map_add_hit(self.h, C.c_void_p(a.ctypes.data)))

    def add_miss(self, p):
        a = np.ascontiguousarray(p, dtype=np.float64)
        _check(self.lib.bnx_map_add_miss(self.h, C.c_void_p(a.ctypes.data)))

    def query(self, xyz, kind: int):
        px, wx, kx = _arg(xyz, np.int32, 3)
        out = np.empty(len(kx), np.uint8)
        _check(self.lib.bnx_map_query(self.h, px, C.c_int64(len(kx)), int(kind), C.c_void_p(out.ctypes.data), BNX_HOST))
        return out.astype(bool)

    def get_voxels(self, kind: int, sort=True):
        cnt = C.c_int64()
        _check(self.lib.bnx_map_get_voxels(self.h, int(kind), None, C.c_int64(0), C.byref(cnt), BNX_HOST))
        xyz = np.empty((cnt.value, 3), np.int32)
        if cnt.value:
            _check(self.lib.bnx_map_get_voxels(self.h, int(kind), C.c_void_p(xyz.ctypes.data), C.c_int64(cnt.value), C.byref(cnt), BNX_HOST))
        if sort and len(xyz):
            xyz = xyz[np.lexsort((xyz[:, 2], xyz[:, 1], xyz[:, 0]))]
        return xyz

    def get_voxel_points(self, kind: int):
        cnt = C.c_int64()
        _check(self.lib.bnx_map_get_voxel_points(self.h, int(kind), None, C.c_int64(0), C.byref(cnt), BNX_HOST))
        xyz = np.empty((cnt.value, 3), np.float64)
        if cnt.value:
            _check(self.lib.bnx_map_get_voxel_points(self.h, int(kind), C.c_void_p(xyz.ctypes.data), C.c_int64(cnt.value), C.byref(cnt), BNX_HOST))
        return xyz

    def publish_occupied(self, z_min: float, z_max: float, stride_floats: int = 3):
        """occupied voxels as float32 points coord*resolution within [z_min, z_max] (the ROS node's publishAll)"""
        cnt = C.c_int64()
        _check(self.lib.bnx_map_publish_occupied_f32(self.h, C.c_double(z_min), C.c_double(z_max), None, C.c_int64(stride_floats), C.c_int64(0), C.byref(cnt), BNX_HOST))
        out = np.empty((cnt.value, stride_floats), np.float32)
        if cnt.value:
            _check(self.lib.bnx_map_publish_occupied_f32(self.h, C.c_double(z_min), C.c_double(z_max), C.c_void_p(out.ctypes.data), C.c_int64(stride_floats),
                                                         C.c_int64(cnt.value), C.byref(cnt), BNX_HOST))
        return out

    def active_count(self) -> int:
        return self._grid.active_count()

    def dump(self, sort=True):
        return self._grid.dump(sort)

    def digest(self):
        return self._grid.digest()

    def shard_stats(self):
        a = (C.c_int64 * 8)()
        _check(self.lib.bnx_map_shard_stats(self.h, a))
        return dict(attempts=a[0], replays=a[1], mailbox_setups=a[2], drains=a[3], sync_retries=a[4], leaf_inbox_cap=a[5],
                    leaf_inbox_max_fill=a[6], scans=a[7])

    def counters(self):
        a = (C.c_int64 * 8)()
        _check(self.lib.bnx_map_counters(self.h, a))
        return dict(N=a[0], E=a[1], V=a[2], U=a[3], leaves_touched=a[4], retries=a[5], rays=a[6], chunks=a[7])

    def update_count(self) -> int:
        v = C.c_int()
        _check(self.lib.bnx_map_update_count(self.h, C.byref(v)))
        return v.value

    def set_marking(self, mode: str):
        """"sparse": per-scan marks inside the leaves; "dense": experimental dense window where the range allows it;
        "default": sparse unless BNX_DENSE=1"""
        _check(self.lib.bnx_map_set_marking(self.h, {"default": 0, "sparse": 1, "dense": 2}[mode]))

    def set_profiling(self, enable=True):
        _check(self.lib.bnx_map_set_profiling(self.h, int(enable)))

    def phase_times(self):
        a = (C.c_double * 8)()
        _check(self.lib.bnx_map_phase_times(self.h, a))
        return dict(h2d=a[0], classify=a[1], resolve=a[2], mark=a[3], apply=a[4], total=a[5], sub6=a[6], sub7=a[7])
