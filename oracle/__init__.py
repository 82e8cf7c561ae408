"""TEST INFRASTRUCTURE ONLY — ctypes loader for the two CPU oracles behind oracle/oracle_api.h.

    load("port")       -> oracle/libbonxai_oracle.so   (plain-C restatement, oracle/bonxai_oracle.c)
    load("reference")  -> oracle/_ref/libbonxai_ref.so (the unmodified reference, oracle/ref_wrap.cpp)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
package. Nothing under bonxai_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "libbonxai_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libbonxai_ref.so")
REFERENCE_TREE = "/root/reference"

_i32p = C.POINTER(C.c_int32)
_u32p = C.POINTER(C.c_uint32)
_u8p = C.POINTER(C.c_uint8)
_f64p = C.POINTER(C.c_double)


def build(which: str = "all", quiet: bool = True) -> None:
    """Compile the oracle libraries (the reference one only where /root/reference exists)."""
    targets = []
    if which in ("all", "port"):
        targets.append("port")
    if which in ("all", "reference") and os.path.isdir(os.path.join(REFERENCE_TREE, "bonxai_core")):
        targets.append("ref")
    for t in targets:
        subprocess.run(["make", "-C", HERE, t], check=True,
                       stdout=subprocess.DEVNULL if quiet else None)


def available(kind: str) -> bool:
    return os.path.exists(PORT_SO if kind == "port" else REF_SO)


def _ptr(a, typ):
    return a.ctypes.data_as(typ)


class OracleLib:
    def __init__(self, path: str):
        self.path = path
        lib = C.CDLL(path)
        self.lib = lib
        lib.orc_kind.restype = C.c_char_p
        lib.orc_logods.restype = C.c_int32
        lib.orc_logods.argtypes = [C.c_float]
        lib.orc_prob.restype = C.c_float
        lib.orc_prob.argtypes = [C.c_int32]
        lib.orc_pos_to_coord.argtypes = [C.c_double, _f64p, C.c_int64, _i32p]
        lib.orc_coord_to_pos.argtypes = [C.c_double, _i32p, C.c_int64, _f64p]
        lib.orc_compute_ray.restype = C.c_int64
        lib.orc_compute_ray.argtypes = [_i32p, _i32p, _i32p, C.c_int64]
        lib.orc_grid_create.restype = C.c_void_p
        lib.orc_grid_create.argtypes = [C.c_double, C.c_int, C.c_int]
        lib.orc_grid_destroy.argtypes = [C.c_void_p]
        lib.orc_grid_set_values.argtypes = [C.c_void_p, _i32p, _u32p, C.c_int64, _u8p]
        lib.orc_grid_get_values.argtypes = [C.c_void_p, _i32p, C.c_int64, _u32p, _u8p]
        lib.orc_grid_get_or_create.argtypes = [C.c_void_p, _i32p, C.c_int64, _u32p]
        lib.orc_grid_set_on.argtypes = [C.c_void_p, _i32p, C.c_int64, C.c_uint32, _u8p]
        lib.orc_grid_set_off.argtypes = [C.c_void_p, _i32p, C.c_int64, _u8p]
        lib.orc_grid_is_on.argtypes = [C.c_void_p, _i32p, C.c_int64, _u8p]
        lib.orc_grid_active_count.restype = C.c_int64
        lib.orc_grid_active_count.argtypes = [C.c_void_p]
        lib.orc_grid_dump.restype = C.c_int64
        lib.orc_grid_dump.argtypes = [C.c_void_p, _i32p, _u32p, C.c_int64]
        lib.orc_grid_clear.argtypes = [C.c_void_p, C.c_int]
        lib.orc_grid_release_unused.argtypes = [C.c_void_p]
        lib.orc_grid_serialize.restype = C.c_int64
        lib.orc_grid_serialize.argtypes = [C.c_void_p, _u8p, C.c_int64]
        lib.orc_grid_deserialize.restype = C.c_void_p
        lib.orc_grid_deserialize.argtypes = [_u8p, C.c_int64]
        lib.orc_map_create.restype = C.c_void_p
        lib.orc_map_create.argtypes = [C.c_double]
        lib.orc_map_destroy.argtypes = [C.c_void_p]
        lib.orc_map_set_options.argtypes = [C.c_void_p, _i32p]
        lib.orc_map_get_options.argtypes = [C.c_void_p, _i32p]
        lib.orc_map_insert_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                           C.POINTER(C.c_float), C.c_double]
        lib.orc_map_insert_f64.argtypes = [C.c_void_p, _f64p, C.c_int64, _f64p, C.c_double]
        lib.orc_map_add_hit.argtypes = [C.c_void_p, _f64p]
        lib.orc_map_add_miss.argtypes = [C.c_void_p, _f64p]
        lib.orc_map_query.argtypes = [C.c_void_p, _i32p, C.c_int64, C.c_int, _u8p]
        lib.orc_map_get_voxels.restype = C.c_int64
        lib.orc_map_get_voxels.argtypes = [C.c_void_p, C.c_int, _i32p, C.c_int64]
        lib.orc_map_active_count.restype = C.c_int64
        lib.orc_map_active_count.argtypes = [C.c_void_p]
        lib.orc_map_dump.restype = C.c_int64
        lib.orc_map_dump.argtypes = [C.c_void_p, _i32p, _u32p, C.c_int64]
        lib.orc_map_counters.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        lib.orc_map_track_updates.argtypes = [C.c_void_p, C.c_int]
        lib.orc_map_last_insert_seconds.restype = C.c_double
        lib.orc_map_last_insert_seconds.argtypes = [C.c_void_p]
        self.kind = lib.orc_kind().decode()

    # ---- scalar helpers
    def logods(self, p: float) -> int:
        return int(self.lib.orc_logods(C.c_float(p)))

    def prob(self, v: int) -> float:
        return float(self.lib.orc_prob(int(v)))

    def pos_to_coord(self, res: float, xyz) -> np.ndarray:
        xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
        out = np.empty((len(xyz), 3), np.int32)
        self.lib.orc_pos_to_coord(res, _ptr(xyz, _f64p), len(xyz), _ptr(out, _i32p))
        return out

    def coord_to_pos(self, res: float, xyz) -> np.ndarray:
        xyz = np.ascontiguousarray(xyz, dtype=np.int32).reshape(-1, 3)
        out = np.empty((len(xyz), 3), np.float64)
        self.lib.orc_coord_to_pos(res, _ptr(xyz, _i32p), len(xyz), _ptr(out, _f64p))
        return out

    def compute_ray(self, a, b) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.int32)
        b = np.ascontiguousarray(b, dtype=np.int32)
        n = int(np.max(np.abs(b.astype(np.int64) - a.astype(np.int64))))
        out = np.empty((max(n, 1), 3), np.int32)
        cnt = self.lib.orc_compute_ray(_ptr(a, _i32p), _ptr(b, _i32p), _ptr(out, _i32p), len(out))
        return out[:cnt]

    def grid(self, voxel_size: float, inner_bits: int = 2, leaf_bits: int = 3) -> "OracleGrid":
        return OracleGrid(self, voxel_size, inner_bits, leaf_bits)

    def map(self, resolution: float) -> "OracleMap":
        return OracleMap(self, resolution)


def digest_pairs(xyz: np.ndarray, words: np.ndarray):
    """(sum, xor, count) digest of an unsorted dump of 4-byte cells — the C spelling of bnx_grid_digest (a helper of
    the port library; capi.digest_of_dump is the numpy spelling, tests/test_digest.py the plain-Python one)"""
    lib = C.CDLL(PORT_SO)
    xyz = np.ascontiguousarray(xyz, np.int32)
    words = np.ascontiguousarray(words).view(np.uint32)
    out = (C.c_uint64 * 3)()
    lib.orc_digest_pairs(_ptr(xyz, _i32p), _ptr(words, _u32p), C.c_int64(len(words)), out)
    return int(out[0]), int(out[1]), int(out[2])


def sort_dump(xyz: np.ndarray, vals: np.ndarray):
    """Canonical order for comparing forEachCell dumps (the reference's order is unspecified)."""
    if len(xyz) == 0:
        return xyz.reshape(0, 3), vals
    order = np.lexsort((xyz[:, 2], xyz[:, 1], xyz[:, 0]))
    return xyz[order], vals[order]


class OracleGrid:
    def __init__(self, lib: OracleLib, voxel_size, inner_bits=2, leaf_bits=3, handle=None):
        self.o = lib
        self.h = handle if handle is not None else lib.lib.orc_grid_create(voxel_size, inner_bits, leaf_bits)
        if not self.h:
            raise RuntimeError("The minimum value of the inner_bits and leaf_bits should be 1")

    def __del__(self):
        if getattr(self, "h", None):
            self.o.lib.orc_grid_destroy(self.h)
            self.h = None

    @staticmethod
    def _c(xyz):
        return np.ascontiguousarray(xyz, dtype=np.int32).reshape(-1, 3)

    def set_values(self, xyz, vals):
        xyz = self._c(xyz)
        vals = np.ascontiguousarray(vals).view(np.uint32)
        was = np.empty(len(xyz), np.uint8)
        self.o.lib.orc_grid_set_values(self.h, _ptr(xyz, _i32p), _ptr(vals, _u32p), len(xyz), _ptr(was, _u8p))
        return was.astype(bool)

    def get_values(self, xyz):
        xyz = self._c(xyz)
        out = np.zeros(len(xyz), np.uint32)
        found = np.empty(len(xyz), np.uint8)
        self.o.lib.orc_grid_get_values(self.h, _ptr(xyz, _i32p), len(xyz), _ptr(out, _u32p), _ptr(found, _u8p))
        return out, found.astype(bool)

    def get_or_create(self, xyz):
        xyz = self._c(xyz)
        out = np.zeros(len(xyz), np.uint32)
        self.o.lib.orc_grid_get_or_create(self.h, _ptr(xyz, _i32p), len(xyz), _ptr(out, _u32p))
        return out

    def set_on(self, xyz, default_value=0):
        xyz = self._c(xyz)
        was = np.empty(len(xyz), np.uint8)
        self.o.lib.orc_grid_set_on(self.h, _ptr(xyz, _i32p), len(xyz), int(default_value), _ptr(was, _u8p))
        return was.astype(bool)

    def set_off(self, xyz):
        xyz = self._c(xyz)
        was = np.empty(len(xyz), np.uint8)
        self.o.lib.orc_grid_set_off(self.h, _ptr(xyz, _i32p), len(xyz), _ptr(was, _u8p))
        return was.astype(bool)

    def is_on(self, xyz):
        xyz = self._c(xyz)
        out = np.empty(len(xyz), np.uint8)
        self.o.lib.orc_grid_is_on(self.h, _ptr(xyz, _i32p), len(xyz), _ptr(out, _u8p))
        return out.astype(bool)

    def active_count(self) -> int:
        return int(self.o.lib.orc_grid_active_count(self.h))

    def dump(self, sort=True):
        n = self.active_count()
        xyz = np.empty((n, 3), np.int32)
        vals = np.empty(n, np.uint32)
        got = self.o.lib.orc_grid_dump(self.h, _ptr(xyz, _i32p), _ptr(vals, _u32p), n)
        assert got == n
        return sort_dump(xyz, vals) if sort else (xyz, vals)

    def digest(self):
        return digest_pairs(*self.dump(sort=False))

    def clear(self, opt: int):
        self.o.lib.orc_grid_clear(self.h, opt)

    def release_unused(self):
        self.o.lib.orc_grid_release_unused(self.h)

    def serialize(self) -> bytes:
        n = self.o.lib.orc_grid_serialize(self.h, None, 0)
        if n < 0:
            raise NotImplementedError("serialisation is only available from the reference build")
        buf = np.empty(n, np.uint8)
        self.o.lib.orc_grid_serialize(self.h, _ptr(buf, _u8p), n)
        return buf.tobytes()

    @classmethod
    def deserialize(cls, lib: OracleLib, data: bytes):
        buf = np.frombuffer(data, np.uint8).copy()
        h = lib.lib.orc_grid_deserialize(_ptr(buf, _u8p), len(buf))
        if not h:
            raise RuntimeError("Header wasn't recognized")
        return cls(lib, 0.0, handle=h)


class OracleMap:
    def __init__(self, lib: OracleLib, resolution: float):
        self.o = lib
        self.h = lib.lib.orc_map_create(resolution)

    def __del__(self):
        if getattr(self, "h", None):
            self.o.lib.orc_map_destroy(self.h)
            self.h = None

    def set_options(self, opts):
        a = np.ascontiguousarray(opts, dtype=np.int32)
        assert a.shape == (5,)
        self.o.lib.orc_map_set_options(self.h, _ptr(a, _i32p))

    def options(self):
        a = np.empty(5, np.int32)
        self.o.lib.orc_map_get_options(self.h, _ptr(a, _i32p))
        return a

    def insert(self, pts, origin, max_range):
        """pts: (n,3)/(n,4) float32 (stride 12/16) or (n,3) float64."""
        pts = np.ascontiguousarray(pts)
        if pts.dtype == np.float32:
            o = np.ascontiguousarray(origin, dtype=np.float32)
            self.o.lib.orc_map_insert_f32(self.h, pts.ctypes.data, pts.shape[1] * 4, len(pts),
                                          o.ctypes.data_as(C.POINTER(C.c_float)), float(max_range))
        elif pts.dtype == np.float64:
            assert pts.shape[1] == 3
            o = np.ascontiguousarray(origin, dtype=np.float64)
            self.o.lib.orc_map_insert_f64(self.h, _ptr(pts, _f64p), len(pts), _ptr(o, _f64p), float(max_range))
        else:
            raise TypeError(pts.dtype)

    def add_hit(self, p):
        a = np.ascontiguousarray(p, dtype=np.float64)
        self.o.lib.orc_map_add_hit(self.h, _ptr(a, _f64p))

    def add_miss(self, p):
        a = np.ascontiguousarray(p, dtype=np.float64)
        self.o.lib.orc_map_add_miss(self.h, _ptr(a, _f64p))

    def query(self, xyz, kind: int):
        xyz = np.ascontiguousarray(xyz, dtype=np.int32).reshape(-1, 3)
        out = np.empty(len(xyz), np.uint8)
        self.o.lib.orc_map_query(self.h, _ptr(xyz, _i32p), len(xyz), kind, _ptr(out, _u8p))
        return out.astype(bool)

    def get_voxels(self, kind: int, sort=True):
        n = self.o.lib.orc_map_get_voxels(self.h, kind, None, 0)
        xyz = np.empty((n, 3), np.int32)
        self.o.lib.orc_map_get_voxels(self.h, kind, _ptr(xyz, _i32p), n)
        if sort and n:
            xyz = xyz[np.lexsort((xyz[:, 2], xyz[:, 1], xyz[:, 0]))]
        return xyz

    def active_count(self) -> int:
        return int(self.o.lib.orc_map_active_count(self.h))

    def dump(self, sort=True):
        n = self.active_count()
        xyz = np.empty((n, 3), np.int32)
        words = np.empty(n, np.uint32)
        got = self.o.lib.orc_map_dump(self.h, _ptr(xyz, _i32p), _ptr(words, _u32p), n)
        assert got == n
        return sort_dump(xyz, words) if sort else (xyz, words)

    def digest(self):
        return digest_pairs(*self.dump(sort=False))

    def counters(self):
        a = (C.c_int64 * 4)()
        self.o.lib.orc_map_counters(self.h, a)
        return dict(N=a[0], E=a[1], V=a[2], U=a[3])

    def track_updates(self, enable=True):
        self.o.lib.orc_map_track_updates(self.h, int(enable))

    def last_insert_seconds(self) -> float:
        return float(self.o.lib.orc_map_last_insert_seconds(self.h))


def load(kind: str = "port") -> OracleLib:
    path = PORT_SO if kind == "port" else REF_SO
    if not os.path.exists(path):
        build("port" if kind == "port" else "reference")
    return OracleLib(path)
