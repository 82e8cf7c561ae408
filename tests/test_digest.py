"""The order-independent map digest (bnx_grid_digest, restated in numpy as capi.digest_of_dump) against a plain-Python
spelling of the same formula. CPU only: the device side is compared with it in tests/test_gpu_fullsize.py."""
import numpy as np

from bonxai_b200 import capi

M = (1 << 64) - 1


def mix64(h):
    h ^= h >> 33
    h = (h * 0xFF51AFD7ED558CCD) & M
    h ^= h >> 33
    h = (h * 0xC4CEB9FE1A85EC53) & M
    return h ^ (h >> 33)


def hash3(x, y, z):
    x, y, z = x & 0xFFFFFFFF, y & 0xFFFFFFFF, z & 0xFFFFFFFF
    h = (x * 0x9E3779B97F4A7C15) & M
    h ^= ((y * 0xC2B2AE3D27D4EB4F) + (h >> 29)) & M
    h ^= ((z * 0x165667B19E3779F9) + ((h << 7) & M)) & M
    return mix64(h)


def slow_digest(xyz, vals):
    s = x = 0
    raw = np.ascontiguousarray(vals).view(np.uint8).reshape(len(xyz), -1)
    for (cx, cy, cz), b in zip(xyz.tolist(), raw.tolist()):
        f = 0xCBF29CE484222325
        for byte in b:
            f = ((f ^ byte) * 0x100000001B3) & M
        h = mix64((hash3(cx, cy, cz) + f * 0x9E3779B97F4A7C15) & M)
        s = (s + h) & M
        x ^= h
    return s, x, len(xyz)


def test_digest_matches_plain_python():
    rng = np.random.default_rng(5)
    xyz = rng.integers(-(1 << 31), 1 << 31, (300, 3), dtype=np.int64).astype(np.int32)
    for vals in (rng.integers(0, 1 << 32, 300, dtype=np.uint64).astype(np.uint32), rng.normal(size=300), rng.integers(0, 255, 300).astype(np.uint8)):
        assert capi.digest_of_dump(xyz, vals) == slow_digest(xyz, vals)
    assert capi.digest_of_dump(xyz[:0], xyz[:0, 0]) == (0, 0, 0)


def test_digest_is_order_independent_and_additive_over_shards():
    rng = np.random.default_rng(6)
    xyz = rng.integers(-5000, 5000, (2000, 3)).astype(np.int32)
    w = rng.integers(0, 1 << 32, 2000, dtype=np.uint64).astype(np.uint32)
    whole = capi.digest_of_dump(xyz, w)
    perm = rng.permutation(2000)
    assert capi.digest_of_dump(xyz[perm], w[perm]) == whole
    parts = [capi.digest_of_dump(xyz[a:b], w[a:b]) for a, b in ((0, 700), (700, 1500), (1500, 2000))]
    assert capi.combine_digests(parts) == whole
    w2 = w.copy()
    w2[17] ^= 1
    assert capi.digest_of_dump(xyz, w2) != whole


def test_c_digest_helper_matches_numpy():
    import oracle
    oracle.build("port")
    rng = np.random.default_rng(7)
    xyz = rng.integers(-(1 << 31), 1 << 31, (5000, 3), dtype=np.int64).astype(np.int32)
    w = rng.integers(0, 1 << 32, 5000, dtype=np.uint64).astype(np.uint32)
    assert oracle.digest_pairs(xyz, w) == capi.digest_of_dump(xyz, w)
    m = oracle.load("port").map(0.1)
    m.insert(rng.normal(0, 2, (500, 3)).astype(np.float32), [0, 0, 0], 5.0)
    assert m.digest() == capi.digest_of_dump(*m.dump())
