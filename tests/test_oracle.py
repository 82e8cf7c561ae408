"""CPU suite: pins the plain-C restatement (oracle/bonxai_oracle.c) against the reference itself — the golden
digests generated from the unmodified reference (tests/golden/map_golden.json) and, where the reference build
is available, live side-by-side runs. Nothing here needs a GPU."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import workloads as W  # noqa: E402

from conftest import assert_same_dump  # noqa: E402

GOLDEN = json.load(open(os.path.join(HERE, "golden", "map_golden.json")))


def test_known_constants(any_oracle):
    o = any_oracle
    assert [o.logods(p) for p in (0.4, 0.7, 0.12, 0.97, 0.5)] == [-405465, 847297, -1992430, 3476099, 0]
    assert abs(o.prob(847297) - 0.7) < 1e-6 and o.prob(0) == 0.5
    m = o.map(0.1)
    assert list(m.options()) == [-405465, 847297, -1992430, 3476099, 0]


def test_golden_inputs_are_reproducible():
    """the synthetic inputs regenerate bit-identically (otherwise the digests below are not comparable)"""
    for name, (res, scans) in W.workloads().items():
        h = hashlib.sha256()
        for pts, origin, max_range in scans:
            h.update(pts.tobytes())
            h.update(origin.tobytes())
        assert h.hexdigest()[:16] == GOLDEN["inputs"][name], name


@pytest.mark.parametrize("name", sorted(GOLDEN["digests"]))
def test_port_matches_reference_golden(port, name):
    got = W.run(port.map, {name})[name]
    assert got == GOLDEN["digests"][name], name


def test_cell_word_layout(any_oracle):
    """(probability_log << 4) | update_id, signed 28-bit / 4-bit fields (SURVEY.md trap 6)"""
    m = any_oracle.map(0.1)
    m.insert(np.array([[1.0, 0, 0]], np.float32), [0, 0, 0], 5.0)
    xyz, w = m.dump()
    cells = {tuple(c): int(v) for c, v in zip(xyz, w)}
    assert cells[(10, 0, 0)] == ((847297 << 4) | 1)
    assert cells[(3, 0, 0)] == ((-405465 << 4) | 1) & 0xFFFFFFFF == 0xFF9D0271


def test_ray_is_closed_form(any_oracle):
    """cell k = origin + sign * floor((2k|d| + m) / 2m): the property the GPU walk relies on"""
    rng = np.random.default_rng(0)
    for _ in range(300):
        a = rng.integers(-50, 50, 3)
        b = a + rng.integers(-120, 120, 3)
        ray = any_oracle.compute_ray(a, b)
        d = b - a
        m = int(np.abs(d).max())
        assert len(ray) == m
        if m:
            k = np.arange(m)[:, None]
            want = a + np.sign(d) * ((2 * k * np.abs(d) + m) // (2 * m))
            assert np.array_equal(ray, want)
    assert len(any_oracle.compute_ray([0, 0, 0], [12, 6, 80])) == 80


def test_port_vs_reference_live_maps(port, ref):
    rng = np.random.default_rng(11)
    for res, rmax in ((0.1, 5.0), (0.25, float("inf")), (0.013, 0.8)):
        pm, rm = port.map(res), ref.map(res)
        opts = [-300000, 900000, -1500000, 3000000, 100000]
        pm.set_options(opts)
        rm.set_options(opts)
        for scan in range(6):
            n = int(rng.integers(1, 3000))
            origin = rng.uniform(-1, 1, 3)
            pts = origin + rng.normal(0, 2.0, (n, 3))
            if scan % 2:
                pts, origin = pts.astype(np.float32), origin.astype(np.float32)
            if scan == 3:
                for m in (pm, rm):
                    m.add_hit(rng.uniform(-1, 1, 3) * 0 + [0.5, 0.25, 0.1])
                    m.add_miss([-0.5, 0.3, 0.2])
            pm.insert(pts, origin, rmax)
            rm.insert(pts, origin, rmax)
            assert_same_dump(pm.dump(), rm.dump(), f"res {res} scan {scan}")
        q = rng.integers(-60, 60, (5000, 3)).astype(np.int32)
        for kind in (0, 1, 2):
            assert np.array_equal(pm.query(q, kind), rm.query(q, kind))
        assert np.array_equal(pm.get_voxels(0), rm.get_voxels(0)) and np.array_equal(pm.get_voxels(2), rm.get_voxels(2))


def test_port_vs_reference_edge_inputs(port, ref):
    """inputs that sit on the corners of the algorithm (bonxai_map/src/probabilistic_map.cpp:79-141 and
    bonxai_core/include/bonxai/grid_coord.hpp posToCoord): points exactly on voxel faces (floor of x / res), negative
    coordinates next to zero, rays of length zero, points further than max_range (clamped to the range sphere and cast as
    misses), the same endpoint many times, voxel coordinates beyond +-2^20, a scan of one point and an empty scan"""
    rng = np.random.default_rng(23)
    res = 0.125  # exactly representable: k * res lands on voxel faces in both float widths
    lattice = (rng.integers(-40, 40, (1500, 3)) * res)
    near_zero = rng.choice([-res, -1e-7, -0.0, 0.0, 1e-7, res], (600, 3))
    far = np.float64([4.0e5, -3.0e5, 2.5e5])  # ~3.2e6 voxels out
    scans = [(lattice, np.zeros(3), 3.0),
             (lattice.astype(np.float32), np.float32([res, res, res]), 1.0),       # origin on a voxel corner, mostly clamped
             (near_zero, np.float64([0.0, 0.0, 0.0]), float("inf")),                # includes zero-length rays
             (np.repeat(lattice[:7], 200, axis=0), np.float64([0.3, -0.2, 0.1]), 2.0),
             (lattice + far, far, 2.5),
             ((lattice + far).astype(np.float32), far.astype(np.float32), float("inf")),
             (lattice[:1], np.float64([1.0, 1.0, 1.0]), 10.0),
             (lattice[:0], np.zeros(3), 10.0),
             (lattice, np.zeros(3), 0.0)]                                           # range 0: every point clamps onto the origin
    pm, rm = port.map(res), ref.map(res)
    for k, (pts, origin, rmax) in enumerate(scans):
        pts = np.ascontiguousarray(pts)
        pm.insert(pts, origin, rmax)
        rm.insert(pts, origin, rmax)
        assert_same_dump(pm.dump(), rm.dump(), f"edge scan {k}")
    assert pm.active_count() > 20000 and len(pm.get_voxels(0)) > 500  # the comparison is not vacuous
    assert np.array_equal(pm.get_voxels(0), rm.get_voxels(0)) and np.array_equal(pm.get_voxels(1), rm.get_voxels(1))


def test_port_counters_match_dump_diff(port, ref):
    """U (cells whose word changed) counted natively by the port == dump diff measured on the reference"""
    from bonxai_b200 import synth
    pm, rm = port.map(0.1), ref.map(0.1)
    rm.track_updates(True)
    for scan in range(3):
        pts, origin = synth.lidar_scan(scan, beams=16, azimuths=256)
        pm.insert(pts, origin, 30.0)
        rm.insert(pts, origin, 30.0)
        assert pm.counters()["U"] == rm.counters()["U"]
        assert pm.counters()["N"] == len(pts)


def test_port_vs_reference_live_grids(port, ref):
    rng = np.random.default_rng(5)
    for bits in ((2, 3), (1, 1), (3, 2)):
        pg, rg = port.grid(0.1, *bits), ref.grid(0.1, *bits)
        for step in range(4):
            xyz = rng.integers(-30, 30, (4000, 3)).astype(np.int32)
            vals = rng.integers(0, 2**32, 4000, dtype=np.uint64).astype(np.uint32)
            assert np.array_equal(pg.set_values(xyz, vals), rg.set_values(xyz, vals))
            off = rng.integers(-30, 30, (1500, 3)).astype(np.int32)
            assert np.array_equal(pg.set_off(off), rg.set_off(off))
            assert np.array_equal(pg.set_on(off[:700], step), rg.set_on(off[:700], step))
            assert np.array_equal(pg.get_or_create(off[700:900]), rg.get_or_create(off[700:900]))
            assert_same_dump(pg.dump(), rg.dump(), f"bits {bits} step {step}")
            assert np.array_equal(pg.is_on(xyz), rg.is_on(xyz))
        pg.release_unused()
        rg.release_unused()
        assert_same_dump(pg.dump(), rg.dump(), "after release")
        pg.clear(1)
        rg.clear(1)
        assert pg.active_count() == rg.active_count() == 0
    assert port.lib.orc_grid_create(0.1, 0, 3) is None and ref.lib.orc_grid_create(0.1, 0, 3) is None
