"""GPU parity of ProbabilisticMap::insertPointCloud (through the C ABI) against the CPU oracles.

Bit-exact bar: after every scan the sorted forEachCell dump (coord, CellT word) and the oracle's work
counters must be identical.
"""
import os

import numpy as np
import pytest

from bonxai_b200 import synth
from conftest import assert_same_dump

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

DEFAULT_OPTS = [-405465, 847297, -1992430, 3476099, 0]


@pytest.fixture(autouse=True, params=["sparse", "dense"])
def marking(request, monkeypatch):
    """every test of this file runs twice: with the per-scan marks in the leaves (the default) and with the experimental
    dense marking window wherever max_range is finite and small enough (bnx_map_set_marking)"""
    if request.param == "dense":
        from bonxai_b200 import capi
        orig = capi.ProbabilisticMap.__init__

        def init(self, *a, **k):
            orig(self, *a, **k)
            self.set_marking("dense")

        monkeypatch.setattr(capi.ProbabilisticMap, "__init__", init)
    return request.param


def check_scan(gm, om, what, counters=True):
    assert_same_dump(gm.dump(), om.dump(), what)
    if counters:
        oc = om.counters()
        if oc["E"] >= 0:  # only the port counts E/V/U
            gc = gm.counters()
            assert (gc["N"], gc["E"], gc["V"], gc["U"]) == (oc["N"], oc["E"], oc["V"], oc["U"]), f"{what}: {gc} vs {oc}"


def test_default_options(bnx, port):
    m = bnx.ProbabilisticMap(0.1)
    assert list(m.options()) == DEFAULT_OPTS == list(port.map(0.1).options())
    assert m.update_count() == 1


def test_quirk_table(bnx, any_oracle):
    """SURVEY.md §3.1 known answers, produced by the reference itself."""
    pts = np.array([[1, 0, 0], [1, .05, 0], [0, 3, 0], [-.55, -.72, .33]], np.float32)
    gm, om = bnx.ProbabilisticMap(0.1), any_oracle.map(0.1)
    gm.insert(pts, [0, 0, 0], 2.0)
    om.insert(pts, [0, 0, 0], 2.0)
    check_scan(gm, om, "quirk scan")
    xyz, w = gm.dump()
    assert len(xyz) == 39
    cells = {tuple(c): int(v) for c, v in zip(xyz, w.view(np.int32) >> 4)}
    assert cells[(10, 0, 0)] == 847297 and cells[(0, 20, 0)] == -405465 and cells[(-6, -8, 3)] == 847297
    assert all(cells[(0, k, 0)] == -405465 for k in range(20))
    assert gm.update_count() == 2
    assert len(gm.get_voxels(bnx.BNX_OCCUPIED)) == 2


def test_stale_update_id_skips_endpoint_and_ray(bnx, any_oracle):
    """trap 2: scan A, B, B, then A again when the counter has wrapped: A's voxel and ray stay untouched."""
    a = np.array([[1.0, 0.02, 0.01]], np.float32)
    b = np.array([[0.0, 1.0, 0.0]], np.float32)
    gm, om = bnx.ProbabilisticMap(0.1), any_oracle.map(0.1)
    for k, pts in enumerate([a, b, b, a, a]):
        gm.insert(pts, [0, 0, 0], 10.0)
        om.insert(pts, [0, 0, 0], 10.0)
        check_scan(gm, om, f"stale scan {k}")
    xyz, w = gm.dump()
    cells = {tuple(c): int(v) for c, v in zip(xyz, w.view(np.int32) >> 4)}
    assert cells[(10, 0, 0)] == 1694594 and cells[(5, 0, 0)] == -810930


def test_first_point_decides_hit_or_miss(bnx, any_oracle):
    """trap 3: same endpoint voxel from an over-range and an in-range point: the lower index wins."""
    for order in (0, 1):
        over = [2.0, 2.0, 0.0]      # beyond max_range 2.0 -> truncated to (1.414.., 1.414.., 0) -> voxel (14,14,0)
        inr = [1.41, 1.41, 0.01]
        pts = np.array([over, inr] if order == 0 else [inr, over], np.float32)
        gm, om = bnx.ProbabilisticMap(0.1), any_oracle.map(0.1)
        gm.insert(pts, [0, 0, 0], 2.0)
        om.insert(pts, [0, 0, 0], 2.0)
        check_scan(gm, om, f"order {order}")
        xyz, w = gm.dump()
        cells = {tuple(c): int(v) for c, v in zip(xyz, w.view(np.int32) >> 4)}
        assert cells[(14, 14, 0)] == (-405465 if order == 0 else 847297)


def test_clamping_over_repeated_scans(bnx, port):
    pts = np.array([[0.75, 0.31, -0.2]], np.float32)
    gm, om = bnx.ProbabilisticMap(0.1), port.map(0.1)
    for k in range(12):
        gm.insert(pts, [0, 0, 0], 5.0)
        om.insert(pts, [0, 0, 0], 5.0)
        check_scan(gm, om, f"clamp scan {k}")
    xyz, w = gm.dump()
    probs = set((w.view(np.int32) >> 4).tolist())
    assert probs == {3476099, -1992430}


def test_empty_and_degenerate_scans(bnx, port):
    gm, om = bnx.ProbabilisticMap(0.05), port.map(0.05)
    empty = np.zeros((0, 3), np.float32)
    gm.insert(empty, [0, 0, 0], 5.0)
    om.insert(empty, [0, 0, 0], 5.0)
    assert gm.active_count() == 0 and gm.update_count() == 2
    same = np.array([[0.01, 0.01, 0.01], [0.02, 0.0, 0.04]], np.float32)  # endpoint voxel == origin voxel: empty ray
    gm.insert(same, [0.0, 0.0, 0.0], 5.0)
    om.insert(same, [0.0, 0.0, 0.0], 5.0)
    check_scan(gm, om, "origin voxel")
    assert gm.active_count() == 1
    one = np.array([[0.07, 0.0, 0.0]], np.float32)  # neighbour voxel: ray = the origin cell only
    gm.insert(one, [0.0, 0.0, 0.0], 5.0)
    om.insert(one, [0.0, 0.0, 0.0], 5.0)
    check_scan(gm, om, "one-cell ray")


@pytest.mark.parametrize("layout", ["f32x3", "f32x4", "f64"])
@pytest.mark.parametrize("res,rng_max", [(0.1, 6.0), (0.02, 1.5), (0.37, float("inf"))])
def test_random_scans(bnx, port, layout, res, rng_max):
    rng = np.random.default_rng(hash((layout, res)) % 2**32)
    gm, om = bnx.ProbabilisticMap(res), port.map(res)
    for scan in range(5):
        n = int(rng.integers(1, 4000))
        origin = rng.uniform(-3, 3, 3)
        pts = origin + rng.normal(0, 4.0, (n, 3))
        pts[: n // 10] = pts[0]  # heavy duplicates
        if layout == "f64":
            p, o = pts.astype(np.float64), origin.astype(np.float64)
        else:
            p, o = pts.astype(np.float32), origin.astype(np.float32)
            if layout == "f32x4":
                p = np.concatenate([p, rng.normal(size=(n, 1)).astype(np.float32)], axis=1)
        gm.insert(p, o, rng_max)
        om.insert(p, o, rng_max)
        check_scan(gm, om, f"{layout} res={res} scan {scan}")


def test_negative_and_exact_boundary_coordinates(bnx, port):
    """floor semantics on negative values and on products that land exactly on / just below integers."""
    res = 0.1
    vals = np.array([-0.3, -0.30000000000000004, -0.2, -0.1, -1e-12, 0.0, 0.1, 0.2, 0.30000000000000004, 0.7, -2.5], np.float64)
    pts = np.stack(np.meshgrid(vals, vals[:4], vals[5:8], indexing="ij"), -1).reshape(-1, 3)
    gm, om = bnx.ProbabilisticMap(res), port.map(res)
    gm.insert(pts, [0.05, -0.05, 0.0], 100.0)
    om.insert(pts, [0.05, -0.05, 0.0], 100.0)
    check_scan(gm, om, "boundary f64")
    p32 = pts.astype(np.float32)
    gm.insert(p32, np.float32([0.05, -0.05, 0.0]), 0.25)
    om.insert(p32, np.float32([0.05, -0.05, 0.0]), 0.25)
    check_scan(gm, om, "boundary f32 clipped")


def test_long_rays_64bit_path(bnx, port):
    """rays longer than 2^15 cells take the 64-bit closed form."""
    res = 0.001
    pts = np.array([[40.0, 3.0, -1.0], [-35.5, 0.2, 0.1], [0.5, 0.5, 39.0]], np.float64)
    gm, om = bnx.ProbabilisticMap(res), port.map(res)
    gm.insert(pts, [0.0, 0.0, 0.0], float("inf"))
    om.insert(pts, [0.0, 0.0, 0.0], float("inf"))
    check_scan(gm, om, "long rays")
    assert gm.counters()["V"] > 100_000


def test_custom_options(bnx, port):
    opts = [-200000, 1500000, -700000, 2500000, 300000]
    gm, om = bnx.ProbabilisticMap(0.2), port.map(0.2)
    gm.set_options(opts)
    om.set_options(opts)
    assert list(gm.options()) == opts
    rng = np.random.default_rng(5)
    for scan in range(6):
        pts = rng.normal(0, 3, (500, 3)).astype(np.float32)
        gm.insert(pts, [0, 0, 0], 4.0)
        om.insert(pts, [0, 0, 0], 4.0)
        check_scan(gm, om, f"opts scan {scan}")
    for kind in (bnx.BNX_OCCUPIED, bnx.BNX_FREE):
        assert np.array_equal(gm.get_voxels(kind), om.get_voxels(kind))


def test_add_hit_miss_are_queued_until_next_insert(bnx, any_oracle):
    gm, om = bnx.ProbabilisticMap(0.1), any_oracle.map(0.1)
    for m in (gm, om):
        m.add_hit([1.0, 1.0, 0.0])
        m.add_miss([-1.0, 0.5, 0.3])
        m.add_hit([1.02, 1.01, 0.0])  # same voxel: ignored
    assert_same_dump(gm.dump(), om.dump(), "after add")
    assert gm.active_count() == 2
    pts = np.array([[0.0, 2.0, 0.0], [1.0, 1.0, 0.0]], np.float32)  # second point hits the queued voxel: skipped
    gm.insert(pts, [0.5, 0.0, 0.0], 10.0)
    om.insert(pts, [0.5, 0.0, 0.0], 10.0)
    check_scan(gm, om, "insert after add", counters=False)
    gm.insert(pts, [0.5, 0.0, 0.0], 10.0)
    om.insert(pts, [0.5, 0.0, 0.0], 10.0)
    check_scan(gm, om, "second insert")


def test_queries_and_voxel_lists(bnx, port):
    pts, origin = synth.room_synth(5000)
    gm, om = bnx.ProbabilisticMap(0.05), port.map(0.05)
    gm.insert(pts, origin, 2.5)
    om.insert(pts, origin, 2.5)
    check_scan(gm, om, "room 5000")
    rng = np.random.default_rng(0)
    q = rng.integers(-70, 70, (20000, 3)).astype(np.int32)
    for kind in (bnx.BNX_OCCUPIED, bnx.BNX_UNKNOWN, bnx.BNX_FREE):
        assert np.array_equal(gm.query(q, kind), om.query(q, kind))
    occ = gm.get_voxels(bnx.BNX_OCCUPIED)
    assert np.array_equal(occ, om.get_voxels(0)) and len(occ) > 0
    assert np.array_equal(gm.get_voxels(bnx.BNX_FREE), om.get_voxels(2))
    pos = gm.get_voxel_points(bnx.BNX_OCCUPIED)  # coord * resolution (voxel corner)
    order = np.lexsort((pos[:, 2], pos[:, 1], pos[:, 0]))
    assert np.array_equal(pos[order], port.coord_to_pos(0.05, occ)[np.lexsort((occ[:, 2], occ[:, 1], occ[:, 0]))])


def test_apple_pcd(bnx, any_oracle):
    """config #1 on the reference's only real cloud (data/apple.pcd), res 0.02, origin (0,0,0)."""
    pts = np.load(os.path.join(GOLDEN, "apple_xyz_f32.npy"))
    for max_range in (float("inf"), 0.74):
        gm, om = bnx.ProbabilisticMap(0.02), any_oracle.map(0.02)
        for scan in range(3):
            gm.insert(pts, [0, 0, 0], max_range)
            om.insert(pts, [0, 0, 0], max_range)
            check_scan(gm, om, f"apple R={max_range} scan {scan}")


def test_room_synth_full(bnx, any_oracle):
    """config #1 stand-in: 50k-point synthetic room, res 0.02, R in {inf, 2.5}."""
    pts, origin = synth.room_synth()
    for max_range in (float("inf"), 2.5):
        gm, om = bnx.ProbabilisticMap(0.02), any_oracle.map(0.02)
        gm.insert(pts, origin, max_range)
        om.insert(pts, origin, max_range)
        check_scan(gm, om, f"room R={max_range}")


def test_lidar_sequence(bnx, port):
    """config #3 at full scan size: 131,072 points per scan, 0.1 m, 50 m; dump after every scan."""
    gm, om = bnx.ProbabilisticMap(0.1), port.map(0.1)
    for scan in range(6):
        pts, origin = synth.lidar_scan(scan)
        gm.insert(pts, origin, 50.0)
        om.insert(pts, origin, 50.0)
        check_scan(gm, om, f"lidar scan {scan}")
        assert gm.counters()["retries"] == 0, "the default pools must hold a LiDAR scan without a retry"


def test_non_finite_points_are_dropped(bnx, any_oracle):
    """NaN / inf coordinates are undefined behaviour in the reference ((int32)floor(NaN)); its callers filter them
    (bonxai_ros/src/bonxai_server.cpp:148-154). Here they leave the cloud inside the classify kernel, in every input
    flavour: the result equals the oracle fed with the filtered cloud (the order of the survivors is unchanged)."""
    rng = np.random.default_rng(17)
    for layout in ("f32x3", "f32x4", "f64"):
        pts = rng.normal(0, 2.0, (6000, 3))
        bad = rng.choice(6000, 300, replace=False)
        pts[bad[:100], 0] = np.nan
        pts[bad[100:200], 1] = np.inf
        pts[bad[200:], 2] = -np.inf
        pts[0] = pts[1]  # a duplicate voxel whose first point survives
        if layout == "f64":
            cloud = pts.copy()
        else:
            cloud = np.zeros((6000, 3 if layout == "f32x3" else 4), np.float32)
            cloud[:, :3] = pts
        good = np.isfinite(cloud[:, :3]).all(axis=1)
        gm, om = bnx.ProbabilisticMap(0.1), any_oracle.map(0.1)
        for k in range(2):
            gm.insert(cloud, [0.05 * k, 0, 0], 4.0)
            om.insert(np.ascontiguousarray(cloud[good]), [0.05 * k, 0, 0], 4.0)
            check_scan(gm, om, f"{layout} scan {k}")
        assert gm.counters()["N"] == int(good.sum())
        gm.insert_async(cloud, [0.2, 0, 0], 4.0)
        om.insert(np.ascontiguousarray(cloud[good]), [0.2, 0, 0], 4.0)
        check_scan(gm, om, f"{layout} pipelined", counters=False)


def test_refused_scan_leaves_the_map_clean(bnx, port):
    """a ray of more chunks than the scan counters can hold (max_range = inf and one far-away garbage point) is refused
    with BNX_ERR_UNSUPPORTED; the marks its first phases left behind must not leak into later scans, and scans queued
    behind it in the pipeline are still applied"""
    rng = np.random.default_rng(23)
    inf = float("inf")
    good = rng.normal(0, 2.0, (70000, 3)).astype(np.float32)
    bad = good.copy()
    bad[5] = [3.0e7, 1.0e7, 0.0]  # 3e8 cells at 0.1 m: more than 2^40 / 70000 chunks of 8 cells
    nxt = rng.normal(0, 2.0, (50000, 3)).astype(np.float32)
    empty = np.zeros((0, 3), np.float32)
    # synchronous call: the refused scan changes nothing but consumes its update id, like an empty scan
    gm, om = bnx.ProbabilisticMap(0.1), port.map(0.1)
    gm.insert(good, [0, 0, 0], inf)
    om.insert(good, [0, 0, 0], inf)
    with pytest.raises(bnx.BonxaiError) as err:
        gm.insert(bad, [0, 0, 0], inf)
    assert err.value.status == 5
    om.insert(empty, [0, 0, 0], 1.0)
    check_scan(gm, om, "right after the refused scan", counters=False)
    for k in range(3):
        gm.insert(nxt, [0.1 * k, 0, 0], 6.0)
        om.insert(nxt, [0.1 * k, 0, 0], 6.0)
        check_scan(gm, om, f"scan {k} after the refused one")
    # pipelined call: the refused scan has consumed its update id (an empty insert does the same to the oracle); the
    # scan queued behind it is applied; the error is reported by the next synchronising call
    gm, om = bnx.ProbabilisticMap(0.1), port.map(0.1)
    gm.insert_async(good, [0, 0, 0], inf)
    gm.insert_async(bad, [0, 0, 0], inf)
    gm.insert_async(nxt, [0.3, 0, 0], inf)
    with pytest.raises(bnx.BonxaiError) as err:
        gm.sync()
    assert err.value.status == 5
    om.insert(good, [0, 0, 0], inf)
    om.insert(empty, [0, 0, 0], 1.0)
    om.insert(nxt, [0.3, 0, 0], inf)
    check_scan(gm, om, "pipeline with a refused scan in the middle", counters=False)
    gm.insert(nxt, [0.1, 0, 0], 6.0)
    om.insert(nxt, [0.1, 0, 0], 6.0)
    check_scan(gm, om, "scan after the refused one (pipelined)")


def test_dense_window_moves_and_resizes(bnx, port):
    """the dense marking window follows the origin (far from 0, negative coordinates, jumps larger than the window) and is
    re-allocated when max_range grows; scans with max_range = inf in between use the leaf-resident marks"""
    rng = np.random.default_rng(31)
    gm, om = bnx.ProbabilisticMap(0.05), port.map(0.05)
    origins = [[0, 0, 0], [-37.3, 12.9, -3.3], [-37.1, 12.9, -3.3], [812.4, -955.6, 40.2], [812.4, -955.2, 40.2], [0.4, 0, 0]]
    ranges = [2.0, 3.0, float("inf"), 3.0, 7.5, 1.0]
    keep = []
    for k, (o, r) in enumerate(zip(origins, ranges)):
        pts = (rng.normal(0, 2.5, (30000, 3)) + o).astype(np.float32)
        keep.append(pts)
        if k % 2:
            gm.insert_async(pts, o, r)
        else:
            gm.insert(pts, o, r)
        om.insert(pts, np.float32(o), r)
        check_scan(gm, om, f"scan {k} origin {o} range {r}", counters=False)
    # exactly max_range away along an axis, and rays that start in the last cell of a leaf block
    for o in ([0.799, 0.799, 0.799], [-0.001, -0.001, -0.001]):
        pts = np.float32([[o[0] + 2.0, o[1], o[2]], [o[0], o[1] - 2.0, o[2]], [o[0], o[1], o[2] + 5.0], [o[0] - 1.99, o[1] + 0.3, o[2]]])
        gm.insert(pts, o, 2.0)
        om.insert(pts, np.float32(o), 2.0)
        check_scan(gm, om, f"boundary rays from {o}")


def test_depth_scan_reduced(bnx, port):
    """config #4 at reduced image size (320x200), 0.01 m, 5 m: long-ray heavy."""
    gm, om = bnx.ProbabilisticMap(0.01), port.map(0.01)
    for scan in range(2):
        pts, origin = synth.depth_scan(scan, width=320, height=200)
        gm.insert(pts, origin, 5.0)
        om.insert(pts, origin, 5.0)
        check_scan(gm, om, f"depth scan {scan}")


def test_pool_growth_retry_is_exact(bnx, port, monkeypatch):
    """a map created with tiny pools must grow (retrying the scan) and still be bit-exact."""
    monkeypatch.setenv("BNX_INIT_LEAF_MB", "1")
    monkeypatch.setenv("BNX_INIT_INNER_MB", "0")
    gm, om = bnx.ProbabilisticMap(0.1), port.map(0.1)
    retries = 0
    for scan in range(3):
        pts, origin = synth.lidar_scan(scan * 7)
        gm.insert(pts, origin, 50.0)
        om.insert(pts, origin, 50.0)
        retries += gm.counters()["retries"]
        check_scan(gm, om, f"growth scan {scan}")
    assert retries > 0


def test_device_resident_input(bnx, port):
    import torch
    pts, origin = synth.lidar_scan(3)
    t = torch.from_numpy(pts).cuda()
    gm, om = bnx.ProbabilisticMap(0.1), port.map(0.1)
    gm.set_stream(torch.cuda.current_stream().cuda_stream)
    gm.insert(bnx.DevPtr(t.data_ptr()), origin, 50.0, n=len(pts), stride_bytes=16)
    om.insert(pts, origin, 50.0)
    check_scan(gm, om, "device input")


def test_async_pipeline_equals_sync(bnx, port):
    """pipelined inserts (no host sync per scan), host and device inputs mixed, equal the oracle scan by scan in total"""
    import torch
    gm, om = bnx.ProbabilisticMap(0.1), port.map(0.1)
    keep, tot = [], dict(N=0, E=0, V=0, U=0)
    for scan in range(12):
        pts, origin = synth.lidar_scan(scan, beams=32, azimuths=1024)
        if scan % 3 == 0:
            t = torch.from_numpy(pts).cuda()
            keep.append(t)
            gm.insert_async(bnx.DevPtr(t.data_ptr()), origin, 45.0, n=len(pts), stride_bytes=16)
        else:
            keep.append(pts)
            gm.insert_async(pts, origin, 45.0)
        om.insert(pts, origin, 45.0)
        for k in tot:
            tot[k] += om.counters()[k]
    gm.sync()
    assert_same_dump(gm.dump(), om.dump(), "after 12 pipelined scans")
    assert gm.totals() == tot
    assert gm.update_count() == 1  # 12 scans: the counter cycled 4 times
    # a synchronous insert after the pipeline continues seamlessly
    pts, origin = synth.lidar_scan(12, beams=32, azimuths=1024)
    gm.insert(pts, origin, 45.0)
    om.insert(pts, origin, 45.0)
    check_scan(gm, om, "sync after async")


def test_async_pipeline_wraps_the_scratch_ring(bnx, port):
    """more scans in one queue than there are scratch sets (34): every set is reused, host and device inputs and two
    scan sizes alternate; the map must still equal the oracle bit for bit"""
    import torch
    gm, om = bnx.ProbabilisticMap(0.1), port.map(0.1)
    keep = []
    for scan in range(90):
        pts, origin = synth.lidar_scan(scan, beams=16, azimuths=256 if scan % 5 else 512)
        if scan % 2:
            t = torch.from_numpy(pts).cuda()
            keep.append(t)
            gm.insert_async(bnx.DevPtr(t.data_ptr()), origin, 30.0, n=len(pts), stride_bytes=16)
        else:
            keep.append(pts)
            gm.insert_async(pts, origin, 30.0)
        om.insert(pts, origin, 30.0)
    gm.sync()
    assert_same_dump(gm.dump(), om.dump(), "after 90 pipelined scans")
    assert gm.totals()["U"] > 0 and gm.update_count() == 1  # 90 scans: the counter cycled 30 times


def test_async_pipeline_freeze_and_replay(bnx, port, monkeypatch):
    """tiny pools: a queued scan runs short, the device freezes the pipeline, sync() grows and replays: still exact"""
    monkeypatch.setenv("BNX_INIT_LEAF_MB", "2")
    monkeypatch.setenv("BNX_INIT_INNER_MB", "0")
    gm, om = bnx.ProbabilisticMap(0.1), port.map(0.1)
    keep = []
    for scan in range(8):
        pts, origin = synth.lidar_scan(scan * 5, beams=32, azimuths=1024)
        keep.append(pts)
        gm.insert_async(pts, origin, 45.0)
        om.insert(pts, origin, 45.0)
        if scan == 4:  # a query in the middle of the queue drains it first
            q = np.random.default_rng(0).integers(-200, 200, (1000, 3)).astype(np.int32)
            assert np.array_equal(gm.query(q, bnx.BNX_OCCUPIED), om.query(q, 0))
    gm.sync()
    assert_same_dump(gm.dump(), om.dump(), "after freeze + replay")
    assert gm.grid().stats()["leaf_capacity"] > 910  # the 2 MB pool had to grow
