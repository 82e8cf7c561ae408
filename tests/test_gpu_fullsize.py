"""Full-size parity of the BASELINE.json configurations, kept affordable by comparing order-independent
digests computed on the device (bnx_grid_digest) with the digest of the CPU oracle's dump, plus sorted dumps at selected
scans. The oracle is the unmodified reference (oracle/_ref) where it is present, else the plain-C port pinned to it."""
import os

import numpy as np
import pytest

from bonxai_b200 import synth
from conftest import assert_same_dump

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def best():
    import oracle
    if os.path.isdir(os.path.join(oracle.REFERENCE_TREE, "bonxai_core")):
        oracle.build("reference")
    if oracle.available("reference"):
        return oracle.load("reference")
    oracle.build("port")
    return oracle.load("port")


def test_device_digest_equals_digest_of_dump(bnx):
    """bnx_grid_digest == the numpy spelling over the grid's own dump, for 4-byte map cells and other cell sizes"""
    rng = np.random.default_rng(21)
    m = bnx.ProbabilisticMap(0.1)
    assert m.digest() == (0, 0, 0)
    for k in range(3):
        m.insert(rng.normal(0, 3, (20000, 3)).astype(np.float32), [0.1 * k, 0, 0], 8.0)
        assert m.digest() == bnx.digest_of_dump(*m.dump())
    for dtype, bits in ((np.uint16, (2, 3)), (np.float64, (1, 2)), (np.uint8, (3, 2))):
        g = bnx.VoxelGrid(0.05, *bits, dtype=dtype)
        xyz = rng.integers(-300, 300, (30000, 3)).astype(np.int32)
        g.set_values(xyz, (np.arange(30000) % 251).astype(dtype))
        assert g.digest() == bnx.digest_of_dump(*g.dump())
        g.set_off(xyz[:5000])
        assert g.digest() == bnx.digest_of_dump(*g.dump())


def test_config3_lidar_200_scans(bnx, best):
    """config #3 at full size: 131,072 points per scan, 0.1 m, 50 m, 200 scans. Digest after EVERY one of the first 16
    scans, every 4th up to 63, every 20th up to 199; sorted dumps at scans 0-3, 8, 16, 32 and 63; the last 136 scans go
    through the pipelined call."""
    gm, om = bnx.ProbabilisticMap(0.1), best.map(0.1)
    dumps = {0, 1, 2, 3, 8, 16, 32, 63}
    keep = []
    for scan in range(200):
        pts, origin = synth.lidar_scan(scan)
        if scan < 64:
            gm.insert(pts, origin, 50.0)
        else:
            keep.append(pts)
            gm.insert_async(pts, origin, 50.0)
        om.insert(pts, origin, 50.0)
        check = scan < 16 or (scan < 64 and scan % 4 == 3) or scan % 20 == 19
        if scan in dumps:
            assert_same_dump(gm.dump(), om.dump(), f"lidar scan {scan}")
        elif check:
            assert gm.digest() == om.digest(), f"lidar scan {scan}: digest differs from the {best.kind} oracle"
        if scan < 64:
            assert gm.counters()["retries"] == 0, f"scan {scan} needed a retry with the default pools"
    assert gm.active_count() == om.active_count()


def test_config4_depth_full_size(bnx, best):
    """config #4 at full size: 1280 x 800 = 1,024,000 points per scan, 0.01 m, 5 m (rays up to 500 cells)."""
    gm, om = bnx.ProbabilisticMap(0.01), best.map(0.01)
    for scan in range(2):
        pts, origin = synth.depth_scan(scan)
        assert len(pts) == 1280 * 800
        gm.insert(pts, origin, 5.0)
        om.insert(pts, origin, 5.0)
        assert gm.digest() == om.digest(), f"depth scan {scan}: digest differs from the {best.kind} oracle"
    assert gm.active_count() == om.active_count()
    occ_g, occ_o = gm.get_voxels(bnx.BNX_OCCUPIED), om.get_voxels(0)
    assert np.array_equal(occ_g, occ_o)


@pytest.mark.parametrize("kind", ["coherent_x", "coherent_z", "random"])
def test_config2_sweep_2p24_dump_parity(bnx, best, kind):
    """config #2: VoxelGrid<float> setValue / value / forEachCell over 2^24 coordinates against the oracle: create, update
    (second pass with other values), read back, full digest; sorted dump on a 2^20 prefix grid is covered in test_gpu_grid."""
    n = 1 << 24
    xyz = synth.random_coords(n) if kind == "random" else synth.coherent_coords(n, "x" if kind == "coherent_x" else "z")
    vals = synth.sweep_values(n)
    g, o = bnx.VoxelGrid(0.1, dtype=np.float32), best.grid(0.1)
    was_g = g.set_values(xyz, vals)
    was_o = o.set_values(xyz, vals.view(np.uint32))
    assert np.array_equal(was_g, was_o)
    assert g.active_count() == o.active_count()
    assert g.digest() == o.digest(), "create"
    vals2 = (vals + 1.0).astype(np.float32)
    g.set_values(xyz, vals2)
    o.set_values(xyz, vals2.view(np.uint32))
    assert g.digest() == o.digest(), "update"
    idx = np.random.default_rng(1).integers(0, n, 1 << 20)
    got, found = g.get_values(xyz[idx])
    want, found_o = o.get_values(xyz[idx])
    assert found.all() and found_o.all() and np.array_equal(got.view(np.uint32), want)


@pytest.mark.parametrize("kind", ["coherent_x", "random"])
def test_config2_sweep_2p27_digest_and_count(bnx, best, kind):
    """config #2 at 2^27 coordinates (1.6 GB of coordinates through the C ABI in sub-batches): digest + count against the
    oracle; duplicates inside the random batch resolve as "last index wins" on both sides."""
    n = 1 << 27
    if kind == "random":  # same distribution as synth.random_coords (bounded cube of side ceil((2n)^(1/3))), cheaper generator
        side = int(np.ceil((2.0 * n) ** (1.0 / 3.0)))
        xyz = np.random.default_rng(42).integers(-(side // 2), side - side // 2, (n, 3), dtype=np.int32)
    else:
        xyz = synth.coherent_coords(n, "x")
    vals = synth.sweep_values(n)
    g, o = bnx.VoxelGrid(0.1, dtype=np.float32), best.grid(0.1)
    g.set_values(xyz, vals)
    o.set_values(xyz, vals.view(np.uint32))
    assert g.active_count() == o.active_count()
    assert g.digest() == o.digest()
