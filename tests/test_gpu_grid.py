"""GPU parity of the VoxelGrid<DataT> bulk operations (config #2) against the CPU oracles."""
import numpy as np
import pytest

from bonxai_b200 import synth
from conftest import assert_same_dump

pytestmark = pytest.mark.gpu


def test_constructor_validation(bnx):
    with pytest.raises(bnx.BonxaiError) as e:
        bnx.VoxelGrid(0.1, inner_bits=0)
    assert e.value.status == 1  # the reference throws std::runtime_error (bonxai.hpp:399-401)
    with pytest.raises(bnx.BonxaiError):
        bnx.VoxelGrid(0.1, leaf_bits=0)
    g = bnx.VoxelGrid(0.25, 3, 2, np.float32)
    assert g.info() == dict(voxel_size=0.25, inner_bits=3, leaf_bits=2, cell_bytes=4)
    assert g.active_count() == 0 and len(g.dump()[0]) == 0


def test_pos_to_coord_floor_semantics(bnx, any_oracle):
    rng = np.random.default_rng(1)
    for res in (0.1, 0.02, 0.37, 1.0):
        xyz = np.concatenate([rng.uniform(-50, 50, (5000, 3)),
                              np.arange(-300, 300).reshape(-1, 3) * res,          # exact multiples
                              np.nextafter(np.arange(-300, 300).reshape(-1, 3) * res, -np.inf)])
        g = bnx.VoxelGrid(res)
        got = g.pos_to_coord(xyz)
        assert np.array_equal(got, any_oracle.pos_to_coord(res, xyz))
        assert np.array_equal(g.coord_to_pos(got), any_oracle.coord_to_pos(res, got))


@pytest.mark.parametrize("bits", [(2, 3), (1, 1), (3, 2), (2, 4), (4, 3)])
def test_set_get_dump_random_with_duplicates(bnx, any_oracle, port, bits):
    ib, lb = bits
    if any_oracle.kind == "reference" and lb >= 4:
        any_oracle = port  # the reference itself double-frees on destruction for leaf_bits >= 4 (heap-backed Mask): the pinned port stands in
    n = 60_000
    xyz = synth.random_coords(n, seed=42 + ib)  # bounded cube -> ~20 % repeated coordinates
    vals = (np.arange(n) * 2654435761 % 2**32).astype(np.uint32)
    g, o = bnx.VoxelGrid(0.1, ib, lb), any_oracle.grid(0.1, ib, lb)
    assert np.array_equal(g.set_values(xyz, vals), o.set_values(xyz, vals))     # create
    assert_same_dump(g.dump(), o.dump(), "create")
    assert g.active_count() == o.active_count()
    assert np.array_equal(g.set_values(xyz[::-1], vals), o.set_values(xyz[::-1], vals))  # update, other order
    assert_same_dump(g.dump(), o.dump(), "update")
    q = np.concatenate([xyz[:5000], synth.random_coords(5000, seed=7) * 3])
    gv, gf = g.get_values(q)
    ov, of = o.get_values(q)
    assert np.array_equal(gf, of) and np.array_equal(gv[gf], ov[of])
    assert np.array_equal(g.is_on(q), o.is_on(q))


def test_coherent_cube_sweep(bnx, port):
    """README Create/Update/Iterate pattern on a dense cube, x-fastest and z-fastest."""
    n = 1 << 18
    for order in ("x", "z"):
        xyz = synth.coherent_coords(n, order)
        vals = synth.sweep_values(n)
        g, o = bnx.VoxelGrid(0.05, dtype=np.float32), port.grid(0.05)
        was = g.set_values(xyz, vals)
        assert not was.any() and not o.set_values(xyz, vals).any()
        assert g.set_values(xyz, vals + 1).all()
        o.set_values(xyz, vals + 1)
        gx, gv = g.dump()
        ox, ov = o.dump()
        assert np.array_equal(gx, ox) and np.array_equal(gv.view(np.uint32), ov)
        v, f = g.get_values(xyz)
        assert f.all() and np.array_equal(v, vals + 1)


def test_set_on_off_get_or_create_semantics(bnx, any_oracle):
    rng = np.random.default_rng(3)
    g, o = bnx.VoxelGrid(0.1), any_oracle.grid(0.1)
    a = rng.integers(-40, 40, (20000, 3)).astype(np.int32)
    b = rng.integers(-40, 40, (20000, 3)).astype(np.int32)
    va = rng.integers(1, 2**31, 20000).astype(np.uint32)
    assert np.array_equal(g.set_values(a, va), o.set_values(a, va))
    assert np.array_equal(g.set_off(b), o.set_off(b))                      # keeps values, repeated coords
    assert_same_dump(g.dump(), o.dump(), "after set_off")
    assert np.array_equal(g.set_on(b[:9000], 77), o.set_on(b[:9000], 77))  # default only where it was off
    assert_same_dump(g.dump(), o.dump(), "after set_on")
    c = rng.integers(-45, 45, (20000, 3)).astype(np.int32)
    assert np.array_equal(g.get_or_create(c), o.get_or_create(c))           # OFF cells are re-created as 0
    assert_same_dump(g.dump(), o.dump(), "after get_or_create")
    g.update_values(c[:100], np.arange(100, dtype=np.uint32) + 5)
    o.set_values(c[:100], np.arange(100, dtype=np.uint32) + 5)
    assert_same_dump(g.dump(), o.dump(), "after update_values")
    missing = np.array([[1000, 1000, 1000]], np.int32)
    g.update_values(missing, np.array([9], np.uint32))                      # no cell there: nothing happens
    assert g.active_count() == o.active_count()
    assert not g.set_off(missing)[0] and not o.set_off(missing)[0]


def test_clear_and_release(bnx, port):
    rng = np.random.default_rng(9)
    xyz = rng.integers(-100, 100, (30000, 3)).astype(np.int32)
    vals = rng.integers(0, 2**32, 30000, dtype=np.uint64).astype(np.uint32)
    g, o = bnx.VoxelGrid(0.1), port.grid(0.1)
    g.set_values(xyz, vals)
    o.set_values(xyz, vals)
    leaves_before = g.stats()["leaves"]
    half = xyz[:15000]
    assert np.array_equal(g.set_off(half), o.set_off(half))
    g.release_unused()
    o.release_unused()
    assert_same_dump(g.dump(), o.dump(), "after release")
    st = g.stats()
    assert st["leaves"] < leaves_before and st["free_leaves"] > 0
    again = rng.integers(-100, 100, (10000, 3)).astype(np.int32)  # recycled leaves must behave like new ones
    assert np.array_equal(g.get_or_create(again), o.get_or_create(again))
    assert_same_dump(g.dump(), o.dump(), "after reuse")
    g.clear(bnx.BNX_SET_ALL_CELLS_OFF)
    o.clear(1)
    assert g.active_count() == 0 == o.active_count()
    assert np.array_equal(g.get_or_create(again[:100]), o.get_or_create(again[:100]))
    g.clear(bnx.BNX_CLEAR_MEMORY)
    o.clear(0)
    assert g.active_count() == 0 and g.stats()["leaves"] == 0
    g.set_values(xyz, vals)
    o.set_values(xyz, vals)
    assert_same_dump(g.dump(), o.dump(), "after clear+refill")


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.float64, np.dtype([("a", "<u4"), ("b", "<u4"), ("c", "<u4"), ("d", "<u4")])])
def test_other_cell_sizes(bnx, dtype):
    dtype = np.dtype(dtype)
    rng = np.random.default_rng(4)
    n = 20000
    xyz = rng.permutation(np.stack(np.meshgrid(np.arange(-20, 20), np.arange(-10, 15), np.arange(20), indexing="ij"), -1).reshape(-1, 3))[:n].astype(np.int32)
    raw = rng.integers(0, 256, (n, dtype.itemsize), dtype=np.uint8)
    vals = raw.view(dtype).reshape(n)
    g = bnx.VoxelGrid(0.1, dtype=dtype)
    assert not g.set_values(xyz, vals).any()
    got, found = g.get_values(xyz)
    assert found.all() and np.array_equal(got.view(np.uint8).reshape(n, -1), raw)
    dx, dv = g.dump()
    order = np.lexsort((xyz[:, 2], xyz[:, 1], xyz[:, 0]))
    assert np.array_equal(dx, xyz[order]) and np.array_equal(dv.view(np.uint8).reshape(n, -1), raw[order])


def test_growth_from_tiny_pools(bnx, port, monkeypatch):
    monkeypatch.setenv("BNX_INIT_LEAF_MB", "1")
    monkeypatch.setenv("BNX_INIT_INNER_MB", "0")
    n = 200_000
    xyz = (synth.random_coords(n, seed=11).astype(np.int64) * 40).astype(np.int32)  # sparse: ~one leaf and root per point
    vals = np.arange(n, dtype=np.uint32)
    g, o = bnx.VoxelGrid(0.1), port.grid(0.1)
    assert np.array_equal(g.set_values(xyz, vals), o.set_values(xyz, vals))
    assert_same_dump(g.dump(), o.dump(), "grown")
    assert g.stats()["root_slots"] > (1 << 14)


def test_large_batch_crosses_sub_batches(bnx, port):
    n = (1 << 22) + 12345  # more than one dedupe sub-batch
    xyz = synth.coherent_coords(n, "x")
    xyz[-5000:] = xyz[:5000]  # repeats across the sub-batch boundary
    vals = synth.sweep_values(n)
    g, o = bnx.VoxelGrid(0.1, dtype=np.float32), port.grid(0.1)
    assert np.array_equal(g.set_values(xyz, vals), o.set_values(xyz, vals))
    assert g.active_count() == o.active_count() == n - 5000
    v, f = g.get_values(xyz[:5000])
    assert f.all() and np.array_equal(v, vals[-5000:])
