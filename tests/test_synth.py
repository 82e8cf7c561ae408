"""CPU suite: the synthetic workload generators are deterministic and have the documented shape."""
import numpy as np

from bonxai_b200 import synth


def test_lidar_scan_shape_and_determinism():
    p, o = synth.lidar_scan(3)
    p2, o2 = synth.lidar_scan(3)
    assert p.shape == (131072, 4) and p.dtype == np.float32 and np.array_equal(p, p2) and np.array_equal(o, o2)
    assert np.allclose(o, [3.0, 0.0, 1.8])
    r = np.linalg.norm(p[:, :3] - o, axis=1)
    assert 500 < (r >= 50.0).sum() < 20000  # some no-return / far beams become truncated miss rays
    assert not np.array_equal(synth.lidar_scan(4)[0], p)


def test_depth_scan_and_room():
    p, o = synth.depth_scan(0, width=64, height=40)
    assert p.shape == (2560, 4) and np.isfinite(p).all()
    room, ro = synth.room_synth(1000)
    assert room.shape == (1000, 3) and np.abs(room).max() < 3.2


def test_sweep_coords():
    c = synth.coherent_coords(1000, "x")
    assert len(np.unique(c, axis=0)) == 1000
    r = synth.random_coords(1000)
    assert r.shape == (1000, 3) and np.abs(r).max() <= 7
    assert synth.sweep_values(70000)[65536] == 0.0
