"""bench.py contract pieces that run without a GPU: the reference arm (the reference's own CPU insertPointCloud,
timed on the host) prints one JSON line with the keys the driver reads, and bounds its sample."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*extra, env=None):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *extra], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, **(env or {})))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run("--steps", "2", "--warmup", "1")
    assert d["impl"] == "reference" and d["unit"] == "points/s" and d["higher_is_better"] is True
    assert d["metric"] == "insertPointCloud points/sec" and d["n_gpus"] == 1 and d["steps_timed"] == 2
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] == 1
    assert d["e2e"] == {"value": d["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert 1e5 < d["value"] < 1e8  # a single host core does about 1-2 M points/s on this workload


def test_reference_arm_other_ranks_stay_silent():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_reference_arm_fleet_step_at_two_gpus():
    """N > 1: the default workload is the fleet step — the reference arm inserts the N vehicles' scans one after the other
    and names the same workload as the GPU arm"""
    sys.path.insert(0, ROOT)
    import bench
    d = _run("--gpus", "2", "--steps", "1", "--warmup", "0")
    assert d["n_gpus"] == 2 and d["config"]["points_per_scan"] == 2 * 131072
    assert d["config"]["workload"] == bench.sharded_workload_name(2, "fleet") and "fleet of 2" in d["config"]["workload"]
    d2 = _run("--gpus", "2", "--steps", "1", "--warmup", "0", "--workload", "dense-scan")
    assert d2["config"]["workload"] == bench.sharded_workload_name(2, "dense-scan") and d2["config"]["points_per_scan"] == 2 * 131072


def test_fleet_scans_are_disjoint_streets():
    sys.path.insert(0, ROOT)
    import numpy as np
    import bench
    a, oa = bench.fleet_scan((3, 0))
    b, ob = bench.fleet_scan((3, 2))
    assert a.shape == b.shape == (131072, 4) and a.dtype == np.float32
    assert np.allclose(ob - oa, [0.0, 2 * bench.FLEET_SPACING, 0.0])
    # reach balls of neighbouring vehicles never overlap: spacing > 2 * max_range + the library's 40-voxel margin
    assert bench.FLEET_SPACING > 2 * bench.MAX_RANGE + 40 * bench.RES
    assert not np.array_equal(a[:, :3] + [0.0, 2 * bench.FLEET_SPACING, 0.0], b[:, :3])  # other seed: other buildings
