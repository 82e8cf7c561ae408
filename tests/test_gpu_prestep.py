"""Row f-2: the ROS caller's pre-step (bonxai_ros/src/bonxai_server.cpp:148-171) fused into the classify kernel:
non-finite filter + float 4x4 sensor->world transform. PCL is not vendored by the reference, so the transform is
pinned to a float32 restatement with the association of pcl::transformPointCloud's SSE kernel (stated in
include/bonxai_b200.h); everything after the transform is checked against the oracle as usual."""
import numpy as np
import pytest

from bonxai_b200 import synth
from conftest import assert_same_dump

pytestmark = pytest.mark.gpu


def restate_prestep(pts, T):
    p = pts[:, :3]
    keep = np.isfinite(p).all(axis=1)
    p = p[keep]
    T = T.astype(np.float32)
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    out = np.empty_like(p)
    for r in range(3):  # x*c0 + (y*c1 + (z*c2 + c3)) in float32, one rounding per operation
        out[:, r] = x * T[r, 0] + (y * T[r, 1] + (z * T[r, 2] + T[r, 3]))
    return out


@pytest.mark.parametrize("use_async", [False, True])
def test_fused_prestep_equals_filter_transform_insert(bnx, port, use_async):
    rng = np.random.default_rng(8)
    gm, om = bnx.ProbabilisticMap(0.1), port.map(0.1)
    keep = []
    for scan in range(4):
        world, origin = synth.lidar_scan(scan, beams=32, azimuths=512)
        yaw = 0.3 * scan + 0.1
        R = np.array([[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1]])
        T = np.eye(4)
        T[:3, :3] = R
        T[:3, 3] = origin
        T = T.astype(np.float32)
        sensor = ((world[:, :3] - origin) @ R).astype(np.float32)      # points as the sensor saw them
        sensor = np.concatenate([sensor, np.ones((len(sensor), 1), np.float32)], axis=1)
        bad = rng.choice(len(sensor), 500, replace=False)
        sensor[bad[:200], 0] = np.nan
        sensor[bad[200:350], 1] = np.inf
        sensor[bad[350:], 2] = -np.inf
        keep.append(sensor)
        gm.insert_transformed(sensor, T, origin, 40.0, use_async=use_async)
        cpu_pts = restate_prestep(sensor, T)
        om.insert(cpu_pts, origin, 40.0)
        if not use_async:
            assert_same_dump(gm.dump(), om.dump(), f"scan {scan}")
            gc, oc = gm.counters(), om.counters()
            assert (gc["N"], gc["E"], gc["V"], gc["U"]) == (oc["N"], oc["E"], oc["V"], oc["U"])
            assert gc["N"] == len(sensor) - 500
    gm.sync()
    assert_same_dump(gm.dump(), om.dump(), "final")
