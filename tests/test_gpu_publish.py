"""Row f-3: the ROS node's publisher post-step (bonxai_ros/src/bonxai_server.cpp:217-251) fused into the
compaction kernel: occupied voxels -> coord*resolution (fp64) -> z window -> float xyz."""
import numpy as np
import pytest

from bonxai_b200 import synth

pytestmark = pytest.mark.gpu


def _restate(om, res, z_min, z_max):
    occ = om.get_voxels(0)
    pts = occ.astype(np.float64) * res            # Point3D coordToPos (voxel corner), bonxai.hpp:412-417
    keep = (pts[:, 2] >= z_min) & (pts[:, 2] <= z_max)
    return pts[keep].astype(np.float32)             # PCLPoint(voxel.x(), voxel.y(), voxel.z())


def _sorted(a):
    return a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))] if len(a) else a


@pytest.mark.parametrize("stride", [3, 4])
def test_publish_occupied_matches_restated_post_step(bnx, port, stride):
    res = 0.1
    gm, om = bnx.ProbabilisticMap(res), port.map(res)
    for scan in range(3):
        pts, origin = synth.lidar_scan(scan, beams=32, azimuths=1024)
        gm.insert(pts, origin, 40.0)
        om.insert(pts, origin, 40.0)
    for z_min, z_max in ((-100.0, 100.0), (0.05, 2.0), (0.3, 0.3), (5.0, 1.0)):
        got = gm.publish_occupied(z_min, z_max, stride)
        want = _restate(om, res, z_min, z_max)
        assert got.shape == (len(want), stride)
        assert np.array_equal(_sorted(got[:, :3]), _sorted(want))
        if stride == 4 and len(got):
            assert (got[:, 3] == 1.0).all()
