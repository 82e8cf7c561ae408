"""Converts the reference's only real point cloud, /root/reference/data/apple.pcd (ASCII PCD, fields
x y z rgb imX imY), to a compact float32 xyz fixture. Run in the build container (the GPU box has no
/root/reference):   python tests/golden/make_apple.py
"""
import os

import numpy as np

SRC = "/root/reference/data/apple.pcd"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "apple_xyz_f32.npy")


def read_ascii_pcd_xyz(path):
    with open(path) as f:
        lines = f.read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith("DATA")) + 1
    assert lines[start - 1].strip() == "DATA ascii"
    npts = int(next(l for l in lines if l.startswith("POINTS")).split()[1])
    xyz = np.array([[np.float32(t) for t in l.split()[:3]] for l in lines[start:start + npts]], dtype=np.float32)
    assert xyz.shape == (npts, 3)
    return xyz


if __name__ == "__main__":
    xyz = read_ascii_pcd_xyz(SRC)
    np.save(DST, xyz)
    print(DST, xyz.shape, xyz.min(0), xyz.max(0))
