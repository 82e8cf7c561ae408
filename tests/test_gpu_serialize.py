"""Row f-1: Serialize / Deserialize of a device-resident grid in the reference's stream format
(bonxai_core/include/bonxai/serialization.hpp:77-199). Streams must be interchangeable with the reference."""
import struct

import numpy as np
import pytest

from conftest import assert_same_dump

pytestmark = pytest.mark.gpu


def _fill(rng, n=20000, span=60):
    xyz = rng.integers(-span, span, (n, 3)).astype(np.int32)
    vals = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    return xyz, vals


@pytest.mark.parametrize("bits", [(2, 3), (1, 2), (3, 2)])
def test_gpu_stream_loads_in_reference_and_back(bnx, ref, bits):
    import oracle
    rng = np.random.default_rng(sum(bits))
    xyz, vals = _fill(rng)
    g, o = bnx.VoxelGrid(0.1, *bits), ref.grid(0.1, *bits)
    g.set_values(xyz, vals)
    o.set_values(xyz, vals)
    g.set_off(xyz[:3000])          # OFF cells are not serialized, their leaves keep mask words
    o.set_off(xyz[:3000])
    blob_g, blob_o = g.serialize("unsigned int"), o.serialize()
    head_g, head_o = blob_g.split(b"\n", 1)[0], blob_o.split(b"\n", 1)[0]
    assert head_g == head_o == f"Bonxai::VoxelGrid<unsigned int,{bits[0]},{bits[1]}>(0.100000)".encode()
    assert len(blob_g) == len(blob_o)                                    # same content, root order may differ
    assert struct.unpack_from("<I", blob_g, len(head_g) + 1) == struct.unpack_from("<I", blob_o, len(head_o) + 1)
    back_in_ref = oracle.OracleGrid.deserialize(ref, blob_g)             # GPU stream -> reference Deserialize
    assert_same_dump(back_in_ref.dump(), o.dump(), "gpu stream in reference")
    from_ref = bnx.VoxelGrid.deserialize(blob_o, np.uint32, "unsigned int")   # reference stream -> GPU Deserialize
    assert_same_dump(from_ref.dump(), g.dump(), "reference stream on gpu")
    assert from_ref.info()["inner_bits"] == bits[0] and from_ref.info()["leaf_bits"] == bits[1]
    round_trip = bnx.VoxelGrid.deserialize(blob_g, np.uint32, "unsigned int")
    assert_same_dump(round_trip.dump(), g.dump(), "gpu round trip")


def test_map_grid_serializes_and_empty_grid(bnx, port):
    from bonxai_b200 import synth
    m = bnx.ProbabilisticMap(0.1)
    pts, origin = synth.lidar_scan(0, beams=16, azimuths=512)
    m.insert(pts, origin, 30.0)
    blob = m.grid().serialize("Bonxai::ProbabilisticMap::CellT")
    back = bnx.VoxelGrid.deserialize(blob, np.uint32, "Bonxai::ProbabilisticMap::CellT")
    assert_same_dump(back.dump(), m.dump(), "map cells round trip")
    e = bnx.VoxelGrid(0.25, dtype=np.float32)
    blob = e.serialize("float")
    assert blob == b"Bonxai::VoxelGrid<float,2,3>(0.250000)\n" + struct.pack("<I", 0)
    assert bnx.VoxelGrid.deserialize(blob, np.float32, "float").active_count() == 0
    with pytest.raises(bnx.BonxaiError):
        bnx.VoxelGrid.deserialize(blob, np.float32, "int")               # "DataT does not match"
    with pytest.raises(bnx.BonxaiError):
        bnx.VoxelGrid.deserialize(b"garbage\n\0\0\0\0", np.float32, "float")
