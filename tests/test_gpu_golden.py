"""GPU path vs the golden digests generated from the UNMODIFIED reference (tests/golden/map_golden.json):
after every scan of every golden workload the order-independent digest of the forEachCell dump must match."""
import json
import os
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import workloads as W  # noqa: E402

GOLDEN = json.load(open(os.path.join(HERE, "golden", "map_golden.json")))


@pytest.mark.parametrize("name", sorted(GOLDEN["digests"]))
def test_gpu_matches_reference_golden(bnx, name):
    got = W.run(bnx.ProbabilisticMap, {name})[name]
    assert got == GOLDEN["digests"][name], name
