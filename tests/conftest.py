import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def port():
    """plain-C restatement oracle (oracle/bonxai_oracle.c)"""
    import oracle
    oracle.build("port")
    return oracle.load("port")


@pytest.fixture(scope="session")
def ref():
    """the unmodified reference behind the oracle API (oracle/_ref); built here, prebuilt on the GPU box"""
    import oracle
    if os.path.isdir(os.path.join(oracle.REFERENCE_TREE, "bonxai_core")):
        oracle.build("reference")
    if not oracle.available("reference"):
        pytest.skip("oracle/_ref/libbonxai_ref.so not available")
    return oracle.load("reference")


@pytest.fixture(scope="session", params=["port", "reference"])
def any_oracle(request):
    import oracle
    kind = request.param
    if kind == "reference":
        if os.path.isdir(os.path.join(oracle.REFERENCE_TREE, "bonxai_core")):
            oracle.build("reference")
        if not oracle.available("reference"):
            pytest.skip("oracle/_ref/libbonxai_ref.so not available")
    else:
        oracle.build("port")
    return oracle.load(kind)


@pytest.fixture(scope="session")
def bnx():
    """the CUDA library through its C ABI; no fallback — fails if it is not built or there is no GPU"""
    from bonxai_b200 import capi
    capi.load_library()
    assert capi.device_count() > 0, "no CUDA device"
    return capi


def assert_same_dump(a, b, what=""):
    (xa, va), (xb, vb) = a, b
    assert len(xa) == len(xb), f"{what}: {len(xa)} vs {len(xb)} active cells"
    if not np.array_equal(xa, xb):
        bad = np.nonzero((xa != xb).any(axis=1))[0][:5]
        raise AssertionError(f"{what}: coordinates differ first at rows {bad}: {xa[bad]} vs {xb[bad]}")
    va = np.asarray(va).view(np.uint32) if np.asarray(va).dtype.itemsize == 4 else np.asarray(va)
    vb = np.asarray(vb).view(np.uint32) if np.asarray(vb).dtype.itemsize == 4 else np.asarray(vb)
    if not np.array_equal(va, vb):
        bad = np.nonzero(va != vb)[0][:5]
        raise AssertionError(f"{what}: values differ at {xa[bad]}: {va[bad]} vs {vb[bad]}")
