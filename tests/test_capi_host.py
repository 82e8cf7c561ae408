"""CPU suite: the C-ABI library loads without a GPU, exports exactly what include/bonxai_b200.h declares, and
refuses to compute without a device (no CPU fallback). The drop-in C++ headers compile."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "bonxai_b200.h")).read()
    return re.findall(r"BNX_API\s+[\w\s\*]+?\b(bnx_\w+)\s*\(", text)


def test_library_exports_every_declared_symbol():
    from bonxai_b200 import build, capi
    build.build()
    lib = capi.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 45 and len(set(declared)) == len(declared)
    assert sorted(declared) == sorted(capi.SYMBOLS), "capi.SYMBOLS and the header disagree"
    for name in declared:
        assert getattr(lib, name) is not None
    out = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert set(declared) <= exported
    assert not [s for s in exported if not s.startswith("bnx_")], "only the C ABI may be exported"
    assert lib.bnx_version() >= 100


def test_no_torch_or_cpp_types_in_signatures():
    text = open(os.path.join(ROOT, "include", "bonxai_b200.h")).read()
    code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)  # declarations only
    assert "std::" not in code and "torch" not in code and "at::" not in code and "&" not in code


def test_library_has_no_link_time_gpu_dependency():
    from bonxai_b200 import capi
    out = subprocess.run(["ldd", capi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    assert "libcuda.so" not in out and "libcudart" not in out and "libtorch" not in out


def test_fails_loudly_without_a_device():
    """no CUDA device in the build container: creation must fail with BNX_ERR_CUDA, never fall back to a CPU path"""
    from bonxai_b200 import capi
    if capi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(capi.BonxaiError) as e:
        capi.ProbabilisticMap(0.1)
    assert e.value.status == 2
    with pytest.raises(capi.BonxaiError):
        capi.VoxelGrid(0.1)
    # argument validation happens before any device work
    lib = capi.load_library()
    h = C.c_void_p()
    assert lib.bnx_grid_create(C.c_double(0.1), 0, 3, 4, C.byref(h)) == 1
    assert b"inner_bits" in lib.bnx_last_error()


def test_product_never_touches_the_oracle():
    """the oracle is test infrastructure: nothing under bonxai_b200/ or include/ may reference it"""
    for base in ("bonxai_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                    text = open(os.path.join(dirpath, f), errors="replace").read()
                    assert "import oracle" not in text and "libbonxai_oracle" not in text and "libbonxai_ref" not in text, os.path.join(dirpath, f)


def test_dropin_headers_compile(tmp_path):
    exe = tmp_path / "dropin"
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests/cpp/dropin_program.cpp"),
                        "-L", os.path.join(ROOT, "bonxai_b200"), "-lbonxai_b200", "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_host_ray_iterator_matches_oracle(port, tmp_path):
    """include/bonxai_map/probabilistic_map.hpp's host-side ComputeRay (closed form) == the reference walk"""
    src = tmp_path / "ray.cpp"
    src.write_text('#include <cstdio>\n#include "bonxai_map/probabilistic_map.hpp"\n'
                   'int main(int c, char** v) { Bonxai::CoordT a{atoi(v[1]),atoi(v[2]),atoi(v[3])}, b{atoi(v[4]),atoi(v[5]),atoi(v[6])};'
                   ' std::vector<Bonxai::CoordT> r; Bonxai::ComputeRay(a, b, r); for (auto& p : r) std::printf("%d %d %d\\n", p.x, p.y, p.z); }\n')
    exe = tmp_path / "ray"
    subprocess.run(["g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), str(src), "-L", os.path.join(ROOT, "bonxai_b200"), "-lbonxai_b200",
                    "-Wl,-rpath," + os.path.join(ROOT, "bonxai_b200"), "-o", str(exe)], check=True, capture_output=True)
    import numpy as np
    for a, b in (((0, 0, 0), (12, 6, 80)), ((5, -3, 2), (-40, 17, -9)), ((1, 1, 1), (1, 1, 1)), ((-7, 3, 0), (-7, 3, 5))):
        out = subprocess.run([str(exe), *map(str, a + b)], capture_output=True, text=True, check=True).stdout
        got = np.array([[int(t) for t in l.split()] for l in out.splitlines()], np.int32).reshape(-1, 3)
        assert np.array_equal(got, port.compute_ray(a, b))


def test_dropin_bench_tool_compiles(tmp_path):
    """tools/cpp/dropin_bench.cpp (the caller bench.py times for e2e_dropin / e2e_insert_publish) builds against include/"""
    exe = tmp_path / "dropin_bench"
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tools/cpp/dropin_bench.cpp"),
                        "-L", os.path.join(ROOT, "bonxai_b200"), "-lbonxai_b200", "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_node_object_api_fails_to_compile_with_a_message(tmp_path):
    """rootMap() & co hand out host node objects upstream (bonxai.hpp:156-161,233-243); here a caller that uses them gets a
    compile-time message naming the replacement instead of 'no member named rootMap'"""
    src = tmp_path / "uses_rootmap.cpp"
    src.write_text('#include "bonxai/bonxai.hpp"\nint main() { Bonxai::VoxelGrid<int> g(0.1); g.rootMap(); return 0; }\n')
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)], capture_output=True, text=True)
    assert r.returncode != 0 and "the nodes live in device pools" in r.stderr, r.stderr[-2000:]
    ok = tmp_path / "plain.cpp"
    ok.write_text('#include "bonxai/bonxai.hpp"\nint main() { Bonxai::VoxelGrid<int> g(0.1); return (int)g.activeCellsCount(); }\n')
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(ok)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
