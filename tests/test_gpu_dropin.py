"""The C++ drop-in headers (include/bonxai/bonxai.hpp, include/bonxai_map/probabilistic_map.hpp): one caller
program, built once against the reference (expected output committed under tests/golden/) and once against
this repo's headers + CUDA library. Outputs must match line for line."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_dropin_program_matches_reference_build(bnx, tmp_path):
    exe = tmp_path / "dropin_b200"
    subprocess.run(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests/cpp/dropin_program.cpp"),
                    "-L", os.path.join(ROOT, "bonxai_b200"), "-lbonxai_b200", "-Wl,-rpath," + os.path.join(ROOT, "bonxai_b200"), "-o", str(exe)],
                   check=True, capture_output=True)
    got = subprocess.run([str(exe)], check=True, capture_output=True, text=True, timeout=600).stdout.splitlines()
    with open(os.path.join(ROOT, "tests/golden/dropin_expected.txt")) as f:
        want = f.read().splitlines()
    assert got == want, "\n".join(f"{'==' if g == w else '!='} got: {g}\n   want: {w}" for g, w in zip(got, want))
