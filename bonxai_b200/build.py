"""Builds bonxai_b200/libbonxai_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libbonxai_b200.so")
SOURCES = ["arena.cu", "grid.cu", "map.cu", "capi.cu"]

NVCC_FLAGS = [
    "-std=c++17", "-O3",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "--fmad=false",  # fp64 classification must round once per operation (no FMA contraction)
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "bonxai_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    objdir = os.path.join(ROOT, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    objs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [_nvcc(), *NVCC_FLAGS, "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd += ["-Xptxas", "-v"]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
    link = [_nvcc(), "-shared", "-cudart", "static", "-Xlinker", "-ldl", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs]
    subprocess.run(link, check=True)
    return LIB


def build_variant(tag: str, defines: list[str]) -> str:
    """a tuning variant of the library (extra -D flags) as build/variants/libbonxai_b200_<tag>.so; load it with
    BNX_LIB=<path>. Used for A/B measurements of kernel parameters on the GPU box."""
    vdir = os.path.join(ROOT, "build", "variants", tag)
    os.makedirs(vdir, exist_ok=True)
    objs, procs = [], []
    for src in SOURCES:
        obj = os.path.join(vdir, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [_nvcc(), *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError(f"nvcc failed on {src} ({tag})")
    lib = os.path.join(ROOT, "build", "variants", f"libbonxai_b200_{tag}.so")
    subprocess.run([_nvcc(), "-shared", "-cudart", "static", "-Xlinker", "-ldl", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib, *objs], check=True)
    return lib


DROPIN_BENCH = os.path.join(ROOT, "build", "dropin_bench")


def build_tools(force: bool = False) -> str:
    """the C++ caller bench.py times for `e2e_dropin` (tools/cpp/dropin_bench.cpp against include/ + the library)"""
    src = os.path.join(ROOT, "tools", "cpp", "dropin_bench.cpp")
    deps = [src, LIB] + [os.path.join(dp, f) for dp, _, fs in os.walk(os.path.join(ROOT, "include")) for f in fs]
    if not force and os.path.exists(DROPIN_BENCH) and all(os.path.getmtime(d) <= os.path.getmtime(DROPIN_BENCH) for d in deps):
        return DROPIN_BENCH
    os.makedirs(os.path.dirname(DROPIN_BENCH), exist_ok=True)
    subprocess.run(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), src, "-L", PKG, "-lbonxai_b200",
                    "-Wl,-rpath,$ORIGIN/../bonxai_b200", "-o", DROPIN_BENCH], check=True)
    return DROPIN_BENCH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_tools(force="--force" in sys.argv))
