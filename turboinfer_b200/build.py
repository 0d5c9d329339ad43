"""Builds libturboinfer_b200.so in-tree with nvcc for sm_100a (no torch, no JIT cache)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libturboinfer_b200.so")
SOURCES = ["capi.cu"]
HEADERS = ["ptx.cuh", "qlayout.cuh", "gemv.cuh", "kernels.cuh", "mega.cuh", "gemm_tc.cuh", "prefill.cuh", "batch.cuh"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
         "-shared", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(HERE, "..", "include", "ti_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    cmd = [NVCC, *FLAGS, "-o", LIB, *[os.path.join(CSRC, s) for s in SOURCES], "-lcudart"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libturboinfer_b200.so")
    if verbose:
        print(log)
    return LIB


HOST_LIB = os.path.join(HERE, "libturboinfer_b200_host.so")
HOST_SRC = os.path.join(HERE, "host", "turboinfer_host.cpp")
CXX = os.environ.get("TI_HOST_CXX", "/usr/bin/g++")


def build_host(force: bool = False) -> str:
    """The C++ class surface (include/turboinfer/*.hpp) above the C ABI: g++ only, no CUDA."""
    build()
    inc = os.path.join(HERE, "..", "include")
    deps = [HOST_SRC, LIB] + [os.path.join(r, f) for r, _, fs in os.walk(inc) for f in fs]
    if not force and os.path.exists(HOST_LIB) and all(os.path.getmtime(d) <= os.path.getmtime(HOST_LIB) for d in deps):
        return HOST_LIB
    cmd = [CXX, "-std=c++20", "-O2", "-fPIC", "-shared", "-Wall", "-Wextra", "-I", inc, "-o", HOST_LIB, HOST_SRC,
           "-L", HERE, "-lturboinfer_b200", "-Wl,-rpath,$ORIGIN"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("g++ failed building libturboinfer_b200_host.so")
    return HOST_LIB


def build_host_test(out_path: str) -> str:
    """tests/cpp/test_host_api.cpp linked against the two libraries."""
    build_host()
    inc = os.path.join(HERE, "..", "include")
    src = os.path.join(HERE, "..", "tests", "cpp", "test_host_api.cpp")
    cmd = [CXX, "-std=c++20", "-O1", "-I", inc, "-o", out_path, src, "-L", HERE, "-lturboinfer_b200_host", "-lturboinfer_b200",
           "-Wl,-rpath," + HERE]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("g++ failed building test_host_api")
    return out_path


if __name__ == "__main__":
    build(force=True, verbose=True)
    build_host(force=True)
