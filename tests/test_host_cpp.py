"""The C++ class surface (include/turboinfer/*.hpp over the C ABI): builds tests/cpp/test_host_api.cpp with g++ and runs
it.  CPU mode: host value types + every device entry point fails loudly without a GPU.  GPU mode: ops, the reference's
quantizer fixtures, and a small decoder through InferenceEngine::generate."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_host_api")


@pytest.fixture(scope="module")
def exe():
    from turboinfer_b200 import build as b
    return b.build_host_test(EXE)


def test_host_api_without_gpu(exe):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")   # also on a GPU box: the library must refuse to fall back
    r = subprocess.run([exe, "cpu"], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
def test_host_api_on_gpu(exe):
    r = subprocess.run([exe, "gpu"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr


def test_packed_layout_host_logic(tmp_path):
    """qlayout.cuh compiled for the host: slab partition, stage bookkeeping and the fragment-order map (bijection, no
    padding bytes) for full and ragged groups -- the logic pack / unpack kernels, producer and consumers share."""
    exe = str(tmp_path / "test_layout")
    src = os.path.join(ROOT, "tests", "cpp", "test_layout.cpp")
    r = subprocess.run([os.environ.get("TI_HOST_CXX", "/usr/bin/g++"), "-std=c++17", "-O1", "-o", exe, src], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "layout test: ok" in r.stdout, r.stdout + r.stderr
