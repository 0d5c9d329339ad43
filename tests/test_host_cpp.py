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


# ---- .tinq files (reference src/optimize/quantization.cpp:120-333, tests/test_quantization_persistence.cpp) ----------
def _run(exe, *args, timeout=300):
    r = subprocess.run([exe, *map(str, args)], capture_output=True, text=True, timeout=timeout,
                       env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.parametrize("quant", ["int8", "int4"])
def test_tinq_round_trip_host(exe, tmp_path, quant):
    """our writer -> our reader: every tensor, the metadata and the REAL (scale, zero_point) come back bit for bit"""
    path = tmp_path / "fixture.tinq"
    _run(exe, "tinq-write", quant, path)
    _run(exe, "tinq-check", quant, path, "ours")


def test_tinq_reader_errors(exe, tmp_path):
    _run(exe, "tinq-errors", tmp_path)


def _need_ref():
    import oracle
    if not oracle.ref_available():
        pytest.skip("the compiled reference (oracle/_ref) is not present")
    return oracle


@pytest.mark.parametrize("quant,qtype", [("int8", 0), ("int4", 1)])
def test_tinq_file_is_accepted_by_the_reference(exe, tmp_path, quant, qtype):
    """our writer -> the reference's reader -> the reference's writer -> our reader: the reference accepts the file, and every
    tensor and metadata field survives its round trip (its writer replaces the parameters by values derived from the integers,
    which our reader recognises and drops)"""
    oracle = _need_ref()
    ours, theirs = tmp_path / "ours.tinq", tmp_path / "theirs.tinq"
    _run(exe, "tinq-write", quant, ours)
    assert oracle.tinq_resave(str(ours), str(theirs), qtype) == 4
    _run(exe, "tinq-check", quant, theirs, "resaved")


@pytest.mark.parametrize("qtype", [0, 1])
def test_tinq_file_written_by_the_reference_loads(exe, tmp_path, qtype):
    """the reference quantizes and saves the model of its own persistence test; our reader returns the same metadata and the
    integers the reference's quantizer produces for those floats"""
    import numpy as np
    oracle = _need_ref()
    path, out = tmp_path / "ref.tinq", tmp_path / "dump"
    out.mkdir()
    oracle.tinq_write_sample(str(path), qtype)
    _run(exe, "tinq-dump", path, out)
    lines = (out / "index.txt").read_text().splitlines()
    assert lines[0].split() == ["meta", "test_model", "transformer", "1.0", "1000", "128", "2", "8", "512", "10000"]
    shapes = {"weight1": ((128, 256), 255), "weight2": ((256, 512), 127), "bias": ((256,), 64)}
    seen = {}
    for ln in lines[1:]:
        _, name, dtype, has_params, *dims = ln.split()
        seen[name] = (int(dtype), int(has_params), tuple(map(int, dims)))
    assert set(seen) == set(shapes)
    ref = oracle.ref()
    for name, (shape, mod) in shapes.items():
        dtype, has_params, dims = seen[name]
        assert dims == shape and dtype == (4 if qtype == 0 else 2) and has_params == 0   # DataType::kInt8 / kInt32
        n = int(np.prod(shape))
        x = ((np.arange(n) % mod).astype(np.float32) / np.float32(mod) - np.float32(0.5)).astype(np.float32)
        scale, zp = ref.quant_info(x, qtype, True)
        want = ref.quantize(x, qtype, scale, zp)
        got = np.fromfile(out / (name + ".bin"), dtype=np.int8 if qtype == 0 else np.int32)
        np.testing.assert_array_equal(got, want.ravel())


@pytest.mark.gpu
def test_tinq_engine_on_gpu(exe, tmp_path):
    r = subprocess.run([exe, "gpu-tinq", str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
