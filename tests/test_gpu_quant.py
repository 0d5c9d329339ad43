"""GPU quantizer vs the oracle: integers, scale and zero-point bits must be exact."""
import os

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz"))


@pytest.fixture(scope="module")
def tb():
    import turboinfer_b200 as t
    t.init(0)
    return t


def _bits(x):
    return np.float32(x).tobytes()


@pytest.mark.parametrize("name", sorted({k.split("/")[1] for k in G.files if k.startswith("quant/")}))
def test_golden_quant_fixtures(tb, name):
    x = G[f"quant/{name}/x"]
    qt, sym = (int(v) for v in G[f"quant/{name}/cfg"])
    s, z = tb.ops.quant_info(x, qt, bool(sym))
    gs, gz = G[f"quant/{name}/scale_zp"]
    assert _bits(s) == _bits(gs) and _bits(z) == _bits(gz)
    q = tb.ops.quantize(x, qt, s, z)
    assert np.array_equal(q.astype(np.int32), G[f"quant/{name}/q"].astype(np.int32))
    assert np.array_equal(tb.ops.dequantize(q, qt, s, z), G[f"quant/{name}/deq"])


@pytest.mark.parametrize("qt", [oracle.QINT8, oracle.QINT4])
@pytest.mark.parametrize("sym", [True, False])
@pytest.mark.parametrize("K,N", [(256, 1024), (300, 77), (1024, 4), (1, 1), (2048, 515), (1500, 1000)])
def test_pack_unpack_equals_reference_ints(tb, port, qt, sym, K, N):
    rng = np.random.default_rng(K * 131 + N)
    w = rng.uniform(-0.03, 0.05, (K, N)).astype(np.float32)
    s, z = port.quant_info(w, qt, sym)
    q_ref = port.quantize(w, qt, s, z).astype(np.int32)
    qw = tb.QWeight(w, qt, sym)
    try:
        assert _bits(qw.scale) == _bits(s) and _bits(qw.zero_point) == _bits(z)
        assert np.array_equal(qw.unpack(), q_ref)
        expected = K * N // 2 if qt == oracle.QINT4 else K * N
        assert qw.packed_bytes >= expected  # padding to 4-column units / whole superchunks only
    finally:
        qw.free()


def test_quantize_edge_cases(tb, port):
    x = np.array([-1e9, 1e9, 0.0, 0.5, 1.5, 2.5, -0.5, -2.5], dtype=np.float32)
    for qt, s, z in [(oracle.QINT8, 1.0, 0.0), (oracle.QINT4, 1.0, 0.0), (oracle.QINT4, 1.0, 3.0), (oracle.QINT8, 0.37, 11.5)]:
        assert np.array_equal(tb.ops.quantize(x, qt, s, z), port.quantize(x, qt, s, z))
    with pytest.raises(tb.B200Error):
        tb.ops.quant_info(np.zeros(0, dtype=np.float32), oracle.QINT8)
    with pytest.raises(tb.B200Error):
        tb.ops.quant_info(np.ones(4, dtype=np.float32), oracle.QNONE)
    # constant tensor: scale 0 -> the reference divides by zero; integers must still agree
    c = np.zeros(64, dtype=np.float32)
    s, z = port.quant_info(c, oracle.QINT8, True)
    gs, gz = tb.ops.quant_info(c, oracle.QINT8, True)
    assert _bits(s) == _bits(gs) and _bits(z) == _bits(gz)


def test_full_size_quantize_matches_oracle(tb, port):
    # one 7B-shape projection (4096 x 4096): element-for-element, both bit widths
    w = np.random.default_rng(42).uniform(-1 / 64, 1 / 64, (4096, 4096)).astype(np.float32)
    for qt in (oracle.QINT4, oracle.QINT8):
        s, z = port.quant_info(w, qt, True)
        qw = tb.QWeight(w, qt, True)
        try:
            assert _bits(qw.scale) == _bits(s)
            assert np.array_equal(qw.unpack(), port.quantize(w, qt, s, z).astype(np.int32))
        finally:
            qw.free()
