"""tcgen05 tensor-core GEMM (prefill / batched decode) vs the streaming GEMV and the CPU oracle."""
import numpy as np
import pytest

import oracle
from helpers import rel_err_inf

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tb():
    import turboinfer_b200 as t
    t.init(0)
    return t


@pytest.mark.parametrize("qt", [oracle.QINT4, oracle.QINT8])
@pytest.mark.parametrize("M,K,N", [(128, 256, 128), (32, 512, 384), (160, 4096, 512), (1, 1024, 256), (200, 1000, 132)])
def test_gemm_rows_equal_gemv_and_match_oracle(tb, port, qt, M, K, N):
    rng = np.random.default_rng(7 * M + K + N)
    w = rng.uniform(-1.0 / np.sqrt(K), 1.0 / np.sqrt(K), (K, N)).astype(np.float32)
    x = rng.standard_normal((M, K)).astype(np.float32)
    x[M // 2] *= 37.0          # rows with very different magnitudes: the activation scale is per row
    qw = tb.QWeight(w, qt)
    try:
        y = qw.gemm(x)
        rows = sorted({0, M // 2, M - 1})
        yv = qw.gemv(x[rows])
    finally:
        qw.free()
    assert y.shape == (M, N)
    assert np.array_equal(y[rows], yv), "a GEMM row must be bit-identical to the GEMV of that row (exact integer accumulation)"
    s, z = port.quant_info(w, qt, True)
    ref = port.matmul(x, port.dequantize(port.quantize(w, qt, s, z), qt, s, z))
    for r in range(M):      # per row: the tolerance of SURVEY 8a is relative to the row's own magnitude
        assert rel_err_inf(y[r], ref[r]) <= 1e-4


def test_gemm_asymmetric_and_zero_rows(tb, port):
    rng = np.random.default_rng(3)
    K, N, M = 512, 256, 64
    w = rng.uniform(0.1, 0.9, (K, N)).astype(np.float32)
    x = rng.standard_normal((M, K)).astype(np.float32)
    x[5] = 0.0
    for qt in (oracle.QINT4, oracle.QINT8):
        qw = tb.QWeight(w, qt, symmetric=False)
        try:
            y = qw.gemm(x)
            yv = qw.gemv(x[[0, 5, 63]])
        finally:
            qw.free()
        assert np.array_equal(y[[0, 5, 63]], yv)
        assert not y[5].any()
