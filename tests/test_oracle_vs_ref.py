"""Pins the C restatement to the compiled reference (oracle/_ref/libti_ref.so), bit for bit, on seeded
random inputs and ragged shapes.  Skipped when neither the built library nor /root/reference exists."""
import numpy as np
import pytest

import oracle
from helpers import SHAPES, make_model, meta_with_layers, prompt_tokens

rng = np.random.default_rng(20261018)


@pytest.mark.parametrize("qt", [oracle.QINT8, oracle.QINT4])
@pytest.mark.parametrize("sym", [True, False])
def test_quant(port, ref, qt, sym):
    w = rng.uniform(-0.03, 0.05, (300, 257)).astype(np.float32)
    sp, zp = port.quant_info(w, qt, sym)
    sr, zr = ref.quant_info(w, qt, sym)
    assert (sp, zp) == (sr, zr)
    q = ref.quantize(w, qt, sr, zr)
    assert np.array_equal(port.quantize(w, qt, sp, zp), q)
    assert np.array_equal(port.dequantize(q, qt, sp, zp), ref.dequantize(q, qt, sr, zr))


@pytest.mark.parametrize("M,K,N", [(1, 256, 1024), (1, 300, 77), (1, 2048, 512), (4, 64, 40), (3, 7, 5),
                                   (40, 64, 48), (33, 100, 45), (32, 300, 39), (35, 33, 32), (1, 3, 1)])
def test_matmul(port, ref, M, K, N):
    a = rng.standard_normal((M, K)).astype(np.float32)
    b = rng.standard_normal((K, N)).astype(np.float32)
    assert np.array_equal(port.matmul(a, b), ref.matmul(a, b))


@pytest.mark.parametrize("rows,H", [(1, 128), (3, 515), (1, 4096), (2, 7), (1, 12), (1, 3), (1, 13)])
def test_rms_norm(port, ref, rows, H):
    x = rng.standard_normal((rows, H)).astype(np.float32)
    w = (1 + 0.1 * rng.standard_normal(H)).astype(np.float32)
    assert np.array_equal(port.rms_norm(x, w), ref.rms_norm(x, w))
    assert np.array_equal(port.rms_norm(x, w, 1e-3), ref.rms_norm(x, w, 1e-3))


def test_rope(port, ref):
    x3 = rng.standard_normal((2, 5, 64)).astype(np.float32)
    pos = np.arange(5, dtype=np.float32) + 100
    assert np.array_equal(port.rope(x3, pos), ref.rope(x3, pos))
    x4 = rng.standard_normal((2, 3, 5, 32)).astype(np.float32)
    pos2 = (np.arange(10, dtype=np.float32) * 37).reshape(2, 5)
    assert np.array_equal(port.rope(x4, pos2), ref.rope(x4, pos2))
    assert np.array_equal(port.rope(x4, pos2, 500000.0), ref.rope(x4, pos2, 500000.0))


def test_elementwise_and_softmax(port, ref):
    v = (rng.standard_normal(1003) * 3).astype(np.float32)
    u = v[::-1].copy()
    assert np.array_equal(port.silu(v), ref.silu(v))
    assert np.array_equal(port.relu(v), ref.relu(v))
    assert np.array_equal(port.add(v, u), ref.add(v, u))
    assert np.array_equal(port.mul(v, u), ref.mul(v, u))
    s = rng.standard_normal((5, 13)).astype(np.float32)
    assert np.array_equal(port.softmax(s, 0.7), ref.softmax(s, 0.7))


def test_softmax_avx2_branch_is_the_references_bug(port, ref):
    # n >= 16 sends the reference down fast_exp_avx2 (src/core/tensor_engine.cpp:286-298), ~12 % off
    # (SURVEY R10).  The restatement follows the exact scalar branch; record the gap, don't chase it.
    s = rng.standard_normal((4, 64)).astype(np.float32)
    exact = np.exp(s.astype(np.float64) - s.max(axis=-1, keepdims=True))
    exact /= exact.sum(axis=-1, keepdims=True)
    assert np.max(np.abs(port.softmax(s) - exact) / exact) < 1e-5
    assert np.max(np.abs(ref.softmax(s) - exact) / exact) > 1e-3


@pytest.mark.parametrize("B,t,H", [(1, 1, 64), (1, 37, 260), (2, 200, 256), (1, 9, 13), (1, 4, 4), (3, 15, 7)])
def test_attention(port, ref, B, t, H):
    q = rng.standard_normal((B, 1, H)).astype(np.float32)
    k = rng.standard_normal((B, t, H)).astype(np.float32)
    v = rng.standard_normal((B, t, H)).astype(np.float32)
    assert np.array_equal(port.attention_fast_incremental(q, k, v), ref.attention_fast_incremental(q, k, v))


@pytest.mark.parametrize("t,H,nh", [(19, 96, 3), (1, 64, 4), (130, 256, 8), (7, 40, 5)])
def test_mha(port, ref, t, H, nh):
    q = rng.standard_normal((1, 1, H)).astype(np.float32)
    k = rng.standard_normal((1, t, H)).astype(np.float32)
    v = rng.standard_normal((1, t, H)).astype(np.float32)
    assert np.array_equal(port.multi_head_attention(q, k, v, nh), ref.multi_head_attention(q, k, v, nh))


@pytest.mark.parametrize("variant", [dict(), dict(gate=False), dict(o_proj=False, norms=False, gate=False)])
@pytest.mark.parametrize("am,rm", [(1, 0), (0, 0), (1, 1), (0, 2)])
def test_decode_level_b(port, ref, variant, am, rm):
    meta = SHAPES["tiny-test"]
    w = make_model(meta, norm_jitter=0.1, **variant)
    p = prompt_tokens(5, meta["vocab"])
    tp, lp = port.decode_greedy(w, meta, p, 12, attn_mode=am, rope_mode=rm)
    tr, lr = ref.decode_greedy(w, meta, p, 12, attn_mode=am, rope_mode=rm)
    assert np.array_equal(tp, tr)
    assert np.array_equal(lp, lr)


@pytest.mark.parametrize("qt", [oracle.QNONE, oracle.QINT8, oracle.QINT4])
@pytest.mark.parametrize("V,H,L,T", [(1000, 256, 4, 4), (1000, 256, 4, 6), (777, 64, 2, 3)])
def test_literal_config1(port, ref, qt, V, H, L, T):
    p = [1, 15, 25, 35, 45, 55][:T]
    tp, _ = port.generate_literal(V, H, L, qt, p, 24)
    tr, _ = ref.generate_literal(V, H, L, qt, p, 24)
    assert np.array_equal(tp, tr)
