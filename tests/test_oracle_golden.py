"""The C restatement (oracle/ti_oracle.c) against golden vectors produced by the COMPILED REFERENCE
(tests/golden/make_golden.py).  Runs anywhere (no GPU, no /root/reference)."""
import os

import numpy as np
import pytest

import oracle
from helpers import SHAPES, make_model, prompt_tokens

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz"))
QUANT_CASES = sorted({k.split("/")[1] for k in G.files if k.startswith("quant/")})


@pytest.mark.parametrize("name", QUANT_CASES)
def test_quantize_bit_exact(port, name):
    x = G[f"quant/{name}/x"]
    qt, sym = (int(v) for v in G[f"quant/{name}/cfg"])
    s, z = port.quant_info(x, qt, bool(sym))
    gs, gz = G[f"quant/{name}/scale_zp"]
    assert np.float32(s).tobytes() == np.float32(gs).tobytes()
    assert np.float32(z).tobytes() == np.float32(gz).tobytes()
    q = port.quantize(x, qt, s, z)
    assert np.array_equal(q.astype(np.int32), G[f"quant/{name}/q"].astype(np.int32))
    assert np.array_equal(port.dequantize(q, qt, s, z), G[f"quant/{name}/deq"])


def test_quantize_roundtrip_error_bounds(port):
    # the bound the reference's own tests state (max error < 1.0 on their linspace fixtures,
    # tests/test_quantization_complete.cpp:74-75) plus the tighter half-step bound for symmetric quantization
    for name in ("int8_sym_linspace", "int4_sym_linspace"):
        x = G[f"quant/{name}/x"]
        qt, sym = (int(v) for v in G[f"quant/{name}/cfg"])
        s, z = port.quant_info(x, qt, bool(sym))
        deq = port.dequantize(port.quantize(x, qt, s, z), qt, s, z)
        assert np.max(np.abs(deq - x)) <= 0.5 * s * (1 + 1e-6)


@pytest.mark.parametrize("t", [10, 50, 100, 200])
def test_fast_attention_fixture(port, t):
    H = 256
    q = (0.1 * (np.arange(H) % 10)).astype(np.float32).reshape(1, 1, H)
    k = (np.float32(0.05) * (np.add.outer(np.arange(t), np.arange(H)) % 20).astype(np.float32)).reshape(1, t, H)
    v = (np.float32(0.02) * (np.add.outer(2 * np.arange(t), np.arange(H)) % 15).astype(np.float32)).reshape(1, t, H)
    assert np.array_equal(port.attention_fast_incremental(q, k, v), G[f"attn/t{t}/out"])
    assert np.array_equal(port.multi_head_attention(q, k, v, 4), G[f"attn/t{t}/mha4"])


def test_mha_and_rope_fixtures(port):
    assert np.array_equal(port.multi_head_attention(G["mha3/q"], G["mha3/k"], G["mha3/v"], 3), G["mha3/out"])
    assert np.array_equal(port.rope(G["rope3/x"], G["rope3/pos"]), G["rope3/out"])
    assert np.array_equal(port.rope(G["rope4/x"], G["rope4/pos"]), G["rope4/out"])
    assert np.array_equal(port.rope(G["rope_dec/x"], G["rope_dec/pos"]), G["rope_dec/out"])


def test_elementwise_matmul_rms_softmax(port):
    assert np.array_equal(port.silu(G["act/x"]), G["act/silu"])
    assert np.array_equal(port.relu(G["act/x"]), G["act/relu"])
    assert np.array_equal(port.matmul(G["mm/a"], G["mm/b"]), G["mm/c"])
    assert np.array_equal(G["mm/c"], np.array([[22, 28], [49, 64]], dtype=np.float32))
    assert np.array_equal(port.matmul(G["gemv/x"], G["gemv/w"]), G["gemv/y"])
    assert np.array_equal(port.rms_norm(G["rms/x"], G["rms/w"]), G["rms/y"])
    y = port.softmax(G["softmax/x"])
    assert np.array_equal(y, G["softmax/y"])
    assert np.allclose(y.sum(axis=-1), 1.0, atol=1e-6)  # the reference's own check (tests/test_tensor_engine.cpp)


@pytest.mark.parametrize("name,qt", [("fp32", oracle.QNONE), ("int8", oracle.QINT8), ("int4", oracle.QINT4)])
def test_literal_config1_tokens(port, name, qt):
    toks, _ = port.generate_literal(1000, 256, 4, qt, [1, 15, 25, 35], 128)
    assert np.array_equal(toks, G[f"literal/{name}/tokens"])


@pytest.mark.parametrize("qname,qt", [("fp32", oracle.QNONE), ("int8", oracle.QINT8), ("int4", oracle.QINT4)])
@pytest.mark.parametrize("am,rm", [(1, 0), (0, 0), (1, 1)])
def test_decode_level_b(port, qname, qt, am, rm):
    meta = SHAPES["tiny-test"]
    w = make_model(meta, norm_jitter=0.1)
    wq = {k: (port.fake_quant(v, qt) if (v.ndim == 2 and "embeddings" not in k) else v) for k, v in w.items()}
    toks, logits = port.decode_greedy(wq, meta, prompt_tokens(5, meta["vocab"]), 24, attn_mode=am, rope_mode=rm)
    key = f"decodeB/{qname}/a{am}r{rm}"
    assert np.array_equal(toks, G[key + "/tokens"])
    assert np.array_equal(logits[0], G[key + "/logits_first"])
    assert np.array_equal(logits[-1], G[key + "/logits_last"])


def test_edge_cases(port):
    # empty input: the reference reads data[0] (UB); the restatement reports an error instead
    with pytest.raises(RuntimeError):
        port.quant_info(np.zeros(0, dtype=np.float32), oracle.QINT8)
    # single-token cache: softmax over one score is exactly 1 -> output is V bit for bit
    rng = np.random.default_rng(0)
    q = rng.standard_normal((1, 1, 24)).astype(np.float32)
    k = rng.standard_normal((1, 1, 24)).astype(np.float32)
    v = rng.standard_normal((1, 1, 24)).astype(np.float32)
    assert np.array_equal(port.attention_fast_incremental(q, k, v), v)
    # clamping: values far outside the scale saturate
    x = np.array([-1e9, 1e9, 0.0], dtype=np.float32)
    assert list(port.quantize(x, oracle.QINT8, 1.0, 0.0)) == [-128, 127, 0]
    assert list(port.quantize(x, oracle.QINT4, 1.0, 0.0)) == [-7, 7, 0]
    assert list(port.quantize(x, oracle.QINT4, 1.0, 3.0)) == [0, 15, 0]
    # round half away from zero (std::round), not half-to-even
    h = np.array([0.5, 1.5, 2.5, -0.5, -2.5], dtype=np.float32)
    assert list(port.quantize(h, oracle.QINT8, 1.0, 0.0)) == [1, 2, 3, -1, -3]
