"""BASELINE.json configs[0] on the GPU: the literal path benchmarks/benchmark_inference runs on the reference today
(create_test_model(1000, 256, 4), prompt {1,15,25,35}, greedy 128 tokens; FP32 and quantize_model INT8 / INT4 with the
unscaled integer weights of SURVEY R8).  Tokens must equal the golden vectors generated from the compiled reference
(tests/golden/make_golden.py, level C) and the C restatement."""
import os

import numpy as np
import pytest

import oracle
from helpers import make_literal_model

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz"))


@pytest.fixture(scope="module")
def tb():
    import turboinfer_b200 as t
    t.init(0)
    return t


@pytest.mark.parametrize("name,qt", [("fp32", oracle.QNONE), ("int8", oracle.QINT8), ("int4", oracle.QINT4)])
def test_literal_config1_tokens_equal_reference(tb, port, name, qt):
    V, H, L = 1000, 256, 4
    meta = dict(vocab=V, hidden=H, layers=L, heads=H // 64, inter=4 * H)
    m = tb.Model(meta, qt, attn_mode=0, rope_mode=0, max_seq=2048, compat_literal=True).load(make_literal_model(V, H, L))
    try:
        assert not m.persistent_engine
        toks, logits, _ = m.generate_greedy([1, 15, 25, 35], 128, want_logits=True)
        tok1, lg1 = m.decode_step(7)
    finally:
        m.free()
    assert np.array_equal(toks, G[f"literal/{name}/tokens"])
    rt, rl = port.generate_literal(V, H, L, qt, [1, 15, 25, 35], 128)
    assert np.array_equal(toks, rt)
    assert np.array_equal(logits[-1], rl), "literal logits are fp32 in the reference's order of roundings: bit-exact"
    assert tok1 == toks[-1] and np.array_equal(lg1, rl)      # a decode step is one row, independent of token and position


def test_literal_rejects_a_complete_attention_block(tb):
    V, H, L = 64, 64, 1
    w = make_literal_model(V, H, L)
    w["layers.0.attention.o_proj.weight"] = w["layers.0.attention.q_proj.weight"]
    meta = dict(vocab=V, hidden=H, layers=L, heads=1, inter=4 * H)
    m = tb.Model(meta, oracle.QNONE, attn_mode=0, max_seq=64, compat_literal=True)
    try:
        with pytest.raises(tb.B200Error, match="compat_literal"):
            m.load(w)
    finally:
        m.free()
