"""The tensor-core paths at the width of the benchmark shapes: the batched decode step with the
split-K 32-row GEMM (the 11008-deep down projection of the 7B shape needs > 48 k-steps, which no small test shape
reaches), the tcgen05 prefill at full width, and the 128-row GEMM on the 7B gate/up and down shapes.  Depth is truncated
to L = 2 so that the CPU oracle finishes in seconds."""
import os

import numpy as np
import pytest

import oracle
from helpers import SHAPES, make_model, meta_with_layers, prompt_tokens, rel_err_inf

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tb():
    import turboinfer_b200 as t
    t.init(0)
    return t


def fake_quant_model(port, w, qt):
    return {k: (port.fake_quant(v, qt) if (v.ndim == 2 and "embeddings" not in k) else v) for k, v in w.items()}


@pytest.fixture(scope="module")
def llama7b_l2():
    meta = meta_with_layers(SHAPES["llama7b"], 2)
    return meta, make_model(meta)


def test_batch32_at_7b_width_runs_the_split_k_gemm(tb, port, llama7b_l2):
    """generate_batch, B = 32, Llama-2-7B width: qkv / o / gate-up go through the 32-row tcgen05 GEMM unsplit, the down
    projection (K = 11008 = 86 k-steps on 32 column tiles) through its split-K variant with 64-bit integer atomics.
    Every row against the single-sequence engine (persistent kernel, streaming GEMV), row 0 against the CPU oracle."""
    meta, w = llama7b_l2
    B, n_prompt, n_new = 32, 4, 3
    prompts = np.array([prompt_tokens(n_prompt, meta["vocab"], offset=b) for b in range(B)], dtype=np.int32)
    m = tb.Model(meta, oracle.QINT4, attn_mode=1, rope_mode=1, max_seq=128).load(w)
    try:
        toks, logits, _ = m.generate_batch_greedy(prompts, n_new, want_logits=True)
        singles = {b: m.generate_greedy(prompts[b], n_new, want_logits=True) for b in (0, 7, 16, 31)}
    finally:
        m.free()
    for b, (st, sl, _) in singles.items():
        assert np.array_equal(toks[b], st), (b, toks[b], st)
        assert rel_err_inf(logits[b], sl[-1]) <= 1e-4
    rt, rl = port.decode_greedy(fake_quant_model(port, w, oracle.QINT4), meta, list(prompts[0]), n_new, attn_mode=1, rope_mode=1)
    assert np.array_equal(toks[0], rt)
    assert rel_err_inf(logits[0], rl[-1]) <= 1e-2


def test_prefill_70_tokens_at_7b_width(tb, llama7b_l2):
    """A 70-token prompt at Llama-2-7B width goes through the tcgen05 prefill (M = 69 rows -> one 128-row tile per column
    tile, K = 4096 / 11008, N up to 22016); tokens and logits must equal the engine's own token-by-token prefill, which
    test_full_width_truncated_depth pins against the oracle at this width."""
    meta, w = llama7b_l2
    prompt = prompt_tokens(70, meta["vocab"])
    m = tb.Model(meta, oracle.QINT4, attn_mode=1, rope_mode=1, max_seq=128).load(w)
    try:
        toks, logits, _ = m.generate_greedy(prompt, 4, want_logits=True)
        os.environ["TURBOINFER_B200_PREFILL"] = "decode"
        try:
            toks_d, logits_d, _ = m.generate_greedy(prompt, 4, want_logits=True)
        finally:
            del os.environ["TURBOINFER_B200_PREFILL"]
    finally:
        m.free()
    assert np.array_equal(toks, toks_d)
    assert rel_err_inf(logits, logits_d) <= 1e-4


@pytest.mark.parametrize("qt", [oracle.QINT4, oracle.QINT8])
def test_prefill_70_tokens_at_tinyllama_width_vs_oracle(tb, port, qt):
    """The same path at TinyLlama width against the CPU oracle (which feeds the prompt token by token)."""
    meta = meta_with_layers(SHAPES["tinyllama"], 2)
    w = make_model(meta)
    prompt = prompt_tokens(70, meta["vocab"])
    m = tb.Model(meta, qt, attn_mode=1, rope_mode=1, max_seq=128).load(w)
    try:
        toks, logits, _ = m.generate_greedy(prompt, 4, want_logits=True)
    finally:
        m.free()
    rt, rl = port.decode_greedy(fake_quant_model(port, w, qt), meta, prompt, 4, attn_mode=1, rope_mode=1)
    assert np.array_equal(toks, rt)
    assert rel_err_inf(logits, rl) <= 1e-2


@pytest.mark.parametrize("M,K,N", [(2048, 4096, 22016), (32, 11008, 4096)])
def test_gemm_at_7b_shapes_rows_equal_gemv(tb, port, M, K, N):
    """The 128-row tcgen05 GEMM on the 7B gate/up shape at the prefill's M = 2048 and on the down shape: sampled rows are
    bit-identical to the streaming GEMV of that row and within 1e-4 of the oracle's dequantize + matmul."""
    rng = np.random.default_rng(M + K + N)
    w = rng.uniform(-1.0 / np.sqrt(K), 1.0 / np.sqrt(K), (K, N)).astype(np.float32)
    x = rng.standard_normal((M, K)).astype(np.float32)
    x[M // 3] *= 19.0
    qw = tb.QWeight(w, oracle.QINT4)
    try:
        y = qw.gemm(x)
        rows = sorted({0, M // 3, M // 2, M - 1})
        yv = qw.gemv(x[rows])
    finally:
        qw.free()
    assert y.shape == (M, N)
    assert np.array_equal(y[rows], yv)
    s, z = port.quant_info(w, oracle.QINT4, True)
    ref = port.matmul(x[rows], port.dequantize(port.quantize(w, oracle.QINT4, s, z), oracle.QINT4, s, z))
    for i, r in enumerate(rows):
        assert rel_err_inf(y[r], ref[i]) <= 1e-4
