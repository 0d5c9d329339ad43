"""Batched decode (generate_batch): B sequences advance in lockstep through the tcgen05 GEMM path,
each with its own KV pages.  Every sequence's greedy tokens must equal what the engine produces for that prompt alone
(persistent-kernel decode path) and the CPU oracle's; last-step logits within 1e-2."""
import numpy as np
import pytest

import oracle
from helpers import SHAPES, make_model, prompt_tokens, rel_err_inf

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tb():
    import turboinfer_b200 as t
    t.init(0)
    return t


def fake_quant_model(port, w, qt):
    return {k: (port.fake_quant(v, qt) if (v.ndim == 2 and "embeddings" not in k) else v) for k, v in w.items()}


@pytest.mark.parametrize("qt", [oracle.QINT8, oracle.QINT4])
@pytest.mark.parametrize("shape,B,n_prompt,n_new,rope", [("tiny-test", 5, 4, 12, 1), ("bench-small", 32, 6, 8, 1), ("bench-small", 3, 70, 6, 0),
                                                         ("tiny-test", 130, 3, 5, 1)])
def test_generate_batch_matches_single_sequence_and_oracle(tb, port, qt, shape, B, n_prompt, n_new, rope):
    meta = SHAPES[shape]
    w = make_model(meta, norm_jitter=0.1)
    prompts = np.array([prompt_tokens(n_prompt, meta["vocab"], offset=b) for b in range(B)], dtype=np.int32)
    m = tb.Model(meta, qt, attn_mode=1, rope_mode=rope, max_seq=256).load(w)
    try:
        toks, logits, ms = m.generate_batch_greedy(prompts, n_new, want_logits=True)
        toks2, _, _ = m.generate_batch_greedy(prompts, n_new)          # graph replay, caches reset
        singles = [m.generate_greedy(prompts[b], n_new, want_logits=True) for b in (0, B // 2, B - 1)]
    finally:
        m.free()
    assert toks.shape == (B, n_new) and ms >= 0.0
    assert np.array_equal(toks, toks2)
    for b, (st, sl, _) in zip((0, B // 2, B - 1), singles):
        assert np.array_equal(toks[b], st), (b, toks[b], st)
        assert rel_err_inf(logits[b], sl[-1]) <= 1e-4
    fq = fake_quant_model(port, w, qt)
    for b in (0, B - 1):
        rt, rl = port.decode_greedy(fq, meta, list(prompts[b]), n_new, attn_mode=1, rope_mode=rope)
        assert np.array_equal(toks[b], rt), (b, toks[b], rt)
        assert rel_err_inf(logits[b], rl[-1]) <= 1e-2


def test_generate_batch_rejects_bad_arguments(tb):
    meta = SHAPES["tiny-test"]
    m = tb.Model(meta, oracle.QINT4, attn_mode=1, rope_mode=1, max_seq=32).load(make_model(meta))
    try:
        with pytest.raises(tb.B200Error):
            m.generate_batch_greedy(np.zeros((2, 30), np.int32), 8)          # KV cache overflow (:100-102)
        with pytest.raises(tb.B200Error):
            m.generate_batch_greedy(np.full((2, 3), meta["vocab"], np.int32), 2)   # token id out of range
    finally:
        m.free()


def test_generate_batch_survives_scratch_reallocation(tb):
    """The step graphs hold pointers into the scratch buffers; a long prompt in between reallocates them."""
    meta = SHAPES["bench-small"]
    w = make_model(meta, norm_jitter=0.1)
    prompts = np.array([prompt_tokens(5, meta["vocab"], offset=b) for b in range(4)], dtype=np.int32)
    m = tb.Model(meta, oracle.QINT4, attn_mode=1, rope_mode=1, max_seq=256).load(w)
    try:
        first, _, _ = m.generate_batch_greedy(prompts, 7)
        m.generate_greedy(prompt_tokens(120, meta["vocab"]), 3)      # batched prefill with M = 119 rows: scratch grows
        again, _, _ = m.generate_batch_greedy(prompts, 7)
    finally:
        m.free()
    assert np.array_equal(first, again)


@pytest.mark.parametrize("qt", [oracle.QINT8, oracle.QINT4])
def test_generate_batch_ragged_prompts(tb, port, qt):
    """Prompts of different lengths (the reference's generate_batch loops over generate(), :804-828, so any lengths go):
    left-aligned lockstep on the device; every sequence's tokens equal the single-sequence engine's and the oracle's."""
    meta = SHAPES["bench-small"]
    w = make_model(meta, norm_jitter=0.1)
    lens = [3, 9, 1, 40, 6, 17]          # 40 > 32: that sequence alone would take the tensor-core prefill; here all go step by step
    prompts = [prompt_tokens(n, meta["vocab"], offset=5 * b) for b, n in enumerate(lens)]
    n_new = 7
    m = tb.Model(meta, qt, attn_mode=1, rope_mode=1, max_seq=128).load(w)
    try:
        toks, _ = m.generate_batch_ragged(prompts, n_new)
        toks2, _ = m.generate_batch_ragged(prompts, n_new)        # graph replay
        singles = [m.generate_greedy(p, n_new)[0] for p in prompts]
        eq, _, _ = m.generate_batch_greedy(np.array([prompts[1], prompts[1]], dtype=np.int32), n_new)   # the equal-length entry still works
    finally:
        m.free()
    assert np.array_equal(toks, toks2)
    for b in range(len(lens)):
        assert np.array_equal(toks[b], singles[b]), (b, toks[b], singles[b])
    assert np.array_equal(eq[0], singles[1]) and np.array_equal(eq[1], singles[1])
    fq = fake_quant_model(port, w, qt)
    for b in (0, 3):
        rt, _ = port.decode_greedy(fq, meta, prompts[b], n_new, attn_mode=1, rope_mode=1)
        assert np.array_equal(toks[b], rt)
