"""Batched prefill on the tensor cores (forward_pass): prompts of >= 33 tokens go through the
tcgen05 GEMM + causal attention path; greedy tokens must equal the CPU oracle's (which feeds the prompt token by token)
and the engine's own token-by-token prefill, logits within 1e-2."""
import os

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
@pytest.mark.parametrize("shape,n_prompt,rope", [("tiny-test", 40, 1), ("bench-small", 97, 1), ("bench-small", 33, 0)])
def test_gemm_prefill_matches_oracle_and_decode_path(tb, port, qt, shape, n_prompt, rope):
    meta = SHAPES[shape]
    w = make_model(meta, norm_jitter=0.1)
    prompt = prompt_tokens(n_prompt, meta["vocab"])
    n_new = 10
    m = tb.Model(meta, qt, attn_mode=1, rope_mode=rope, max_seq=256).load(w)
    try:
        toks, logits, _ = m.generate_greedy(prompt, n_new, want_logits=True)
        os.environ["TURBOINFER_B200_PREFILL"] = "decode"
        try:
            toks_d, logits_d, _ = m.generate_greedy(prompt, n_new, want_logits=True)
        finally:
            del os.environ["TURBOINFER_B200_PREFILL"]
    finally:
        m.free()
    assert np.array_equal(toks, toks_d)
    assert rel_err_inf(logits, logits_d) <= 1e-4
    rt, rl = port.decode_greedy(fake_quant_model(port, w, qt), meta, prompt, n_new, attn_mode=1, rope_mode=rope)
    assert np.array_equal(toks, rt)
    assert rel_err_inf(logits, rl) <= 1e-2
