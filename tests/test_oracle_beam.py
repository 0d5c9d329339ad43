"""Beam search restatement (oracle/ti_oracle.c: tio_beam_expand / beam_search_core) pinned against the compiled reference's own
InferenceEngine::generate_beam_search (src/model/inference_engine.cpp:830-871, :1912-2069) on the literal benchmark model, where
the reference engine runs as it stands (SURVEY.md 8c level C).  The literal model's logits repeat with the period of its ramp
fills, so equal probabilities are common and the reference's std::sort / heap order among them is unspecified: the
log-probabilities must agree to the bit everywhere, the token ids wherever the winners are not tied."""
import numpy as np
import pytest

import oracle

CASES = [dict(), dict(top_k=5, top_p=1.0), dict(temperature=0.7, top_k=0, top_p=0.95, length_penalty=0.6),
         dict(temperature=1.3, top_k=20, top_p=0.5, length_penalty=2.0)]


@pytest.mark.parametrize("qtype", [3, 0, 1])   # fp32 / int8 / int4 variants of benchmark_inference
@pytest.mark.parametrize("case", range(len(CASES)))
def test_literal_beam_search_port_vs_reference(qtype, case):
    if not oracle.ref_available():
        pytest.skip("the compiled reference (oracle/_ref) is not present")
    kw = CASES[case]
    a = oracle.port().beam_search_literal(1000, 256, 4, qtype, [1, 15, 25, 35], 5, 3, **kw)
    b = oracle.ref().beam_search_literal(1000, 256, 4, qtype, [1, 15, 25, 35], 5, 3, **kw)
    assert len(a) == len(b) >= 1
    for x, y in zip(a, b):
        assert np.float32(x["avg_logprob"]) == np.float32(y["avg_logprob"])     # bit for bit
        assert len(x["tokens"]) == len(y["tokens"]) and x["finished"] == y["finished"]
    if qtype == 0 and case < 3:   # no ties among the winners in these runs: the sequences are the reference's
        assert [x["tokens"] for x in a] == [y["tokens"] for y in b]


def test_beam_expand_properties():
    """the expansion on random logits: probabilities descend, sum <= 1, tokens are the arg-sort of the logits, top-k / top-p limits"""
    port = oracle.port()
    lg = (np.random.default_rng(5).normal(size=2000) * 2.5).astype(np.float32)
    order = np.argsort(-lg, kind="stable")
    for T, k, p, beam in [(1.0, 50, 0.9, 4), (0.7, 0, 1.0, 8), (1.0, 3, 1.0, 8), (1.0, 0, 0.3, 8), (2.0, 1, 0.9, 4)]:
        e = port.beam_expand(lg, beam, T, k, p)
        probs = [x[0] for x in e]
        toks = [x[1] for x in e]
        assert toks == [int(t) for t in order[: len(toks)]]
        assert all(probs[i] >= probs[i + 1] for i in range(len(probs) - 1)) and sum(probs) <= 1.0 + 1e-5
        if 0 < k < beam:
            assert len(e) == k and abs(sum(probs) - 1.0) < 1e-5     # the whole filtered distribution fits in the beam
        if k == 0 and p == 1.0:
            ref = np.exp(lg[toks] / T - np.max(lg / T)) / np.sum(np.exp(lg / T - np.max(lg / T)))
            np.testing.assert_allclose(probs, ref, rtol=1e-5)


def test_level_b_beam_search_consistency():
    """beam_size 1 with top_k 1 is greedy decoding; a wider beam never returns a worse best score than a narrower one's greedy path"""
    import sys, os
    sys.path.insert(0, os.path.dirname(__file__))
    from helpers import SHAPES, make_model, prompt_tokens
    port = oracle.port()
    meta = SHAPES["tiny-test"]
    w = make_model(meta, norm_jitter=0.1)
    prompt = prompt_tokens(5, meta["vocab"])
    greedy, _ = port.decode_greedy(w, meta, prompt, 6)
    one = port.beam_search(w, meta, prompt, 6, 1, top_k=1, top_p=1.0, eos_token=-1)
    assert len(one) == 1 and one[0]["tokens"] == [int(t) for t in greedy] and one[0]["log_prob"] == 0.0
    b1 = port.beam_search(w, meta, prompt, 6, 1, top_k=0, top_p=1.0, eos_token=-1)
    b4 = port.beam_search(w, meta, prompt, 6, 4, top_k=0, top_p=1.0, eos_token=-1)
    assert b1[0]["tokens"] == [int(t) for t in greedy]
    assert len(b4) == 4 and b4[0]["score"] >= b1[0]["score"] - 1e-6
    assert all(b4[i]["score"] >= b4[i + 1]["score"] for i in range(3))
    # EOS ends a candidate early; max_new = 0 returns the bare prompt
    eos = b4[0]["tokens"][2]
    cut = port.beam_search(w, meta, prompt, 6, 4, top_k=0, top_p=1.0, eos_token=eos)
    assert any(r["tokens"][-1] == eos and len(r["tokens"]) < 6 for r in cut)
    empty = port.beam_search(w, meta, prompt, 0, 4)
    assert len(empty) == 1 and empty[0]["tokens"] == [] and empty[0]["finished"]
