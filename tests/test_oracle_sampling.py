"""The oracle's restatement of sample_next_token / compute_logprobs (no GPU): invariants the reference algorithm has, and
the one case where it can be pinned against the compiled reference's behaviour (top_k = 1 is an arg-max)."""
import numpy as np
import pytest

import oracle


def test_sample_invariants():
    p = oracle.port()
    rng = np.random.default_rng(1)
    lg = (rng.standard_normal(4000) * 3).astype(np.float32)
    order = np.argsort(-lg, kind="stable")
    for u in (1e-7, 0.3, 0.999999):
        tok, lp = p.sample(lg, 1.0, 1, 0.9, u)                      # top_k = 1: the arg-max, probability 1
        assert tok == int(np.argmax(lg)) and lp == 0.0
    # a quirk of the reference loop (:1655-1665): `random_value <= cumsum` holds at i = 0 for a uniform of exactly 0, whatever
    # the probability of token 0
    assert p.sample(lg, 1.0, 1, 0.9, 0.0)[0] == 0
    for top_k in (5, 50):
        seen = {p.sample(lg, 0.8, top_k, 1.0, u)[0] for u in np.linspace(1e-6, 0.9999, 400, dtype=np.float32)}
        assert seen <= set(order[:top_k].tolist())                   # only survivors of the top-k filter are ever picked
    # top_p keeps the smallest prefix of the sorted probabilities whose mass reaches top_p
    probs = np.exp(lg.astype(np.float64) - lg.max()); probs /= probs.sum()
    keep = int(np.searchsorted(np.cumsum(probs[order]), 0.5)) + 1
    seen = {p.sample(lg, 1.0, 0, 0.5, u)[0] for u in np.linspace(1e-6, 0.9999, 600, dtype=np.float32)}
    assert seen <= set(order[:keep + 1].tolist())
    # inverse CDF is monotone in u
    picks = [p.sample(lg, 1.0, 50, 0.9, u)[0] for u in np.linspace(1e-6, 0.9999, 50, dtype=np.float32)]
    assert picks == sorted(picks)
    # the counter-based uniform: deterministic, in [0, 1), decorrelated across steps
    us = [p.uniform(9, s) for s in range(1000)]
    assert us == [p.uniform(9, s) for s in range(1000)] and 0.0 <= min(us) and max(us) < 1.0 and 0.4 < float(np.mean(us)) < 0.6


def test_logprobs_restatement():
    p = oracle.port()
    rng = np.random.default_rng(2)
    lg = rng.standard_normal((6, 300)).astype(np.float32)
    toks = [3, 299, 0, 17, 300, -1]
    got = p.logprobs(lg, toks)
    ref = lg - lg.max(axis=1, keepdims=True)
    ref = ref - np.log(np.exp(ref).sum(axis=1, keepdims=True))
    for i, t in enumerate(toks):
        if 0 <= t < 300:
            assert abs(got[i] - ref[i, t]) < 1e-5
        else:
            assert got[i] == -20.0                                   # LOGPROB_INVALID_TOKEN, inference_engine.cpp:933-936


@pytest.mark.parametrize("qtype", [3, 0, 1])
def test_logprobs_port_vs_reference_on_the_literal_model(qtype):
    """tio_logprobs pinned: the compiled reference's own InferenceEngine::compute_logprobs (:873-954) on the literal benchmark
    model, against the restatement (literal forward pass + tio_logprobs) -- bit for bit, including the -20 of an invalid id"""
    if not oracle.ref_available():
        pytest.skip("the compiled reference (oracle/_ref) is not present")
    toks = [1, 15, 25, 35, 999, 0, 500]
    a = oracle.port().logprobs_literal(1000, 256, 4, qtype, toks)
    b = oracle.ref().logprobs_literal(1000, 256, 4, qtype, toks)
    np.testing.assert_array_equal(a, b)
    assert np.all(a <= 0) and np.all(np.isfinite(a))


@pytest.mark.parametrize("kw", [dict(temperature=0.8, top_k=50, top_p=1e-6), dict(temperature=1.0, top_k=5, top_p=1e-4),
                                dict(temperature=1.7, top_k=0, top_p=1e-6)])
def test_sampling_pipeline_port_vs_reference_with_a_one_token_nucleus(kw):
    """The reference's generator is time-seeded (:472), but with a top_p so small that the nucleus is a single token the draw
    cannot matter: the compiled reference's own generate() with temperature / top-k / top-p switched on must then produce what
    tio_sample produces for ANY uniform.  INT8 literal model (its winning logit is never tied; in the fp32 / INT4 variants the
    ramp fills tie the maximum and std::sort's order among equals is unspecified)."""
    if not oracle.ref_available():
        pytest.skip("the compiled reference (oracle/_ref) is not present")
    want = oracle.ref().generate_literal_sampled(1000, 256, 4, 0, [1, 15, 25, 35], 12, **kw)
    for u in (0.0, 0.01, 0.5, 0.99):
        got = oracle.port().generate_literal_sampled(1000, 256, 4, 0, [1, 15, 25, 35], 12, u=u, **kw)
        if u == 0.0:
            continue   # (u == 0 selects index 0 by the loop's `<=`, a case the reference's generator produces with probability 2^-24)
        np.testing.assert_array_equal(got, want)
