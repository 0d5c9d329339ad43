"""On-device sampling (src/model/inference_engine.cpp:1554-1673) and compute_logprobs (:873-954) against
the oracle's restatement of the reference algorithm on the SAME logits and the SAME uniform (the engine's counter-based
RNG; the reference's own generator is seeded from the clock and cannot be replayed)."""
import numpy as np
import pytest

import oracle
from helpers import SHAPES, make_model, prompt_tokens

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tb():
    import turboinfer_b200 as t
    t.init(0)
    return t


def logits_rows(rows, V, seed, spread=4.0):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((rows, V)) * spread).astype(np.float32)


@pytest.mark.parametrize("V", [1000, 32000])
@pytest.mark.parametrize("temperature,top_k,top_p", [(1.0, 50, 0.9), (0.7, 50, 0.9), (1.3, 8, 1.0), (1.0, 1, 0.5), (0.9, 1024, 0.95),
                                                     (1.0, 40, 0.2)])
def test_sampler_exact_path_equals_oracle(tb, port, V, temperature, top_k, top_p):
    """top_k <= 1024: every sequential sum of the reference is reproduced in its order over the survivors -> the same token
    for the same uniform (row r of the call uses the seed of row r), log-probability within the last bits of expf."""
    rows = 24
    lg = logits_rows(rows, V, 11 * V + top_k)
    for step in (0, 5):
        toks, lps = tb.ops.sample(lg, temperature, top_k, top_p, seed=1234, step=step)
        for r in range(rows):
            u = port.uniform(1234 + r * 0x51ED27, step)
            rt, rl = port.sample(lg[r], temperature, top_k, top_p, u)
            assert toks[r] == rt, (r, step, toks[r], rt, u)
            assert abs(lps[r] - rl) <= 2e-5 * max(1.0, abs(rl))


@pytest.mark.parametrize("top_k,top_p", [(0, 0.9), (0, 1.0), (5000, 0.8)])
def test_sampler_wide_path_matches_oracle_distribution(tb, port, top_k, top_p):
    """No top-k filter (or one wider than the exact path holds): block-wide sums in another order.  The token equals the
    oracle's unless the uniform falls within rounding of a CDF boundary; the log-probability agrees to 1e-4."""
    V, rows = 32000, 32
    lg = logits_rows(rows, V, 99 + top_k, spread=3.0)
    toks, lps = tb.ops.sample(lg, 1.0, top_k, top_p, seed=77, step=3)
    same = 0
    for r in range(rows):
        u = port.uniform(77 + r * 0x51ED27, 3)
        rt, rl = port.sample(lg[r], 1.0, top_k, top_p, u)
        if toks[r] == rt:
            same += 1
            assert abs(lps[r] - rl) <= 1e-4 * max(1.0, abs(rl))
    assert same >= rows - 1, same


def test_sampler_top_k_one_is_greedy_and_argument_checks(tb):
    lg = logits_rows(4, 777, 5)
    toks, lps = tb.ops.sample(lg, 1.0, 1, 0.9, seed=3)
    assert np.array_equal(toks, np.argmax(lg, axis=1))
    assert np.allclose(lps, 0.0)
    with pytest.raises(tb.B200Error, match="Temperature must be positive"):
        tb.ops.sample(lg, 0.0, 10, 0.9)


def test_generate_sampled_reproducible_and_consistent_with_sampler(tb, port):
    """The device loop (prefill -> sample -> decode -> sample ...) is reproducible for a seed, differs between seeds, and
    each pick is what the sampler returns for that step's logits (replayed with decode_step on the same token history)."""
    meta = SHAPES["bench-small"]
    w = make_model(meta, norm_jitter=0.05)
    prompt = prompt_tokens(5, meta["vocab"])
    m = tb.Model(meta, oracle.QINT8, attn_mode=1, rope_mode=1, max_seq=128).load(w)
    try:
        a, la, _ = m.generate_sampled(prompt, 20, temperature=0.9, top_k=50, top_p=0.9, seed=42)
        b, lb, _ = m.generate_sampled(prompt, 20, temperature=0.9, top_k=50, top_p=0.9, seed=42)
        c, _, _ = m.generate_sampled(prompt, 20, temperature=0.9, top_k=50, top_p=0.9, seed=43)
        g, _, _ = m.generate_greedy(prompt, 20)
        k1, _, _ = m.generate_sampled(prompt, 20, temperature=1.0, top_k=1, top_p=1.0, seed=7)
        # replay: feed prompt + sampled tokens one by one, sample each step's logits on the host oracle
        m.reset()
        logits = None
        for t in prompt:
            _, logits = m.decode_step(t)
        replay = []
        for i in range(len(a)):
            tok, lp = port.sample(logits, 0.9, 50, 0.9, port.uniform(42, i))
            replay.append(tok)
            assert abs(lp - la[i]) <= 1e-4 * max(1.0, abs(lp))
            if i + 1 < len(a):
                _, logits = m.decode_step(int(a[i]))
    finally:
        m.free()
    assert np.array_equal(a, b) and np.array_equal(la, lb)
    assert not np.array_equal(a, c)
    assert np.array_equal(k1, g)                # top_k = 1 is the greedy path
    assert list(a) == replay


def test_compute_logprobs_vs_oracle(tb, port):
    meta = SHAPES["bench-small"]
    w = make_model(meta, norm_jitter=0.05)
    toks = prompt_tokens(9, meta["vocab"])
    m = tb.Model(meta, oracle.QINT8, attn_mode=1, rope_mode=1, max_seq=64).load(w)
    try:
        got = m.compute_logprobs(toks)
        # the logits of every position, replayed step by step
        m.reset()
        rows = np.stack([m.decode_step(t)[1] for t in toks])
    finally:
        m.free()
    ref = port.logprobs(rows, toks)
    assert np.allclose(got, ref, rtol=1e-5, atol=1e-5)
    fq = {k: (port.fake_quant(v, oracle.QINT8) if (v.ndim == 2 and "embeddings" not in k) else v) for k, v in w.items()}
    # and against the oracle's own forward pass (level B logits, position by position)
    ol = []
    for i in range(len(toks)):
        _, l = port.decode_greedy(fq, meta, toks[: i + 1], 1, attn_mode=1, rope_mode=1)
        ol.append(l[0])
    ref2 = port.logprobs(np.stack(ol), toks)
    assert np.allclose(got, ref2, rtol=1e-3, atol=1e-3)
