"""Beam search on the cached engine (reference generate_beam_search :830-871, beam_search_decode :1912-2069):
the device-side expansion against the CPU restatement on the same logits, the whole search against the restatement on a small
decoder, and the page-table forks against a cache-free replay that uses the engine's own logits."""
import numpy as np
import pytest

import oracle
from helpers import SHAPES, make_model, meta_with_layers, prompt_tokens

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tb():
    import turboinfer_b200 as t
    t.init(0)
    return t


@pytest.mark.parametrize("V", [512, 32000])
@pytest.mark.parametrize("T,k,p", [(1.0, 50, 0.9), (0.7, 50, 0.9), (1.0, 0, 1.0), (1.3, 0, 0.9), (1.0, 2000, 0.95), (1.0, 1, 0.9),
                                   (1.0, 3, 1.0), (0.5, 1024, 0.5)])
def test_beam_expand_vs_oracle(tb, port, V, T, k, p):
    rng = np.random.default_rng(V + k)
    lg = (rng.normal(size=(3, V)) * 3.0).astype(np.float32)
    lg[1, 7] = lg[1, 300] = lg[1].max() + 1.0          # a tie at the top: lower token id first
    for beam in (1, 4, 9):
        got = tb.ops.beam_expand(lg, beam, T, k, p)
        for r in range(3):
            want = port.beam_expand(lg[r], beam, T, k, p)
            assert [t for _, t in got[r]] == [t for _, t in want]
            # (rtol) The softmax denominator over the whole vocabulary is a block-wide sum here and a sequential fp32 sum in the
            # reference, whose rounding error grows with V (~3e-5 at 32000 terms); a top-k / top-p filter renormalises over the
            # survivors, where it cancels.
            np.testing.assert_allclose([q for q, _ in got[r]], [q for q, _ in want], rtol=1e-4 if (k == 0 and p == 1.0) else 2e-5)


def sharpen(w, f):
    """a more decided next-token distribution (random-init logits are nearly flat, which makes every beam a near-tie)"""
    w = dict(w)
    w["lm_head.weight"] = (w["lm_head.weight"] * f).astype(np.float32)
    return w


def same_results(got, want, lp_tol):
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert g["tokens"] == w["tokens"] and g["finished"] == w["finished"]
        assert abs(g["log_prob"] - w["log_prob"]) <= lp_tol * max(1.0, abs(w["log_prob"]))
        assert abs(g["score"] - w["score"]) <= lp_tol * max(1.0, abs(w["score"]))


@pytest.mark.parametrize("qt", [oracle.QINT8, oracle.QINT4])
@pytest.mark.parametrize("kw", [dict(beam=3, top_k=50, top_p=0.9), dict(beam=4, top_k=0, top_p=1.0, length_penalty=0.7),
                                dict(beam=2, top_k=5, top_p=0.8, temperature=0.8), dict(beam=1, top_k=1, top_p=1.0)])
def test_beam_search_vs_oracle(tb, port, qt, kw):
    kw = dict(kw)
    beam = kw.pop("beam")
    meta = SHAPES["tiny-test"]
    w = sharpen(make_model(meta, norm_jitter=0.1), 12.0)
    fq = {k: (port.fake_quant(v, qt) if (v.ndim == 2 and "embeddings" not in k) else v) for k, v in w.items()}
    prompt = prompt_tokens(6, meta["vocab"])
    m = tb.Model(meta, qt, max_seq=64).load(w)
    try:
        got = m.beam_search(prompt, 7, beam, eos_token=-1, **kw)
        want = port.beam_search(fq, meta, prompt, 7, beam, eos_token=-1, **kw)
        same_results(got, want, 2e-3)     # logits agree to <= 1e-2 relative; log-probabilities inherit that
        if beam == 1 and kw["top_k"] == 1:
            toks, _, _ = m.generate_greedy(prompt, 7)
            assert got[0]["tokens"] == [int(t) for t in toks]
    finally:
        m.free()


def replay_beam_search(logits_of, expand, prompt, max_new, beam, T, k, p, pen, eos):
    """beam_search_decode as the reference runs it -- every candidate's WHOLE sequence through the model at every step, no cache
    (:1961) -- with the engine's own logits: what the page-table forks must reproduce."""
    f32 = np.float32
    active = [dict(toks=[], lp=f32(0), score=f32(0), fin=False)]
    done = []
    for _ in range(max_new):
        if not active:
            break
        active.sort(key=lambda c: -c["lp"])          # stable: the earlier candidate first among equals
        nxt = []
        for c in active:
            for prob, tok in expand(logits_of(list(prompt) + c["toks"]), beam, T, k, p):
                toks = c["toks"] + [tok]
                lp = f32(c["lp"] + np.log(f32(prob)))
                score = f32(lp / np.power(f32(len(prompt) + len(toks)), f32(pen)))
                nxt.append(dict(toks=toks, lp=lp, score=score, fin=(tok == eos or len(toks) >= max_new)))
        nxt.sort(key=lambda c: -c["score"])
        active = []
        for c in nxt[:beam]:
            (done if c["fin"] else active).append(c)
        if len(done) >= beam:
            break
    active.sort(key=lambda c: -c["lp"])
    for c in active:
        c["fin"] = True
        done.append(c)
    done.sort(key=lambda c: -c["score"])
    return [dict(tokens=c["toks"], log_prob=float(c["lp"]), score=float(c["score"]), finished=c["fin"]) for c in done[:beam]]


@pytest.mark.parametrize("page_tokens,n_prompt,max_new,beam", [(4, 5, 14, 4), (8, 17, 20, 3), (64, 3, 10, 5)])
def test_page_table_forks_vs_cache_free_replay(tb, port, page_tokens, n_prompt, max_new, beam):
    """small pages: the search crosses many page boundaries, shares the prompt's pages between all beams and copies partly
    filled pages on every fork; the result must be the one a cache-free search over the same engine's logits finds"""
    meta = SHAPES["tiny-test"]
    w = sharpen(make_model(meta, norm_jitter=0.1), 12.0)
    prompt = prompt_tokens(n_prompt, meta["vocab"], offset=3)
    m = tb.Model(meta, oracle.QINT8, max_seq=64, kv_page_tokens=page_tokens).load(w)
    try:
        eos = -1
        kw = dict(temperature=0.9, top_k=40, top_p=0.95, length_penalty=1.0)
        got = m.beam_search(prompt, max_new, beam, eos_token=eos, **kw)

        def logits_of(seq):
            return m.prefill(seq)

        want = replay_beam_search(logits_of, port.beam_expand, prompt, max_new, beam, kw["temperature"], kw["top_k"], kw["top_p"],
                                  kw["length_penalty"], eos)
        same_results(got, want, 1e-4)     # the lockstep step and the single-sequence engine agree to fp32 rounding
        # an EOS id taken from the winning beam: candidates finish early, the search stops once `beam` of them have
        eos = got[0]["tokens"][max_new // 2]
        got = m.beam_search(prompt, max_new, beam, eos_token=eos, **kw)
        want = replay_beam_search(logits_of, port.beam_expand, prompt, max_new, beam, kw["temperature"], kw["top_k"], kw["top_p"],
                                  kw["length_penalty"], eos)
        same_results(got, want, 1e-4)
        assert any(r["tokens"][-1] == eos for r in got)
    finally:
        m.free()


def test_beam_search_arguments(tb):
    meta = SHAPES["tiny-test"]
    m = tb.Model(meta, oracle.QINT8, max_seq=32).load(make_model(meta))
    try:
        with pytest.raises(RuntimeError, match="Beam size"):
            m.beam_search([1, 2], 4, 0)
        with pytest.raises(RuntimeError):
            m.beam_search([], 4, 2)
        with pytest.raises(RuntimeError, match="overflow"):
            m.beam_search([1] * 30, 8, 2)
        with pytest.raises(RuntimeError):
            m.beam_search([1, meta["vocab"]], 4, 2)
        r = m.beam_search([1, 2, 3], 0, 3)
        assert len(r) == 1 and r[0]["tokens"] == [] and r[0]["finished"]
        # the search leaves the model usable for ordinary generation, and is repeatable
        a = m.beam_search([1, 2, 3], 5, 3)
        toks, _, _ = m.generate_greedy([1, 2, 3], 5)
        assert m.beam_search([1, 2, 3], 5, 3) == a and len(toks) == 5
    finally:
        m.free()


def test_beam_search_full_width(tb):
    """7B-wide layers (2 of them): the lockstep engine's tensor-core GEMM path under the beams; beam 1 / top_k 1 is greedy"""
    meta = meta_with_layers(SHAPES["llama7b"], 2)
    m = tb.Model(meta, oracle.QINT4, max_seq=128).load_synthetic()
    try:
        prompt = prompt_tokens(9, meta["vocab"])
        toks, _, _ = m.generate_greedy(prompt, 12)
        one = m.beam_search(prompt, 12, 1, top_k=1, top_p=1.0, eos_token=-1)
        assert one[0]["tokens"] == [int(t) for t in toks]
        four = m.beam_search(prompt, 12, 4, top_k=0, top_p=1.0, eos_token=-1)
        assert len(four) == 4 and all(four[i]["score"] >= four[i + 1]["score"] for i in range(3))
        assert four[0]["score"] >= m.beam_search(prompt, 12, 1, top_k=0, top_p=1.0, eos_token=-1)[0]["score"] - 1e-5
    finally:
        m.free()
