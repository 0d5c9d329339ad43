"""End-to-end decode on the GPU vs oracle level B: greedy token ids bit-exact, logits <= 1e-2."""
import os

import numpy as np
import pytest

import oracle
from helpers import SHAPES, make_model, meta_with_layers, prompt_tokens, rel_err_inf

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz"))


@pytest.fixture(scope="module")
def tb():
    import turboinfer_b200 as t
    t.init(0)
    return t


def fake_quant_model(port, w, qt):
    return {k: (port.fake_quant(v, qt) if (v.ndim == 2 and "embeddings" not in k) else v) for k, v in w.items()}


def run_pair(tb, port, meta, w, qt, prompt, n_new, am, rm, max_seq=256, engine=None):
    m = tb.Model(meta, qt, attn_mode=am, rope_mode=rm, max_seq=max_seq).load(w)
    try:
        if engine is not None:   # which engine must have been picked: True = persistent kernel, False = per-op graph
            assert m.persistent_engine is engine
        toks, logits, ms = m.generate_greedy(prompt, n_new, want_logits=True)
    finally:
        m.free()
    rt, rl = port.decode_greedy(fake_quant_model(port, w, qt), meta, prompt, n_new, attn_mode=am, rope_mode=rm)
    return toks, logits, rt, rl


@pytest.mark.parametrize("qname,qt", [("int8", oracle.QINT8), ("int4", oracle.QINT4)])
@pytest.mark.parametrize("am,rm", [(1, 0), (0, 0), (1, 1)])
def test_tiny_model_vs_golden_reference(tb, qname, qt, am, rm):
    meta = SHAPES["tiny-test"]
    w = make_model(meta, norm_jitter=0.1)
    m = tb.Model(meta, qt, attn_mode=am, rope_mode=rm, max_seq=128).load(w)
    try:
        toks, logits, _ = m.generate_greedy(prompt_tokens(5, meta["vocab"]), 24, want_logits=True)
    finally:
        m.free()
    key = f"decodeB/{qname}/a{am}r{rm}"
    assert np.array_equal(toks, G[key + "/tokens"])
    assert rel_err_inf(logits[0], G[key + "/logits_first"]) <= 1e-2
    assert rel_err_inf(logits[-1], G[key + "/logits_last"]) <= 1e-2


@pytest.mark.parametrize("qt", [oracle.QINT8, oracle.QINT4])
@pytest.mark.parametrize("am,rm", [(1, 0), (0, 0), (1, 1), (0, 2)])
def test_bench_small_shape_vs_oracle(tb, port, qt, am, rm):
    meta = SHAPES["bench-small"]
    w = make_model(meta, norm_jitter=0.05)
    toks, logits, rt, rl = run_pair(tb, port, meta, w, qt, prompt_tokens(4, meta["vocab"]), 48, am, rm)
    assert np.array_equal(toks, rt)
    assert rel_err_inf(logits, rl) <= 1e-2


@pytest.mark.parametrize("variant", [dict(gate=False), dict(o_proj=False), dict(norms=False), dict(o_proj=False, norms=False, gate=False)])
def test_null_weight_fallbacks(tb, port, variant):
    meta = SHAPES["tiny-test"]
    w = make_model(meta, norm_jitter=0.1, **variant)
    # a layer without attention (o_proj absent: x <- x + norm(x)) has no phase in the persistent kernel: such a model must
    # be decoded by the per-op engine; a missing gate (relu(up)) or missing norms are phases / prologues the persistent
    # kernel does have, so those variants test it with absent tensors
    toks, logits, rt, rl = run_pair(tb, port, meta, w, oracle.QINT8, prompt_tokens(3, meta["vocab"]), 10, 1, 0,
                                    engine=variant.get("o_proj") is not False)
    assert np.array_equal(toks, rt)
    assert rel_err_inf(logits, rl) <= 1e-2


def test_decode_step_api_and_kv_semantics(tb, port):
    meta = SHAPES["tiny-test"]
    w = make_model(meta, norm_jitter=0.1)
    m = tb.Model(meta, oracle.QINT8, attn_mode=1, max_seq=8).load(w)
    try:
        prompt = prompt_tokens(4, meta["vocab"])
        assert m.kv_length == 0
        last = None
        for tok in prompt:
            last = m.decode_step(tok)
        assert m.kv_length == 4
        rt, rl = port.decode_greedy(fake_quant_model(port, w, oracle.QINT8), meta, prompt, 1, attn_mode=1)
        assert last[0] == rt[0]
        assert rel_err_inf(last[1], rl[0]) <= 1e-2
        for _ in range(4):
            m.decode_step(1)
        with pytest.raises(tb.B200Error, match="KV cache overflow"):      # inference_engine.cpp:100-102
            m.decode_step(1)
        m.reset()
        assert m.kv_length == 0
        again = m.decode_step(prompt[0])
        m.reset()
        assert np.array_equal(m.decode_step(prompt[0])[1], again[1])       # reset really forgets the cache
        with pytest.raises(tb.B200Error):
            m.generate_greedy([], 4)
        with pytest.raises(tb.B200Error):
            m.generate_greedy([meta["vocab"]], 4)
    finally:
        m.free()


def test_eos_stop(tb, port):
    meta = SHAPES["tiny-test"]
    w = make_model(meta, norm_jitter=0.1)
    # make token 2 irresistible: huge lm_head column
    w = dict(w)
    lm = w["lm_head.weight"].copy()
    m0 = tb.Model(meta, oracle.QINT8, max_seq=64).load(w)
    try:
        full, _, _ = m0.generate_greedy([5, 6], 12)
        stopped, _, _ = m0.generate_greedy([5, 6], 12, stop_on_eos=True)
    finally:
        m0.free()
    if 2 in list(full):
        assert list(stopped) == list(full[: list(full).index(2) + 1])
    else:
        assert np.array_equal(stopped, full)


@pytest.mark.parametrize("shape,qt,n_new,layers", [("tinyllama", oracle.QINT4, 6, 2), ("llama7b", oracle.QINT4, 3, 2), ("llama7b", oracle.QINT8, 3, 2),
                                                   ("llama13b", oracle.QINT8, 2, 1)])
def test_full_width_truncated_depth(tb, port, shape, qt, n_new, layers):
    """Full-width, truncated depth: the end-to-end parity case for the big shapes (13B: the widths of configs[3])."""
    meta = meta_with_layers(SHAPES[shape], layers)
    w = make_model(meta)
    prompt = prompt_tokens(4, meta["vocab"])
    toks, logits, rt, rl = run_pair(tb, port, meta, w, qt, prompt, n_new, 1, 1, max_seq=64, engine=True)
    top2 = np.sort(rl, axis=-1)[:, -2:]
    margin = (top2[:, 1] - top2[:, 0]) / np.abs(rl).max()
    err = rel_err_inf(logits, rl)
    print(f"{shape} q{qt}: logits rel err {err:.3e}, min top-2 margin {margin.min():.3e}")
    assert err <= 1e-2
    assert np.array_equal(toks, rt)


def _quantized_tensors(port, w, qt):
    """Quantizer::quantize_model on the CPU oracle: integers + (scale, zero_point) per projection / lm_head tensor"""
    out = {}
    for name, v in w.items():
        if v.ndim == 2 and "embeddings" not in name:
            scale, zp = port.quant_info(v, qt, True)
            out[name] = (port.quantize(v, qt, scale, zp), float(scale), float(zp))
    return out


@pytest.mark.parametrize("qt", [oracle.QINT8, oracle.QINT4])
def test_model_from_quantized_integers(tb, port, qt):
    """ti_b200_model_set_tensor_q (the .tinq path): integers quantized elsewhere are packed as they are -- the model
    is bit-identical to the one the device quantizes from the same float32 tensors"""
    meta = SHAPES["tiny-test"]
    w = make_model(meta, norm_jitter=0.1)
    prompt = prompt_tokens(5, meta["vocab"])
    a = tb.Model(meta, qt, max_seq=64).load(w)
    b = tb.Model(meta, qt, max_seq=64)
    try:
        qs = _quantized_tensors(port, w, qt)
        for name, v in w.items():
            if name in qs:
                q, scale, zp = qs[name]
                b.set_tensor_q(name, q, qt, scale, zp)
            else:
                b.set_tensor(name, v)
        b.finalize()
        ta, la, _ = a.generate_greedy(prompt, 12, want_logits=True)
        tb_, lb, _ = b.generate_greedy(prompt, 12, want_logits=True)
        assert list(ta) == list(tb_)
        np.testing.assert_array_equal(la, lb)
        # wrong element type / wrong quantization type / non-positive scale / float-only slots are refused
        c = tb.Model(meta, qt, max_seq=64)
        name = "layers.0.attention.q_proj.weight"
        q, scale, zp = qs[name]
        with pytest.raises(TypeError):
            c.set_tensor_q(name, q.astype(np.int16), qt, scale, zp)
        other = oracle.QINT4 if qt == oracle.QINT8 else oracle.QINT8
        with pytest.raises(RuntimeError):
            c.set_tensor_q(name, q.astype(np.int32 if other == oracle.QINT4 else np.int8), other, scale, zp)
        with pytest.raises(RuntimeError):
            c.set_tensor_q(name, q, qt, 0.0, zp)
        with pytest.raises(RuntimeError):
            c.set_tensor_q("norm.weight", q[:1], qt, scale, zp)
        if qt == oracle.QINT4:   # integers outside the nibble code are refused instead of wrapping
            bad = q.copy()
            bad[0, 0] = 9
            with pytest.raises(RuntimeError, match="INT4 integers"):
                c.set_tensor_q(name, bad, qt, scale, zp)
        c.free()
    finally:
        a.free()
        b.free()


def test_model_from_asymmetric_int4_integers(tb, port):
    """INT4 integers made WITH a zero-point live in [0, 15] (quantize_to_int4, quantization.cpp:683-693).  Re-coding a symmetric
    tensor as (q + 8, zero_point - 8) describes the same real values, so the logits agree to rounding of the zero-point term."""
    meta = SHAPES["tiny-test"]
    w = make_model(meta, norm_jitter=0.1)
    prompt = prompt_tokens(5, meta["vocab"])
    a = tb.Model(meta, oracle.QINT4, max_seq=64).load(w)
    b = tb.Model(meta, oracle.QINT4, max_seq=64)
    try:
        qs = _quantized_tensors(port, w, oracle.QINT4)
        for name, v in w.items():
            if name in qs:
                q, scale, zp = qs[name]
                assert zp == 0.0 and q.min() >= -7 and q.max() <= 7
                b.set_tensor_q(name, (q + 8).astype(np.int32), oracle.QINT4, scale, -8.0)
            else:
                b.set_tensor(name, v)
        b.finalize()
        ta, la, _ = a.generate_greedy(prompt, 8, want_logits=True)
        tb_, lb, _ = b.generate_greedy(prompt, 8, want_logits=True)
        assert rel_err_inf(lb, la) <= 1e-4     # the same real weights through the zero-point term: fp32 rounding only
        assert list(ta) == list(tb_)
    finally:
        a.free()
        b.free()


def test_eos_stop_ends_the_device_loop_early(tb):
    """stop_on_eos: the persistent launch is issued in chunks of 32 steps and the host stops issuing once token 2 has appeared, so
    a generation that ends early does not decode all n_new tokens.  The EOS position is planted by relabelling: swapping vocabulary
    entries T <-> 2 (embedding rows and lm_head columns) turns the first occurrence of T into the first EOS."""
    meta = SHAPES["tiny-test"]
    n_new = 120
    found = None
    for seed in range(12):
        w = make_model(meta, norm_jitter=0.1)
        w["lm_head.weight"] = (w["lm_head.weight"] * 6.0).astype(np.float32)
        prompt = prompt_tokens(5, meta["vocab"], offset=17 * seed)
        m = tb.Model(meta, oracle.QINT8, rope_mode=1, max_seq=160).load(w)
        try:
            full = [int(t) for t in m.generate_greedy(prompt, n_new)[0]]
        finally:
            m.free()
        first = {}
        for i, t in enumerate(full):
            first.setdefault(t, i)
        cands = [(i, t) for t, i in first.items() if 34 <= i <= 60 and t != 2 and 2 not in full[: i + 1] and t not in prompt and 2 not in prompt]
        if cands:
            found = (w, prompt, full, min(cands))
            break
    if found is None:
        pytest.skip("no token with a first occurrence between steps 34 and 60 in these generations")
    w, prompt, full, (idx, T) = found
    w2 = dict(w)
    emb, lm = w["token_embeddings.weight"].copy(), w["lm_head.weight"].copy()
    emb[[2, T]] = emb[[T, 2]]
    lm[:, [2, T]] = lm[:, [T, 2]]
    w2["token_embeddings.weight"], w2["lm_head.weight"] = emb, lm
    relabel = {2: T, T: 2}
    m = tb.Model(meta, oracle.QINT8, rope_mode=1, max_seq=160).load(w2)
    try:
        whole = [int(t) for t in m.generate_greedy(prompt, n_new)[0]]
        assert [relabel.get(t, t) for t in whole] == full          # the relabelled model walks the same path
        l0 = tb.launch_count()
        stopped = [int(t) for t in m.generate_greedy(prompt, n_new, stop_on_eos=True)[0]]
        launches = tb.launch_count() - l0
        assert stopped == whole[: idx + 1] and stopped[-1] == 2
        # prompt launch + ceil(idx / 32) chunks instead of the 4 chunks of 119 steps
        assert launches <= 1 + (idx + 31) // 32 + 1 < 1 + 4
    finally:
        m.free()
