"""The stand-alone entry points SURVEY.md 8b lists beside the model: the paged KV cache manager (reference KVCache,
src/model/inference_engine.cpp:25-172), the fused GEMV (RMSNorm prologue, residual / ReLU / SwiGLU epilogues) and
ti_b200_prefill, each against the oracle composing the reference's separate ops."""
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


@pytest.mark.parametrize("heads,hd,page", [(4, 32, 0), (1, 256, 16), (32, 128, 64)])
def test_kv_cache_append_read_attention(tb, port, heads, hd, page):
    rng = np.random.default_rng(heads + hd)
    layers, max_seq = 2, 200
    kv = tb.KVCache(layers, heads, hd, max_seq, page)
    try:
        ks = [np.zeros((heads, 0, hd), np.float32) for _ in range(layers)]
        vs = [np.zeros((heads, 0, hd), np.float32) for _ in range(layers)]
        for step, nt in enumerate((1, 7, 1, 64, 1, 90)):                     # single tokens and multi-token appends, across pages
            for l in range(layers):
                if l == 1 and step % 2:                                       # layers advance independently
                    continue
                k = rng.standard_normal((heads, nt, hd)).astype(np.float32)
                v = rng.standard_normal((heads, nt, hd)).astype(np.float32)
                kv.append(l, k, v)
                ks[l] = np.concatenate([ks[l], k], axis=1)
                vs[l] = np.concatenate([vs[l], v], axis=1)
        for l in range(layers):
            assert kv.length(l) == (ks[l].shape[1], max_seq)
            rk, rv = kv.read(l)
            assert np.array_equal(rk, ks[l]) and np.array_equal(rv, vs[l])  # update_incremental's (full_keys, full_values)
            q = rng.standard_normal(heads * hd).astype(np.float32)
            got = kv.attention(l, q)
            t = ks[l].shape[1]
            kf = np.ascontiguousarray(ks[l].transpose(1, 0, 2)).reshape(1, t, heads * hd)   # [B, t, H] as the tensor engine takes it
            vf = np.ascontiguousarray(vs[l].transpose(1, 0, 2)).reshape(1, t, heads * hd)
            ref = (port.attention_fast_incremental(q.reshape(1, 1, -1), kf, vf) if heads == 1
                   else port.multi_head_attention(q.reshape(1, 1, -1), kf, vf, heads))
            assert rel_err_inf(got, ref.ravel()) <= 1e-4
        with pytest.raises(tb.B200Error, match="KV cache overflow"):         # inference_engine.cpp:100-102
            kv.append(0, np.zeros((heads, 40, hd), np.float32), np.zeros((heads, 40, hd), np.float32))
        with pytest.raises(tb.B200Error, match="Layer index out of bounds"):  # :82-84
            kv.append(5, np.zeros((heads, 1, hd), np.float32), np.zeros((heads, 1, hd), np.float32))
        kv.reset()
        assert kv.length(0)[0] == 0 and kv.length(1)[0] == 0
        with pytest.raises(tb.B200Error):
            kv.attention(0, np.zeros(heads * hd, np.float32))                # nothing cached
    finally:
        kv.free()


@pytest.mark.parametrize("qt", [oracle.QINT4, oracle.QINT8])
def test_fused_gemv_prologue_and_epilogues(tb, port, qt):
    rng = np.random.default_rng(3 + qt)
    K, N = 1024, 768
    x = rng.standard_normal(K).astype(np.float32) * 3
    nw = (1 + 0.1 * rng.standard_normal(K)).astype(np.float32)
    resid = rng.standard_normal(N).astype(np.float32)
    wa = (rng.uniform(-1, 1, (K, N)) / np.sqrt(K)).astype(np.float32)
    wb = (rng.uniform(-2, 2, (K, N)) / np.sqrt(K)).astype(np.float32)     # a different range: its own scale

    def fq(w):
        s, z = port.quant_info(w, qt, True)
        return port.dequantize(port.quantize(w, qt, s, z), qt, s, z)

    n = port.rms_norm(x.reshape(1, -1), nw)
    single = tb.FusedQWeight([wa], qt)
    pair = tb.FusedQWeight([wa, wb], qt, interleave=True)                  # (gate_i, up_i) interleaved
    cat = tb.FusedQWeight([wa, wb], qt)                                    # concatenated, like q | k | v
    try:
        y = single.gemv_ex(x, tb.EPI_STORE, norm_w=nw)
        assert rel_err_inf(y, port.matmul(n, fq(wa)).ravel()) <= 1e-4      # rms_norm fused as the prologue
        y = single.gemv_ex(x, tb.EPI_RESIDUAL, resid=resid)
        assert rel_err_inf(y, port.add(resid.reshape(1, -1), port.matmul(x.reshape(1, -1), fq(wa))).ravel()) <= 1e-4
        y = single.gemv_ex(x, tb.EPI_RELU)
        assert rel_err_inf(y, port.relu(port.matmul(x.reshape(1, -1), fq(wa))).ravel()) <= 1e-4
        y = pair.gemv_ex(x, tb.EPI_SWIGLU, norm_w=nw)                      # multiply(up, silu(gate)), compute_ffn :389-391
        ref = port.mul(port.matmul(n, fq(wb)), port.silu(port.matmul(n, fq(wa))))
        assert y.shape == (N,) and rel_err_inf(y, ref.ravel()) <= 1e-4
        y = cat.gemv_ex(x)
        ref = np.concatenate([port.matmul(x.reshape(1, -1), fq(wa)).ravel(), port.matmul(x.reshape(1, -1), fq(wb)).ravel()])
        assert rel_err_inf(y, ref) <= 1e-4                                  # each source keeps its own per-tensor scale
        with pytest.raises(tb.B200Error, match="residual"):
            single.gemv_ex(x, tb.EPI_RESIDUAL)
    finally:
        single.free(); pair.free(); cat.free()


@pytest.mark.parametrize("n_prompt", [5, 50])
def test_prefill_entry_then_decode_steps(tb, port, n_prompt):
    meta = SHAPES["bench-small"]
    w = make_model(meta, norm_jitter=0.05)
    prompt = prompt_tokens(n_prompt, meta["vocab"])
    m = tb.Model(meta, oracle.QINT8, attn_mode=1, rope_mode=1, max_seq=128).load(w)
    try:
        logits = m.prefill(prompt)                      # n_prompt = 50: the tensor-core GEMM path
        assert m.kv_length == n_prompt
        toks = [int(np.argmax(logits))]
        for _ in range(5):
            toks.append(m.decode_step(toks[-1])[0])
        g, _, _ = m.generate_greedy(prompt, 6)
    finally:
        m.free()
    assert toks == list(g)
    fq = {k: (port.fake_quant(v, oracle.QINT8) if (v.ndim == 2 and "embeddings" not in k) else v) for k, v in w.items()}
    rt, rl = port.decode_greedy(fq, meta, prompt, 6, attn_mode=1, rope_mode=1)
    assert toks == list(rt) and rel_err_inf(logits, rl[0]) <= 1e-2
