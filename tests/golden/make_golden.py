"""Regenerates tests/golden/*.npz from the COMPILED REFERENCE (oracle/_ref/libti_ref.so).

Run in the build container (needs /root/reference to build oracle/_ref):
    python tests/golden/make_golden.py
Inputs are the closed-form fixtures of the reference's own tests plus small seeded
cases; outputs are whatever the reference computes.  The reference's tests assert none of these
values themselves, so the compiled reference is the pin.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from helpers import SHAPES, make_model, prompt_tokens  # noqa: E402


def main():
    R = oracle.ref()
    out = {}

    # --- quantization fixtures -------------------------------------------------------------
    quant_inputs = {
        # tests/test_quantization_complete.cpp:26-29, :88-91, :141-144
        "int8_sym_linspace": (np.linspace(-10, 10, 16, dtype=np.float32), oracle.QINT8, True),
        "int4_sym_linspace": (np.linspace(-2, 2, 9, dtype=np.float32), oracle.QINT4, True),
        "int8_asym_linspace": (np.linspace(1, 6, 6, dtype=np.float32), oracle.QINT8, False),
        "int4_asym_linspace": (np.linspace(1, 6, 6, dtype=np.float32), oracle.QINT4, False),
        # tests/test_quantization.cpp:79-86
        "int8_mod256": (((np.arange(100 * 200) % 256) - 128).astype(np.float32), oracle.QINT8, True),
        # tests/test_quantization_persistence.cpp:50-65
        "int8_mod100": (((np.arange(256 * 256) % 100) / np.float32(100) - np.float32(0.5)).astype(np.float32), oracle.QINT8, True),
        "int4_mod127": (((np.arange(256 * 512) % 127) / np.float32(127) - np.float32(0.5)).astype(np.float32), oracle.QINT4, True),
        "int4_random": (np.random.default_rng(5).uniform(-0.02, 0.02, 4096).astype(np.float32), oracle.QINT4, True),
        "int8_random_asym": (np.random.default_rng(6).uniform(-0.01, 0.03, 4096).astype(np.float32), oracle.QINT8, False),
    }
    for name, (x, qt, sym) in quant_inputs.items():
        s, z = R.quant_info(x, qt, sym)
        q = R.quantize(x, qt, s, z)
        out[f"quant/{name}/x"] = x
        out[f"quant/{name}/cfg"] = np.array([qt, int(sym)], dtype=np.int32)
        out[f"quant/{name}/scale_zp"] = np.array([s, z], dtype=np.float32)
        out[f"quant/{name}/q"] = q.astype(np.int8)  # all values fit int8
        out[f"quant/{name}/deq"] = R.dequantize(q, qt, s, z)

    # --- fast incremental attention, tests/test_fast_attention.cpp:58-68 ---------------------
    H = 256
    for t in (10, 50, 100, 200):
        q = (0.1 * (np.arange(H) % 10)).astype(np.float32).reshape(1, 1, H)
        sh = np.add.outer(np.arange(t), np.arange(H))
        k = (np.float32(0.05) * (sh % 20).astype(np.float32)).reshape(1, t, H)
        v = (np.float32(0.02) * ((np.add.outer(2 * np.arange(t), np.arange(H))) % 15).astype(np.float32)).reshape(1, t, H)
        out[f"attn/t{t}/out"] = R.attention_fast_incremental(q, k, v)
        out[f"attn/t{t}/mha4"] = R.multi_head_attention(q, k, v, 4)

    # --- MHA / RoPE ramps, tests/test_advanced_math.cpp:106-124, :130-173 ---------------------
    q = (0.01 * np.arange(12, dtype=np.float32)).reshape(1, 1, 12)
    k = (0.02 * np.arange(36, dtype=np.float32)).reshape(1, 3, 12)
    v = (0.03 * np.arange(36, dtype=np.float32)).reshape(1, 3, 12)
    out["mha3/q"], out["mha3/k"], out["mha3/v"] = q, k, v
    out["mha3/out"] = R.multi_head_attention(q, k, v, 3)
    x3 = (0.1 * np.arange(32, dtype=np.float32)).reshape(1, 4, 8)
    out["rope3/x"], out["rope3/pos"] = x3, np.arange(4, dtype=np.float32)
    out["rope3/out"] = R.rope(x3, np.arange(4, dtype=np.float32))
    x4 = (0.1 * np.arange(24, dtype=np.float32)).reshape(1, 2, 3, 4)
    out["rope4/x"], out["rope4/pos"] = x4, np.arange(3, dtype=np.float32)
    out["rope4/out"] = R.rope(x4, np.arange(3, dtype=np.float32))
    xr = np.random.default_rng(11).standard_normal((1, 32, 1, 128)).astype(np.float32)
    out["rope_dec/x"], out["rope_dec/pos"] = xr, np.array([1777.0], dtype=np.float32)
    out["rope_dec/out"] = R.rope(xr, np.array([1777.0], dtype=np.float32))

    # --- activations and the 2x3 . 3x2 matmul, tests/test_math_ops.cpp:80-129 -----------------
    a = np.array([-2.0, -0.5, 0.5, 2.0], dtype=np.float32)
    out["act/x"] = a
    out["act/silu"], out["act/relu"] = R.silu(a), R.relu(a)
    m1 = np.arange(1, 7, dtype=np.float32).reshape(2, 3)
    m2 = np.arange(1, 7, dtype=np.float32).reshape(3, 2)
    out["mm/a"], out["mm/b"], out["mm/c"] = m1, m2, R.matmul(m1, m2)
    rng = np.random.default_rng(3)
    xa = rng.standard_normal((1, 512)).astype(np.float32)
    wb = rng.uniform(-0.05, 0.05, (512, 96)).astype(np.float32)
    out["gemv/x"], out["gemv/w"], out["gemv/y"] = xa, wb, R.matmul(xa, wb)
    xn = rng.standard_normal((2, 320)).astype(np.float32)
    wn = (1 + 0.1 * rng.standard_normal(320)).astype(np.float32)
    out["rms/x"], out["rms/w"], out["rms/y"] = xn, wn, R.rms_norm(xn, wn)
    sm = rng.standard_normal((3, 12)).astype(np.float32)
    out["softmax/x"], out["softmax/y"] = sm, R.softmax(sm)  # n < 16: the reference's scalar branch

    # --- level C: literal benchmark path, config 1 --------------------------------------------
    for name, qt in (("fp32", oracle.QNONE), ("int8", oracle.QINT8), ("int4", oracle.QINT4)):
        toks, _ = R.generate_literal(1000, 256, 4, qt, [1, 15, 25, 35], 128)
        out[f"literal/{name}/tokens"] = toks

    # --- level B: tiny decoder, fp32 / fake-quant weights ---------------------------------------
    meta = SHAPES["tiny-test"]
    w = make_model(meta, norm_jitter=0.1)
    prompt = prompt_tokens(5, meta["vocab"])
    for qname, qt in (("fp32", oracle.QNONE), ("int8", oracle.QINT8), ("int4", oracle.QINT4)):
        wq = {k_: (R.fake_quant(v_, qt) if (v_.ndim == 2 and "embeddings" not in k_) else v_) for k_, v_ in w.items()}
        for am, rm in ((1, 0), (0, 0), (1, 1)):
            toks, logits = R.decode_greedy(wq, meta, prompt, 24, attn_mode=am, rope_mode=rm)
            out[f"decodeB/{qname}/a{am}r{rm}/tokens"] = toks
            out[f"decodeB/{qname}/a{am}r{rm}/logits_last"] = logits[-1]
            out[f"decodeB/{qname}/a{am}r{rm}/logits_first"] = logits[0]

    path = os.path.join(HERE, "reference_vectors.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
