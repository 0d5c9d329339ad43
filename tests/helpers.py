"""Shared fixtures for the parity tests: synthetic decoder weights and error metrics."""
from __future__ import annotations

import numpy as np

SHAPES = {
    # name: vocab, hidden, layers, heads, inter          (SURVEY.md section 8 "Config shapes")
    "bench-small": dict(vocab=1000, hidden=256, layers=4, heads=4, inter=1024),
    "tiny-test": dict(vocab=512, hidden=128, layers=2, heads=4, inter=384),
    "tinyllama": dict(vocab=32000, hidden=2048, layers=22, heads=32, inter=5632),
    "llama7b": dict(vocab=32000, hidden=4096, layers=32, heads=32, inter=11008),
    "llama13b": dict(vocab=32000, hidden=5120, layers=40, heads=40, inter=13824),
    "llama70b": dict(vocab=32000, hidden=8192, layers=80, heads=64, inter=28672),
}

_SLOT = {"q": 1, "k": 2, "v": 3, "up": 4, "down": 5, "o": 6, "gate": 7}


def uniform(seed: int, shape, a: float) -> np.ndarray:
    return np.random.default_rng(seed).uniform(-a, a, size=shape).astype(np.float32)


def make_model(meta: dict, *, layers: int | None = None, gate: bool = True, norms: bool = True,
               o_proj: bool = True, norm_jitter: float = 0.0) -> dict:
    """Random-init decoder with the reference tensor names (src/model/inference_engine.cpp:483-563),
    all weights [in, out]; seeds follow SURVEY.md 8d (1000*layer + slot, embeddings 777, lm_head 999)."""
    V, H, I = meta["vocab"], meta["hidden"], meta["inter"]
    L = meta["layers"] if layers is None else layers
    w = {"token_embeddings.weight": uniform(777, (V, H), 0.1),
         "lm_head.weight": uniform(999, (H, V), 1.0 / np.sqrt(H))}

    def norm(seed):
        base = np.ones(H, dtype=np.float32)
        if norm_jitter:
            base = base + uniform(seed, (H,), norm_jitter)
        return base

    if norms:
        w["norm.weight"] = norm(555)
    for l in range(L):
        p = f"layers.{l}."
        s = 1000 * l
        w[p + "attention.q_proj.weight"] = uniform(s + 1, (H, H), 1.0 / np.sqrt(H))
        w[p + "attention.k_proj.weight"] = uniform(s + 2, (H, H), 1.0 / np.sqrt(H))
        w[p + "attention.v_proj.weight"] = uniform(s + 3, (H, H), 1.0 / np.sqrt(H))
        if o_proj:
            w[p + "attention.o_proj.weight"] = uniform(s + 6, (H, H), 1.0 / np.sqrt(H))
        w[p + "mlp.up_proj.weight"] = uniform(s + 4, (H, I), 1.0 / np.sqrt(H))
        if gate:
            w[p + "mlp.gate_proj.weight"] = uniform(s + 7, (H, I), 1.0 / np.sqrt(H))
        w[p + "mlp.down_proj.weight"] = uniform(s + 5, (I, H), 1.0 / np.sqrt(I))
        if norms:
            w[p + "attention_norm.weight"] = norm(s + 8)
            w[p + "ffn_norm.weight"] = norm(s + 9)
    return w


def meta_with_layers(meta: dict, layers: int) -> dict:
    m = dict(meta)
    m["layers"] = layers
    return m


def prompt_tokens(n: int, vocab: int, offset: int = 0) -> list:
    """p_i = (7919*i + 1 + offset) mod V  """
    return [int((7919 * i + 1 + offset) % vocab) for i in range(n)]


def rel_err_inf(got: np.ndarray, ref: np.ndarray) -> float:
    """max|got - ref| / max(|ref|_inf, tiny) -- the tolerance rule of SURVEY.md 8a"""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    denom = max(float(np.max(np.abs(ref))) if ref.size else 0.0, 1e-30)
    return float(np.max(np.abs(got - ref))) / denom if ref.size else 0.0


def make_literal_model(vocab: int, hidden: int, layers: int) -> dict:
    """create_test_model of benchmarks/benchmark_inference.cpp:145-225, fp32 arithmetic as written there:
    (float(i % m) / float(m) - 0.5f) * amp.  q/k/v are present, o_proj is not (the attention fall-back), no gate, no norms."""
    f = np.float32

    def ramp(n, shift, mod, amp):
        i = (np.arange(n, dtype=np.int64) + shift) % mod
        return ((i.astype(np.float32) / f(mod) - f(0.5)) * f(amp)).astype(np.float32)

    H, I = hidden, hidden * 4
    w = {"token_embeddings.weight": ramp(vocab * H, 0, 1000, 0.1).reshape(vocab, H),
         "lm_head.weight": ramp(H * vocab, 0, 500, 0.01).reshape(H, vocab)}
    for l in range(layers):
        p = f"layers.{l}."
        w[p + "attention.q_proj.weight"] = ramp(H * H, 0, 100, 0.05).reshape(H, H)
        w[p + "attention.k_proj.weight"] = ramp(H * H, 1, 100, 0.05).reshape(H, H)
        w[p + "attention.v_proj.weight"] = ramp(H * H, 2, 100, 0.05).reshape(H, H)
        w[p + "mlp.up_proj.weight"] = ramp(H * I, 0, 200, 0.02).reshape(H, I)
        w[p + "mlp.down_proj.weight"] = ramp(I * H, 0, 200, 0.02).reshape(I, H)
    return w
