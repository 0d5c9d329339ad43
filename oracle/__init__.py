"""CPU oracles for the decode hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package, and only as the checker or the reported CPU baseline.  The product
(``turboinfer_b200``) never imports it and has no CPU fallback.

Two libraries export the C ABI of ``ti_oracle.h``:

* ``port()``  -> ``oracle/libti_oracle.so``: plain-C restatement (``ti_oracle.c``)
* ``ref()``   -> ``oracle/_ref/libti_ref.so``: the unmodified reference sources compiled from
  ``/root/reference/src`` by ``oracle/Makefile`` (build container only; the built ``.so`` travels to
  the GPU box with the repo snapshot, the sources do not)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "libti_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libti_ref.so")
REFERENCE_ROOT = "/root/reference"

QINT8, QINT4, QNONE = 0, 1, 3

_f = C.POINTER(C.c_float)
_fpp = C.POINTER(_f)


class TioModel(C.Structure):
    _fields_ = [
        ("vocab", C.c_int32), ("hidden", C.c_int32), ("layers", C.c_int32), ("heads", C.c_int32),
        ("inter", C.c_int32), ("rope_theta", C.c_float), ("rms_eps", C.c_float),
        ("attn_mode", C.c_int32), ("rope_mode", C.c_int32),
        ("tok_emb", _f), ("out_norm", _f), ("lm_head", _f),
        ("attn_norm", _fpp), ("wq", _fpp), ("wk", _fpp), ("wv", _fpp), ("wo", _fpp),
        ("ffn_norm", _fpp), ("w_up", _fpp), ("w_gate", _fpp), ("w_down", _fpp),
    ]


def build(which: str = "all") -> None:
    """Compile the oracles.  ``ref`` is only attempted when /root/reference exists."""
    targets = []
    if which in ("all", "port"):
        targets.append("port")
    if which in ("all", "ref") and os.path.isdir(os.path.join(REFERENCE_ROOT, "src")):
        targets.append("ref")
    if targets:
        subprocess.run(["make", "-s", "-C", HERE, "-j8", *targets], check=True)


def _fp(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_f)


class Oracle:
    """Thin numpy front end over one oracle library."""

    def __init__(self, path: str):
        self.path = path
        self.lib = C.CDLL(path)
        L = self.lib
        L.tio_kind.restype = C.c_int
        L.tio_quant_info.restype = C.c_int
        L.tio_quant_info.argtypes = [_f, C.c_size_t, C.c_int, C.c_int, _f, _f]
        L.tio_quantize_int8.argtypes = [_f, C.c_void_p, C.c_size_t, C.c_float, C.c_float]
        L.tio_quantize_int4.argtypes = [_f, C.c_void_p, C.c_size_t, C.c_float, C.c_float]
        L.tio_dequantize_int8.argtypes = [C.c_void_p, _f, C.c_size_t, C.c_float, C.c_float]
        L.tio_dequantize_int4.argtypes = [C.c_void_p, _f, C.c_size_t, C.c_float, C.c_float]
        L.tio_matmul.argtypes = [_f, _f, _f, C.c_size_t, C.c_size_t, C.c_size_t]
        L.tio_rms_norm.argtypes = [_f, _f, _f, C.c_size_t, C.c_size_t, C.c_float]
        L.tio_rope.argtypes = [_f, _f, _f, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_float]
        for n in ("tio_silu", "tio_relu"):
            getattr(L, n).argtypes = [_f, _f, C.c_size_t]
        for n in ("tio_add", "tio_mul"):
            getattr(L, n).argtypes = [_f, _f, _f, C.c_size_t]
        L.tio_softmax.argtypes = [_f, _f, C.c_size_t, C.c_size_t, C.c_float]
        L.tio_attention_fast_incremental.argtypes = [_f, _f, _f, _f, C.c_size_t, C.c_size_t, C.c_size_t]
        L.tio_multi_head_attention.argtypes = [_f, _f, _f, _f, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t]
        L.tio_decode_greedy.restype = C.c_int
        L.tio_decode_greedy.argtypes = [C.POINTER(TioModel), C.POINTER(C.c_int32), C.c_int, C.c_int, C.c_int,
                                        C.POINTER(C.c_int32), _f]
        L.tio_decode_greedy_timed.restype = C.c_int
        L.tio_decode_greedy_timed.argtypes = [C.POINTER(TioModel), C.POINTER(C.c_int32), C.c_int, C.c_int, C.POINTER(C.c_int32),
                                              C.POINTER(C.c_double), C.c_int]
        if hasattr(L, "tio_sample"):   # the C restatement only (the reference's RNG is time-seeded, see ti_oracle.h)
            L.tio_sample.restype = C.c_int
            L.tio_sample.argtypes = [_f, C.c_size_t, C.c_float, C.c_int, C.c_float, C.c_float, _f]
            L.tio_uniform.restype = C.c_float
            L.tio_uniform.argtypes = [C.c_uint64, C.c_uint64]
            L.tio_logprobs.argtypes = [_f, C.c_size_t, C.c_size_t, C.POINTER(C.c_int32), _f]
            i32p = C.POINTER(C.c_int32)
            L.tio_beam_expand.restype = C.c_int
            L.tio_beam_expand.argtypes = [_f, C.c_size_t, C.c_float, C.c_int, C.c_float, C.c_int, _f, i32p]
            L.tio_beam_search.restype = C.c_int
            L.tio_beam_search.argtypes = [C.c_void_p, i32p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_float, C.c_float, C.c_int,
                                          i32p, i32p, _f, _f, i32p]
        L.tio_generate_literal.restype = C.c_int
        L.tio_generate_literal.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32), C.c_int,
                                           C.c_int, C.POINTER(C.c_int32), _f]

    @property
    def kind(self) -> str:
        return "reference" if self.lib.tio_kind() == 1 else "port"

    # ---- level A ----
    def quant_info(self, x: np.ndarray, qtype: int, symmetric: bool = True):
        x = np.ascontiguousarray(x, dtype=np.float32).ravel()
        s, z = C.c_float(), C.c_float()
        rc = self.lib.tio_quant_info(_fp(x), x.size, qtype, int(symmetric), C.byref(s), C.byref(z))
        if rc != 0:
            raise RuntimeError("tio_quant_info failed")
        return np.float32(s.value), np.float32(z.value)

    def quantize(self, x: np.ndarray, qtype: int, scale: float, zp: float) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32)
        if qtype == QINT8:
            q = np.empty(x.shape, dtype=np.int8)
            self.lib.tio_quantize_int8(_fp(x.ravel()), q.ctypes.data, x.size, scale, zp)
        else:
            q = np.empty(x.shape, dtype=np.int32)
            self.lib.tio_quantize_int4(_fp(x.ravel()), q.ctypes.data, x.size, scale, zp)
        return q

    def dequantize(self, q: np.ndarray, qtype: int, scale: float, zp: float) -> np.ndarray:
        out = np.empty(q.shape, dtype=np.float32)
        if qtype == QINT8:
            q = np.ascontiguousarray(q, dtype=np.int8)
            self.lib.tio_dequantize_int8(q.ctypes.data, _fp(out.ravel()), q.size, scale, zp)
        else:
            q = np.ascontiguousarray(q, dtype=np.int32)
            self.lib.tio_dequantize_int4(q.ctypes.data, _fp(out.ravel()), q.size, scale, zp)
        return out

    def fake_quant(self, w: np.ndarray, qtype: int, symmetric: bool = True) -> np.ndarray:
        """dequantize_tensor(quantize_tensor(w)) -- the weights the quantized oracle-B model uses."""
        if qtype == QNONE:
            return np.ascontiguousarray(w, dtype=np.float32)
        s, z = self.quant_info(w, qtype, symmetric)
        return self.dequantize(self.quantize(w, qtype, s, z), qtype, s, z)

    def matmul(self, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.float32)
        b = np.ascontiguousarray(b, dtype=np.float32)
        M, K = a.shape
        K2, N = b.shape
        assert K == K2
        c = np.empty((M, N), dtype=np.float32)
        self.lib.tio_matmul(_fp(a), _fp(b), _fp(c), M, K, N)
        return c

    def rms_norm(self, x: np.ndarray, w: np.ndarray, eps: float = 1e-5) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32)
        w = np.ascontiguousarray(w, dtype=np.float32)
        H = x.shape[-1]
        y = np.empty_like(x)
        self.lib.tio_rms_norm(_fp(x), _fp(w), _fp(y), x.size // H, H, eps)
        return y

    def rope(self, x: np.ndarray, pos: np.ndarray, theta: float = 10000.0) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32)
        pos = np.ascontiguousarray(pos, dtype=np.float32)
        y = np.empty_like(x)
        if x.ndim == 3:
            B, T, D = x.shape
            nh = 1
        else:
            B, nh, T, D = x.shape
        self.lib.tio_rope(_fp(x), _fp(pos), _fp(y), B, nh, T, D, x.ndim, int(pos.ndim == 2), theta)
        return y

    def _unary(self, name: str, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32)
        y = np.empty_like(x)
        getattr(self.lib, name)(_fp(x), _fp(y), x.size)
        return y

    def silu(self, x):
        return self._unary("tio_silu", x)

    def relu(self, x):
        return self._unary("tio_relu", x)

    def _binary(self, name: str, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.float32)
        b = np.ascontiguousarray(b, dtype=np.float32)
        assert a.shape == b.shape
        y = np.empty_like(a)
        getattr(self.lib, name)(_fp(a), _fp(b), _fp(y), a.size)
        return y

    def add(self, a, b):
        return self._binary("tio_add", a, b)

    def mul(self, a, b):
        return self._binary("tio_mul", a, b)

    def softmax(self, x: np.ndarray, temperature: float = 1.0) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32)
        n = x.shape[-1]
        y = np.empty_like(x)
        self.lib.tio_softmax(_fp(x), _fp(y), x.size // n, n, temperature)
        return y

    def attention_fast_incremental(self, q: np.ndarray, k: np.ndarray, v: np.ndarray) -> np.ndarray:
        q = np.ascontiguousarray(q, dtype=np.float32)
        k = np.ascontiguousarray(k, dtype=np.float32)
        v = np.ascontiguousarray(v, dtype=np.float32)
        B, one, H = q.shape
        t = k.shape[1]
        out = np.empty((B, 1, H), dtype=np.float32)
        self.lib.tio_attention_fast_incremental(_fp(q), _fp(k), _fp(v), _fp(out), B, t, H)
        return out

    def multi_head_attention(self, q, k, v, num_heads: int) -> np.ndarray:
        q = np.ascontiguousarray(q, dtype=np.float32)
        k = np.ascontiguousarray(k, dtype=np.float32)
        v = np.ascontiguousarray(v, dtype=np.float32)
        B, one, H = q.shape
        t = k.shape[1]
        out = np.empty((B, 1, H), dtype=np.float32)
        self.lib.tio_multi_head_attention(_fp(q), _fp(k), _fp(v), _fp(out), B, t, H, num_heads)
        return out

    # ---- level B ----
    def _marshal(self, weights: dict, meta: dict, attn_mode: int, rope_mode: int):
        """weights: name -> fp32 array using the reference tensor names of
        InferenceEngineImpl::initialize_model (src/model/inference_engine.cpp:483-563), all [in, out]."""
        L = meta["layers"]
        keep = []

        def arr(name):
            a = weights.get(name)
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=np.float32)
            keep.append(a)
            return a

        def per_layer(fmt):
            ptrs = (_f * L)()
            for l in range(L):
                a = arr(fmt.format(l))
                ptrs[l] = _fp(a) if a is not None else C.cast(None, _f)
            keep.append(ptrs)
            return C.cast(ptrs, _fpp)

        m = TioModel()
        m.vocab, m.hidden, m.layers, m.heads, m.inter = (meta["vocab"], meta["hidden"], L, meta["heads"], meta["inter"])
        m.rope_theta = meta.get("rope_theta", 10000.0)
        m.rms_eps = meta.get("rms_eps", 1e-5)
        m.attn_mode, m.rope_mode = attn_mode, rope_mode
        m.tok_emb = _fp(arr("token_embeddings.weight"))
        on = arr("norm.weight")
        m.out_norm = _fp(on) if on is not None else C.cast(None, _f)
        m.lm_head = _fp(arr("lm_head.weight"))
        m.attn_norm = per_layer("layers.{}.attention_norm.weight")
        m.wq = per_layer("layers.{}.attention.q_proj.weight")
        m.wk = per_layer("layers.{}.attention.k_proj.weight")
        m.wv = per_layer("layers.{}.attention.v_proj.weight")
        m.wo = per_layer("layers.{}.attention.o_proj.weight")
        m.ffn_norm = per_layer("layers.{}.ffn_norm.weight")
        m.w_up = per_layer("layers.{}.mlp.up_proj.weight")
        m.w_gate = per_layer("layers.{}.mlp.gate_proj.weight")
        m.w_down = per_layer("layers.{}.mlp.down_proj.weight")
        return m, keep

    def decode_greedy(self, weights: dict, meta: dict, prompt: Sequence[int], n_new: int, *,
                      attn_mode: int = 1, rope_mode: int = 0, stop_on_eos: bool = False,
                      want_logits: bool = True):
        m, keep = self._marshal(weights, meta, attn_mode, rope_mode)
        p = np.ascontiguousarray(prompt, dtype=np.int32)
        out = np.zeros(max(n_new, 1), dtype=np.int32)
        logits = np.zeros((max(n_new, 1), meta["vocab"]), dtype=np.float32) if want_logits else None
        n = self.lib.tio_decode_greedy(C.byref(m), p.ctypes.data_as(C.POINTER(C.c_int32)), p.size, n_new,
                                       int(stop_on_eos), out.ctypes.data_as(C.POINTER(C.c_int32)),
                                       _fp(logits) if logits is not None else C.cast(None, _f))
        del keep
        if n < 0:
            raise RuntimeError(f"tio_decode_greedy failed ({n})")
        return out[:n].copy(), (logits[:n].copy() if logits is not None else None)

    def decode_greedy_timed(self, weights: dict, meta: dict, prompt: Sequence[int], n_new: int, *,
                            attn_mode: int = 1, rope_mode: int = 0):
        """(tokens, seconds of each of the n_prompt + n_new - 1 forward passes) -- bench.py's CPU arm."""
        m, keep = self._marshal(weights, meta, attn_mode, rope_mode)
        p = np.ascontiguousarray(prompt, dtype=np.int32)
        out = np.zeros(max(n_new, 1), dtype=np.int32)
        cap = p.size + max(n_new, 1)
        secs = np.zeros(cap, dtype=np.float64)
        n = self.lib.tio_decode_greedy_timed(C.byref(m), p.ctypes.data_as(C.POINTER(C.c_int32)), p.size, n_new,
                                             out.ctypes.data_as(C.POINTER(C.c_int32)), secs.ctypes.data_as(C.POINTER(C.c_double)), cap)
        del keep
        if n < 0:
            raise RuntimeError(f"tio_decode_greedy_timed failed ({n})")
        return out[:max(n_new, 0)].copy(), secs[:n].copy()

    # ---- sampling ----
    def uniform(self, seed: int, step: int) -> float:
        return float(self.lib.tio_uniform(seed, step))

    def sample(self, logits: np.ndarray, temperature: float, top_k: int, top_p: float, u: float):
        """(token, logprob) of sample_next_token on one row of logits with the uniform u."""
        lg = np.ascontiguousarray(logits, dtype=np.float32).ravel()
        lp = C.c_float()
        tok = self.lib.tio_sample(_fp(lg), lg.size, temperature, top_k, top_p, u, C.byref(lp))
        return int(tok), float(lp.value)

    def logprobs(self, logits: np.ndarray, tokens) -> np.ndarray:
        lg = np.ascontiguousarray(logits, dtype=np.float32)
        t = np.ascontiguousarray(tokens, dtype=np.int32)
        out = np.zeros(t.size, dtype=np.float32)
        self.lib.tio_logprobs(_fp(lg.reshape(-1)), t.size, lg.shape[-1], t.ctypes.data_as(C.POINTER(C.c_int32)), _fp(out))
        return out

    # ---- beam search (port only) ----
    def beam_expand(self, logits: np.ndarray, beam_size: int, temperature: float, top_k: int, top_p: float):
        """[(probability, token)] best first: the expansion of one candidate, beam_search_decode :1964-2005"""
        lg = np.ascontiguousarray(logits, dtype=np.float32).ravel()
        pr = np.zeros(beam_size, dtype=np.float32)
        tk = np.zeros(beam_size, dtype=np.int32)
        n = self.lib.tio_beam_expand(_fp(lg), lg.size, temperature, top_k, top_p, beam_size, _fp(pr), tk.ctypes.data_as(C.POINTER(C.c_int32)))
        return [(float(pr[i]), int(tk[i])) for i in range(n)]

    def beam_search(self, weights: dict, meta: dict, prompt: Sequence[int], max_new: int, beam_size: int, *, temperature: float = 1.0,
                    top_k: int = 50, top_p: float = 0.9, length_penalty: float = 1.0, eos_token: int = 2,
                    attn_mode: int = 1, rope_mode: int = 0):
        """generate_beam_search (:830-871) -> [dict(tokens (new only), log_prob, score, finished)], best score first"""
        m, keep = self._marshal(weights, meta, attn_mode, rope_mode)
        p = np.ascontiguousarray(prompt, dtype=np.int32)
        i32p = C.POINTER(C.c_int32)
        out = np.zeros((beam_size, max(max_new, 1)), dtype=np.int32)
        lens = np.zeros(beam_size, dtype=np.int32)
        lp = np.zeros(beam_size, dtype=np.float32)
        sc = np.zeros(beam_size, dtype=np.float32)
        fin = np.zeros(beam_size, dtype=np.int32)
        n = self.lib.tio_beam_search(C.cast(C.byref(m), C.c_void_p), p.ctypes.data_as(i32p), p.size, max_new, beam_size, temperature, top_k, top_p,
                                     length_penalty, eos_token, out.ctypes.data_as(i32p), lens.ctypes.data_as(i32p), _fp(lp), _fp(sc),
                                     fin.ctypes.data_as(i32p))
        del keep
        if n < 0:
            raise RuntimeError("tio_beam_search failed")
        return [dict(tokens=[int(t) for t in out[i, : lens[i]]], log_prob=float(lp[i]), score=float(sc[i]), finished=bool(fin[i]))
                for i in range(n)]

    def generate_literal_sampled(self, vocab: int, hidden: int, layers: int, qtype: int, prompt: Sequence[int], n_new: int, *,
                                 temperature: float, top_k: int, top_p: float, u: float = 0.5) -> np.ndarray:
        """generate() on the literal model through the sampling pipeline (both oracles; u matters only to the restatement)"""
        L = self.lib
        i32p = C.POINTER(C.c_int32)
        L.tio_generate_literal_sampled.restype = C.c_int
        L.tio_generate_literal_sampled.argtypes = [C.c_int] * 4 + [i32p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_float, C.c_float, i32p]
        p = np.ascontiguousarray(prompt, dtype=np.int32)
        out = np.zeros(max(n_new, 1), dtype=np.int32)
        n = L.tio_generate_literal_sampled(vocab, hidden, layers, qtype, p.ctypes.data_as(i32p), p.size, n_new, temperature, top_k, top_p, u,
                                           out.ctypes.data_as(i32p))
        if n < 0:
            raise RuntimeError("tio_generate_literal_sampled failed")
        return out[:n].copy()

    def logprobs_literal(self, vocab: int, hidden: int, layers: int, qtype: int, tokens: Sequence[int]) -> np.ndarray:
        """compute_logprobs on the literal benchmark model (both oracles)"""
        L = self.lib
        L.tio_logprobs_literal.restype = C.c_int
        L.tio_logprobs_literal.argtypes = [C.c_int] * 4 + [C.POINTER(C.c_int32), C.c_int, _f]
        t = np.ascontiguousarray(tokens, dtype=np.int32)
        out = np.zeros(t.size, dtype=np.float32)
        if L.tio_logprobs_literal(vocab, hidden, layers, qtype, t.ctypes.data_as(C.POINTER(C.c_int32)), t.size, _fp(out)) != t.size:
            raise RuntimeError("tio_logprobs_literal failed")
        return out

    def beam_search_literal(self, vocab: int, hidden: int, layers: int, qtype: int, prompt: Sequence[int], max_new: int, beam_size: int, *,
                            temperature: float = 1.0, top_k: int = 50, top_p: float = 0.9, length_penalty: float = 1.0):
        """generate_beam_search on the literal benchmark model (both oracles) -> [dict(tokens, avg_logprob, finished)]"""
        L = self.lib
        i32p = C.POINTER(C.c_int32)
        L.tio_beam_search_literal.restype = C.c_int
        L.tio_beam_search_literal.argtypes = [C.c_int] * 4 + [i32p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_float, C.c_float,
                                              i32p, i32p, _f, i32p]
        p = np.ascontiguousarray(prompt, dtype=np.int32)
        out = np.zeros((beam_size, max(max_new, 1)), dtype=np.int32)
        lens = np.zeros(beam_size, dtype=np.int32)
        lp = np.zeros(beam_size, dtype=np.float32)
        fin = np.zeros(beam_size, dtype=np.int32)
        n = L.tio_beam_search_literal(vocab, hidden, layers, qtype, p.ctypes.data_as(i32p), p.size, max_new, beam_size, temperature, top_k, top_p,
                                      length_penalty, out.ctypes.data_as(i32p), lens.ctypes.data_as(i32p), _fp(lp), fin.ctypes.data_as(i32p))
        if n < 0:
            raise RuntimeError("tio_beam_search_literal failed")
        return [dict(tokens=[int(t) for t in out[i, : lens[i]]], avg_logprob=float(lp[i]), finished=bool(fin[i])) for i in range(n)]

    # ---- level C ----
    def generate_literal(self, vocab: int, hidden: int, layers: int, qtype: int, prompt: Sequence[int], n_new: int):
        p = np.ascontiguousarray(prompt, dtype=np.int32)
        out = np.zeros(max(n_new, 1), dtype=np.int32)
        logits = np.zeros(vocab, dtype=np.float32)
        n = self.lib.tio_generate_literal(vocab, hidden, layers, qtype, p.ctypes.data_as(C.POINTER(C.c_int32)),
                                          p.size, n_new, out.ctypes.data_as(C.POINTER(C.c_int32)), _fp(logits))
        if n < 0:
            raise RuntimeError("tio_generate_literal failed")
        return out[:n].copy(), logits


_cache: dict = {}


def port() -> Oracle:
    if "port" not in _cache:
        if not os.path.exists(PORT_SO):
            build("port")
        _cache["port"] = Oracle(PORT_SO)
    return _cache["port"]


def tinq_write_sample(path: str, qtype: int) -> None:
    """reference-only: quantize the model of tests/test_quantization_persistence.cpp with the reference and save it as .tinq"""
    lib = ref().lib
    lib.tio_tinq_write_sample.restype = C.c_int
    lib.tio_tinq_write_sample.argtypes = [C.c_char_p, C.c_int]
    if lib.tio_tinq_write_sample(path.encode(), qtype) != 0:
        raise RuntimeError("reference save_quantized_model failed")


def tinq_resave(src: str, dst: str, qtype: int) -> int:
    """reference-only: load_quantized_model(src) then save_quantized_model(dst); returns the tensor count"""
    lib = ref().lib
    lib.tio_tinq_resave.restype = C.c_int
    lib.tio_tinq_resave.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
    n = lib.tio_tinq_resave(src.encode(), dst.encode(), qtype)
    if n < 0:
        raise RuntimeError("the reference could not load " + src)
    return n


def ref_available() -> bool:
    return os.path.exists(REF_SO)


def ref() -> Oracle:
    if "ref" not in _cache:
        if not os.path.exists(REF_SO):
            build("ref")
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(f"{REF_SO} missing and {REFERENCE_ROOT} not present to build it")
        _cache["ref"] = Oracle(REF_SO)
    return _cache["ref"]
