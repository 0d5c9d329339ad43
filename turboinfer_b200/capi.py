"""ctypes bindings of include/ti_b200.h (numpy in / numpy out).  No torch, no fallback."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
Q_INT8, Q_INT4, Q_NONE = 0, 1, 3

_f = C.POINTER(C.c_float)
_i32 = C.POINTER(C.c_int32)


class B200Error(RuntimeError):
    pass


class ModelConfig(C.Structure):
    _fields_ = [("vocab", C.c_int32), ("hidden", C.c_int32), ("layers", C.c_int32), ("heads", C.c_int32),
                ("inter", C.c_int32), ("rope_theta", C.c_float), ("rms_eps", C.c_float), ("qtype", C.c_int32),
                ("attn_mode", C.c_int32), ("rope_mode", C.c_int32), ("max_seq", C.c_int32),
                ("kv_page_tokens", C.c_int32), ("compat_literal", C.c_int32), ("reserved", C.c_int32 * 7)]


# every symbol include/ti_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "ti_b200_abi_version": (C.c_int, []),
    "ti_b200_init": (C.c_int, [C.c_int]),
    "ti_b200_shutdown": (C.c_int, []),
    "ti_b200_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "ti_b200_device_info": (C.c_int, [C.c_char_p, C.c_size_t]),
    "ti_b200_last_error": (C.c_char_p, []),
    "ti_b200_sync": (C.c_int, []),
    "ti_b200_tp_unique_id": (C.c_int, [C.c_char_p, C.c_size_t]),
    "ti_b200_tp_init": (C.c_int, [C.c_int, C.c_int, C.c_char_p, C.c_size_t]),
    "ti_b200_malloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "ti_b200_free": (C.c_int, [C.c_void_p]),
    "ti_b200_memcpy_h2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "ti_b200_memcpy_d2h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "ti_b200_quant_info": (C.c_int, [_f, C.c_size_t, C.c_int, C.c_int, _f, _f]),
    "ti_b200_quantize": (C.c_int, [_f, C.c_size_t, C.c_int, C.c_float, C.c_float, C.c_void_p]),
    "ti_b200_dequantize": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.c_float, C.c_float, _f]),
    "ti_b200_quantize_pack": (C.c_int, [_f, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_uint64)]),
    "ti_b200_qweight_info": (C.c_int, [C.c_uint64, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(C.c_int), _f, _f,
                                       C.POINTER(C.c_size_t)]),
    "ti_b200_qweight_unpack": (C.c_int, [C.c_uint64, _i32]),
    "ti_b200_qweight_free": (C.c_int, [C.c_uint64]),
    "ti_b200_gemv_q": (C.c_int, [C.c_uint64, _f, _f, C.c_size_t]),
    "ti_b200_gemm_q": (C.c_int, [C.c_uint64, _f, _f, C.c_size_t]),
    "ti_b200_bench_gemm": (C.c_int, [C.c_uint64, C.c_size_t, C.c_size_t, _f, C.POINTER(C.c_double)]),
    "ti_b200_matmul_f32": (C.c_int, [_f, _f, _f, C.c_size_t, C.c_size_t, C.c_size_t]),
    "ti_b200_rms_norm": (C.c_int, [_f, _f, _f, C.c_size_t, C.c_size_t, C.c_float]),
    "ti_b200_rope": (C.c_int, [_f, _f, _f, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_float]),
    "ti_b200_silu": (C.c_int, [_f, _f, C.c_size_t]),
    "ti_b200_relu": (C.c_int, [_f, _f, C.c_size_t]),
    "ti_b200_add": (C.c_int, [_f, _f, _f, C.c_size_t]),
    "ti_b200_mul": (C.c_int, [_f, _f, _f, C.c_size_t]),
    "ti_b200_silu_mul": (C.c_int, [_f, _f, _f, C.c_size_t]),
    "ti_b200_softmax": (C.c_int, [_f, _f, C.c_size_t, C.c_size_t, C.c_float]),
    "ti_b200_attention_decode": (C.c_int, [_f, _f, _f, _f, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t]),
    "ti_b200_quantize_pack_fused": (C.c_int, [C.POINTER(_f), C.POINTER(C.c_size_t), C.c_int32, C.c_int32, C.c_size_t, C.c_int, C.POINTER(C.c_uint64)]),
    "ti_b200_gemv_q_ex": (C.c_int, [C.c_uint64, _f, _f, C.c_int32, _f, _f, C.c_float]),
    "ti_b200_kv_create": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_uint64)]),
    "ti_b200_kv_destroy": (C.c_int, [C.c_uint64]),
    "ti_b200_kv_reset": (C.c_int, [C.c_uint64]),
    "ti_b200_kv_length": (C.c_int, [C.c_uint64, C.c_int32, _i32, _i32]),
    "ti_b200_kv_append": (C.c_int, [C.c_uint64, C.c_int32, _f, _f, C.c_int32]),
    "ti_b200_kv_read": (C.c_int, [C.c_uint64, C.c_int32, _f, _f]),
    "ti_b200_kv_attention": (C.c_int, [C.c_uint64, C.c_int32, _f, _f]),
    "ti_b200_prefill": (C.c_int, [C.c_uint64, _i32, C.c_int32, _f]),
    "ti_b200_gemv_q_dev": (C.c_int, [C.c_uint64, C.c_void_p, C.c_void_p]),
    "ti_b200_model_new": (C.c_int, [C.POINTER(ModelConfig), C.POINTER(C.c_uint64)]),
    "ti_b200_model_set_tensor": (C.c_int, [C.c_uint64, C.c_char_p, _f, C.c_size_t, C.c_size_t]),
    "ti_b200_model_set_tensor_q": (C.c_int, [C.c_uint64, C.c_char_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_float, C.c_float]),
    "ti_b200_model_set_tensor_synthetic": (C.c_int, [C.c_uint64, C.c_char_p, C.c_size_t, C.c_size_t, C.c_uint64, C.c_float]),
    "ti_b200_model_finalize": (C.c_int, [C.c_uint64]),
    "ti_b200_model_free": (C.c_int, [C.c_uint64]),
    "ti_b200_model_reset": (C.c_int, [C.c_uint64]),
    "ti_b200_model_kv_length": (C.c_int, [C.c_uint64, _i32]),
    "ti_b200_model_engine": (C.c_int, [C.c_uint64, _i32]),
    "ti_b200_model_step_bytes": (C.c_int, [C.c_uint64, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "ti_b200_decode_step": (C.c_int, [C.c_uint64, C.c_int32, _f, _i32]),
    "ti_b200_generate_greedy": (C.c_int, [C.c_uint64, _i32, C.c_int32, C.c_int32, C.c_int32, _i32, _i32, _f, _f]),
    "ti_b200_generate_batch_greedy": (C.c_int, [C.c_uint64, _i32, C.c_int32, C.c_int32, C.c_int32, _i32, _f, _f]),
    "ti_b200_sample_logits": (C.c_int, [_f, C.c_size_t, C.c_size_t, C.c_float, C.c_int32, C.c_float, C.c_uint64, C.c_int32, _i32, _f]),
    "ti_b200_generate_sampled": (C.c_int, [C.c_uint64, _i32, C.c_int32, C.c_int32, C.c_float, C.c_int32, C.c_float, C.c_uint64, C.c_int32, _i32, _i32, _f, _f]),
    "ti_b200_compute_logprobs": (C.c_int, [C.c_uint64, _i32, C.c_int32, _f]),
    "ti_b200_generate_batch_ragged": (C.c_int, [C.c_uint64, _i32, _i32, C.c_int32, C.c_int32, C.c_int32, _i32, _f]),
    "ti_b200_beam_search": (C.c_int, [C.c_uint64, _i32, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_int32, C.c_float, C.c_float, C.c_int32,
                                      _i32, _i32, _f, _f, _i32, _i32]),
    "ti_b200_beam_expand": (C.c_int, [_f, C.c_size_t, C.c_size_t, C.c_float, C.c_int32, C.c_float, C.c_int32, _f, _i32, _i32]),
    "ti_b200_model_last_prefill_ms": (C.c_int, [C.c_uint64, _f]),
    "ti_b200_launch_count": (C.c_int, [C.POINTER(C.c_uint64)]),
    "ti_b200_bench_gemv": (C.c_int, [C.POINTER(C.c_uint64), C.c_size_t, C.c_size_t, _f]),
    "ti_b200_model_bench_gemv": (C.c_int, [C.c_uint64, C.c_int, C.c_size_t, _f, C.POINTER(C.c_double)]),
    "ti_b200_debug_timeline": (C.c_int, [C.c_uint64, C.c_int32, C.POINTER(C.c_int64), C.c_size_t, C.POINTER(C.c_size_t)]),
    "ti_b200_debug_timeline_all": (C.c_int, [C.c_uint64, C.c_int32, C.POINTER(C.c_int64), C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
}

_lib: Optional[C.CDLL] = None
_inited = False


def library_path() -> str:
    # TURBOINFER_B200_LIB: an alternative build of the same library (A/B experiments with compile-time switches)
    return os.environ.get("TURBOINFER_B200_LIB") or os.path.join(HERE, "libturboinfer_b200.so")


def lib() -> C.CDLL:
    """Loads the extension (building it with nvcc if the in-tree .so is missing or stale)."""
    global _lib
    if _lib is None:
        path = library_path()
        srcdir = os.path.join(HERE, "csrc")
        if os.path.isdir(srcdir) and os.path.exists("/usr/local/cuda/bin/nvcc") and not os.environ.get("TURBOINFER_B200_LIB"):
            from . import build as _build
            _build.build()
        if not os.path.exists(path):
            raise B200Error(f"{path} is missing: build it with `python -m turboinfer_b200.build` (no CPU fallback exists)")
        L = C.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _ck(rc: int) -> None:
    if rc != 0:
        raise B200Error(lib().ti_b200_last_error().decode("utf-8", "replace"))


def init(device: int = 0) -> None:
    global _inited
    _ck(lib().ti_b200_init(device))
    _inited = True


def shutdown() -> None:
    global _inited
    if _lib is not None:
        _ck(_lib.ti_b200_shutdown())
    _inited = False


def _need() -> C.CDLL:
    if not _inited:
        init(int(os.environ.get("LOCAL_RANK", "0")))
    return lib()


def tp_unique_id() -> bytes:
    """Rank 0: the NCCL id of a new tensor-parallel group; hand it to the other ranks, then tp_init() everywhere."""
    buf = C.create_string_buffer(128)
    _ck(_need().ti_b200_tp_unique_id(buf, 128))
    return buf.raw


def tp_init(nranks: int, rank: int, unique_id: bytes) -> None:
    _ck(_need().ti_b200_tp_init(nranks, rank, unique_id, len(unique_id)))


def device_info() -> str:
    buf = C.create_string_buffer(512)
    _ck(_need().ti_b200_device_info(buf, 512))
    return buf.value.decode()


def launch_count() -> int:
    n = C.c_uint64()
    _ck(lib().ti_b200_launch_count(C.byref(n)))
    return n.value


def _fp(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_f)


def _c(a, dtype=np.float32) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=dtype)


class QWeight:
    """A packed INT4 / INT8 weight resident in HBM (Quantizer::quantize_tensor + device re-tiling)."""

    def __init__(self, w: np.ndarray, qtype: int, symmetric: bool = True):
        w = _c(w)
        assert w.ndim == 2
        h = C.c_uint64()
        _ck(_need().ti_b200_quantize_pack(_fp(w), w.shape[0], w.shape[1], qtype, int(symmetric), C.byref(h)))
        self.handle = h.value
        K, N, qt, by = C.c_size_t(), C.c_size_t(), C.c_int(), C.c_size_t()
        s, z = C.c_float(), C.c_float()
        _ck(lib().ti_b200_qweight_info(self.handle, C.byref(K), C.byref(N), C.byref(qt), C.byref(s), C.byref(z), C.byref(by)))
        self.K, self.N, self.qtype, self.packed_bytes = K.value, N.value, qt.value, by.value
        self.scale, self.zero_point = np.float32(s.value), np.float32(z.value)

    def unpack(self) -> np.ndarray:
        q = np.empty((self.K, self.N), dtype=np.int32)
        _ck(lib().ti_b200_qweight_unpack(self.handle, q.ctypes.data_as(_i32)))
        return q

    def gemv(self, x: np.ndarray) -> np.ndarray:
        x = _c(x)
        rows = x.size // self.K
        y = np.empty((rows, self.N), dtype=np.float32)
        _ck(lib().ti_b200_gemv_q(self.handle, _fp(x), _fp(y), rows))
        return y

    def gemm(self, x: np.ndarray) -> np.ndarray:
        """Y[rows, N] on the tcgen05 tensor cores (prefill / batched decode); rows are bit-identical to gemv()."""
        x = _c(x)
        rows = x.size // self.K
        y = np.empty((rows, self.N), dtype=np.float32)
        _ck(lib().ti_b200_gemm_q(self.handle, _fp(x), _fp(y), rows))
        return y

    def bench_gemm(self, rows: int, reps: int):
        """(avg ms per launch of the GEMM kernel alone, integer ops per launch)."""
        ms, ops = C.c_float(), C.c_double()
        _ck(lib().ti_b200_bench_gemm(self.handle, rows, reps, C.byref(ms), C.byref(ops)))
        return ms.value, ops.value

    def free(self) -> None:
        if self.handle:
            _ck(lib().ti_b200_qweight_free(self.handle))
            self.handle = 0


EPI_STORE, EPI_RESIDUAL, EPI_SWIGLU, EPI_RELU = 0, 1, 2, 3


class FusedQWeight(QWeight):
    """Several [K, n_i] matrices that share x, packed as one streaming weight (q | k | v, or gate / up interleaved)."""

    def __init__(self, mats, qtype: int, interleave: bool = False):
        mats = [_c(m) for m in mats]
        K = mats[0].shape[0]
        ptrs = (_f * len(mats))(*[_fp(m) for m in mats])
        ncols = (C.c_size_t * len(mats))(*[m.shape[1] for m in mats])
        h = C.c_uint64()
        _ck(_need().ti_b200_quantize_pack_fused(ptrs, ncols, len(mats), int(interleave), K, qtype, C.byref(h)))
        self.handle = h.value
        self.K, self.N, self.qtype = K, sum(m.shape[1] for m in mats), qtype

    def gemv_ex(self, x, epilogue: int = EPI_STORE, resid=None, norm_w=None, eps: float = 1e-5) -> np.ndarray:
        x = _c(x).ravel()
        n_out = self.N // 2 if epilogue == EPI_SWIGLU else self.N
        y = np.empty(n_out, dtype=np.float32)
        r = _c(resid).ravel() if resid is not None else None
        nw = _c(norm_w).ravel() if norm_w is not None else None
        _ck(lib().ti_b200_gemv_q_ex(self.handle, _fp(x), _fp(y), epilogue, _fp(r) if r is not None else C.cast(None, _f),
                                    _fp(nw) if nw is not None else C.cast(None, _f), eps))
        return y


class KVCache:
    """Stand-alone paged KV cache (reference KVCache, src/model/inference_engine.cpp:25-172), device resident."""

    def __init__(self, layers: int, heads: int, head_dim: int, max_seq: int, page_tokens: int = 0):
        h = C.c_uint64()
        _ck(_need().ti_b200_kv_create(layers, heads, head_dim, max_seq, page_tokens, C.byref(h)))
        self.handle, self.layers, self.heads, self.head_dim = h.value, layers, heads, head_dim

    def append(self, layer: int, k_new, v_new) -> None:
        k, v = _c(k_new), _c(v_new)           # [heads, new_tokens, head_dim]
        _ck(lib().ti_b200_kv_append(self.handle, layer, _fp(k.reshape(-1)), _fp(v.reshape(-1)), k.shape[1]))

    def length(self, layer: int = 0):
        cur, mx = C.c_int32(), C.c_int32()
        _ck(lib().ti_b200_kv_length(self.handle, layer, C.byref(cur), C.byref(mx)))
        return cur.value, mx.value

    def read(self, layer: int):
        n = self.length(layer)[0]
        k = np.zeros((self.heads, n, self.head_dim), dtype=np.float32)
        v = np.zeros_like(k)
        _ck(lib().ti_b200_kv_read(self.handle, layer, _fp(k.reshape(-1)), _fp(v.reshape(-1))))
        return k, v

    def attention(self, layer: int, q) -> np.ndarray:
        q = _c(q).ravel()
        out = np.empty_like(q)
        _ck(lib().ti_b200_kv_attention(self.handle, layer, _fp(q), _fp(out)))
        return out

    def reset(self) -> None:
        _ck(lib().ti_b200_kv_reset(self.handle))

    def free(self) -> None:
        if self.handle:
            _ck(lib().ti_b200_kv_destroy(self.handle))
            self.handle = 0


class _Ops:
    """TensorEngine / Quantizer entry points on host arrays (upload -> kernel -> download)."""

    def quant_info(self, x, qtype: int, symmetric: bool = True):
        x = _c(x).ravel()
        s, z = C.c_float(), C.c_float()
        _ck(_need().ti_b200_quant_info(_fp(x), x.size, qtype, int(symmetric), C.byref(s), C.byref(z)))
        return np.float32(s.value), np.float32(z.value)

    def quantize(self, x, qtype: int, scale: float, zp: float) -> np.ndarray:
        x = _c(x)
        q = np.empty(x.shape, dtype=np.int8 if qtype == Q_INT8 else np.int32)
        _ck(_need().ti_b200_quantize(_fp(x.ravel()), x.size, qtype, scale, zp, q.ctypes.data))
        return q

    def dequantize(self, q, qtype: int, scale: float, zp: float) -> np.ndarray:
        q = _c(q, np.int8 if qtype == Q_INT8 else np.int32)
        x = np.empty(q.shape, dtype=np.float32)
        _ck(_need().ti_b200_dequantize(q.ctypes.data, q.size, qtype, scale, zp, _fp(x.ravel())))
        return x

    def matmul(self, a, b) -> np.ndarray:
        a, b = _c(a), _c(b)
        M, K = a.shape
        N = b.shape[1]
        c = np.empty((M, N), dtype=np.float32)
        _ck(_need().ti_b200_matmul_f32(_fp(a), _fp(b), _fp(c), M, K, N))
        return c

    def rms_norm(self, x, w, eps: float = 1e-5) -> np.ndarray:
        x, w = _c(x), _c(w)
        H = x.shape[-1]
        y = np.empty_like(x)
        _ck(_need().ti_b200_rms_norm(_fp(x), _fp(w), _fp(y), x.size // H, H, eps))
        return y

    def rope(self, x, pos, theta: float = 10000.0) -> np.ndarray:
        x, pos = _c(x), _c(pos)
        y = np.empty_like(x)
        if x.ndim == 3:
            B, T, D = x.shape
            nh = 1
        else:
            B, nh, T, D = x.shape
        _ck(_need().ti_b200_rope(_fp(x), _fp(pos), _fp(y), B, nh, T, D, x.ndim, int(pos.ndim == 2), theta))
        return y

    def _unary(self, fn, x):
        x = _c(x)
        y = np.empty_like(x)
        _ck(fn(_fp(x), _fp(y), x.size))
        return y

    def silu(self, x):
        return self._unary(_need().ti_b200_silu, x)

    def relu(self, x):
        return self._unary(_need().ti_b200_relu, x)

    def _binary(self, fn, a, b):
        a, b = _c(a), _c(b)
        y = np.empty_like(a)
        _ck(fn(_fp(a), _fp(b), _fp(y), a.size))
        return y

    def add(self, a, b):
        return self._binary(_need().ti_b200_add, a, b)

    def mul(self, a, b):
        return self._binary(_need().ti_b200_mul, a, b)

    def silu_mul(self, gate, up):
        return self._binary(_need().ti_b200_silu_mul, gate, up)

    def softmax(self, x, temperature: float = 1.0) -> np.ndarray:
        x = _c(x)
        n = x.shape[-1]
        y = np.empty_like(x)
        _ck(_need().ti_b200_softmax(_fp(x), _fp(y), x.size // n, n, temperature))
        return y

    def sample(self, logits, temperature: float = 1.0, top_k: int = 50, top_p: float = 0.9, seed: int = 0, step: int = 0):
        """sample_next_token on [rows, vocab] logits -> (tokens [rows], logprobs [rows])."""
        lg = _c(logits)
        lg = lg.reshape(-1, lg.shape[-1])
        tok = np.zeros(lg.shape[0], dtype=np.int32)
        lp = np.zeros(lg.shape[0], dtype=np.float32)
        _ck(_need().ti_b200_sample_logits(_fp(lg.reshape(-1)), lg.shape[0], lg.shape[1], temperature, top_k, top_p, seed, step,
                                          tok.ctypes.data_as(_i32), _fp(lp)))
        return tok, lp

    def beam_expand(self, logits, beam_size: int, temperature: float = 1.0, top_k: int = 50, top_p: float = 0.9):
        """beam_search_decode's expansion (:1964-2005) on [rows, vocab] logits -> per row a list of (probability, token), best first."""
        lg = _c(logits)
        lg = lg.reshape(-1, lg.shape[-1])
        rows = lg.shape[0]
        pr = np.zeros((rows, beam_size), dtype=np.float32)
        tk = np.zeros((rows, beam_size), dtype=np.int32)
        cn = np.zeros(rows, dtype=np.int32)
        _ck(_need().ti_b200_beam_expand(_fp(lg.reshape(-1)), rows, lg.shape[1], temperature, top_k, top_p, beam_size, _fp(pr.reshape(-1)),
                                        tk.ctypes.data_as(_i32), cn.ctypes.data_as(_i32)))
        return [[(float(pr[r, i]), int(tk[r, i])) for i in range(cn[r])] for r in range(rows)]

    def attention_decode(self, q, k, v, num_heads: int = 1) -> np.ndarray:
        q, k, v = _c(q), _c(k), _c(v)
        B, _, H = q.shape
        t = k.shape[1]
        out = np.empty((B, 1, H), dtype=np.float32)
        _ck(_need().ti_b200_attention_decode(_fp(q), _fp(k), _fp(v), _fp(out), B, t, H, num_heads))
        return out


ops = _Ops()


class Model:
    """Device-resident decoder (InferenceEngine's weights + KV cache + incremental forward + greedy generate)."""

    def __init__(self, meta: dict, qtype: int, *, attn_mode: int = 1, rope_mode: int = 0, max_seq: int = 2048,
                 kv_page_tokens: int = 0, compat_literal: bool = False, tp: int = 1, per_op_engine: bool = False):
        cfg = ModelConfig()
        cfg.vocab, cfg.hidden, cfg.layers, cfg.heads, cfg.inter = (meta["vocab"], meta["hidden"], meta["layers"],
                                                                   meta["heads"], meta["inter"])
        cfg.rope_theta = meta.get("rope_theta", 10000.0)
        cfg.rms_eps = meta.get("rms_eps", 1e-5)
        cfg.qtype, cfg.attn_mode, cfg.rope_mode = qtype, attn_mode, rope_mode
        cfg.max_seq, cfg.kv_page_tokens, cfg.compat_literal = max_seq, kv_page_tokens, int(compat_literal)
        cfg.reserved[0] = 1 if per_op_engine else 0
        cfg.reserved[1] = tp if tp > 1 else 0   # tensor-parallel degree: needs tp_init() on every rank first
        self.meta = dict(meta)
        h = C.c_uint64()
        _ck(_need().ti_b200_model_new(C.byref(cfg), C.byref(h)))
        self.handle = h.value

    def set_tensor(self, name: str, data: np.ndarray) -> None:
        data = _c(data)
        rows, cols = (1, data.size) if data.ndim == 1 else data.shape
        _ck(lib().ti_b200_model_set_tensor(self.handle, name.encode(), _fp(data), rows, cols))

    def set_tensor_q(self, name: str, q: np.ndarray, qtype: int, scale: float, zero_point: float = 0.0) -> None:
        """An already quantized tensor (Quantizer::quantize_model / a .tinq file): int8 for INT8, int32 for INT4."""
        want = np.int8 if qtype == Q_INT8 else np.int32
        if q.dtype != want:
            raise TypeError(f"quantized tensor '{name}' must be {np.dtype(want).name}")
        q = np.ascontiguousarray(q)
        rows, cols = (1, q.size) if q.ndim == 1 else q.shape
        _ck(lib().ti_b200_model_set_tensor_q(self.handle, name.encode(), q.ctypes.data_as(C.c_void_p), rows, cols, qtype,
                                             float(scale), float(zero_point)))

    def set_tensor_synthetic(self, name: str, rows: int, cols: int, seed: int, amp: float) -> None:
        _ck(lib().ti_b200_model_set_tensor_synthetic(self.handle, name.encode(), rows, cols, seed, amp))

    def load(self, weights: dict) -> "Model":
        for name, arr in weights.items():
            self.set_tensor(name, arr)
        self.finalize()
        return self

    def load_synthetic(self, *, gate: bool = True, norms: bool = True) -> "Model":
        """Random-init weights generated ON THE DEVICE (uniform, amplitude 1/sqrt(fan_in); norm weights 1.0;
        seeds per SURVEY.md 8d) -- for benchmark-size models whose fp32 weights would not fit host RAM."""
        V, H, I, L = (self.meta[k] for k in ("vocab", "hidden", "inter", "layers"))
        self.set_tensor_synthetic("token_embeddings.weight", V, H, 777, 0.1)
        self.set_tensor_synthetic("lm_head.weight", H, V, 999, 1.0 / np.sqrt(H))
        ones = np.ones(H, dtype=np.float32)
        if norms:
            self.set_tensor("norm.weight", ones)
        for l in range(L):
            p, s = f"layers.{l}.", 1000 * l
            a_h, a_i = 1.0 / np.sqrt(H), 1.0 / np.sqrt(I)
            self.set_tensor_synthetic(p + "attention.q_proj.weight", H, H, s + 1, a_h)
            self.set_tensor_synthetic(p + "attention.k_proj.weight", H, H, s + 2, a_h)
            self.set_tensor_synthetic(p + "attention.v_proj.weight", H, H, s + 3, a_h)
            self.set_tensor_synthetic(p + "attention.o_proj.weight", H, H, s + 6, a_h)
            self.set_tensor_synthetic(p + "mlp.up_proj.weight", H, I, s + 4, a_h)
            if gate:
                self.set_tensor_synthetic(p + "mlp.gate_proj.weight", H, I, s + 7, a_h)
            self.set_tensor_synthetic(p + "mlp.down_proj.weight", I, H, s + 5, a_i)
            if norms:
                self.set_tensor(p + "attention_norm.weight", ones)
                self.set_tensor(p + "ffn_norm.weight", ones)
        self.finalize()
        return self

    def debug_timeline(self, token: int = 1) -> np.ndarray:
        """[phases, 32] SM-clock stamps of CTA 0 for one decode step on the persistent-kernel engine: 6 phase-level
        stamps (start, barrier passed, x staged, weights consumed, epilogue done, arrived) + 6 inside the prologue."""
        buf = np.zeros(32 * 4096, dtype=np.int64)
        n = C.c_size_t()
        _ck(lib().ti_b200_debug_timeline(self.handle, token, buf.ctypes.data_as(C.POINTER(C.c_int64)), buf.size, C.byref(n)))
        return buf[: 32 * n.value].reshape(n.value, 32).copy()

    def debug_timeline_all(self, token: int = 1) -> np.ndarray:
        """[ctas, phases, 32]: the same stamps from thread 0 of EVERY CTA, plus the barrier warp's [25..28]."""
        buf = np.zeros(32 * 4096 * 160, dtype=np.int64)
        n, c = C.c_size_t(), C.c_size_t()
        _ck(lib().ti_b200_debug_timeline_all(self.handle, token, buf.ctypes.data_as(C.POINTER(C.c_int64)), buf.size, C.byref(n), C.byref(c)))
        return buf[: 32 * n.value * c.value].reshape(c.value, n.value, 32).copy()

    def last_prefill_ms(self) -> float:
        ms = C.c_float()
        _ck(lib().ti_b200_model_last_prefill_ms(self.handle, C.byref(ms)))
        return ms.value

    def bench_gemv(self, slot: int, reps: int):
        """(avg ms per launch, algorithmic bytes per launch) of the model's own GEMVs of one kind, back to back."""
        ms, by = C.c_float(), C.c_double()
        _ck(lib().ti_b200_model_bench_gemv(self.handle, slot, reps, C.byref(ms), C.byref(by)))
        return ms.value / reps, by.value

    def finalize(self) -> None:
        _ck(lib().ti_b200_model_finalize(self.handle))

    def reset(self) -> None:
        _ck(lib().ti_b200_model_reset(self.handle))

    @property
    def kv_length(self) -> int:
        n = C.c_int32()
        _ck(lib().ti_b200_model_kv_length(self.handle, C.byref(n)))
        return n.value

    @property
    def persistent_engine(self) -> bool:
        """True: the persistent decode kernel runs this model; False: the per-op graph engine (null-weight fall-backs)."""
        k = C.c_int32()
        _ck(lib().ti_b200_model_engine(self.handle, C.byref(k)))
        return bool(k.value)

    def step_bytes(self, t: int):
        w, k = C.c_double(), C.c_double()
        _ck(lib().ti_b200_model_step_bytes(self.handle, t, C.byref(w), C.byref(k)))
        return w.value, k.value

    def prefill(self, tokens: Sequence[int], want_logits: bool = True):
        """forward_pass over a prompt (KV cache reset first); logits of the last position."""
        t = _c(tokens, np.int32)
        logits = np.empty(self.meta["vocab"], dtype=np.float32) if want_logits else None
        _ck(lib().ti_b200_prefill(self.handle, t.ctypes.data_as(_i32), t.size, _fp(logits) if want_logits else C.cast(None, _f)))
        return logits

    def decode_step(self, token: int, want_logits: bool = True):
        V = self.meta["vocab"]
        logits = np.empty(V, dtype=np.float32) if want_logits else None
        am = C.c_int32()
        _ck(lib().ti_b200_decode_step(self.handle, token, _fp(logits) if want_logits else C.cast(None, _f), C.byref(am)))
        return am.value, logits

    def generate_greedy(self, prompt: Sequence[int], n_new: int, *, stop_on_eos: bool = False, want_logits: bool = False):
        p = _c(prompt, np.int32)
        out = np.zeros(max(n_new, 1), dtype=np.int32)
        n_out, ms = C.c_int32(), C.c_float()
        logits = np.empty((max(n_new, 1), self.meta["vocab"]), dtype=np.float32) if want_logits else None
        _ck(lib().ti_b200_generate_greedy(self.handle, p.ctypes.data_as(_i32), p.size, n_new, int(stop_on_eos),
                                          out.ctypes.data_as(_i32), C.byref(n_out),
                                          _fp(logits) if want_logits else C.cast(None, _f), C.byref(ms)))
        n = n_out.value
        return out[:n].copy(), (logits[:n].copy() if want_logits else None), ms.value

    def generate_sampled(self, prompt: Sequence[int], n_new: int, *, temperature: float = 1.0, top_k: int = 50, top_p: float = 0.9,
                         seed: int = 0, stop_on_eos: bool = False):
        """generate() with the on-device sampler -> (tokens, logprobs of the picks, decode ms)."""
        p = _c(prompt, np.int32)
        out = np.zeros(max(n_new, 1), dtype=np.int32)
        lp = np.zeros(max(n_new, 1), dtype=np.float32)
        n_out, ms = C.c_int32(), C.c_float()
        _ck(lib().ti_b200_generate_sampled(self.handle, p.ctypes.data_as(_i32), p.size, n_new, temperature, top_k, top_p, seed,
                                           int(stop_on_eos), out.ctypes.data_as(_i32), C.byref(n_out), _fp(lp), C.byref(ms)))
        return out[: n_out.value].copy(), lp[: n_out.value].copy(), ms.value

    def compute_logprobs(self, tokens: Sequence[int]) -> np.ndarray:
        t = _c(tokens, np.int32)
        out = np.zeros(t.size, dtype=np.float32)
        _ck(lib().ti_b200_compute_logprobs(self.handle, t.ctypes.data_as(_i32), t.size, _fp(out)))
        return out

    def generate_batch_greedy(self, prompts, n_new: int, *, want_logits: bool = False):
        """generate_batch: prompts [B][n_prompt] (equal lengths) -> tokens [B][n_new], logits of the last step, decode ms."""
        p = np.ascontiguousarray(np.asarray(prompts, dtype=np.int32))
        assert p.ndim == 2
        B, n_prompt = p.shape
        out = np.zeros((B, n_new), dtype=np.int32)
        logits = np.empty((B, self.meta["vocab"]), dtype=np.float32) if want_logits else None
        ms = C.c_float()
        _ck(lib().ti_b200_generate_batch_greedy(self.handle, p.ctypes.data_as(_i32), B, n_prompt, n_new, out.ctypes.data_as(_i32),
                                                _fp(logits) if want_logits else C.cast(None, _f), C.byref(ms)))
        return out, logits, ms.value

    def generate_batch_ragged(self, prompts, n_new: int):
        """generate_batch for prompts of different lengths (list of token lists) -> tokens [B][n_new], decode ms."""
        lens = np.array([len(p) for p in prompts], dtype=np.int32)
        B, mx = len(prompts), int(lens.max())
        flat = np.zeros((B, mx), dtype=np.int32)
        for b, p in enumerate(prompts):
            flat[b, : len(p)] = p
        out = np.zeros((B, n_new), dtype=np.int32)
        ms = C.c_float()
        _ck(lib().ti_b200_generate_batch_ragged(self.handle, flat.ctypes.data_as(_i32), lens.ctypes.data_as(_i32), B, mx, n_new,
                                                out.ctypes.data_as(_i32), C.byref(ms)))
        return out, ms.value

    def beam_search(self, prompt, max_new: int, beam_size: int = 4, *, temperature: float = 1.0, top_k: int = 50, top_p: float = 0.9,
                    length_penalty: float = 1.0, eos_token: int = 2):
        """generate_beam_search (:830-871): list of dicts {tokens (new only), log_prob, score, finished}, best score first."""
        p = np.ascontiguousarray(prompt, dtype=np.int32)
        out = np.zeros((beam_size, max(max_new, 1)), dtype=np.int32)
        lens = np.zeros(beam_size, dtype=np.int32)
        lp = np.zeros(beam_size, dtype=np.float32)
        sc = np.zeros(beam_size, dtype=np.float32)
        fin = np.zeros(beam_size, dtype=np.int32)
        n = C.c_int32()
        _ck(lib().ti_b200_beam_search(self.handle, p.ctypes.data_as(_i32), p.size, max_new, beam_size, temperature, top_k, top_p, length_penalty,
                                      eos_token, out.ctypes.data_as(_i32), lens.ctypes.data_as(_i32), _fp(lp), _fp(sc), fin.ctypes.data_as(_i32),
                                      C.byref(n)))
        if max_new > 0:
            out = out.reshape(beam_size, max_new)
        return [dict(tokens=[int(t) for t in out[i, : lens[i]]], log_prob=float(lp[i]), score=float(sc[i]), finished=bool(fin[i]))
                for i in range(n.value)]

    def free(self) -> None:
        if self.handle:
            _ck(lib().ti_b200_model_free(self.handle))
            self.handle = 0
