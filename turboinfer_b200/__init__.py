"""turboinfer_b200 -- B200 (sm_100a) implementation of TurboInfer's token-generation hot path.

The product is ``libturboinfer_b200.so`` (hand-written CUDA behind the C ABI of ``include/ti_b200.h``).
This package is the thin Python loader used by the tests and ``bench.py``; the C++ host classes that mirror
the reference API live in ``include/turboinfer`` + ``turboinfer_b200/host``.

There is no CPU fallback: importing works anywhere (so the symbol table can be checked), but every compute
entry point raises unless a CUDA device is present and the extension is built.
"""
from .capi import (Q_INT4, Q_INT8, Q_NONE, EPI_STORE, EPI_RESIDUAL, EPI_SWIGLU, EPI_RELU, B200Error, FusedQWeight, KVCache, Model, QWeight, lib, library_path, init, shutdown,  # noqa: F401
                   device_info, launch_count, ops, tp_init, tp_unique_id)

__all__ = ["Q_INT4", "Q_INT8", "Q_NONE", "B200Error", "Model", "QWeight", "lib", "library_path", "init", "shutdown",
           "device_info", "launch_count", "ops", "tp_init", "tp_unique_id"]
