"""Tensor-core GEMM micro-benchmark: kernel-only time (CUDA events) of ti_b200_gemm_q's kernel on Llama-2-7B prefill shapes."""
import sys, os, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import turboinfer_b200 as tb
tb.init(0)
rng = np.random.default_rng(0)
out = []
for qname, qt in (("int4", tb.Q_INT4), ("int8", tb.Q_INT8)):
    for (K, N) in ((4096, 4096), (4096, 12288), (4096, 22016), (11008, 4096)):
        w = rng.uniform(-0.02, 0.02, (K, N)).astype(np.float32)
        qw = tb.QWeight(w, qt)
        for M in (32, 256, 2048):
            ms, ops = qw.bench_gemm(M, 5 if len(sys.argv) > 1 else 20)
            r = {"weights": qname, "M": M, "K": K, "N": N, "us": round(ms * 1e3, 2), "int8_TOPS_issued": round(ops / (ms * 1e-3) / 1e12, 1),
                 "effective_TFLOPS_2MNK": round(2.0 * M * K * N / (ms * 1e-3) / 1e12, 1)}
            out.append(r)
            print(json.dumps(r), flush=True)
        qw.free()
