import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import turboinfer_b200 as tb
tb.init(0)
rng = np.random.default_rng(0)
K, N, M = 4096, 22016, 2048
qw = tb.QWeight(rng.uniform(-0.02, 0.02, (K, N)).astype(np.float32), tb.Q_INT4)
print(qw.bench_gemm(M, 3))
