"""A/B: decode tokens/s of the 7B INT4 workload for the library named by TURBOINFER_B200_LIB (one line)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import turboinfer_b200 as tb
from helpers import SHAPES, prompt_tokens
shape = sys.argv[1] if len(sys.argv) > 1 else "llama7b"
n_new = int(sys.argv[2]) if len(sys.argv) > 2 else 256
tb.init(0)
meta = SHAPES[shape]
m = tb.Model(meta, tb.Q_INT4, attn_mode=1, rope_mode=1, max_seq=1024); m.load_synthetic()
p = prompt_tokens(4, meta["vocab"])
for _ in range(3): m.generate_greedy(p, n_new)
ms = []
for _ in range(3):
    toks, _, t = m.generate_greedy(p, n_new); ms.append(t)
ms = sorted(ms)[1]
print(os.environ.get("TAG", ""), shape, f"{(n_new - 1) / (ms * 1e-3):.1f} tok/s  {1e3 * ms / (n_new - 1):.1f} us/token", [int(x) for x in toks[-3:]], flush=True)
m.free()
