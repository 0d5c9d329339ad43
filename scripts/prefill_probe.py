import sys, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo")); sys.path.insert(0, os.path.join(os.environ.get("GRAFT_REPO_ROOT", "/root/repo"), "tests"))
import numpy as np
import turboinfer_b200 as tb
from helpers import SHAPES, meta_with_layers, prompt_tokens
tb.init(0)
meta = meta_with_layers(SHAPES["llama7b"], 2)
m = tb.Model(meta, tb.Q_INT4, attn_mode=1, rope_mode=1, max_seq=2304)
m.load_synthetic()
p = prompt_tokens(2048, meta["vocab"])
for _ in range(2):
    toks, _, ms = m.generate_greedy(p, 2)
print("prefill ms", m.last_prefill_ms())
