import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import turboinfer_b200 as tb
from helpers import SHAPES, make_model, prompt_tokens
tb.init(0)
meta = SHAPES["tiny-test"]
w = make_model(meta, norm_jitter=0.1)
m = tb.Model(meta, tb.Q_INT8, attn_mode=1, rope_mode=0, max_seq=128).load(w)
print(m.generate_greedy(prompt_tokens(5, meta["vocab"]), 8)[0])
