"""Sampled generation (on-device sampler, one decode launch + one sampler launch per token) and beam search against the greedy
persistent launch on the same model: tokens/s of each, CUDA-event timed decode."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import turboinfer_b200 as tb
from helpers import SHAPES, prompt_tokens
shape = sys.argv[1] if len(sys.argv) > 1 else "llama7b"
n_new = int(sys.argv[2]) if len(sys.argv) > 2 else 256
tb.init(0)
meta = SHAPES[shape]
m = tb.Model(meta, tb.Q_INT4, attn_mode=1, rope_mode=1, max_seq=1024).load_synthetic()
prompt = prompt_tokens(4, meta["vocab"])
out = {"shape": shape, "new_tokens": n_new}
m.generate_greedy(prompt, n_new)
_, _, ms = m.generate_greedy(prompt, n_new)
out["greedy_tok_s"] = (n_new - 1) / (ms * 1e-3)
for name, kw in (("sampled_top50_p0.9", dict(temperature=0.8, top_k=50, top_p=0.9)), ("sampled_no_topk", dict(temperature=1.0, top_k=0, top_p=0.95))):
    m.generate_sampled(prompt, n_new, seed=1, **kw)
    t0 = time.perf_counter()
    toks, lps, ms = m.generate_sampled(prompt, n_new, seed=1, **kw)
    out[name + "_tok_s"] = (n_new - 1) / (ms * 1e-3)
    out[name + "_e2e_tok_s"] = n_new / (time.perf_counter() - t0)
for beam in (4,):
    m.beam_search(prompt, 64, beam, eos_token=-1)
    t0 = time.perf_counter()
    r = m.beam_search(prompt, 64, beam, eos_token=-1)
    dt = time.perf_counter() - t0
    out[f"beam{beam}_steps_per_s"] = 64 / dt
    out[f"beam{beam}_ms_per_step"] = dt / 64 * 1e3
print(json.dumps(out))
m.free()
