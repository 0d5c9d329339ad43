"""Per-launch overhead of the persistent decode kernel: the same 255 decode steps as one launch and as launches of c steps
(stop_on_eos issues chunks of TURBOINFER_B200_EOS_CHUNK steps and looks at the tokens in between)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import turboinfer_b200 as tb
from helpers import SHAPES, prompt_tokens
tb.init(0)
meta = SHAPES["llama7b"]
m = tb.Model(meta, tb.Q_INT4, attn_mode=1, rope_mode=1, max_seq=1024).load_synthetic()
prompt = prompt_tokens(4, meta["vocab"])
m.generate_greedy(prompt, 256)
for eos in (False, True):
    ms = min(m.generate_greedy(prompt, 256, stop_on_eos=eos)[2] for _ in range(3))
    print("chunk", os.environ.get("TURBOINFER_B200_EOS_CHUNK", "32") if eos else "none", "decode ms", round(ms, 2), "us/token", round(ms / 255 * 1e3, 1))
m.free()
