"""Per-phase timeline of one decode step on the persistent-kernel engine (debug)."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import turboinfer_b200 as tb
from helpers import SHAPES, meta_with_layers

shape = sys.argv[1] if len(sys.argv) > 1 else "tinyllama"
layers = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 8
tb.init(0)
meta = meta_with_layers(SHAPES[shape], layers)
m = tb.Model(meta, tb.Q_INT4, attn_mode=1, rope_mode=1, max_seq=2048)
m.load_synthetic()
for t in range(ctx):
    m.decode_step(t + 1, want_logits=False)
for rep in range(2):
    ts = m.debug_timeline(5)
names = ["qkv", "attn", "o", "gateup", "down"]
print(f"{shape} L={layers} t={ctx}: phase  wait  stage_x  consume  epilogue  arrive  total (us @1.965GHz)")
f = 1 / 1965.0
for i, r in enumerate(ts):
    nm = names[i % 5] if i < len(ts) - 1 else "lm_head"
    d = np.diff(r[:6]) * f
    sub = ""
    if r[6] > 0:
        pts = [r[1], r[6], r[7], r[8], r[9], r[10]]
        cons = os.environ.get("TURBOINFER_B200_DBG_NOMATH") == "3"
        if cons:
            pts = [r[2], r[8], r[9], r[10], r[3]]
        sub = "   | prologue: " + " ".join(f"{(b - a) * f:5.2f}" for a, b in zip(pts[:-1], pts[1:])) + ("" if cons else f" | ready stages {r[11]}")
    print(f"{i:3d} {nm:7s} " + " ".join(f"{x:8.2f}" for x in d) + f"  {(r[5]-r[0])*f:8.2f}" + sub)
print("prologue columns: stats gathered | rms+scale | x,w loads issued..arrived | (gap) | digits stored | final barrier")
print("step total us:", (ts[-1, 5] - ts[0, 0]) * f)
