"""Skew of the persistent kernel's phases across ALL CTAs (debug): for every phase, when each CTA finished it and how
long it then waited at the grid barrier (global nanosecond timer + per-SM clocks)."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import turboinfer_b200 as tb
from helpers import SHAPES, meta_with_layers

shape = sys.argv[1] if len(sys.argv) > 1 else "llama7b"
layers = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 16
tb.init(0)
meta = meta_with_layers(SHAPES[shape], layers)
m = tb.Model(meta, tb.Q_INT4, attn_mode=1, rope_mode=1, max_seq=2048)
m.load_synthetic()
for t in range(ctx):
    m.decode_step(t + 1, want_logits=False)
for rep in range(3):
    ts = m.debug_timeline_all(5)
C, P, _ = ts.shape
names = ["qkv", "attn", "o", "gateup", "down"]
f = 1 / 1965.0
print(f"{shape} L={layers} t={ctx}: {C} CTAs; per phase, us")
print("phase      | finish skew (global timer): p50-min  max-p50  max-min | wait at barrier (SM clk): min   p50   max | phase length: p50  max | loop: p50 max | epi p50 | stage p50")
g_done, g_pass = ts[:, :, 27].astype(np.float64) / 1e3, ts[:, :, 28].astype(np.float64) / 1e3
for ph in range(P):
    nm = names[ph % 5] if ph < P - 1 else "lm_head"
    d = g_done[:, ph]
    wait = (ts[:, ph, 26] - ts[:, ph, 25]) * f
    length = (ts[:, ph, 5] - ts[:, ph, 0]) * f
    loop = (ts[:, ph, 18] - ts[:, ph, 17]) * f if nm != "attn" else np.zeros(C)
    epi = (ts[:, ph, 4] - ts[:, ph, 3]) * f
    stage = (ts[:, ph, 2] - ts[:, ph, 1]) * f
    print(f"{ph:3d} {nm:7s}|  {np.median(d) - d.min():6.2f} {d.max() - np.median(d):6.2f} {d.max() - d.min():6.2f} | {wait.min():5.2f} {np.median(wait):5.2f} {wait.max():5.2f} | "
          f"{np.median(length):5.2f} {length.max():5.2f} | {np.median(loop):5.2f} {loop.max():5.2f} | {np.median(epi):5.2f} | {np.median(stage):5.2f}")
    if ph < 6:
        late = np.argsort(-d)[:6]
        print("      latest CTAs:", [(int(c), round(float(d[c] - np.median(d)), 2)) for c in late], " earliest:", [(int(c), round(float(d[c] - np.median(d)), 2)) for c in np.argsort(d)[:4]])
print("step us (CTA 0):", (ts[0, -1, 5] - ts[0, 0, 0]) * f, " global timer span:", g_pass[:, -1].max() - g_done[:, 0].min())
