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
print(f"{shape} L={layers} t={ctx}: us @1.965GHz, CTA 0 thread 0")
print("phase        | hdr   wait | issue stats conv  bar  =stage | setup+lds+wait   loop flush  bar =consume | epi  | bar  slot  red =arrive | total  ready")
f = 1 / 1965.0
tot = {}
for i, r in enumerate(ts):
    nm = names[i % 5] if i < len(ts) - 1 else "lm_head"
    u = lambda a, b: (r[b] - r[a]) * f if r[a] > 0 and r[b] > 0 else 0.0
    if nm == "attn":
        line = f"{i:3d} {nm:7s} | {u(0,12):5.2f} {u(12,1):5.2f} | {'':29s} | {'':29s} | {u(1,4):4.2f} | {u(4,21):4.2f}  0.00 0.00 ={u(4,5):5.2f} | {u(0,5):6.2f}   attn: entry {u(1,11):.2f} issue {u(11,16):.2f} loads+dot {u(16,22):.2f} softmax+rest {u(22,7):.2f} | kv+softmax {u(1,7):.2f} warp-merge+store {u(7,8):.2f} atomic {u(8,9):.2f} head-merge {u(9,10):.2f}"
    else:
        line = (f"{i:3d} {nm:7s} | {u(0,12):5.2f} {u(12,1):5.2f} | {u(6,13):5.2f} {u(13,14):5.2f} {u(14,15):4.2f} {u(15,2):4.2f} ={u(1,2):5.2f} | "
                f"{u(2,23):4.2f}+{u(23,24):4.2f}+{u(24,17):4.2f} {u(17,18):5.2f} {u(18,19):5.2f} {u(19,3):4.2f} ={u(2,3):6.2f} | {u(3,4):4.2f} | {u(4,21):4.2f}  0.00 0.00 ={u(4,5):5.2f} | {u(0,5):6.2f}  "
                f"ready {r[16]}/{r[22]} rounds: {u(17,7):.2f} {u(7,8):.2f} {u(8,9):.2f} {u(9,10):.2f} {u(10,11):.2f}")
    print(line)
print("step total us:", (ts[-1, 5] - ts[0, 0]) * f)
