"""Per-phase timeline of one decode step of the fused tensor-parallel engine (rank 0, CTA 0): under torchrun, every rank runs
the same step.  Phases per layer: qkv | attn | o (+ peer stores, barrier across the GPUs) | reduce | gateup | down (+ barrier
across the GPUs) | reduce, then lm_head (+ key exchange when the lm_head is sharded)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import turboinfer_b200 as tb
from helpers import SHAPES, meta_with_layers
import torch.distributed as dist

shape = sys.argv[1] if len(sys.argv) > 1 else "llama70b"
layers = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 16
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
tb.init(int(os.environ.get("LOCAL_RANK", "0")))
if world > 1:
    dist.init_process_group("gloo")
    box = [tb.tp_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    tb.tp_init(world, rank, box[0])
meta = meta_with_layers(SHAPES[shape], layers)
m = tb.Model(meta, tb.Q_INT4, attn_mode=1, rope_mode=1, max_seq=256, tp=world)
m.load_synthetic()
for t in range(ctx):
    m.decode_step(t + 1, want_logits=False)
for rep in range(2):
    ts = m.debug_timeline(5)
if rank == 0:
    f = 1 / 1965.0
    per_layer = 7 if world > 1 else 5
    names = ["qkv", "attn", "o+xgpu", "reduce", "gateup", "down+xgpu", "reduce"] if world > 1 else ["qkv", "attn", "o", "gateup", "down"]
    print(f"{shape} L={layers} t={ctx} TP={world}: us, rank 0 CTA 0 thread 0; phases {len(ts)}")
    print("phase          | before barrier | barrier wait | work after barrier | arrive | total")
    agg = {}
    for i, r in enumerate(ts):
        nm = names[i % per_layer] if i < layers * per_layer else f"head{i - layers * per_layer}"
        u = lambda a, b: (r[b] - r[a]) * f if r[a] > 0 and r[b] > 0 else 0.0
        row = (u(0, 12), u(12, 1), u(1, 4), u(4, 5), u(0, 5))
        print(f"{i:3d} {nm:10s} | {row[0]:6.2f} | {row[1]:6.2f} | {row[2]:6.2f} | {row[3]:5.2f} | {row[4]:6.2f}")
        if i >= per_layer and i < layers * per_layer:   # skip the first layer (cold) in the averages
            a = agg.setdefault(nm + f"#{i % per_layer}", [0, 0.0, 0.0, 0.0, 0.0, 0.0])
            a[0] += 1
            for j in range(5):
                a[1 + j] += row[j]
    print("averages over layers 1.. (us):")
    tot = 0.0
    for k, a in agg.items():
        print(f"  {k:12s} before {a[1]/a[0]:5.2f}  wait {a[2]/a[0]:5.2f}  work {a[3]/a[0]:5.2f}  arrive {a[4]/a[0]:5.2f}  total {a[5]/a[0]:6.2f}")
        tot += a[5] / a[0]
    print(f"  layer total {tot:.2f} us; step total {(ts[-1, 5] - ts[0, 0]) * f:.2f} us")
m.free()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
