"""Batched decode (BASELINE.json configs[2], batch 32): B sequences in lockstep through the tcgen05 GEMM path.
Prints one JSON line: aggregate decode tokens/s (CUDA-event time of the decode steps), e2e wall-clock rate, and the HBM
bytes a step has to move (the GEMM reads the weights as one byte per element -- INT4 is stored unpacked for the tensor
cores -- plus every sequence's fp32 KV cache)."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import turboinfer_b200 as tb
from helpers import SHAPES, meta_with_layers, prompt_tokens

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="llama7b")
ap.add_argument("--qtype", default="int4", choices=["int4", "int8"])
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--prompt", type=int, default=4)
ap.add_argument("--new", type=int, default=256)
ap.add_argument("--layers", type=int, default=0)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--no-warmup", action="store_true", help="skip the untimed warm-up generation (long prompts: the step graphs are built during the prompt phase, outside the timed decode steps anyway)")
ap.add_argument("--no-single", action="store_true", help="skip the single-sequence cross-check of row 0")
ap.add_argument("--tp", action="store_true", help="under torchrun: the ranks form one tensor-parallel group (NCCL all-reduce after o / down)")
args = ap.parse_args()
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
tp = world if (args.tp and world > 1) else 1
tb.init(int(os.environ.get("LOCAL_RANK", "0")))
if tp > 1:
    import torch.distributed as dist
    dist.init_process_group("gloo")
    box = [tb.tp_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    tb.tp_init(world, rank, box[0])
meta = SHAPES[args.shape] if not args.layers else meta_with_layers(SHAPES[args.shape], args.layers)
m = tb.Model(meta, tb.Q_INT4 if args.qtype == "int4" else tb.Q_INT8, attn_mode=1, rope_mode=1, max_seq=args.prompt + args.new + 64, tp=tp)
m.load_synthetic()
prompts = np.array([prompt_tokens(args.prompt, meta["vocab"], offset=b) for b in range(args.batch)], dtype=np.int32)
l0 = tb.launch_count()
if not args.no_warmup:
    m.generate_batch_greedy(prompts, args.new)      # warm-up: builds the K-major weight copies and the step graphs
dev, wall = [], []
for _ in range(args.reps):
    t0 = time.perf_counter()
    toks, _, ms = m.generate_batch_greedy(prompts, args.new)
    wall.append(time.perf_counter() - t0)
    dev.append(ms)
single = toks[0][:0] if args.no_single else m.generate_greedy(prompts[0], min(args.new, 16))[0]
H, L, I, V = meta["hidden"], meta["layers"], meta["inter"], meta["vocab"]
w_elems = L * (4 * H * H + 3 * H * I) + H * V
t_mid = args.prompt + args.new // 2
kv = 2 * L * args.batch * t_mid * H * 4
ms = float(np.median(dev))
steps = args.new - 1
out = {"metric": "decode_tokens_per_s", "value": args.batch * steps / (ms * 1e-3), "unit": "tokens/s", "n_gpus": world, "parallelism": f"tp{tp}" if tp > 1 else "single",
       "config": {"workload": f"{args.shape}-{args.qtype}-batch{args.batch}-decode{args.new}", "batch": args.batch, "prompt_tokens": args.prompt,
                  "new_tokens": args.new, "path": "tcgen05 INT8 GEMM (three digit planes) + flash-decoding attention per sequence, CUDA graph per step"},
       "ms_per_step": ms / steps, "e2e": {"value": args.batch * args.new / float(np.median(wall)), "unit": "tokens/s"},
       "step_bytes": {"weights_as_read_by_the_gemm": w_elems, "kv_mid_run": kv, "GBps": (w_elems + kv) / (ms / steps * 1e-3) / 1e9,
                      "per_gpu": {"weights_GB": w_elems / tp / 1e9, "kv_GB": kv / tp / 1e9, "weights_GBps": w_elems / tp / (ms / steps * 1e-3) / 1e9,
                                  "kv_GBps": kv / tp / (ms / steps * 1e-3) / 1e9}},
       "row0_equals_single_sequence_engine": bool(np.array_equal(toks[0][: len(single)], single)),
       "gpu_launches": tb.launch_count() - l0, "tokens_tail_row0": [int(x) for x in toks[0][-4:]]}
if rank == 0:
    print(json.dumps(out))
m.free()
if tp > 1:
    dist.barrier()
    dist.destroy_process_group()
