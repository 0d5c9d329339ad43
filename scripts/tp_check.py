"""Tensor-parallel decode vs the single-GPU engine on the same weights (run under torchrun, one rank per GPU).
Every rank builds the TP model (its shards) AND a private single-GPU model, generates greedily with both and compares."""
import os, sys, json, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch.distributed as dist
import turboinfer_b200 as tb
from helpers import SHAPES, make_model, meta_with_layers, prompt_tokens, rel_err_inf

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
dist.init_process_group("gloo")
tb.init(local)
box = [tb.tp_unique_id() if rank == 0 else None]
dist.broadcast_object_list(box, src=0)
tb.tp_init(world, rank, box[0])
ok = True
cases = [("tiny-test", 0, tb.Q_INT8, 12), ("tiny-test", 0, tb.Q_INT4, 12), ("bench-small", 0, tb.Q_INT4, 24), ("tinyllama", 2, tb.Q_INT4, 16)]
for shape, layers, qt, n_new in cases:
    meta = dict(SHAPES[shape])
    if layers:
        meta = meta_with_layers(meta, layers)
    if meta["heads"] % world or meta["inter"] % world:
        continue
    w = make_model(meta, norm_jitter=0.1)
    prompt = prompt_tokens(4, meta["vocab"])
    prompts = np.array([prompt_tokens(4, meta["vocab"], offset=b) for b in range(5)], dtype=np.int32)   # batched decode, B = 5
    ref = tb.Model(meta, qt, attn_mode=1, rope_mode=1, max_seq=128).load(w)
    rt, rl, _ = ref.generate_greedy(prompt, n_new, want_logits=True)
    rbt, rbl, _ = ref.generate_batch_greedy(prompts, 6, want_logits=True)
    rs, rsl, _ = ref.generate_sampled(prompt, 10, temperature=0.9, top_k=40, top_p=0.9, seed=5)   # on-device sampling
    rlp = ref.compute_logprobs(prompt + [int(x) for x in rt[:4]])
    ref.free()
    m = tb.Model(meta, qt, attn_mode=1, rope_mode=1, max_seq=128, tp=world).load(w)
    t0 = time.perf_counter()
    tt, tl, ms = m.generate_greedy(prompt, n_new, want_logits=True)
    dt = time.perf_counter() - t0
    tbt, tbl, _ = m.generate_batch_greedy(prompts, 6, want_logits=True)
    ts_, tsl, _ = m.generate_sampled(prompt, 10, temperature=0.9, top_k=40, top_p=0.9, seed=5)
    tlp = m.compute_logprobs(prompt + [int(x) for x in rt[:4]])
    m.free()
    same = bool(np.array_equal(rt, tt)) and bool(np.array_equal(rbt, tbt)) and bool(np.array_equal(rs, ts_))
    err = max(float(rel_err_inf(tl, rl)), float(rel_err_inf(tbl, rbl)), float(np.max(np.abs(tlp - rlp))), float(np.max(np.abs(tsl - rsl))))
    ok &= same and err <= 1e-4
    print(json.dumps({"rank": rank, "case": f"{shape}/L{meta['layers']}/q{qt}", "tokens_equal": same, "logits_rel_err": err,
                      "tp_decode_ms_per_token": ms / max(1, n_new - 1)}), flush=True)
flags = [None] * world
dist.all_gather_object(flags, ok)
if rank == 0:
    print("TP CHECK", "PASSED" if all(flags) else "FAILED", flush=True)
dist.destroy_process_group()
sys.exit(0 if all(flags) else 1)
