#!/usr/bin/env python
"""bench.py -- decode tokens/s of the B200 hot path (BASELINE.json metric), one JSON line on stdout.

N = 1 (default): the north-star target of BASELINE.json -- Llama-2-7B shape, random-init, INT4 weights, batch-1 greedy
decode of 256 tokens after a 4-token prompt (configs[2], batch 1).  A "step" is one such generation.
  value        tokens/s over the K timed steps, decode loop timed on the device with CUDA events (weights, KV cache
               and the prompt already resident in HBM when the timed region starts), max over ranks
  e2e          the same metric through the public C-ABI call ti_b200_generate_greedy with HOST buffers: prompt H2D,
               prefill, decode, tokens D2H, wall clock around the call
  roofline     the dominant kernel = the persistent decode kernel (ONE launch per generation): algorithmic bytes of the
               launch (sum over the decoded tokens of packed weights + scales + KV read/write) / its
               duration from CUDA events on the launching stream; `traffic` = DRAM bytes of the same launch from the
               committed ncu --set full capture (profiles/).  per_kernel: the stand-alone GEMV launches (TensorEngine-
               level entry point) timed back to back over all layers' weights, for reference
  extras       the other single-GPU configurations of BASELINE.json, measured in the same run with fewer repetitions
               (TinyLlama INT4, 7B INT8, 7B prefill 2048, 7B batch 32, 70B single-GPU base of the TP curve); --no-extras skips
  cpu_baseline the reference's own CPU implementation (oracle/_ref, else the C restatement) on a bounded sample
--impl reference  times only that CPU arm (rank 0), same metric / config.

--gpus N > 1 (launched by torchrun): the N ranks form ONE tensor-parallel group decoding one sequence of the
Llama-2-70B shape (INT4, BASELINE.json configs[4] / the north-star's TP curve), scaling = strong.  Before the timed region
every such run (a) decodes a truncated (L = 2) model of the same width on rank 0's private single-GPU engine and on the TP
group and reports tp_tokens_equal / tp_logits_rel_err, (b) times the full model on rank 0's single-GPU engine so that the
line carries its own strong-scaling base (strong_scaling.single_gpu_tokens_per_s).  --dp keeps the old behaviour
(N independent replicas, no data-path collective, scaling = weak).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import SHAPES, make_model, meta_with_layers, prompt_tokens, rel_err_inf  # noqa: E402

WORKLOADS = {
    # name: (shape, qtype name, prompt tokens, new tokens, max_seq)
    "tinyllama-int4-decode512": ("tinyllama", "int4", 4, 512, 1024),
    "llama7b-int4-decode256": ("llama7b", "int4", 4, 256, 1024),
    "llama7b-int8-decode256": ("llama7b", "int8", 4, 256, 1024),
    # BASELINE.json configs[2]: prefill 2048 (tcgen05 GEMM path) + decode 256
    "llama7b-int4-prefill2048-decode256": ("llama7b", "int4", 2048, 256, 2560),
    "bench-small-int8-decode128": ("bench-small", "int8", 4, 128, 256),
    # tensor-parallel parity / scaling cases of BASELINE.json
    "llama13b-int8-decode128": ("llama13b", "int8", 4, 128, 512),
    "llama70b-int4-decode64": ("llama70b", "int4", 4, 64, 512),
    # configs[4] as written: 4k-context paged KV cache (the prompt is prefilled by the tensor-core path on one GPU,
    # token by token under tensor parallelism)
    "llama70b-int4-ctx4096-decode64": ("llama70b", "int4", 4032, 64, 4160),
}
DEFAULT_SINGLE = "llama7b-int4-decode256"
DEFAULT_TP = "llama70b-int4-decode64"
QT = {"int8": 0, "int4": 1}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                    if r[col].lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---- the CPU arm: the reference's own implementation on a bounded sample -------------------------------------------
class CpuArm:
    """One forward pass (one decoded token, lm_head included) of a full-WIDTH, truncated-DEPTH model on the reference's CPU
    implementation (oracle/_ref = the unmodified reference sources compiled here; else the C restatement).  The full
    depth cannot run on the CPU in minutes (7B: ~115 s per token, 26 GB of fp32 weights), so the per-token time of the
    full model is extrapolated: layers * per_layer + lm_head, with per_layer from the difference between an L = 2 and
    an L = 1 sample.  Every number derived that way is labelled `extrapolated`."""

    def __init__(self, shape: str, qname: str):
        import oracle
        self.use_ref = oracle.ref_available()
        self.orc = oracle.ref() if self.use_ref else oracle.port()
        self.full = SHAPES[shape]
        self.shape, self.qname = shape, qname
        self.qt = {"int8": oracle.QINT8, "int4": oracle.QINT4}[qname]
        self.models = {}

    def _model(self, layers: int):
        if layers not in self.models:
            meta = meta_with_layers(self.full, layers)
            w = make_model(meta)
            w = {k: (self.orc.fake_quant(v, self.qt) if (v.ndim == 2 and "embeddings" not in k) else v) for k, v in w.items()}
            self.models[layers] = (meta, w)
        return self.models[layers]

    def forward_passes_s(self, layers: int, n: int) -> list:
        """seconds of n consecutive forward passes (one token each, lm_head included), timed inside the oracle around
        each pass -- the engine set-up (deep copies of the weights, initialize_model) is outside the timed passes"""
        meta, w = self._model(layers)
        _, secs = self.orc.decode_greedy_timed(w, meta, [1], n, attn_mode=1, rope_mode=1)
        return [float(x) for x in secs]

    def summary(self, t1: float, t2: float) -> dict:
        per_layer = max(t2 - t1, 1e-9)
        per_head = max(t1 - per_layer, 0.0)
        per_token = per_layer * self.full["layers"] + per_head
        cores = 1 if self.use_ref else int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1))
        return {
            "value": 1.0 / per_token, "unit": "tokens/s", "cores": cores, "kind": "reference" if self.use_ref else "port",
            "extrapolated": True,
            "sample": (f"{self.shape} width, {self.qname} fake-quant weights, one forward pass (one token, lm_head included, weight set-up excluded) of an "
                       f"L=1 and an L=2 truncated model: {t1 * 1e3:.0f} / {t2 * 1e3:.0f} ms -> per layer {per_layer * 1e3:.1f} ms, lm_head {per_head * 1e3:.1f} ms, "
                       f"extrapolated to L={self.full['layers']}" + (" (the reference's decode GEMV is single-threaded, SURVEY R9)" if self.use_ref else "")),
            "host_cores_available": os.cpu_count(),
        }


def cpu_baseline(shape: str, qname: str) -> dict:
    arm = CpuArm(shape, qname)
    t1 = statistics.median(arm.forward_passes_s(1, 2))
    t2 = statistics.median(arm.forward_passes_s(2, 2))
    return arm.summary(t1, t2)


def run_reference_arm(args, config, shape, qname, n_new):
    """--impl reference: every step is ONE measured forward pass (one decoded token) of the full-width model truncated to
    L = 1, a bounded sample of the workload; `ms_per_step` is that measured time, `value` the full-depth tokens/s
    extrapolated from it (labelled).  The arm is time-boxed (TI_REF_BUDGET_S, default 240 s of timed passes): on the big
    shapes a single pass takes tens of seconds on one core, so fewer steps than asked may run -- `steps` / `warmup`
    report what actually ran."""
    t0 = time.perf_counter()
    budget = float(os.environ.get("TI_REF_BUDGET_S", "240"))
    arm = CpuArm(shape, qname)
    t2s = arm.forward_passes_s(2, 2)      # once: separates the per-layer time from the lm_head
    t2 = min(t2s)
    fit = max(2, int(budget / max(t2 * 0.6, 1e-3)))
    warm = max(1, min(args.warmup, fit // 6))
    steps = max(1, min(args.steps, fit - warm))
    ts = arm.forward_passes_s(1, warm + steps)[warm:]
    t1 = statistics.median(ts)
    base = arm.summary(t1, t2)
    v = base["value"]
    print(json.dumps({"impl": "reference", "metric": "decode_tokens_per_s", "value": v, "unit": "tokens/s", "n_gpus": args.gpus,
                      "steps": steps, "warmup": warm, "steps_requested": args.steps, "warmup_requested": args.warmup,
                      "ms_per_step": 1e3 * sum(ts) / len(ts), "higher_is_better": True,
                      "scaling": "strong" if config["parallelism"].startswith("tp") else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                      "extrapolated": True,
                      "step_definition": "one forward pass (one token) of the full-width model truncated to L=1, measured inside the oracle; value = 1 / (L * per_layer + lm_head)",
                      "cpu_baseline": base,
                      "e2e": {"value": v, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                      "wall_s": time.perf_counter() - t0}))
    return 0


# ---- GPU arm helpers ------------------------------------------------------------------------------------------------
def time_generation(model, prompt, n_new, steps, warmup):
    """K timed generations through the public call (host prompt in, host tokens out)."""
    for _ in range(warmup):
        model.generate_greedy(prompt, n_new)
    dev_ms, wall, prefill_ms, toks = [], [], [], None
    for _ in range(steps):
        t0 = time.perf_counter()
        toks, _, ms = model.generate_greedy(prompt, n_new)
        wall.append(time.perf_counter() - t0)
        dev_ms.append(ms)
        prefill_ms.append(model.last_prefill_ms())
    return dev_ms, wall, prefill_ms, toks


def launch_alg_bytes(model, n_prompt, n_new):
    """algorithmic bytes of one decode launch: tokens 2..n_new, token i runs at cache length n_prompt + i - 1"""
    total = 0.0
    for i in range(1, n_new):
        wb, kb = model.step_bytes(n_prompt + i - 1)
        total += wb + kb
    return total


def quick_single(tb, workload, steps=2, warmup=1, peak=6550.0):
    """An extra single-GPU workload measured beside the headline (fewer repetitions)."""
    shape, qname, n_prompt, n_new, max_seq = WORKLOADS[workload]
    meta = dict(SHAPES[shape])
    m = tb.Model(meta, QT[qname], attn_mode=1, rope_mode=1, max_seq=max_seq)
    m.load_synthetic()
    prompt = prompt_tokens(n_prompt, meta["vocab"])
    dev_ms, wall, prefill_ms, toks = time_generation(m, prompt, n_new, steps, warmup)
    by = launch_alg_bytes(m, n_prompt, n_new)
    us = 1e3 * sum(dev_ms) / steps
    out = {"value": steps * (n_new - 1) / (sum(dev_ms) / 1e3), "unit": "tokens/s", "steps": steps, "warmup": warmup,
           "ms_per_step": sum(dev_ms) / steps, "e2e_tokens_per_s": steps * n_new / sum(wall),
           "roofline_frac": by / (us * 1e-6) / 1e9 / peak, "achieved_GBps": by / (us * 1e-6) / 1e9,
           "prefill_ms": statistics.median(prefill_ms), "prompt_tokens": n_prompt, "new_tokens": n_new, "tokens_tail": [int(x) for x in toks[-4:]]}
    m.free()
    return out


def quick_batch(tb, model, meta, batch, n_prompt, n_new, reps=2):
    """BASELINE.json configs[2], batch 32: B sequences in lockstep through the tensor-core GEMM path (generate_batch)."""
    prompts = np.array([prompt_tokens(n_prompt, meta["vocab"], offset=b) for b in range(batch)], dtype=np.int32)
    model.generate_batch_greedy(prompts, n_new)   # warm-up: builds the step graphs
    dev, wall = [], []
    for _ in range(reps):
        t0 = time.perf_counter()
        toks, _, ms = model.generate_batch_greedy(prompts, n_new)
        wall.append(time.perf_counter() - t0)
        dev.append(ms)
    ms = statistics.median(dev)
    return {"value": batch * (n_new - 1) / (ms * 1e-3), "unit": "tokens/s (aggregate over the batch)", "batch": batch, "ms_per_decode_step": ms / (n_new - 1),
            "e2e_tokens_per_s": batch * n_new / statistics.median(wall), "prompt_tokens": n_prompt, "new_tokens": n_new,
            "tokens_tail_row0": [int(x) for x in toks[0][-4:]]}


def top2_margin(logits_row):
    s = np.sort(np.asarray(logits_row, dtype=np.float64))
    return float((s[-1] - s[-2]) / max(abs(s[-1]), 1e-30))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra single-GPU workloads measured beside the headline at N = 1")
    ap.add_argument("--tp", action="store_true", help="(default for N > 1) the N ranks form ONE tensor-parallel group decoding one sequence (strong scaling)")
    ap.add_argument("--dp", action="store_true", help="N > 1: N data-parallel replicas instead (no data-path collective, weak scaling)")
    ap.add_argument("--layers", type=int, default=0, help="debug: truncate depth (the number is then NOT a bench value)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # the reference arm is launched like ours (torchrun for N > 1) and must describe the same config: decide from --gpus
    n_ranks = max(world, args.gpus)
    use_tp = n_ranks > 1 and not args.dp
    workload = args.workload or (DEFAULT_TP if use_tp else DEFAULT_SINGLE)
    shape, qname, n_prompt, n_new, max_seq = WORKLOADS[workload]
    full = SHAPES[shape]
    w_elems = full["layers"] * (4 * full["hidden"] ** 2 + 3 * full["hidden"] * full["inter"]) + full["hidden"] * full["vocab"]
    w_mb = w_elems * (0.5 if qname == "int4" else 1.0) / 1e6
    per_gpu_mb = w_mb / n_ranks if use_tp else w_mb
    config = {"workload": workload, "shape": dict(full), "weights": qname + " symmetric per-tensor (reference default)",
              "batch": 1, "prompt_tokens": n_prompt, "new_tokens": n_new, "kv_cache": "fp32 paged, 64 tokens/page",
              "attention": "multi-head (mode B), RoPE per head",
              "parallelism": (f"tp{n_ranks}" if use_tp else f"dp{n_ranks}") if n_ranks > 1 else "single",
              "l2": f"packed weights {per_gpu_mb:.0f} MB per GPU > 126 MB L2: every decode step streams them from HBM (inputs larger than L2, no flush needed)"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        return run_reference_arm(args, config, shape, qname, n_new)

    import turboinfer_b200 as tb

    dist = None
    side = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl")
        side = dist.new_group(backend="gloo")   # host-side channel: NCCL id, parity vectors, barriers while rank 0 works alone
    tb.init(local_rank)
    tp = world if (use_tp and world > 1) else 1
    peak, peak_src = measured_peak_gbs()
    meta = dict(full)
    if args.layers:
        meta = meta_with_layers(meta, args.layers)

    tp_parity = None
    strong = None
    extras = {}
    if tp > 1:
        box = [tb.tp_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0, group=side)
        tb.tp_init(world, rank, box[0])
        # (a) TP parity inside the bench: a truncated (L = 2) model of the same width, the same synthetic weights, on rank
        # 0's private single-GPU engine and on the TP group -- greedy tokens must be equal, logits within 1e-4
        pmeta = meta_with_layers(full, 2)
        pprompt = prompt_tokens(4, pmeta["vocab"])
        ref_t = ref_l = None
        if rank == 0:
            r = tb.Model(pmeta, QT[qname], attn_mode=1, rope_mode=1, max_seq=128)
            r.load_synthetic()
            ref_t, ref_l, _ = r.generate_greedy(pprompt, 16, want_logits=True)
            r.free()
        tm = tb.Model(pmeta, QT[qname], attn_mode=1, rope_mode=1, max_seq=128, tp=tp)
        tm.load_synthetic()
        tt, tl, _ = tm.generate_greedy(pprompt, 16, want_logits=True)
        tm.free()
        if rank == 0:
            tp_parity = {"model": f"{shape} width, L=2, {qname}, prompt 4 + 16 greedy tokens", "tp_tokens_equal": bool(np.array_equal(ref_t, tt)),
                         "tp_logits_rel_err": float(rel_err_inf(tl, ref_l)), "top2_margin_min": min(top2_margin(r_) for r_ in ref_l),
                         "tokens": [int(x) for x in tt]}
        dist.barrier(group=side)
        # (b) the strong-scaling base: the same full workload on rank 0's single-GPU engine (the others wait on the host)
        if rank == 0 and not args.layers:
            try:
                strong = quick_single(tb, workload, steps=2, warmup=1, peak=peak)
            except Exception as e:   # e.g. a shape that does not fit one GPU: the curve then has no single-GPU point
                strong = {"error": repr(e)}
        dist.barrier(group=side)

    model = tb.Model(meta, QT[qname], attn_mode=1, rope_mode=1, max_seq=max_seq, tp=tp)
    model.load_synthetic()
    prompt = prompt_tokens(n_prompt, meta["vocab"], offset=0 if tp > 1 else rank)

    def barrier():
        tb.lib().ti_b200_sync()
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    warm = max(args.warmup, 3)
    for _ in range(warm):
        model.generate_greedy(prompt, n_new)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = tb.launch_count()
    t_all0 = time.perf_counter()
    dev_ms, wall, prefill_ms, toks = time_generation(model, prompt, n_new, args.steps, 0)
    barrier()
    t_all = time.perf_counter() - t_all0
    launches = tb.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None

    # the decode loop covers tokens 2..n_new (the first token comes out of the last prefill step)
    dev_total_ms = sum(dev_ms)
    wall_total = sum(wall)
    if dist is not None:
        import torch
        t = torch.tensor([dev_total_ms, wall_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_total_ms, wall_total = float(t[0]), float(t[1])
    seqs = 1 if tp > 1 else world   # tensor parallel: the ranks share one sequence
    tokens_dev = seqs * args.steps * (n_new - 1)
    tokens_e2e = seqs * args.steps * n_new
    value = tokens_dev / (dev_total_ms / 1e3)
    e2e = tokens_e2e / wall_total

    out = None
    if rank == 0:
        # the stand-alone GEMV launches, timed back to back over all layers' weights (> L2)
        per_kernel = {}
        for slot, name in ((2, "gemv_gate_up"), (0, "gemv_qkv"), (1, "gemv_o"), (3, "gemv_down"), (4, "gemv_lm_head")):
            ms, by = model.bench_gemv(slot, 20 * max(1, meta["layers"]))
            per_kernel[name] = {"us": ms * 1e3, "alg_bytes": by, "GBps": by / (ms * 1e-3) / 1e9}
        launch_bytes = launch_alg_bytes(model, n_prompt, n_new)
        launch_us = 1e3 * dev_total_ms / args.steps
        launch_gbs = launch_bytes / (launch_us * 1e-6) / 1e9
        wb, kb = model.step_bytes(n_prompt + n_new // 2)
        step_gbs = (wb + kb) * value / seqs / 1e9   # per GPU (step_bytes counts this rank's shards)
        traffic = None
        for name in ("r02_ncu_decode_kernel_summary.json", "r01_ncu_decode_kernel_summary.json"):
            try:
                with open(os.path.join(ROOT, "profiles", name)) as f:
                    traffic = json.load(f).get(workload, {}).get("dram_bytes_per_launch")
                if traffic is not None:
                    break
            except Exception:
                traffic = None
        # a near-tie at the top of the logits is where an arg-max could flip against the fp32 reference: keep the smallest
        # relative top-2 margin of a short greedy run visible (tests/test_gpu_decode.py checks tokens against the oracle)
        margin = None
        if tp == 1:
            _, lg, _ = model.generate_greedy(prompt, 8, want_logits=True)
            margin = min(top2_margin(r_) for r_ in lg)
        bits = 4 if qname == "int4" else 8
        out = {
            "metric": "decode_tokens_per_s", "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": dev_total_ms / args.steps, "higher_is_better": True, "scaling": "strong" if tp > 1 else "weak",
            "vs_baseline": None, "dtype": "int32 accumulate (IMMA.16832 warp MMAs) of " + qname + " weights x 24-bit fixed-point activations, f32 elsewhere",
            "data": "synthetic", "config": config,
            "e2e": {"value": e2e, "unit": "tokens/s", "h2d_bytes_per_step": 4 * n_prompt, "d2h_bytes_per_step": 4 * n_new},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": (("gemv_kernel<%d> + NCCL all-reduce per layer (per-op engine, tensor parallel; per-GPU bytes)" if os.environ.get("TURBOINFER_B200_TP_ENGINE") == "nccl"
                                                     else "mega_decode_kernel<%d> with the all-reduce fused in (peer stores over NVLink + barrier across the GPUs; per-GPU bytes)") % bits) if tp > 1 else
                         "mega_decode_kernel<%d> (persistent: one launch decodes %d tokens)" % (bits, n_new - 1),
                         "achieved": launch_gbs, "peak": peak, "unit": "GB/s", "frac": launch_gbs / peak,
                         "traffic": traffic, "peak_source": peak_src + ", sustained (the launch lasts hundreds of ms)",
                         "frac_of_8TBps_spec": launch_gbs / 8000.0, "alg_bytes_per_launch": launch_bytes, "us_per_launch": launch_us},
            "per_kernel": per_kernel,
            "whole_step": {"alg_bytes_per_token": wb + kb, "weight_bytes": wb, "kv_bytes_mid_run": kb, "GBps": step_gbs,
                           "frac_of_measured_peak": step_gbs / peak, "frac_of_8TBps_spec": step_gbs / 8000.0,
                           "us_per_token": 1e6 / (value / seqs)},
            "prefill": {"prompt_tokens": n_prompt, "ms": statistics.median(prefill_ms), "tokens_per_s": n_prompt / (statistics.median(prefill_ms) * 1e-3),
                        "path": "tcgen05 INT8 GEMM, batched" if (n_prompt > 32 and tp == 1) else "decode engine, token by token"},
            "clocks": clocks, "tokens_tail": [int(x) for x in toks[-4:]], "top2_margin_min_first8": margin, "wall_s_timed_region": t_all,
            "device": tb.device_info(),
        }
        if tp > 1:
            out["tp_parity"] = tp_parity
            out["tp_tokens_equal"] = tp_parity["tp_tokens_equal"]
            out["tp_logits_rel_err"] = tp_parity["tp_logits_rel_err"]
            if strong and "value" in strong:
                out["strong_scaling"] = {"workload": workload, "single_gpu_tokens_per_s": strong["value"], "speedup": value / strong["value"],
                                         "efficiency": value / strong["value"] / world, "single_gpu_roofline_frac": strong["roofline_frac"],
                                         "tokens_equal_single_gpu": strong["tokens_tail"] == [int(x) for x in toks[-4:]],
                                         "note": "the driver's N = 1 run measures the 7B north-star workload; this is the 1-GPU base of THIS workload, measured in this run on rank 0"}
            else:
                out["strong_scaling"] = strong
        if tp == 1 and world == 1 and not args.no_extras and not args.layers and workload == DEFAULT_SINGLE:
            try:
                extras["llama7b-int4-batch32-decode256"] = quick_batch(tb, model, meta, 32, 4, 256)
            except Exception as e:
                extras["llama7b-int4-batch32-decode256"] = {"error": repr(e)}
    model.free()
    if rank == 0:
        if tp == 1 and world == 1 and not args.no_extras and not args.layers and workload == DEFAULT_SINGLE:
            for wl in ("llama7b-int4-prefill2048-decode256", "llama7b-int8-decode256", "tinyllama-int4-decode512", "llama70b-int4-decode64"):
                try:
                    extras[wl] = quick_single(tb, wl, peak=peak)
                except Exception as e:
                    extras[wl] = {"error": repr(e)}
            out["extras"] = extras
        if not args.no_cpu_baseline:
            try:
                out["cpu_baseline"] = cpu_baseline(shape, qname)
            except Exception as e:  # the oracle is a reported baseline, never the product path
                out["cpu_baseline"] = {"value": None, "unit": "tokens/s", "cores": 0, "kind": "unavailable", "sample": repr(e)}
        print(json.dumps(out))
    if dist is not None:
        dist.barrier(group=side)
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
