#!/usr/bin/env python
"""bench.py -- decode tokens/s of the B200 hot path (BASELINE.json metric), one JSON line on stdout.

Default (N = 1): BASELINE.json configs[1] -- TinyLlama-1.1B shape, random-init, INT4 weights, batch-1 greedy decode of
512 tokens.  A "step" is one such generation (4-token prompt prefilled, 512 tokens decoded).
  value        tokens/s over the K timed steps, decode loop timed on the device with CUDA events (weights, KV cache
               and the prompt already resident in HBM when the timed region starts), max over ranks
  e2e          the same metric through the public C-ABI call ti_b200_generate_greedy with HOST buffers: prompt H2D,
               prefill, decode, tokens D2H, wall clock around the call
  roofline     the dominant kernel = the persistent decode kernel (ONE launch per generation): algorithmic bytes of the
               launch (sum over the decoded tokens of packed weights + scales + KV read/write) / its
               duration from CUDA events on the launching stream; `traffic` = DRAM bytes of the same launch from the
               committed ncu --set full capture (profiles/).  per_kernel: the stand-alone GEMV launches (TensorEngine-
               level entry point) timed back to back over all layers' weights, for reference
  cpu_baseline the reference's own CPU implementation (oracle/_ref, else the C restatement) on a bounded sample
--impl reference  times only that CPU arm (rank 0), same metric / config.
--gpus N > 1: data-parallel replicas -- each rank decodes an independent sequence on its own GPU (SURVEY.md 8e "DP:
batched-sequence path"), no data-path collective, scaling = weak; launched by torchrun, NCCL only for the barrier
and the max-over-ranks reduction.
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

from helpers import SHAPES, make_model, meta_with_layers, prompt_tokens  # noqa: E402

WORKLOADS = {
    # name: (shape, qtype name, prompt tokens, new tokens, max_seq)
    "tinyllama-int4-decode512": ("tinyllama", "int4", 4, 512, 1024),
    "llama7b-int4-decode256": ("llama7b", "int4", 4, 256, 1024),
    "llama7b-int8-decode256": ("llama7b", "int8", 4, 256, 1024),
    # BASELINE.json configs[2]: prefill 2048 (tcgen05 GEMM path) + decode 256
    "llama7b-int4-prefill2048-decode256": ("llama7b", "int4", 2048, 256, 2560),
    "bench-small-int8-decode128": ("bench-small", "int8", 4, 128, 256),
    # tensor-parallel parity / scaling cases of BASELINE.json (run with --tp under torchrun)
    "llama13b-int8-decode128": ("llama13b", "int8", 4, 128, 512),
    "llama70b-int4-decode64": ("llama70b", "int4", 4, 64, 512),
}
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


def cpu_reference_arm(shape: str, qname: str, n_prompt: int, sample_layers: int = 2, sample_tokens: int = 3):
    """The reference's CPU implementation on a bounded sample: full-width model truncated to `sample_layers` layers,
    `sample_tokens` greedy tokens after the prompt, then extrapolated to the full depth from per-layer and lm_head
    times measured separately (the full shape cannot run on the CPU in minutes)."""
    import oracle
    use_ref = oracle.ref_available()
    orc = oracle.ref() if use_ref else oracle.port()
    full = SHAPES[shape]
    qt = {"int8": oracle.QINT8, "int4": oracle.QINT4}[qname]
    prompt = prompt_tokens(n_prompt, full["vocab"])

    def timed(layers):
        meta = meta_with_layers(full, layers)
        w = make_model(meta)
        w = {k: (orc.fake_quant(v, qt) if (v.ndim == 2 and "embeddings" not in k) else v) for k, v in w.items()}
        t0 = time.perf_counter()
        orc.decode_greedy(w, meta, prompt, sample_tokens, attn_mode=1, rope_mode=1, want_logits=False)
        dt = time.perf_counter() - t0
        # forward passes executed: n_prompt (only the last with lm_head) + sample_tokens - 1
        return dt, n_prompt + sample_tokens - 1, sample_tokens

    t_lo, fw_lo, heads_lo = timed(1)
    t_hi, fw_hi, heads_hi = timed(sample_layers)
    per_layer = max((t_hi - t_lo) / ((sample_layers - 1) * fw_hi), 1e-9)
    per_head = max((t_lo - per_layer * fw_lo) / heads_lo, 0.0)
    per_token = per_layer * full["layers"] + per_head
    cores = 1 if use_ref else int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1))
    return {
        "value": 1.0 / per_token, "unit": "tokens/s", "cores": cores, "kind": "reference" if use_ref else "port",
        "sample": (f"{shape} width, {qname} fake-quant weights, L=1 and L={sample_layers} truncated models, prompt {n_prompt} + "
                   f"{sample_tokens} greedy tokens each; per-layer {per_layer * 1e3:.1f} ms, lm_head {per_head * 1e3:.1f} ms, "
                   f"extrapolated to L={full['layers']}" + (" (reference decode GEMV is single-threaded, SURVEY R9)" if use_ref else "")),
        "host_cores_available": os.cpu_count(),
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="tinyllama-int4-decode512", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--tp", action="store_true", help="the N ranks form ONE tensor-parallel group decoding one sequence (strong scaling) "
                                                      "instead of N data-parallel replicas")
    ap.add_argument("--layers", type=int, default=0, help="debug: truncate depth (the number is then NOT a bench value)")
    args = ap.parse_args()

    shape, qname, n_prompt, n_new, max_seq = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": args.workload, "shape": dict(SHAPES[shape]), "weights": qname + " symmetric per-tensor (reference default)",
              "batch": 1, "prompt_tokens": n_prompt, "new_tokens": n_new, "kv_cache": "fp32 paged, 64 tokens/page",
              "attention": "multi-head (mode B), RoPE per head",
              "parallelism": (f"tp{world}" if args.tp else f"dp{world}") if world > 1 else "single",
              "l2": "weights 598 MB > 126 MB L2: every decode step streams them from HBM"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        t0 = time.perf_counter()
        vals = []
        base = None
        for _ in range(max(1, min(args.steps, 2))):
            base = cpu_reference_arm(shape, qname, n_prompt)
            vals.append(base["value"])
        v = statistics.median(vals)
        base["value"] = v
        print(json.dumps({"impl": "reference", "metric": "decode_tokens_per_s", "value": v, "unit": "tokens/s", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * n_new / v, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                          "cpu_baseline": base,
                          "e2e": {"value": v, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "wall_s": time.perf_counter() - t0}))
        return 0

    import turboinfer_b200 as tb

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl")
    tb.init(local_rank)
    tp = world if (args.tp and world > 1) else 1
    if tp > 1:
        # the NCCL id of the tensor-parallel group travels over a gloo side channel
        side = dist.new_group(backend="gloo")
        box = [tb.tp_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0, group=side)
        tb.tp_init(world, rank, box[0])
    meta = dict(SHAPES[shape])
    if args.layers:
        meta = meta_with_layers(meta, args.layers)
    model = tb.Model(meta, QT[qname], attn_mode=1, rope_mode=1, max_seq=max_seq, tp=tp)
    model.load_synthetic()
    prompt = prompt_tokens(n_prompt, meta["vocab"], offset=0 if tp > 1 else rank)

    def barrier():
        tb.lib().ti_b200_sync()
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        model.generate_greedy(prompt, n_new)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = tb.launch_count()
    dev_ms, wall, prefill_ms = [], [], []
    toks = None
    t_all0 = time.perf_counter()
    for _ in range(args.steps):
        t0 = time.perf_counter()
        toks, _, ms = model.generate_greedy(prompt, n_new)      # host prompt in, host tokens out
        wall.append(time.perf_counter() - t0)
        dev_ms.append(ms)
        prefill_ms.append(model.last_prefill_ms())
    barrier()
    t_all = time.perf_counter() - t_all0
    launches = tb.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None

    # decode loop covers n_new - 1 graph launches (the first token comes out of the last prefill step)
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
        # roofline of the dominant kernel, timed alone over all layers' weights (> L2) -> burst HBM peak
        per_kernel = {}
        for slot, name in ((2, "gemv_gate_up"), (0, "gemv_qkv"), (1, "gemv_o"), (3, "gemv_down"), (4, "gemv_lm_head")):
            ms, by = model.bench_gemv(slot, 20 * max(1, meta["layers"]))
            per_kernel[name] = {"us": ms * 1e3, "alg_bytes": by, "GBps": by / (ms * 1e-3) / 1e9}
        peak, peak_src = measured_peak_gbs()
        # algorithmic bytes of one decode launch: tokens 2..n_new, token i runs at cache length n_prompt + i - 1
        launch_bytes = 0.0
        for i in range(1, n_new):
            wb_i, kb_i = model.step_bytes(n_prompt + i - 1)
            launch_bytes += wb_i + kb_i
        launch_us = 1e3 * dev_total_ms / args.steps
        launch_gbs = launch_bytes / (launch_us * 1e-6) / 1e9
        wb, kb = model.step_bytes(n_prompt + n_new // 2)
        step_gbs = (wb + kb) * value / seqs / 1e9   # per GPU (step_bytes counts this rank's shards)
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "r01_ncu_decode_kernel_summary.json")) as f:
                traffic = json.load(f).get(args.workload, {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
        out = {
            "metric": "decode_tokens_per_s", "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_total_ms / args.steps, "higher_is_better": True, "scaling": "strong" if tp > 1 else "weak",
            "vs_baseline": None, "dtype": "int32 accumulate (IMMA.16832 warp MMAs) of " + qname + " weights x 24-bit fixed-point activations, f32 elsewhere",
            "data": "synthetic", "config": config,
            "e2e": {"value": e2e, "unit": "tokens/s", "h2d_bytes_per_step": 4 * n_prompt, "d2h_bytes_per_step": 4 * n_new},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": (("gemv_kernel<%d> + NCCL all-reduce per layer (per-op engine, tensor parallel; per-GPU bytes)" if os.environ.get("TURBOINFER_B200_TP_ENGINE") == "nccl"
                                                     else "mega_decode_kernel<%d> with the all-reduce fused in (peer stores over NVLink + barrier across the GPUs; per-GPU bytes)") % (4 if qname == "int4" else 8)) if tp > 1 else
                         "mega_decode_kernel<%d> (persistent: one launch decodes %d tokens)" % (4 if qname == "int4" else 8, n_new - 1),
                         "achieved": launch_gbs, "peak": peak, "unit": "GB/s", "frac": launch_gbs / peak,
                         "traffic": traffic, "peak_source": peak_src + ", sustained (the launch lasts hundreds of ms)",
                         "frac_of_8TBps_spec": launch_gbs / 8000.0, "alg_bytes_per_launch": launch_bytes, "us_per_launch": launch_us},
            "per_kernel": per_kernel,
            "whole_step": {"alg_bytes_per_token": wb + kb, "weight_bytes": wb, "kv_bytes_mid_run": kb, "GBps": step_gbs,
                           "frac_of_measured_peak": step_gbs / peak, "frac_of_8TBps_spec": step_gbs / 8000.0,
                           "us_per_token": 1e6 / (value / seqs)},
            "prefill": {"prompt_tokens": n_prompt, "ms": statistics.median(prefill_ms), "tokens_per_s": n_prompt / (statistics.median(prefill_ms) * 1e-3),
                        "path": "tcgen05 INT8 GEMM, batched" if n_prompt > 32 else "decode engine, token by token"},
            "clocks": clocks, "tokens_tail": [int(x) for x in toks[-4:]], "wall_s_timed_region": t_all,
            "device": tb.device_info(),
        }
    model.free()
    if rank == 0:
        if not args.no_cpu_baseline:
            try:
                out["cpu_baseline"] = cpu_reference_arm(shape, qname, n_prompt)
            except Exception as e:  # the oracle is a reported baseline, never the product path
                out["cpu_baseline"] = {"value": None, "unit": "tokens/s", "cores": 0, "kind": "unavailable", "sample": repr(e)}
        print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
