"""Summarise the ncu artefacts of scripts/gpu_r02_final.sh (gpurun_out/, scratch) into profiles/ (tracked):
launch-list shares of the bench command, and the raw metrics of the three --set full captures (decode launch, tcgen05 GEMM,
tensor-core prefill attention).  Usage: python scripts/ncu_summary_r02.py"""
import collections, csv, json, os, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

rows = [r for r in csv.reader(l for l in open(os.path.join(G, "r02_launches.csv")) if not l.startswith("=="))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}[r[ui]]
    except Exception:
        continue
    a = agg.setdefault(r[ki][:80], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
with open(os.path.join(P, "r02_ncu_launch_list_summary.txt"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 over `python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline`\n")
    f.write("# (7B INT4 decode-256; cold-cache, serialised: compare shares, not absolutes)\n")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write(f"{a[0]:5d} launches {a[1]:14.1f} us {100 * a[1] / tot:6.2f}%  {k}\n")

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput", "gpu__dram_throughput", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct", "smsp__issue_active.avg.pct",
        "sm__throughput.avg.pct", "lts__t_bytes.sum ", "smsp__average_warp", "smsp__pcsamp_warps_issue_stalled", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe", "smsp__inst_executed.sum ", "sm__pipe_tensor", "sm__cycles_active.avg", "smsp__cycles_active.avg", "lts__t_sector_hit_rate")


def capture(rep, out, note):
    raw = subprocess.run(["ncu", "-i", os.path.join(G, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h, units, vals = rr[0], rr[1], rr[2]
    keep = {"_capture": note, "Kernel Name": vals[h.index("Kernel Name")]}
    for name in h:
        if any(s in name for s in KEEP):
            keep[name] = (vals[h.index(name)] + " " + units[h.index(name)]).strip()
    json.dump(keep, open(os.path.join(P, out), "w"), indent=1)

    def num(name):
        x = float(vals[h.index(name)].replace(",", ""))
        return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}.get(units[h.index(name)], 1.0)
    return keep, num


keep, num = capture("r02_prof_mega.ncu-rep", "r02_ncu_decode_kernel_raw.json",
                    "ncu --set full --clock-control none -k regex:mega_decode_kernel -s 7 -c 1 over `python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline`: "
                    "the 255-token decode launch of the timed generation (7B INT4)")
summ = {"llama7b-int4-decode256": {"kernel": keep["Kernel Name"], "dram_bytes_per_launch": num("dram__bytes_read.sum") + num("dram__bytes_write.sum"),
                                   "dram_bytes_read": num("dram__bytes_read.sum"), "dram_bytes_write": num("dram__bytes_write.sum"),
                                   "gpu_time_duration": keep["gpu__time_duration.sum"], "source": "profiles/r02_ncu_decode_kernel_raw.json"}}
json.dump(summ, open(os.path.join(P, "r02_ncu_decode_kernel_summary.json"), "w"), indent=1)
print(json.dumps(summ, indent=1))
for rep, out, note in (("r02_prof_gemm.ncu-rep", "r02_ncu_gemm_tc_raw.json", "ncu --set full -k regex:gemm_i8_tc_kernel -s 2 -c 1 over `python scripts/gemm_one.py` (M 2048, K 4096, N 22016, INT4 weights)"),
                       ("r02_prof_attn_tc.ncu-rep", "r02_ncu_attn_tc_raw.json", "ncu --set full -k regex:causal_attention_tc -s 3 -c 1 over the 7B prefill-2048 workload (one layer's attention: 32 heads x 2048 queries)")):
    k2, n2 = capture(rep, out, note)
    print(out, k2.get("gpu__time_duration.sum"), {k: v for k, v in k2.items() if "pipe_tensor" in k and "pct" in k})
stalls = sorted(((float(v.split()[0].replace(",", "")), k) for k, v in keep.items() if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k), reverse=True)
tots = sum(s for s, _ in stalls) or 1
for s, k in stalls[:10]:
    print(f"{100 * s / tots:6.2f}%  {k}")
