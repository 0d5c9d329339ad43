"""Summarise the ncu artefacts of scripts/gpu_ncu.sh into profiles/ (tracked): launch list shares, and the decode
launch's DRAM traffic / throughput / stall picture from the --set full capture."""
import collections, csv, json, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
workload = sys.argv[2] if len(sys.argv) > 2 else "tinyllama-int4-decode512"
rows = [r for r in csv.reader(l for l in open(os.path.join(ROOT, "gpurun_out", "launches.csv")) if not l.startswith("=="))]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except Exception:
        continue
    a = agg.setdefault(r[ki][:70], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_launch_list_summary.txt"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -k regex:mega_decode|gemv_kernel|attn_ over `python bench.py --steps 1 --warmup 3`\n")
    f.write("# (cold-cache, serialised: compare shares, not absolutes).  gpu__time_duration in ns.\n")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write(f"{a[0]:5d} launches {a[1]/1e3:14.1f} us {100*a[1]/tot:6.2f}%  {k}\n")
raw = subprocess.run(["ncu", "-i", os.path.join(ROOT, "gpurun_out", "prof_mega.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, units, vals = rr[0], rr[1], rr[2]
def get(name):
    return vals[h.index(name)], units[h.index(name)]
keep = {}
for name in h:
    if any(s in name for s in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput", "gpu__dram_throughput",
                               "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
                               "sm__warps_active.avg.pct", "smsp__issue_active.avg.pct", "sm__throughput.avg.pct", "lts__t_bytes.sum ",
                               "smsp__average_warp", "smsp__pcsamp_warps_issue_stalled", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
                               "sm__inst_executed_pipe", "smsp__inst_executed.sum ")):
        keep[name] = " ".join(get(name))
with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_decode_kernel_raw.json"), "w") as f:
    json.dump(keep, f, indent=1)
def num(name):
    v, u = get(name)
    x = float(v.replace(",", ""))
    return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}.get(u, 1.0)
summ = {}
p = os.path.join(ROOT, "profiles", "r01_ncu_decode_kernel_summary.json")
if os.path.exists(p):
    summ = json.load(open(p))
summ[workload] = {"kernel": get("Kernel Name")[0], "dram_bytes_per_launch": num("dram__bytes_read.sum") + num("dram__bytes_write.sum"),
                  "dram_bytes_read": num("dram__bytes_read.sum"), "dram_bytes_write": num("dram__bytes_write.sum"),
                  "gpu_time_duration": " ".join(get("gpu__time_duration.sum")), "source": f"profiles/{tag}_ncu_decode_kernel_raw.json (ncu --set full, decode launch of bench.py --steps 1 --warmup 3)"}
json.dump(summ, open(p, "w"), indent=1)
print(json.dumps(summ[workload], indent=1))
stalls = sorted(((float(v.split()[0].replace(",", "")), k) for k, v in keep.items() if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k), reverse=True)
tots = sum(s for s, _ in stalls) or 1
for s, k in stalls[:12]:
    print(f"{100*s/tots:6.2f}%  {k}")
