"""Turn the two ncu captures of a join (launch list CSV, `--set full` report) into the committed summaries under profiles/:
  python tools/ncu_summary.py <launches.csv> <full.ncu-rep> <round-tag>
writes profiles/launches_<tag>.csv (copy), profiles/ncu_traffic_<tag>.json and prints the markdown tables for profiles/README.md."""
import csv
import json
import shutil
import subprocess
import sys
from collections import OrderedDict

launch_csv, rep, tag = sys.argv[1:4]


def short(name):
    n = name.split("(")[0].replace("void ", "").replace("ejb::", "")
    return n[:60]


rows = list(csv.reader(open(launch_csv, errors="replace")))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hi]
kn, mv, idc = h.index("Kernel Name"), h.index("Metric Value"), h.index("ID")
seq = [(int(r[idc]), short(r[kn]), float(r[mv].replace(",", "")) / 1000) for r in rows[hi + 1:] if len(r) > mv]
starts = [i for i, s in enumerate(seq) if s[1] == "k_heads"]
last = seq[starts[-1]:] if starts else seq          # the last join of the process
agg = OrderedDict()
for _, k, us in last:
    a = agg.setdefault(k, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += us
    a[2] = max(a[2], us)
total = sum(a[1] for a in agg.values())
print(f"Sum of kernel time in one join: {total / 1000:.2f} ms ({len(last)} launches).\n")
print("| kernel | launches | total µs | share | max µs |\n|---|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
    print(f"| `{k}` | {a[0]} | {a[1]:.1f} | {100 * a[1] / total:.1f}% | {a[2]:.1f} |")
shutil.copy(launch_csv, f"profiles/launches_{tag}.csv")

# a report exported on the GPU box (`ncu -i x.ncu-rep --page raw --csv > x_raw.csv`: gpurun only brings back 64 MiB) or the report itself
raw = open(rep, errors="replace").read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hh = rr[0]
ix = {n: i for i, n in enumerate(hh)}
want = {"t": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum", "inst": "smsp__inst_executed.sum",
        "regs": "launch__registers_per_thread", "warps": "sm__warps_active.avg.pct_of_peak_sustained_active", "issue": "sm__inst_issued.avg.pct_of_peak_sustained_active",
        "xu": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "alu": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "l1": "l1tex__t_sector_hit_rate.pct", "l2": "lts__t_sector_hit_rate.pct"}
units = rr[1]


def val(r, key):
    i = ix.get(want[key])
    if i is None or not r[i]:
        return 0.0
    v = float(r[i].replace(",", ""))
    u = units[i]
    if key == "t":
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1e-3)
    if key in ("rd", "wr"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    return v


out = OrderedDict()
big = {}
for r in rr[2:]:
    k = short(r[ix["Kernel Name"]])
    o = out.setdefault(k, {"launches": 0, "time_us": 0.0, "dram_bytes": 0.0, "warp_inst": 0.0})
    o["launches"] += 1
    o["time_us"] += val(r, "t")
    o["dram_bytes"] += val(r, "rd") + val(r, "wr")
    o["warp_inst"] += val(r, "inst")
    if k not in big or val(r, "t") > big[k][0]:
        big[k] = (val(r, "t"), {m: round(val(r, m), 1) for m in ("regs", "warps", "issue", "xu", "alu", "l1", "l2")}, val(r, "rd") + val(r, "wr"))
for o in out.values():
    o["dram_bytes_per_launch"] = o["dram_bytes"] / max(o["launches"], 1)
# keys the bench looks up
alias = {}
for k in out:
    if k.startswith("k_stage1<"):
        a = alias.setdefault("k_stage1", {"launches": 0, "time_us": 0.0, "dram_bytes": 0.0, "warp_inst": 0.0})
        for f in a:
            a[f] += out[k][f]
if alias:
    alias["k_stage1"]["dram_bytes_per_launch"] = alias["k_stage1"]["dram_bytes"] / max(alias["k_stage1"]["launches"], 1)
    out.update(alias)
json.dump(out, open(f"profiles/ncu_traffic_{tag}.json", "w"), indent=1)
print("\n| kernel | launches | time µs | DRAM read+write | DRAM bytes / launch | warp instructions |\n|---|---|---|---|---|---|")
for k, o in out.items():
    print(f"| `{k}` | {o['launches']} | {o['time_us']:.0f} | {o['dram_bytes'] / 1e6:.0f} MB | {o['dram_bytes_per_launch'] / 1e6:.1f} MB | {o['warp_inst']:.3g} |")
print("\nLargest launches (regs / warps active % / issue active % / xu(popc) pipe % / alu pipe % / L1 hit % / L2 hit %):\n")
for k, (t, m, d) in big.items():
    print(f"* `{k}` {t:.0f} µs, {d / 1e6:.0f} MB DRAM: {m['regs']:.0f} regs / {m['warps']} / {m['issue']} / {m['xu']} / {m['alu']} / {m['l1']} / {m['l2']}")
