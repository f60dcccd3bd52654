"""Stall-reason summary per kernel from an `ncu --page source --csv` export (possibly gzipped, several launches concatenated):
  python tools/ncu_stalls.py <source.csv[.gz]> [top_lines]
Per kernel (the launch with the most samples): sampled stall reasons as shares of all samples, and the SASS lines with the
most samples."""
import csv
import gzip
import re
import sys
from collections import Counter

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 6
f = gzip.open(path, "rt", errors="replace") if path.endswith(".gz") else open(path, errors="replace")
launches, cur, hdr = [], None, None
for r in csv.reader(f):
    if not r:
        continue
    if r[0] == "Kernel Name":
        name = re.sub(r"\(const.*", "", r[1]).replace("void ", "").replace("ejb::", "").replace("(int)", "").replace("(bool)", "")
        cur = {"name": name, "rows": [], "hdr": None}
        launches.append(cur)
        continue
    if r[0] == "Address" and cur is not None:
        cur["hdr"] = r
        continue
    if cur is not None and cur["hdr"] is not None and len(r) == len(cur["hdr"]):
        cur["rows"].append(r)
best = {}
for L in launches:
    if not L["rows"]:
        continue
    ix = {n: i for i, n in enumerate(L["hdr"])}
    L["ix"] = ix
    stall_cols = [(n, i) for n, i in ix.items() if n.startswith("stall_") and "Not Issued" not in n]
    tot = Counter()
    for r in L["rows"]:
        for n, i in stall_cols:
            tot[n] += int(r[i] or 0)
    L["samples"] = sum(int(r[ix["# Samples"]] or 0) for r in L["rows"])
    L["inst"] = sum(int(r[ix["Instructions Executed"]] or 0) for r in L["rows"])
    L["tot"] = tot
    if L["samples"] and (L["name"] not in best or L["samples"] > best[L["name"]]["samples"]):
        best[L["name"]] = L
for name, L in best.items():
    ix = L["ix"]
    s = sum(L["tot"].values())
    print(f"\n{name}: largest of its launches, {L['samples']} samples, {L['inst']:.3g} warp instructions")
    print("  stalls: " + "  ".join(f"{n[6:]} {100.0 * v / s:.1f}%" for n, v in L["tot"].most_common(8)))
    for r in sorted(L["rows"], key=lambda r: -int(r[ix["# Samples"]] or 0))[:top]:
        print(f"    {int(r[ix['# Samples']]):6d} samples  {r[ix['Source']].strip()[:90]}")
