"""Replay the N-rank bucket partition of bench.py --gpus N on ONE GPU: every rank's share is joined in turn and its
device time reported, so that the balance of multi.partition_buckets' cost model can be checked without an N-GPU box.
Prints one JSON line per foreign_weight tried."""
import argparse
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import enclone_ranger_b200 as E
from enclone_ranger_b200 import multi

ap = argparse.ArgumentParser()
ap.add_argument("--world", type=int, default=8)
ap.add_argument("--cells-per-rank", type=int, default=1000000)
ap.add_argument("--weights", type=str, default="45")
ap.add_argument("--no-gpu", action="store_true", help="only time the host-side partitioning")
a = ap.parse_args()
t = time.time()
rep = E.generate(1, n_cells=a.cells_per_rank * a.world)
full = rep.to_join_input()
rep.close()
print(f"generated {full.n_rows} rows in {time.time() - t:.1f} s", file=sys.stderr)
p = E.default_params(True)
ctx = None if a.no_gpu else E.JoinContext()
for w in [float(x) for x in a.weights.split(",")]:
    t = time.time()
    multi.W_CROSS = w
    t0 = time.time()
    live = multi.probe_live_fractions(full, ctx, p) if ctx is not None else {}
    t_probe = time.time() - t0
    parts, split = multi.partition_work(full, a.world, live_fraction=live)
    t_part = time.time() - t
    per = []
    for part in parts:
        rows = part.rows
        t = time.time()
        sub, _ = full.subset_rows(rows)
        pk = sub.to_packed()
        t_sub = time.time() - t
        if ctx is None:
            per.append({"rows": int(len(rows)), "subset_s": round(t_sub, 2), "column_only": 0 if part.roles is None else int((part.roles & 1).sum())})
            continue
        ctx.set_row_roles(part.roles)
        ctx.upload(p, pk)
        ctx.run()
        s = ctx.run()
        per.append({"rows": int(len(rows)), "device_ms": round(s["ms_device_total"], 2), "pairs_scored": int(s["pairs_scored"]),
                    "stage1_candidates": int(s["stage1_candidates"]), "stage1_ref_kills": int(s["stage1_ref_kills"]),
                    "ms_stage1": round(s["ms_stage1"], 2), "ms_stage2a": round(s["ms_stage2a"], 2), "rounds": int(s["rounds"]), "subset_s": round(t_sub, 2)})
    out = {"world": a.world, "w_cross": w, "split_rows": int(split.sum()), "probe_s": round(t_probe, 2), "live": {str(k): round(v, 4) for k, v in sorted(live.items(), key=lambda kv: -kv[1])[:12]}, "partition_s": round(t_part, 2), "per_rank": per}
    if ctx is not None:
        ms = [x["device_ms"] for x in per]
        out["max_ms"] = max(ms)
        out["mean_ms"] = round(sum(ms) / len(ms), 2)
        out["balance"] = round(sum(ms) / len(ms) / max(ms), 3)
    print(json.dumps(out), flush=True)
if ctx is not None:
    ctx.close()
