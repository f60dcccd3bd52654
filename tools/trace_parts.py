"""Replay the N-rank partition on ONE GPU with EJB_TRACE on for the slowest parts (dev tool)."""
import json
import os
import sys
import time

sys.path.insert(0, ".")
import enclone_ranger_b200 as E
from enclone_ranger_b200 import multi

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rep = E.generate(1, n_cells=1000000 * world)
full = rep.to_join_input()
rep.close()
p = E.default_params(True)
ctx = E.JoinContext()
live = multi.probe_live_fractions(full, ctx, p)
print("live", sorted(live.items(), key=lambda kv: -kv[1])[:8], file=sys.stderr)
parts, split = multi.partition_work(full, world, live_fraction=live)
res = []
for r, part in enumerate(parts):
    sub, _ = full.subset_rows(part.rows)
    pk = sub.to_packed()
    ctx.set_row_roles(part.roles)
    ctx.upload(p, pk)
    ctx.run()
    s = ctx.run()
    res.append((s["ms_device_total"], r))
    print(json.dumps({"rank": r, "rows": len(part.rows), "device_ms": round(s["ms_device_total"], 2), "cands": int(s["stage1_candidates"]),
                      "kills": int(s["stage1_ref_kills"]), "donor_kills": int(s["stage1_donor_kills"]), "two_ref": int(s["stage2_two_ref"]), "memo_dead": int(s["stage2_memo_dead"]), "degr_dead": int(s["stage2_degr_dead"]),
                      "surv": int(s["stage2_survivors"]), "pass": int(s["pass_edges"]), "ms_stage2a": round(s["ms_stage2a"], 2), "rounds": int(s["rounds"]),
                      "pairs_eval": int(s["pairs_evaluated"]), "pairs": int(s["pairs_scored"])}), flush=True)
res.sort(reverse=True)
os.environ["EJB_TRACE"] = "1"
for _, r in res[:2]:
    part = parts[r]
    sub, _ = full.subset_rows(part.rows)
    pk = sub.to_packed()
    ctx.set_row_roles(part.roles)
    ctx.upload(p, pk)
    print(f"---- trace of rank {r}", file=sys.stderr, flush=True)
    ctx.run()
ctx.close()
