"""Scheduler sweep on one GPU: strict allowance divisor x growth factor, on one config.
  python tools/sweep_sched.py <config> [cells]"""
import json
import sys

sys.path.insert(0, ".")
import enclone_ranger_b200 as E

config = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cells = int(sys.argv[2]) if len(sys.argv) > 2 else None
rep = E.generate(config, n_cells=cells)
inp = rep.to_join_input().to_packed().to_packed_cdr3()
p = E.default_params(bool(rep.config.is_bcr))
rep.close()
ctx = E.JoinContext()
ctx.upload(p, inp)
ctx.run()
for growth in (12, 6):
    for sd in (64, 256, 1024, 4096):
        ctx.set_option("strict_div", sd)
        ctx.set_option("growth", growth)
        ctx.run()
        s = ctx.run()
        print(json.dumps({"config": config, "growth": growth, "strict_div": sd, "ms": round(s["ms_device_total"], 3), "rounds": s["rounds"],
                          "stage1": round(s["ms_stage1"], 2), "stage2": round(s["ms_stage2"], 2), "forest": round(s["ms_forest"], 2),
                          "cand": s["stage1_candidates"], "pass": s["pass_edges"], "joins": s["joins"]}), flush=True)
ctx.close()
