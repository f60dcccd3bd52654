"""Sweep of the unit-schedule knobs (EJB_FIRST_UNIT, EJB_GROWTH) on BASELINE configs[1] (dev tool)."""
import json
import os
import sys

sys.path.insert(0, ".")
import enclone_ranger_b200 as E

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cells = int(sys.argv[2]) if len(sys.argv) > 2 else None
rep = E.generate(cfg, n_cells=cells)
inp = rep.to_join_input().to_packed()
p = E.default_params(bool(rep.config.is_bcr))
rep.close()
ctx = E.JoinContext()
ctx.upload(p, inp)
for first, growth in ((1, 8), (2, 12), (1, 16), (2, 16), (4, 16), (1, 8)):
    os.environ["EJB_FIRST_UNIT"] = str(first)
    os.environ["EJB_GROWTH"] = str(growth)
    ctx.run()
    ss = [ctx.run() for _ in range(3)]
    s = min(ss, key=lambda x: x["ms_device_total"])
    keys = ("ms_device_total", "ms_pack", "ms_stage1", "ms_stage2", "ms_stage2a", "ms_forest", "ms_finalize", "pairs_evaluated", "stage1_candidates",
            "stage1_ref_kills", "stage1_donor_kills", "stage2_survivors", "pass_edges", "rounds", "kernel_launches")
    print(json.dumps(dict({"first": first, "growth": growth}, **{k: (round(s[k], 3) if isinstance(s[k], float) else int(s[k])) for k in keys})), flush=True)
ctx.close()
