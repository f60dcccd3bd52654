"""A/B of the stage-1 kernel variants (EJB_S1_MODE bit 0: uniform-label column vote, bit 1: reference-set tags) on
BASELINE configs[1] (dev tool): device time and counters per variant."""
import json
import os
import sys

sys.path.insert(0, ".")
import enclone_ranger_b200 as E

rep = E.generate(1)
inp = rep.to_join_input().to_packed()
rep.close()
p = E.default_params(True)
ctx = E.JoinContext()
ctx.upload(p, inp)
for mode in (0, 1, 2, 3, 0, 3):
    os.environ["EJB_S1_MODE"] = str(mode)
    ctx.run()
    ss = [ctx.run() for _ in range(3)]
    s = min(ss, key=lambda x: x["ms_device_total"])
    keys = ("ms_device_total", "ms_pack", "ms_stage1", "ms_stage2", "ms_stage2a", "ms_forest", "ms_finalize", "pairs_evaluated", "stage1_candidates",
            "stage1_ref_kills", "stage2_two_ref", "stage2_memo_dead", "stage2_degr_dead", "stage2_survivors", "pass_edges", "forest_edges", "rounds", "kernel_launches")
    print(json.dumps(dict({"s1_mode": mode}, **{k: (round(s[k], 3) if isinstance(s[k], float) else int(s[k])) for k in keys})), flush=True)
os.environ["EJB_TRACE"] = "1"
ctx.run()
ctx.close()
