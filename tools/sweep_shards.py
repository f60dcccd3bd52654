"""Unit-schedule sweep on the whole join and on two emulated shards of an 8-rank plan (dev tool, one GPU).
  python tools/sweep_shards.py [config]"""
import json
import sys

sys.path.insert(0, ".")
import enclone_ranger_b200 as E

config = int(sys.argv[1]) if len(sys.argv) > 1 else 4
rep = E.generate(config)
inp = rep.to_join_input().to_packed().to_packed_cdr3()
p = E.default_params(bool(rep.config.is_bcr))
rep.close()
ctx = E.JoinContext()
ctx.upload(p, inp)
for rank, world in ((0, 1), (1, 8), (5, 8)):
    ctx.set_shard(rank, world)
    for fu, gr in ((2, 6), (2, 8), (4, 8), (2, 12), (4, 12), (8, 12), (4, 16)):
        ctx.set_option("first_unit", fu)
        ctx.set_option("growth", gr)
        ctx.run()
        s = ctx.run()
        print(json.dumps({"rank": rank, "world": world, "first_unit": fu, "growth": gr, "ms": round(s["ms_device_total"], 3), "rounds": s["rounds"],
                          "pack": round(s["ms_pack"], 2), "stage1": round(s["ms_stage1"], 2), "stage2": round(s["ms_stage2"], 2), "forest": round(s["ms_forest"], 2),
                          "cand": s["stage1_candidates"], "pass": s["pass_edges"]}), flush=True)
ctx.close()
