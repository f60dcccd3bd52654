"""Replay the library's N-rank shard plan on ONE GPU (dev tool): every rank's share of configs[4] is joined in turn through
ejb_set_shard, its stats printed; with trace=1 the per-phase / per-round lines of each share go to stderr.
  python tools/trace_shards.py [world] [config] [trace]"""
import json
import sys

sys.path.insert(0, ".")
import enclone_ranger_b200 as E

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
config = int(sys.argv[2]) if len(sys.argv) > 2 else 4
trace = int(sys.argv[3]) if len(sys.argv) > 3 else 1
rep = E.generate(config)
inp = rep.to_join_input().to_packed().to_packed_cdr3()
p = E.default_params(bool(rep.config.is_bcr))
rep.close()
ctx = E.JoinContext()
ctx.upload(p, inp)
keys = ("ms_device_total", "ms_pack", "ms_stage1", "ms_stage2", "ms_stage2a", "ms_forest", "pairs_scored", "pairs_evaluated", "stage1_candidates",
        "stage2_survivors", "stage2_two_ref", "stage2_memo_dead", "stage2_degr_dead", "pass_edges", "forest_edges", "rounds", "host_syncs", "boruvka_iterations")
for r in range(world):
    ctx.set_shard(r, world)
    ctx.set_option("trace", 0)
    ctx.run()
    if trace:
        print(f"---- trace of rank {r} of {world}", file=sys.stderr, flush=True)
        ctx.set_option("trace", 1)
    s = ctx.run()
    print(json.dumps(dict({"rank": r}, **{k: (round(s[k], 2) if isinstance(s[k], float) else s[k]) for k in keys})), flush=True)
ctx.close()
