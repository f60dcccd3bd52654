"""The workload the ncu captures under profiles/ are taken on: BASELINE configs[1] (1M cells, 2-bit packed contigs and CDR3s as
bench.py feeds them), `joins` joins in one process (default 2: the launch list is cut to the last one by tools/ncu_summary.py).
  python tools/profile_join.py [joins] [config] [cells]"""
import json
import sys

sys.path.insert(0, ".")
import enclone_ranger_b200 as E

joins = int(sys.argv[1]) if len(sys.argv) > 1 else 2
config = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cells = int(sys.argv[3]) if len(sys.argv) > 3 else None
rep = E.generate(config, n_cells=cells)
inp = rep.to_join_input().to_packed().to_packed_cdr3()
p = E.default_params(bool(rep.config.is_bcr))
rep.close()
ctx = E.JoinContext()
ctx.upload(p, inp)
for _ in range(joins):
    s = ctx.run()
keys = ("ms_device_total", "ms_pack", "ms_stage1", "ms_stage2", "ms_stage2a", "ms_forest", "ms_finalize", "pairs_scored", "pairs_evaluated",
        "pairs_distance_computed", "pairs_distance_full", "popc_executed_stage1", "stage1_candidates", "stage2_survivors", "pass_edges", "forest_edges",
        "joins", "rounds", "kernel_launches", "host_syncs", "boruvka_iterations", "h2d_bytes")
print(json.dumps({k: s[k] for k in keys}))
print("create_ms", ctx.create_ms())
ctx.close()
