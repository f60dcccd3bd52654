"""The workload the ncu captures under profiles/ are taken on: BASELINE configs[1] (1M cells, 2-bit packed contigs as bench.py
feeds them), `joins` joins in one process (default 2: the launch list is cut to the last one by tools/ncu_summary.py)."""
import sys

sys.path.insert(0, ".")
import enclone_ranger_b200 as E

joins = int(sys.argv[1]) if len(sys.argv) > 1 else 2
rep = E.generate(1)
inp = rep.to_join_input().to_packed()
rep.close()
p = E.default_params(True)
ctx = E.JoinContext()
ctx.upload(p, inp)
for _ in range(joins):
    s = ctx.run()
print({k: s[k] for k in ("ms_device_total", "pairs_scored", "stage1_candidates", "rounds", "kernel_launches")})
ctx.close()
