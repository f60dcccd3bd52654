"""Run one BASELINE.json config at FULL size on the GPU and check size-independent properties (the CPU oracle is too
slow there): sorted forest, every edge inside one sub-bucket, labels == components of edges + clonotype_id merge,
EquivRel consistent, idempotent under a re-run.  Prints one JSON line."""
import argparse
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import enclone_ranger_b200 as E

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, required=True)
ap.add_argument("--cells", type=int, default=None)
a = ap.parse_args()
t = time.time()
rep = E.generate(a.config, n_cells=a.cells)
inp = rep.to_join_input()
rep.close()
pk = inp.to_packed()
t_gen = time.time() - t
p = E.default_params(bool(rep.config.is_bcr))
ctx = E.JoinContext()
res = ctx.join(p, pk)
t = time.time()
res2 = ctx.join(p, pk)
t_join = time.time() - t
k1, k2 = res.edge_k1.astype(np.int64), res.edge_k2.astype(np.int64)
key = k1 << 32 | k2
ok = {}
ok["sorted_forest"] = bool(np.all(key[1:] > key[:-1]) and np.all(k1 < k2))
lens = inp.a["row_len"].reshape(-1, 2)
c3 = inp.a["row_cdr3_len"].reshape(-1, 2)
ok["edges_inside_sub_buckets"] = bool(np.array_equal(lens[k1], lens[k2]) and np.array_equal(c3[k1], c3[k2]))
eq, lab = E.equiv_replay(inp.n_rows, res.edge_k1, res.edge_k2, inp.a["row_exact"])
ok["labels_match_replay"] = bool(np.array_equal(lab, res.labels) and np.array_equal(eq.y, res.eq_y) and np.array_equal(eq.x, res.eq_x))
ok["forest_minus_two_cell"] = bool(res.stats["forest_edges"] - res.stats["two_cell_deleted"] == len(k1))
ok["idempotent"] = bool(np.array_equal(res.edge_k1, res2.edge_k1) and np.array_equal(res.edge_k2, res2.edge_k2) and np.array_equal(res.edge_score, res2.edge_score))
ok["score_threshold"] = bool(np.all((res.edge_score <= p.max_score) | (res.edge_d >= p.auto_share) | (res.edge_cd <= 1000)))
if not p.is_bcr:
    ok["tcr_cd_zero"] = bool(np.all(res.edge_cd == 0))
s = res2.stats
out = {"config": a.config, "cells": int(inp.n_cells), "rows": int(inp.n_rows), "generate_s": round(t_gen, 1), "join_wall_s": round(t_join, 4),
       "device_ms": round(s["ms_device_total"], 2), "pairs_scored": int(s["pairs_scored"]), "pairs_lens": int(s["pairs_lens"]),
       "pairs_per_s_device": s["pairs_scored"] / max(s["ms_device_total"], 1e-9) * 1e3, "edges": int(len(k1)), "orbits": int(eq.norbits()),
       "stage_ms": {k: round(s[k], 2) for k in ("ms_h2d", "ms_pack", "ms_stage1", "ms_stage2", "ms_forest", "ms_finalize", "ms_d2h", "ms_replay")},
       "counts": {k: int(s[k]) for k in ("n_buckets", "n_subbuckets", "pairs_evaluated", "stage1_candidates", "stage2_slow", "pass_edges", "rounds", "overflow_retries")},
       "properties": ok}
print(json.dumps(out))
sys.exit(0 if all(ok.values()) else 1)
