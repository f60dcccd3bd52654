"""Compare the CUDA join with the CPU oracle on a synthetic config (dev tool; tests/ do the same)."""
import argparse
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import enclone_ranger_b200 as E
from oracle import pyoracle


def compare(res, ora, verbose=True):
    ok = True
    def chk(name, a, b, exact=True, tol=0.0):
        nonlocal ok
        a = np.asarray(a); b = np.asarray(b)
        if a.shape != b.shape:
            ok = False
            print(f"  MISMATCH {name}: shape {a.shape} vs {b.shape}")
            return
        if exact:
            bad = np.nonzero(a != b)[0] if a.ndim == 1 else np.nonzero((a != b).any(axis=1))[0]
        else:
            den = np.maximum(np.abs(b), 1e-300)
            bad = np.nonzero((np.abs(a - b) / den > tol) & (a != b))[0]
        if len(bad):
            ok = False
            print(f"  MISMATCH {name}: {len(bad)} of {len(a)}; first idx {bad[:5]} got {a[bad[:5]]} want {b[bad[:5]]}")
        elif verbose:
            print(f"  ok {name} ({len(a)})")
    chk("edge_k1", res.edge_k1, ora["edge_k1"]); chk("edge_k2", res.edge_k2, ora["edge_k2"])
    if len(res.edge_k1) == len(ora["edge_k1"]):
        for f in ("cd", "nrefs", "shares", "indeps", "k", "d", "n", "err"):
            chk("edge_" + f, getattr(res, "edge_" + f), ora["edge_" + f])
        for f in ("p1", "mult", "score"):
            chk("edge_" + f + " (bit-exact)", getattr(res, "edge_" + f), ora["edge_" + f])
    chk("labels", res.labels, ora["labels"])
    chk("eq_x", res.eq_x, ora["eq_x"]); chk("eq_y", res.eq_y, ora["eq_y"]); chk("eq_z", res.eq_z, ora["eq_z"])
    return ok


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=0)
    ap.add_argument("--cells", type=int, default=None)
    ap.add_argument("--seed", type=int, default=None)
    ap.add_argument("--cand-cap", type=int, default=0)
    ap.add_argument("--pair-budget", type=int, default=0)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--no-oracle", action="store_true")
    a = ap.parse_args()
    t = time.time()
    rep = E.generate(a.config, n_cells=a.cells, seed=a.seed)
    print("generated", rep.summary, f"{time.time()-t:.1f}s")
    p = E.default_params(bool(rep.config.is_bcr))
    ctx = E.JoinContext(cand_cap=a.cand_cap, pair_budget=a.pair_budget)
    t = time.time(); res = ctx.join(p, rep); t_gpu = time.time() - t
    t = time.time(); res = ctx.join(p, rep); t_gpu2 = time.time() - t
    print(json.dumps(res.stats, indent=None))
    print(f"gpu join wall: first {t_gpu:.3f}s second {t_gpu2:.3f}s  edges {len(res.edge_k1)}")
    if not a.no_oracle:
        t = time.time(); ora = pyoracle.run(p, rep.c_ptr, threads=a.threads); t_cpu = time.time() - t
        print(f"oracle: {t_cpu:.2f}s join {ora['seconds_join']:.2f}s threads {ora['threads']} join_one_calls {ora['join_one_calls']} "
              f"pairs_scored {ora['pairs_scored']} edges {ora['joins']} two_cell_deleted {ora['two_cell_deleted']}")
        ok = compare(res, ora)
        assert res.stats["pairs_scored"] == ora["pairs_scored"], (res.stats["pairs_scored"], ora["pairs_scored"])
        assert res.stats["pairs_lens"] == ora["pairs_lens"]
        print("PARITY", "OK" if ok else "FAILED")
        sys.exit(0 if ok else 1)
