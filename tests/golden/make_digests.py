"""Full-size oracle digests of the BASELINE.json configs (generated OFFLINE in the CPU container, committed as
tests/golden/digests.json, compared on the GPU box by tests/test_gpu_fullsize.py).

    python tests/golden/make_digests.py --config 1 [--cells N] [--threads T]

The CPU oracle (oracle/liboracle.so, the restatement of join_exacts) joins the WHOLE synthetic repertoire of the config
(generator seed = the preset's) and this script records sha256 digests of everything the join returns: raw_joins
(edge_k1, edge_k2), every PotentialJoin member of every edge, the EquivRel vectors x / y / z and the canonical labels,
plus the counters.  ORACLE_P1_MEMO=1 is set: p1 is a pure function of (k, d, n), the memo returns identical bits.
"""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
OUT = os.path.join(HERE, "digests.json")

ARRAYS = ("edge_k1", "edge_k2", "edge_cd", "edge_nrefs", "edge_shares", "edge_indeps", "edge_k", "edge_d", "edge_n",
          "edge_p1", "edge_mult", "edge_score", "edge_err", "eq_x", "eq_y", "eq_z", "labels")
COUNTERS = ("n_buckets", "pairs_lens", "pairs_scored", "two_cell_deleted", "joins", "errors")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def digest_of(res, n_rows):
    """res: dict (oracle) or JoinResult-like object with the ARRAYS members.  Shared by the generator and the GPU test."""
    g = (lambda f: np.asarray(res[f])) if isinstance(res, dict) else (lambda f: np.asarray(getattr(res, f)))
    d = {f: sha(g(f)) for f in ARRAYS}
    k1 = g("edge_k1").astype(np.int64)
    d["n_edges"] = int(len(k1))
    d["n_rows"] = int(n_rows)
    # where the edges are: count per block of 2^16 first rows (localises a mismatch without shipping the edge list)
    d["edges_per_64k_rows"] = np.bincount(k1 >> 16, minlength=(int(n_rows) >> 16) + 1).tolist()
    d["orbits"] = int(len(np.unique(g("labels"))))
    return d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True)
    ap.add_argument("--cells", type=int, default=None)
    ap.add_argument("--threads", type=int, default=0)
    a = ap.parse_args()
    os.environ["ORACLE_P1_MEMO"] = "1"
    import enclone_ranger_b200 as E
    from oracle import pyoracle
    t = time.time()
    rep = E.generate(a.config, n_cells=a.cells)
    p = E.default_params(bool(rep.config.is_bcr))
    n_rows = int(rep.c.n_rows)
    print(f"config {a.config}: {rep.c.n_cells} cells, {n_rows} rows generated in {time.time() - t:.1f} s", flush=True)
    t = time.time()
    res = pyoracle.run(p, rep.c_ptr, threads=a.threads)
    wall = time.time() - t
    d = digest_of(res, n_rows)
    d.update({c: int(res[c]) for c in COUNTERS})
    d["join_one_calls"] = int(res["join_one_calls"])
    d["cells"] = int(rep.c.n_cells)
    d["seed"] = int(rep.config.seed)
    d["oracle_seconds"] = round(float(res["seconds_join"]), 1)
    d["oracle_threads"] = int(res["threads"])
    d["p1_memo"] = True
    print(f"oracle: {wall:.1f} s wall, {d['n_edges']} edges, {d['orbits']} orbits, {d['join_one_calls']} join_one calls", flush=True)
    key = f"config{a.config}" + (f"_cells{a.cells}" if a.cells else "")
    allj = json.load(open(OUT)) if os.path.exists(OUT) else {}
    allj[key] = d
    with open(OUT + ".tmp", "w") as f:
        json.dump(allj, f, indent=1, sort_keys=True)
    os.replace(OUT + ".tmp", OUT)
    print(f"wrote {key} -> {OUT}", flush=True)


if __name__ == "__main__":
    main()
