"""Multi-rank host logic on CPU: two gloo ranks each hold the edges of the buckets they own; rank 0
gathers, merges by (k1, k2) and replays — the result must equal the single-process oracle."""
import os
import socket
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

import enclone_ranger_b200 as E
from enclone_ranger_b200.multi import EDGE_FIELDS, gather_edges, merge_edge_parts


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import pyoracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rep = E.generate(0, n_cells=3000, seed=21)
    p = E.default_params(True)
    full = pyoracle.run(p, rep.c_ptr, threads=1)
    # emulate bucket ownership: an edge belongs to the rank that owns its bucket (here: bucket id parity)
    inp = rep.to_join_input()
    lens = inp.a["row_len"].reshape(-1, 2)
    nch = inp.a["row_nchains"]
    head = np.ones(inp.n_rows, dtype=bool)
    head[1:] = (lens[1:] != lens[:-1]).any(axis=1) | (nch[1:] != nch[:-1])
    bucket = np.cumsum(head) - 1
    mine = (bucket[full["edge_k1"]] % world) == rank
    local = {f: full[f][mine] for f in EDGE_FIELDS}
    merged = gather_edges(local)
    if rank == 0:
        ok = all(np.array_equal(merged[f], full[f]) for f in EDGE_FIELDS)
        eq, lab = E.equiv_replay(inp.n_rows, merged["edge_k1"], merged["edge_k2"], inp.a["row_exact"])
        ok = ok and np.array_equal(eq.x, full["eq_x"]) and np.array_equal(eq.y, full["eq_y"]) and np.array_equal(eq.z, full["eq_z"])
        ok = ok and np.array_equal(lab, full["labels"])
        q.put(bool(ok))
    else:
        assert merged is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_merge_replay():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok


def test_merge_is_order_independent():
    rng = np.random.default_rng(1)
    n = 50
    k1 = np.sort(rng.integers(0, 1000, n)).astype(np.int32)
    k2 = (k1 + 1 + rng.integers(0, 10, n)).astype(np.int32)
    d = {f: np.arange(n) for f in EDGE_FIELDS}
    d["edge_k1"], d["edge_k2"] = k1, k2
    d["edge_shares"] = np.arange(2 * n).reshape(-1, 2)
    d["edge_indeps"] = np.arange(2 * n).reshape(-1, 2)
    key = k1.astype(np.int64) << 32 | k2
    order = np.argsort(key, kind="stable")
    want = {f: v[order] for f, v in d.items()}
    idx = rng.permutation(n)
    parts = [{f: v[idx[:20]] for f, v in d.items()}, {f: v[idx[20:]] for f, v in d.items()}]
    got = merge_edge_parts(parts)
    uniq = len(np.unique(key)) == n
    if uniq:
        for f in EDGE_FIELDS:
            assert np.array_equal(got[f], want[f])
