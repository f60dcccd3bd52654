"""Multi-GPU host logic: one process per GPU, buckets sharded by the library (ejb_set_shard), per-rank
edge lists gathered to rank 0 and merged there.

The join itself needs no collective (buckets are independent, enclone/src/join.rs:70-88 and
join2.rs:47-54 only concatenates per-bucket join lists); the only exchange is this final gather.
Works over NCCL (device tensors) and gloo (CPU tensors, used by the CPU tests).
"""
import io

import numpy as np
import torch
import torch.distributed as dist

EDGE_FIELDS = ("edge_k1", "edge_k2", "edge_cd", "edge_nrefs", "edge_shares", "edge_indeps", "edge_k", "edge_d", "edge_n",
               "edge_p1", "edge_mult", "edge_score", "edge_err")


def merge_edge_parts(parts):
    """parts: list of dicts of per-rank edge arrays.  Returns one dict sorted by (k1, k2) — the order of
    raw_joins (bucket order, then lexicographic; join.rs:584-588)."""
    cat = {f: np.concatenate([np.asarray(p[f]) for p in parts]) for f in EDGE_FIELDS}
    key = cat["edge_k1"].astype(np.int64) << 32 | cat["edge_k2"].astype(np.int64)
    order = np.argsort(key, kind="stable")
    return {f: v[order] for f, v in cat.items()}


def _pack(d):
    buf = io.BytesIO()
    np.savez(buf, **d)
    return np.frombuffer(buf.getvalue(), dtype=np.uint8)


def _unpack(a):
    with np.load(io.BytesIO(a.tobytes())) as z:
        return {k: z[k] for k in z.files}


def gather_edges(local, device=None, group=None, dst=0):
    """Gather every rank's edge arrays to `dst`; returns the merged dict there, None elsewhere.
    `local` is a JoinResult or a dict with EDGE_FIELDS."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    d = {f: np.asarray(getattr(local, f) if not isinstance(local, dict) else local[f]) for f in EDGE_FIELDS}
    if world == 1:
        return merge_edge_parts([d])
    dev = torch.device(device) if device is not None else torch.device("cpu")
    payload = torch.from_numpy(_pack(d).copy()).to(dev)
    n = torch.tensor([payload.numel()], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes)
    padded = torch.zeros(mx, dtype=torch.uint8, device=dev)
    padded[: payload.numel()] = payload
    bufs = [torch.zeros(mx, dtype=torch.uint8, device=dev) for _ in range(world)] if rank == dst else None
    dist.gather(padded, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    parts = [_unpack(bufs[r][: sizes[r]].cpu().numpy()) for r in range(world)]
    return merge_edge_parts(parts)


# ------------------------------------------------------------------------------------------------------------
# Host-side bucket partitioning: each rank receives (and uploads) only the rows of the buckets it owns.
# ------------------------------------------------------------------------------------------------------------
def bucket_ids(inp):
    """Bucket id of every row: maximal runs of equal `lens` (enclone/src/join.rs:68-88)."""
    a = inp.a
    n = inp.n_rows
    lens = a["row_len"].reshape(-1, 2)
    nch = a["row_nchains"]
    head = np.ones(n, dtype=bool)
    if n > 1:
        head[1:] = (lens[1:] != lens[:-1]).any(axis=1) | (nch[1:] != nch[:-1])
    return np.cumsum(head) - 1


def partition_buckets(inp, world, foreign_weight=100.0):
    """Longest-processing-time assignment of whole buckets to `world` ranks.
    Cost of a bucket = sum over its sub-buckets (equal CDR3 lengths) of
        n(n-1)/2                      pairs the CDR3 pre-filter evaluates
      + foreign_weight * f * n        stage-2 candidates: f rows whose V/J reference set differs from the sub-bucket's
                                      majority (e.g. indel-edited references) are scored against every clone-mate through
                                      the two-reference path and never get skipped by the class test
      + 64 n                          per-row work.
    Returns a list of ascending row-index arrays (rows of 1-chain buckets, which never join, are dropped)."""
    bucket = bucket_ids(inp)
    nb = int(bucket[-1]) + 1 if inp.n_rows else 0
    two = np.nonzero(inp.a["row_nchains"] == 2)[0]
    c3 = inp.a["row_cdr3_len"].reshape(-1, 2)
    key = np.stack([bucket[two], c3[two, 0], c3[two, 1]], axis=1)
    uniq, sub_of, cnt = np.unique(key, axis=0, return_inverse=True, return_counts=True)
    sub_of = sub_of.reshape(-1)
    # rows off their sub-bucket's majority reference set
    vs = inp.a["row_vs"].reshape(-1, 2)[two].astype(np.int64)
    js = inp.a["row_js"].reshape(-1, 2)[two].astype(np.int64)
    refkey = ((vs[:, 0] * 1000003 + vs[:, 1]) * 1000003 + js[:, 0]) * 1000003 + js[:, 1]
    pair = np.stack([sub_of.astype(np.int64), refkey], axis=1)
    pu, pc = np.unique(pair, axis=0, return_counts=True)
    major = np.zeros(len(uniq), dtype=np.int64)
    np.maximum.at(major, pu[:, 0], pc)
    foreign = cnt - major
    cost = np.zeros(nb, dtype=np.float64)
    np.add.at(cost, uniq[:, 0], cnt.astype(np.float64) * (cnt - 1) / 2 + foreign_weight * foreign.astype(np.float64) * cnt + 64.0 * cnt)
    # buckets taken out of their neighbourhood must not fuse with another run of the same lens
    lens = inp.a["row_len"].reshape(-1, 2)
    first = np.nonzero(np.diff(np.concatenate([[-1], bucket])))[0]
    lk = (lens[first, 0].astype(np.int64) << 32) | lens[first, 1].astype(np.int64) | (inp.a["row_nchains"][first].astype(np.int64) << 60)
    if len(np.unique(lk)) != len(lk):
        raise ValueError("equal lens occur in several runs (unsorted info); use in-library sharding (ejb_set_shard) instead")
    order = np.argsort(-cost, kind="stable")
    load = np.zeros(world)
    owner = np.full(nb, -1, dtype=np.int64)
    for b in order:
        if cost[b] <= 0:
            continue
        r = int(np.argmin(load))
        owner[b] = r
        load[r] += cost[b]
    row_owner = owner[bucket]
    return [np.nonzero(row_owner == r)[0] for r in range(world)]


def scatter_parts(parts, device, src=0, group=None):
    """rank `src` holds `parts` = list (one per rank) of (JoinInput, global_rows); every rank gets its own.
    Arrays travel as byte tensors over the process group (NCCL: device tensors)."""
    from .inputs import JoinInput, _FIELDS
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    names = list(_FIELDS.keys()) + ["__rows__"]
    mine = None
    for r in range(world):
        if rank == src:
            inp, rows = parts[r]
            arrs = {n: inp.a[n] for n in _FIELDS}
            arrs["__rows__"] = np.asarray(rows, dtype=np.int64)
            sizes = torch.tensor([arrs[n].nbytes for n in names], dtype=torch.int64, device=device)
        else:
            sizes = torch.zeros(len(names), dtype=torch.int64, device=device)
        if r == src:
            if rank == src:
                mine = {n: arrs[n] for n in names}
            continue
        if rank == src:
            dist.send(sizes, dst=r, group=group)
            for n in names:
                if arrs[n].nbytes:
                    dist.send(torch.from_numpy(np.ascontiguousarray(arrs[n]).view(np.uint8).reshape(-1).copy()).to(device), dst=r, group=group)
        elif rank == r:
            dist.recv(sizes, src=src, group=group)
            mine = {}
            for n, sz in zip(names, sizes.tolist()):
                dt = np.int64 if n == "__rows__" else _FIELDS[n]
                if sz:
                    t = torch.empty(int(sz), dtype=torch.uint8, device=device)
                    dist.recv(t, src=src, group=group)
                    mine[n] = t.cpu().numpy().view(dt).copy()
                else:
                    mine[n] = np.zeros(0, dtype=dt)
    rows = mine.pop("__rows__")
    return JoinInput(**mine), rows
