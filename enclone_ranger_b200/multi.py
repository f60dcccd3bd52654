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
    """rank `src` holds `parts` = list (one per rank) of (JoinInput, global_rows[, roles]); every rank gets its own as
    (JoinInput, rows, roles) — roles is None unless a sub-bucket was split (partition_work).
    Arrays travel as byte tensors over the process group (NCCL: device tensors)."""
    from .inputs import JoinInput, _FIELDS
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    names = list(_FIELDS.keys()) + ["__rows__", "__roles__"]
    mine = None
    for r in range(world):
        if rank == src:
            inp, rows = parts[r][0], parts[r][1]
            roles = parts[r][2] if len(parts[r]) > 2 else None
            arrs = {n: inp.a[n] for n in _FIELDS}
            arrs["__rows__"] = np.asarray(rows, dtype=np.int64)
            arrs["__roles__"] = np.zeros(0, dtype=np.uint8) if roles is None else np.asarray(roles, dtype=np.uint8)
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
                dt = np.int64 if n == "__rows__" else (np.uint8 if n == "__roles__" else _FIELDS[n])
                if sz:
                    t = torch.empty(int(sz), dtype=torch.uint8, device=device)
                    dist.recv(t, src=src, group=group)
                    mine[n] = t.cpu().numpy().view(dt).copy()
                else:
                    mine[n] = np.zeros(0, dtype=dt)
    rows = mine.pop("__rows__")
    roles = mine.pop("__roles__")
    return JoinInput(**mine), rows, (roles if roles.size else None)


# ------------------------------------------------------------------------------------------------------------
# Sub-bucket granular partitioning with row-range splits of the heaviest sub-buckets (ejb_set_row_roles).
# ------------------------------------------------------------------------------------------------------------
ROLE_COLUMN_ONLY, ROLE_SPLIT = 1, 2
W_KILL = 3.0             # cost of a CDR3-close pair that stage 1 drops from its per-row tags (measured: ~6 ps)
W_CROSS = 60.0           # cost of one fully scored stage-2 candidate in units of one CDR3 pair (fit on the 8 x 1M-cell run: 0.12 ns vs 1.9 ps)


def _sub_bucket_table(inp):
    """(rows2, sub_of, n_sub): the 2-chain rows (ascending), the sub-bucket index of each, the number of sub-buckets.
    Sub-bucket = run of equal lens (join.rs:68-88) x (heavy, light) CDR3 length."""
    bucket = bucket_ids(inp)
    two = np.nonzero(inp.a["row_nchains"] == 2)[0]
    c3 = inp.a["row_cdr3_len"].reshape(-1, 2)
    key = (bucket[two].astype(np.int64) << 32) | (c3[two, 0].astype(np.int64) << 16) | c3[two, 1].astype(np.int64)
    uniq, sub_of = np.unique(key, return_inverse=True)
    # a partial bucket must not fuse with another run of the same lens once rows in between are taken away
    lens = inp.a["row_len"].reshape(-1, 2)
    first = np.nonzero(np.diff(np.concatenate([[-1], bucket])))[0]
    lk = (lens[first, 0].astype(np.int64) << 32) | lens[first, 1].astype(np.int64) | (inp.a["row_nchains"][first].astype(np.int64) << 60)
    if len(np.unique(lk)) != len(lk):
        raise ValueError("equal lens occur in several runs (unsorted info); use in-library sharding (ejb_set_shard) instead")
    return two, sub_of.reshape(-1), len(uniq)


def _seg(inp, i):
    off = inp.a["seg_off"]
    return inp.a["seg_bytes"][off[i]:off[i + 1]]


def _refsets_compatible(inp, ra, rb, max_degradation):
    """Can a row on reference set ra survive the degradation test (join_one.rs:434-440) against rb?  The count of
    differences of one contig to two references differs by at most their Hamming distance, so references further apart than
    max_degradation (different genes) kill the pair — once per row, then the stage-1 tags drop it — while near-identical
    ones (two alleles of one gene) let every CDR3-close pair through to the full two-reference scoring."""
    d = 0
    for x, y in zip(ra, rb):
        if x == y:
            continue
        sx, sy = _seg(inp, x), _seg(inp, y)
        if len(sx) != len(sy):
            return False
        d += int(np.count_nonzero(sx != sy))
        if d > max_degradation:
            return False
    return True


def _donor_masks(inp, rows):
    """64-bit donor-set mask of the exact subclonotype of each given row (join_one.rs:301-315)."""
    off = inp.a["ex_cell_off"]
    don = inp.a["cell_donor"]
    ex = inp.a["row_exact"][rows]
    out = np.zeros(len(rows), dtype=np.uint64)
    for i, e in enumerate(ex):
        d = don[off[e]:off[e + 1]]
        d = d[d >= 0]
        if len(d):
            out[i] = np.bitwise_or.reduce(np.uint64(1) << d.astype(np.uint64))
    return out


def probe_live_fractions(inp, ctx, params, min_rows=4096, sample=256):
    """Estimate, for every large sub-bucket, the fraction of its pairs that reach the full stage-2 scoring and do NOT join
    (two big clones on one V/J gene pair with similar junctions, or on two alleles of one gene): such pairs cost ~W_CROSS
    times a CDR3 comparison each and decide how evenly ranks are loaded, but nothing short of the join itself tells them
    from pairs inside one clone.  So the join is run on a SAMPLE: `sample` evenly spaced rows of each sub-bucket with at
    least `min_rows` rows go through `ctx` (one small join for all of them), and a sampled pair counts as live when it is
    CDR3-close (join_one.rs:216-234), ends in different clonotypes, its reference sets are equal or compatible and its
    donor sets do not already reject it (join_one.rs:301-323; stage 1 drops those pairs for free).
    CDR3-close pairs that stage 1 rejects from its per-row tags (other donor set, incompatible reference sets) are counted
    with the small weight W_KILL / W_CROSS.
    Returns {sub-bucket index (as in partition_work): effective fraction}."""
    two, sub_of, n_sub = _sub_bucket_table(inp)
    cnt = np.bincount(sub_of, minlength=n_sub)
    order = np.argsort(sub_of, kind="stable")
    starts = np.concatenate([[0], np.cumsum(cnt)])
    big = np.nonzero(cnt >= min_rows)[0]
    if len(big) == 0:
        return {}
    picks = {}
    for s in big:
        idx = two[order[starts[s]:starts[s + 1]]]
        picks[s] = idx[np.unique(np.linspace(0, len(idx) - 1, sample).astype(np.int64))]
    rows = np.sort(np.concatenate(list(picks.values())))
    probe, _ = inp.subset_rows(rows)
    ctx.set_row_roles(None)
    labels = np.asarray(ctx.join(params, probe).labels)
    lab_of = np.full(inp.n_rows, -1, dtype=np.int64)
    lab_of[rows] = labels
    thr = 1.0 - params.join_cdr3_ident / 100.0
    c3off = inp.a["row_cdr3_off"].reshape(-1, 2)
    c3len = inp.a["row_cdr3_len"].reshape(-1, 2)
    c3 = inp.a["cdr3_bytes"]
    vs = inp.a["row_vs"].reshape(-1, 2)
    js = inp.a["row_js"].reshape(-1, 2)
    out = {}
    for s, r in picks.items():
        m = len(r)
        ch, cl = int(c3len[r[0], 0]), int(c3len[r[0], 1])
        T = ch + cl
        if m < 2 or T == 0:
            continue
        seq = np.empty((m, T), dtype=np.uint8)
        for i, k in enumerate(r):
            seq[i, :ch] = c3[c3off[k, 0]:c3off[k, 0] + ch]
            seq[i, ch:] = c3[c3off[k, 1]:c3off[k, 1] + cl]
        cd = (seq[:, None, :] != seq[None, :, :]).sum(axis=2)
        close = (cd.astype(np.float64) / T <= thr) if params.is_bcr else (cd == 0)
        lab = lab_of[r]
        live = close & (lab[:, None] != lab[None, :])
        killed = np.zeros_like(live)          # CDR3-close pairs stage 1 rejects from its per-row tags (cheap, but not free)
        if not params.mix_donors:
            dm = _donor_masks(inp, r)
            rej = (dm[:, None] != 0) & (dm[None, :] != 0) & (dm[:, None] != dm[None, :])
            killed |= live & rej
            live &= ~rej
        refs = np.concatenate([vs[r], js[r]], axis=1)
        u, inv = np.unique(refs, axis=0, return_inverse=True)
        inv = inv.reshape(-1)
        if len(u) > 1:
            comp = np.eye(len(u), dtype=bool)
            cntu = np.bincount(inv)
            for a in range(len(u)):
                for b in range(a + 1, len(u)):
                    if cntu[a] * cntu[b] * 50 >= m:      # rare pairings cannot move the estimate
                        comp[a, b] = comp[b, a] = _refsets_compatible(inp, u[a], u[b], int(params.max_degradation))
            cm = comp[inv[:, None], inv[None, :]]
            killed |= live & ~cm
            live &= cm
        out[int(s)] = (float(np.triu(live, 1).sum()) + (W_KILL / W_CROSS) * float(np.triu(killed, 1).sum())) / (m * (m - 1) / 2)
    return out


class WorkPart:
    """What one rank joins: global row ids (ascending) and, when a sub-bucket was split, the row roles."""

    def __init__(self, rows, roles):
        self.rows = rows
        self.roles = roles if roles is not None and roles.any() else None


def partition_work(inp, world, live_fraction=None, split_above=0.5, piece=0.34):
    """Longest-processing-time assignment of sub-buckets to `world` ranks; a sub-bucket whose cost exceeds `split_above` x the
    per-rank share is cut into row ranges of about `piece` x that share (each part = the pairs whose FIRST row lies in the
    range; it needs the rows of the range and every row after it).
    Cost of a row acting as first row = (rows to its right) x (1 + W_CROSS x live fraction of its sub-bucket) + 64;
    live_fraction = probe_live_fractions(...) (None: pairs only).
    Returns (parts, split_rows): one WorkPart per rank and a boolean mask over all rows marking the split sub-buckets
    (their partial forests are merged by merge_split_edges)."""
    two, sub_of, n_sub = _sub_bucket_table(inp)
    order = np.argsort(sub_of, kind="stable")           # rows grouped by sub-bucket, ascending inside
    starts = np.concatenate([[0], np.cumsum(np.bincount(sub_of, minlength=n_sub))])
    live_fraction = live_fraction or {}
    costs = [None] * n_sub
    total = 0.0
    for s in range(n_sub):
        n = int(starts[s + 1] - starts[s])
        if n < 2:
            continue
        scale = 1.0 + W_CROSS * live_fraction.get(s, 0.0)
        if n < 1024:
            costs[s] = (scale * n * (n - 1) / 2 + 64.0 * n, None)
        else:
            per_row = scale * (n - 1 - np.arange(n)).astype(np.float64) + 64.0
            costs[s] = (float(per_row.sum()), per_row)
        total += costs[s][0]
    share = total / max(world, 1)
    units = []    # (cost, sub, lo, hi, n_parts)
    for s in range(n_sub):
        if costs[s] is None:
            continue
        c, per_row = costs[s]
        n = int(starts[s + 1] - starts[s])
        if world > 1 and per_row is not None and c > split_above * share:
            parts = int(min(world, np.ceil(c / (piece * share))))
            cum = np.cumsum(per_row)
            cuts = [0] + [int(np.searchsorted(cum, c * i / parts)) + 1 for i in range(1, parts)] + [n]
            cuts = sorted(set(min(max(x, 0), n) for x in cuts))
            for lo, hi in zip(cuts[:-1], cuts[1:]):
                if hi > lo:
                    units.append((float(per_row[lo:hi].sum()), s, lo, hi, len(cuts) - 1))
        else:
            units.append((c, s, 0, n, 1))
    units.sort(key=lambda u: -u[0])
    load = np.zeros(world)
    holder = {}                                          # sub -> ranks already holding one of its parts
    take = [[] for _ in range(world)]
    for c, s, lo, hi, nparts in units:
        cand = np.argsort(load, kind="stable")
        r = next(int(x) for x in cand if int(x) not in holder.get(s, ()))
        holder.setdefault(s, set()).add(r)
        load[r] += c
        take[r].append((s, lo, hi, nparts))
    split_rows = np.zeros(inp.n_rows, dtype=bool)
    parts = []
    for r in range(world):
        rows, roles = [], []
        for s, lo, hi, nparts in take[r]:
            idx = two[order[starts[s]:starts[s + 1]]]
            if nparts == 1:
                rows.append(idx)
                roles.append(np.zeros(len(idx), dtype=np.uint8))
            else:
                split_rows[idx] = True
                rows.append(idx[lo:])
                ro = np.full(len(idx) - lo, ROLE_SPLIT, dtype=np.uint8)
                ro[hi - lo:] |= ROLE_COLUMN_ONLY
                roles.append(ro)
        rows = np.concatenate(rows) if rows else np.zeros(0, dtype=np.int64)
        roles = np.concatenate(roles) if roles else np.zeros(0, dtype=np.uint8)
        o = np.argsort(rows, kind="stable")
        parts.append(WorkPart(rows[o].astype(np.int64), roles[o]))
    return parts, split_rows


def merge_split_edges(n_rows, k1, k2, cd, min_sh, split_rows, ncells_row, easy=False):
    """Edges of all ranks sorted by (k1, k2) (global row ids) -> boolean keep mask.  Edges of split sub-buckets are the union
    of partial forests: keep their minimum spanning forest (ejb_forest_merge, Kruskal in (k1, k2) order), then apply the
    two-cell filter (join.rs:120-178: an orbit of exactly two rows holding two cells in total loses its edge when
    cd > min(shares) / 2, unless EASY) that the parts had to skip.  Edges of whole sub-buckets are kept as they are."""
    from .join import forest_merge
    k1 = np.asarray(k1)
    k2 = np.asarray(k2)
    keep = np.ones(len(k1), dtype=bool)
    sel = np.nonzero(split_rows[k1])[0]
    if len(sel) == 0:
        return keep
    f = forest_merge(n_rows, k1[sel], k2[sel])
    keep[sel[~f]] = False
    sel = sel[f]
    a, b = k1[sel], k2[sel]
    deg = np.bincount(np.concatenate([a, b]), minlength=n_rows)
    two_cell = (deg[a] == 1) & (deg[b] == 1) & (ncells_row[a] + ncells_row[b] == 2) & (np.asarray(cd)[sel] > np.asarray(min_sh)[sel] // 2)
    if not easy:
        keep[sel[two_cell]] = False
    return keep


def ncells_per_row(inp):
    """exact_clonotypes[info[k].clonotype_id].ncells() for every row."""
    off = inp.a["ex_cell_off"]
    ex = inp.a["row_exact"]
    return (off[ex + 1] - off[ex]).astype(np.int64)
