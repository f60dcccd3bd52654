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
