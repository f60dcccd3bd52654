"""Host logic of the in-library multi-GPU sharding (ejb_shard_plan == the shard_assign every rank runs inside ejb_run): CPU only.
The pair domain of every sub-bucket must be partitioned exactly (each first row on exactly one rank), deterministically, with
the heaviest sub-buckets cut into row ranges on different ranks; and two gloo ranks must agree on the plan."""
import ctypes as C
import os

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from enclone_ranger_b200 import _abi


def plan(sizes, world, rank):
    L = _abi.load_join()
    L.ejb_shard_plan.argtypes = [C.POINTER(C.c_int32), C.c_int32, C.c_int, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.ejb_shard_plan.restype = C.c_int
    s = np.ascontiguousarray(sizes, dtype=np.int32)
    lo, hi = np.zeros(len(s), np.int32), np.zeros(len(s), np.int32)
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    assert L.ejb_shard_plan(ip(s), len(s), world, rank, ip(lo), ip(hi)) == 0
    return lo, hi


def pairs(n, lo, hi):   # pairs whose first row lies in [lo, hi)
    m = hi - lo
    return m * (n - 1) - (lo + hi - 1) * m // 2


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_plan_partitions_the_pair_domain(world):
    rng = np.random.default_rng(world)
    sizes = np.concatenate([rng.integers(0, 300, 4000), rng.integers(1000, 9000, 60), [35000, 31000, 52000, 1, 2]]).astype(np.int64)
    rng.shuffle(sizes)
    cover = [np.zeros(int(n), dtype=np.int32) for n in sizes]
    load = np.zeros(world)
    split = 0
    for r in range(world):
        lo, hi = plan(sizes, world, r)
        for s, n in enumerate(sizes):
            if hi[s] > lo[s]:
                cover[s][lo[s]:hi[s]] += 1
                load[r] += pairs(int(n), int(lo[s]), int(hi[s]))
                split += int(hi[s] - lo[s] < n)
    for s, n in enumerate(sizes):
        if n >= 2:
            assert (cover[s] == 1).all(), f"sub-bucket {s} ({n} rows): first rows not covered exactly once"
        else:
            assert cover[s].sum() == 0
    total = sum(int(n) * (int(n) - 1) // 2 for n in sizes)
    assert load.sum() == total
    if world > 1:
        # pair counts are what the cost model balances (the merging rank gets a full share: the merge only starts when the
        # slowest part is in, so an early rank 0 would shorten nothing)
        assert load.max() / load.mean() < 1.05, load
        assert split > 0, "a 52k-row sub-bucket must be cut into row ranges"


def test_plan_is_deterministic_and_rejects_bad_arguments():
    sizes = np.arange(1, 3000, 7)
    a = plan(sizes, 8, 3)
    b = plan(sizes, 8, 3)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    L = _abi.load_join()
    s = np.ascontiguousarray(sizes, dtype=np.int32)
    out = np.zeros(len(s), np.int32)
    ip = lambda x: x.ctypes.data_as(C.POINTER(C.c_int32))
    assert L.ejb_shard_plan(ip(s), len(s), 4, 4, ip(out), ip(out)) == _abi.EJB_ERR_INVALID
    assert L.ejb_shard_plan(ip(s), len(s), 0, 0, ip(out), ip(out)) == _abi.EJB_ERR_INVALID


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch
    sizes = np.random.default_rng(11).integers(0, 20000, 500)
    lo, hi = plan(sizes, world, rank)
    mine = torch.from_numpy(np.stack([lo, hi]).astype(np.int64))
    allp = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allp, mine)                      # what a rank would need to check the others: it does not — the plan is a pure function
    cover = np.zeros((len(sizes), 20000), dtype=np.int8)
    for t in allp:
        l, h = t.numpy()
        for s in range(len(sizes)):
            cover[s, l[s]:h[s]] += 1
    ok = all((cover[s, :n] == 1).all() for s, n in enumerate(sizes) if n >= 2)
    # and every rank computes every other rank's part identically
    for r in range(world):
        l2, h2 = plan(sizes, world, r)
        ok = ok and np.array_equal(allp[r][0].numpy(), l2) and np.array_equal(allp[r][1].numpy(), h2)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_two_gloo_ranks_agree_on_the_plan():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
