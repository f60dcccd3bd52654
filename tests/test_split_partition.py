"""Host logic of the multi-GPU partitioner with row-range splits (multi.partition_work / merge_split_edges); CPU only."""
import numpy as np

import enclone_ranger_b200 as E
from enclone_ranger_b200 import multi
from enclone_ranger_b200.join import forest_merge


def _sub_keys(inp):
    bucket = multi.bucket_ids(inp)
    c3 = inp.a["row_cdr3_len"].reshape(-1, 2)
    return (bucket.astype(np.int64) << 32) | (c3[:, 0].astype(np.int64) << 16) | c3[:, 1].astype(np.int64)


def test_partition_work_covers_every_pair_exactly_once():
    rep = E.generate(1, n_cells=60000, seed=3)
    inp = rep.to_join_input()
    world = 4
    parts, split = multi.partition_work(inp, world)
    assert split.any(), "a heavy-tailed repertoire over 4 ranks must split its largest sub-bucket"
    key = _sub_keys(inp)
    two = inp.a["row_nchains"] == 2
    owner_count = np.zeros(inp.n_rows, dtype=np.int64)      # how many parts hold the row as a FIRST row
    for p in parts:
        assert np.all(np.diff(p.rows) > 0)
        roles = p.roles if p.roles is not None else np.zeros(len(p.rows), dtype=np.uint8)
        active = (roles & multi.ROLE_COLUMN_ONLY) == 0
        owner_count[p.rows[active]] += 1
        # split flag <=> row of a split sub-bucket
        assert np.array_equal((roles & multi.ROLE_SPLIT) != 0, split[p.rows])
        # inside the part, per sub-bucket: column-only rows form a suffix, and every row of the sub-bucket after the
        # first row held is present (a first row needs all of its partners to the right)
        k = key[p.rows]
        for s in np.unique(k[(roles & multi.ROLE_SPLIT) != 0]):
            mine = p.rows[k == s]
            all_rows = np.nonzero((key == s) & two)[0]
            assert np.array_equal(mine, all_rows[all_rows >= mine[0]])
            r = roles[k == s] & multi.ROLE_COLUMN_ONLY
            assert np.all(np.diff(r.astype(int)) >= 0)
    # every 2-chain row of a sub-bucket with >= 2 rows is a first row in exactly one part
    _, inv, cnt = np.unique(key[two], return_inverse=True, return_counts=True)
    need = np.zeros(inp.n_rows, dtype=bool)
    need[np.nonzero(two)[0]] = cnt[inv.reshape(-1)] >= 2
    assert np.array_equal(owner_count[need], np.ones(int(need.sum()), dtype=np.int64))
    assert owner_count[~need].sum() == 0
    # parts of one sub-bucket sit on distinct ranks
    for s in np.unique(key[split]):
        holders = [r for r, p in enumerate(parts) if np.any(key[p.rows] == s)]
        assert len(holders) == len(set(holders)) >= 2


def test_partition_work_single_rank_is_the_whole_input():
    rep = E.generate(0, n_cells=3000, seed=5)
    inp = rep.to_join_input()
    parts, split = multi.partition_work(inp, 1)
    assert not split.any() and parts[0].roles is None
    key = _sub_keys(inp)
    two = np.nonzero(inp.a["row_nchains"] == 2)[0]
    _, inv, cnt = np.unique(key[two], return_inverse=True, return_counts=True)
    assert np.array_equal(parts[0].rows, two[cnt[inv.reshape(-1)] >= 2])


def _kruskal(n, edges):
    parent = list(range(n))

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x
    out = []
    for a, b in sorted(edges):
        ra, rb = find(a), find(b)
        if ra != rb:
            parent[max(ra, rb)] = min(ra, rb)
            out.append((a, b))
    return out


def test_forest_of_partial_forests_is_the_forest():
    """MSF(G) == MSF(union of the MSFs of an edge partition of G), with the partition the row-range split induces (edges
    grouped by their first row) and the lexicographic (k1, k2) order join_core.rs:23-50 scans in."""
    rng = np.random.default_rng(7)
    for trial in range(20):
        n = int(rng.integers(5, 60))
        m = int(rng.integers(0, 4 * n))
        e = set()
        for _ in range(m):
            a, b = sorted(rng.integers(0, n, size=2).tolist())
            if a != b:
                e.add((a, b))
        want = _kruskal(n, e)
        cuts = sorted(set([0, n] + rng.integers(0, n, size=int(rng.integers(1, 4))).tolist()))
        union = []
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            union += _kruskal(n, [x for x in e if lo <= x[0] < hi])
        union.sort()
        k1 = np.array([x[0] for x in union], dtype=np.int32)
        k2 = np.array([x[1] for x in union], dtype=np.int32)
        keep = forest_merge(n, k1, k2)
        assert [(int(a), int(b)) for a, b in zip(k1[keep], k2[keep])] == want


def test_merge_split_edges_two_cell_filter():
    # rows 0..5 split sub-bucket; forest edges (0,1), (2,3), (4,5) after the merge, plus a redundant (0,1)-closing edge
    split = np.array([True] * 6 + [False] * 2)
    k1 = np.array([0, 0, 1, 2, 4, 6], dtype=np.int32)
    k2 = np.array([1, 2, 2, 3, 5, 7], dtype=np.int32)        # (1,2) closes a cycle with (0,1),(0,2): dropped by Kruskal
    cd = np.array([0, 0, 0, 3, 3, 9], dtype=np.int32)
    msh = np.array([9, 9, 9, 4, 7, 0], dtype=np.int32)       # (2,3): 3 > 4//2 -> dropped if two cells; (4,5): 3 > 7//2 is false
    ncells = np.array([1, 1, 1, 1, 1, 1, 1, 1], dtype=np.int64)
    keep = multi.merge_split_edges(8, k1, k2, cd, msh, split, ncells)
    # (0,1),(0,2) are in an orbit of three rows: kept; (1,2) cycle: dropped; (2,3): rows 2 has degree 2 -> kept;
    assert keep.tolist() == [True, True, False, True, True, True]
    # an isolated two-row, two-cell orbit with cd > min_sh // 2 loses its edge; the unsplit edge (6,7) is never touched
    k1 = np.array([0, 2, 6], dtype=np.int32)
    k2 = np.array([1, 3, 7], dtype=np.int32)
    cd = np.array([3, 3, 9], dtype=np.int32)
    msh = np.array([4, 7, 0], dtype=np.int32)
    keep = multi.merge_split_edges(8, k1, k2, cd, msh, split, ncells)
    assert keep.tolist() == [False, True, True]
    ncells[0] = 2
    assert multi.merge_split_edges(8, k1, k2, cd, msh, split, ncells).tolist() == [True, True, True]
    ncells[0] = 1
    assert multi.merge_split_edges(8, k1, k2, cd, msh, split, ncells, easy=True).tolist() == [True, True, True]


def test_probe_live_fractions_with_a_stub_context():
    """The partitioner's sample join, with the join replaced by a stub that returns the generator's true clone of every sampled
    row: pairs inside one clone never count, the stub that joins nothing makes every CDR3-close, not donor-rejected pair
    count, and a partition built from the fractions still gives every row exactly one first-row owner."""
    rep = E.generate(1, n_cells=120000, seed=41)
    inp = rep.to_join_input()
    clone = np.asarray(rep.row_clone()).copy()
    p = E.default_params(True)

    class Stub:
        def __init__(self, labels_of):
            self.labels_of = labels_of
            self.rows = None

        def set_row_roles(self, roles):
            assert roles is None

        def join(self, params, probe):
            class R:
                pass
            r = R()
            r.labels = self.labels_of(self.rows)
            assert len(r.labels) == probe.n_rows
            return r

    orig = inp.subset_rows

    def run(stub):
        def capture(rows):
            stub.rows = rows
            return orig(rows)
        inp.subset_rows = capture
        try:
            return multi.probe_live_fractions(inp, stub, p, min_rows=1024, sample=128)
        finally:
            inp.subset_rows = orig

    true = run(Stub(lambda rows: clone[rows]))
    none = run(Stub(lambda rows: np.arange(len(rows))))
    assert len(true) >= 1 and set(true) == set(none)
    assert all(0.0 <= v <= 1.0 for v in true.values()) and all(0.0 <= v <= 1.0 + 1e-9 for v in none.values())
    assert max(none.values()) > 0.5 and all(none[s] >= true[s] for s in true)
    parts, split = multi.partition_work(inp, 4, live_fraction=none)
    first = np.zeros(inp.n_rows, dtype=np.int64)
    for part in parts:
        roles = part.roles if part.roles is not None else np.zeros(len(part.rows), dtype=np.uint8)
        first[part.rows[(roles & multi.ROLE_COLUMN_ONLY) == 0]] += 1
    assert first.max() == 1
