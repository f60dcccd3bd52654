"""The oracle against hand-derived expectations, one micro-case per decision row of join_one
 plus the order-dependent pieces of join_core / join_exacts / finish_join."""
import numpy as np
import pytest

import cases
import enclone_ranger_b200 as E
from oracle import pyoracle
from util import assert_same_result, load_golden

CASES = cases.all_cases()


@pytest.mark.parametrize("name,params,inp,expected", CASES, ids=[c[0] for c in CASES])
def test_micro_case(name, params, inp, expected):
    r = pyoracle.run(params, inp.c_ptr(), threads=1)
    got = list(zip(r["edge_k1"].tolist(), r["edge_k2"].tolist()))
    if expected is not None:
        assert got == expected
    # invariants of every result
    assert r["joins"] == len(got)
    x, y, z = r["eq_x"], r["eq_y"], r["eq_z"]
    for k1, k2 in got:
        assert k1 < k2 and y[k1] == y[k2]
    assert np.array_equal(np.sort(x), np.arange(len(x)))  # x is a permutation (disjoint cycles)


def case(name):
    for c in CASES:
        if c[0] == name:
            return c
    raise KeyError(name)


def test_err_flag_semantics():
    _, p, inp, _ = case("row8_mix_donors_join_err")
    assert pyoracle.run(p, inp.c_ptr(), 1)["edge_err"].tolist() == [1]
    _, p, inp, _ = case("base_join")
    assert pyoracle.run(p, inp.c_ptr(), 1)["edge_err"].tolist() == [0]


def test_two_refs_tuple():
    _, p, inp, _ = case("row9_two_refs_join")
    r = pyoracle.run(p, inp.c_ptr(), 1)
    # position 200 is the allele difference: against the alt reference both contigs carry one more shared mutation
    assert r["edge_nrefs"].tolist() == [2]
    assert r["edge_shares"].tolist() == [[16, 17]] and r["edge_indeps"].tolist() == [[4, 4]]
    assert r["edge_d"].tolist() == [16] and r["edge_k"].tolist() == [4 + 2 * 16]
    assert r["edge_n"].tolist() == [3 * (362 + 332)]


def test_finish_join_merges_rows_of_one_exact():
    _, p, inp, _ = case("finish_join_three_chain_merge")
    r = pyoracle.run(p, inp.c_ptr(), 1)
    rows_of_exact0 = np.nonzero(inp.a["row_exact"] == 0)[0]
    assert len(rows_of_exact0) == 2
    assert r["eq_y"][rows_of_exact0[0]] == r["eq_y"][rows_of_exact0[1]]
    assert len(r["edge_k1"]) == 0  # raw_joins does not contain the clonotype_id merge (join2.rs:69-84)


def test_equivrel_tie_rule_and_sizes():
    # equiv/src/lib.rs:64-93: equal sizes -> b's orbit is relabelled to a's class; size stored at the class id
    x, y, z = pyoracle.equiv_replay(6, [3, 0, 1], [1, 2, 0])
    assert y.tolist() == [3, 3, 3, 3, 4, 5]     # join(3,1): 1 -> class 3; join(0,2): 2 -> 0; join(1,0): {0,2} (size 2) == {3,1}: b's orbit moves
    assert z[3] == 4
    orbit = [3]
    b = x[3]
    while b != 3:
        orbit.append(int(b)); b = x[b]
    assert sorted(orbit) == [0, 1, 2, 3]


def test_unsupported_modes_are_rejected():
    _, p, inp, _ = case("base_join")
    p2 = E.default_params(True, unsupported_modes=E._abi.EJB_MODE_BCJOIN)
    with pytest.raises(RuntimeError):
        pyoracle.run(p2, inp.c_ptr(), 1)


def test_stirling_range_is_an_error_not_a_wrong_answer():
    # k = indeps + 2*shares > 3000 would index past the reference's table (join_one.rs:30)
    w = cases.World()
    rng = w.rng
    w.vh = cases.rand_seq(rng, 2400); w.VH = 2400
    w.seg_vh = w.b.seg(w.vh); w.ref_vh = w.b.ref(w.vh, 77)
    h0, l0 = w.heavy(), w.light()
    shared = list(range(10, 2380))[:1600]   # k = 2 + 2 * 1600 > 3000
    w.member(cases.mutate(h0, shared + [3]), l0, feats=False); w.member(cases.mutate(h0, shared + [7]), l0, feats=False)
    with pytest.raises(RuntimeError, match="panic"):
        pyoracle.run(E.default_params(True), w.build().c_ptr(), 1)


@pytest.mark.parametrize("name", ["hs_bcr_vdjrefs7.npz", "synth_bcr_2k.npz", "synth_tcr_20k.npz"])
def test_oracle_reproduces_golden(name):
    inp, want, p = load_golden(name)
    got = pyoracle.run(p, inp.c_ptr(), threads=0)
    assert_same_result(got, want, name)


def test_oracle_thread_count_does_not_matter():
    rep = E.generate(0, n_cells=3000, seed=99)
    p = E.default_params(True)
    a = pyoracle.run(p, rep.c_ptr, threads=1)
    b = pyoracle.run(p, rep.c_ptr, threads=4)
    assert_same_result(a, b)


def test_forest_equals_sequential_replay():
    """raw_joins replayed through EquivRel must never hit an already-joined pair (it is a forest) and the
    partition must equal the partition of the ground-truth-agnostic union of all edges."""
    rep = E.generate(1, n_cells=6000, seed=5)
    r = pyoracle.run(E.default_params(True), rep.c_ptr, threads=0)
    parent = list(range(rep.c.n_rows))

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a
    last = (-1, -1)
    for k1, k2 in zip(r["edge_k1"].tolist(), r["edge_k2"].tolist()):
        assert (k1, k2) > last          # bucket order, then lexicographic == globally sorted
        last = (k1, k2)
        a, b = find(k1), find(k2)
        assert a != b                   # forest
        parent[a] = b


def test_fuzz_generator_is_sane():
    """The random worlds used by the GPU differential test exercise joins, two-reference joins and err joins."""
    import fuzz
    edges = nrefs2 = errs = 0
    for seed in range(40):
        p, inp = fuzz.random_world(seed)
        r = pyoracle.run(p, inp.c_ptr(), 1)
        edges += len(r["edge_k1"]); nrefs2 += int((r["edge_nrefs"] == 2).sum()); errs += int(r["edge_err"].sum())
    assert edges > 50 and nrefs2 > 5 and errs > 5
