"""Parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the same inputs.
Integer / index results bit-exact; fp64 p1 / mult / score bit-exact as well (=> within 1e-12 relative)."""
import numpy as np
import pytest

import cases
import enclone_ranger_b200 as E
from enclone_ranger_b200 import _abi
from oracle import pyoracle
from util import assert_same_result, load_golden

pytestmark = pytest.mark.gpu
CASES = cases.all_cases()


@pytest.mark.parametrize("name,params,inp,expected", CASES, ids=[c[0] for c in CASES])
def test_micro_case_matches_oracle(gpu_ctx, name, params, inp, expected):
    want = pyoracle.run(params, inp.c_ptr(), threads=1)
    got = gpu_ctx.join(params, inp)
    assert_same_result(got, want, name)
    if expected is not None:
        assert list(zip(got.edge_k1.tolist(), got.edge_k2.tolist())) == expected


@pytest.mark.parametrize("name", ["hs_bcr_vdjrefs7.npz", "synth_bcr_2k.npz", "synth_tcr_20k.npz"])
def test_golden_fixture(gpu_ctx, name):
    inp, want, p = load_golden(name)
    assert_same_result(gpu_ctx.join(p, inp), want, name)


@pytest.mark.parametrize("config,cells,seed", [(0, 10000, None), (0, 30000, 3), (1, 40000, None), (2, 150000, None), (3, 30000, None)])
def test_synthetic_config_matches_oracle(gpu_ctx, config, cells, seed):
    rep = E.generate(config, n_cells=cells, seed=seed)
    p = E.default_params(bool(rep.config.is_bcr))
    want = pyoracle.run(p, rep.c_ptr, threads=0)
    got = gpu_ctx.join(p, rep)
    assert_same_result(got, want, f"config {config}")
    assert got.stats["pairs_scored"] == want["pairs_scored"] and got.stats["pairs_lens"] == want["pairs_lens"]
    assert got.stats["n_buckets"] == want["n_buckets"] and got.stats["two_cell_deleted"] == want["two_cell_deleted"]
    assert got.stats["errors"] == want["errors"] and got.stats["joins"] == want["joins"]
    # the device labels (union-find) agree with the replayed EquivRel
    assert np.array_equal(got.labels, want["labels"])


@pytest.mark.parametrize("name,params,inp,expected", CASES, ids=[c[0] for c in CASES])
def test_micro_case_packed_contigs(gpu_ctx, name, params, inp, expected):
    """Same cases with every eligible contig supplied 2-bit packed (the tigsp path of the ABI)."""
    want = pyoracle.run(params, inp.c_ptr(), threads=1)
    assert_same_result(gpu_ctx.join(params, inp.to_packed()), want, name + " (packed)")


@pytest.mark.parametrize("config,cells", [(0, 20000), (1, 40000), (2, 100000)])
def test_synthetic_packed_contigs(gpu_ctx, config, cells):
    rep = E.generate(config, n_cells=cells, seed=77)
    p = E.default_params(bool(rep.config.is_bcr))
    inp = rep.to_join_input()
    pk = inp.to_packed()
    assert pk.a["tig_bytes"].size < inp.a["tig_bytes"].size // 10
    want = pyoracle.run(p, inp.c_ptr(), threads=0)
    got = gpu_ctx.join(p, pk)
    assert_same_result(got, want, f"config {config} packed")
    assert got.stats["h2d_bytes"] < 0.6 * gpu_ctx.join(p, inp).stats["h2d_bytes"]


def test_tiny_work_buffers_force_round_splitting():
    """A candidate buffer far smaller than one round's candidates: the scheduler must split rounds / row-block
    groups and still produce the identical forest (the order-dependent part of join_core)."""
    rep = E.generate(1, n_cells=20000, seed=8)
    p = E.default_params(True)
    want = pyoracle.run(p, rep.c_ptr, threads=0)
    # tiny pair budget: many scheduler rounds (>= 128 x the largest sub-bucket must still fit the buffers)
    small = E.JoinContext(cand_cap=1 << 22, pair_budget=1 << 20)
    got = small.join(p, rep)
    assert got.stats["rounds"] > 20
    assert_same_result(got, want, "small budget")
    small.close()
    # buffers smaller than a round's candidates: overflow -> split -> retry
    tight = E.JoinContext(cand_cap=1 << 21, pair_budget=1 << 30)
    got = tight.join(p, rep)
    assert got.stats["overflow_retries"] >= 0   # whether a retry happens depends on the candidate rate prediction
    assert_same_result(got, want, "overflow retry")
    tight.close()


def test_run_is_deterministic_and_idempotent(gpu_ctx):
    rep = E.generate(3, n_cells=20000, seed=2)
    p = E.default_params(True)
    a = gpu_ctx.join(p, rep)
    b = gpu_ctx.join(p, rep)
    for f in ("edge_k1", "edge_k2", "edge_score", "labels", "eq_x", "eq_y", "eq_z"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    gpu_ctx.upload(p, rep)
    gpu_ctx.run()
    gpu_ctx.run()
    c = gpu_ctx.download()
    assert np.array_equal(a.edge_k1, c.edge_k1) and np.array_equal(a.edge_score, c.edge_score)


def test_p1_kernel_bit_exact(gpu_ctx):
    rng = np.random.default_rng(0)
    k = rng.integers(0, 400, 300)
    d = (rng.random(300) * (k + 1)).astype(np.int64)
    n = rng.integers(1500, 3000, 300)
    k[:4] = [0, 5, 30, 3000]; d[:4] = [0, 0, 3, 40]; n[:4] = [2000, 2000, 2500, 9000]
    got = gpu_ctx.p1_batch(k, d, n)
    want = np.array([pyoracle.p1(int(a), int(b), int(c)) for a, b, c in zip(k, d, n)])
    assert np.array_equal(got, want)
    assert f"{got[2]:.7f}" == "0.0005953"   # stirling_numbers/src/lib.rs:315


def _sharded_join(p, inp, world, cand_cap=1 << 23):
    """`world` contexts (here: all on one GPU, run back to back) each upload the whole input and join their shard; the partial
    forests are concatenated in the first context's merge buffers and merged there."""
    import torch
    ctxs = [E.JoinContext(rank=r, world=world, cand_cap=cand_cap) for r in range(world)]
    try:
        parts, pairs = [], 0
        for c in ctxs:
            c.upload(p, inp)
            st = c.run()
            pairs += st["pairs_scored"]
            parts.append(c.part_forest())
        total = sum(n for n, _, _ in parts)
        ew, tw = ctxs[0].merge_buffers(total)
        off = 0
        for n, e, t in parts:
            ew[off:off + n].copy_(e)
            tw[32 * off:32 * (off + n)].copy_(t)
            off += n
        torch.cuda.synchronize()
        ctxs[0].merge_parts(total)
        return ctxs[0].download(), pairs, [n for n, _, _ in parts]
    finally:
        for c in ctxs:
            c.close()


@pytest.mark.parametrize("config,cells,world,seed", [(1, 30000, 3, 17), (1, 60000, 8, 32), (3, 30000, 4, 33), (2, 100000, 2, 5)])
def test_sharded_contexts_merge_to_the_full_result(config, cells, world, seed):
    """In-library sharding (ejb_set_shard): whole sub-buckets by LPT, the heaviest cut into row ranges; the merge (forest of
    the partial forests, two-cell filter, finalize) must give the oracle's edges, tuples and EquivRel of the WHOLE join."""
    rep = E.generate(config, n_cells=cells, seed=seed)
    inp = rep.to_join_input()
    p = E.default_params(bool(rep.config.is_bcr))
    want = pyoracle.run(p, inp.c_ptr(), threads=0)
    got, pairs, sizes = _sharded_join(p, inp.to_packed().to_packed_cdr3(), world)
    assert pairs == want["pairs_scored"], "the shards' pair domains must partition the whole domain"
    assert_same_result(got, want, f"config {config} sharded x{world}")
    assert np.array_equal(got.labels, want["labels"])
    if world == 8:
        assert sum(sizes) > len(want["edge_k1"]) + want["two_cell_deleted"], "split sub-buckets: the partial forests overlap"


def test_multi_device_context_single_call(gpu_ctx):
    """ejb_create with n_devices > 1 (here: three contexts on the one GPU of the test box): ejb_join shards, gathers GPU-to-GPU,
    merges and returns one result."""
    rep = E.generate(1, n_cells=50000, seed=44)
    inp = rep.to_join_input()
    p = E.default_params(True)
    want = pyoracle.run(p, inp.c_ptr(), threads=0)
    multi = E.JoinContext(device=[0, 0, 0], cand_cap=1 << 23)
    try:
        got = multi.join(p, inp.to_packed().to_packed_cdr3())
        assert got.stats["n_devices"] == 3 and got.stats["pairs_scored"] == want["pairs_scored"]
        assert_same_result(got, want, "3-context ejb_join")
        assert np.array_equal(got.labels, want["labels"])
    finally:
        multi.close()


def test_device_resident_edges_and_host_partitioning(gpu_ctx):
    """The multi-GPU data path on one GPU: whole buckets partitioned on the host (LPT), each part joined on its own,
    device-resident (k1,k2) mapped back to global rows, merged -> equals the oracle on the whole input."""
    import torch
    from enclone_ranger_b200.multi import partition_buckets
    rep = E.generate(1, n_cells=30000, seed=23)
    inp = rep.to_join_input()
    p = E.default_params(True)
    want = pyoracle.run(p, inp.c_ptr(), threads=0)
    keys = []
    for rows in partition_buckets(inp, 3):
        sub, _ = inp.subset_rows(rows)
        gpu_ctx.upload(p, sub)
        gpu_ctx.run()
        k1, k2 = gpu_ctx.device_edges()
        host = gpu_ctx.download()
        assert np.array_equal(k1.cpu().numpy(), host.edge_k1) and np.array_equal(k2.cpu().numpy(), host.edge_k2)
        r = torch.from_numpy(rows).to(k1.device)
        keys.append((r[k1.long()] << 32) | r[k2.long()])
    k = torch.sort(torch.cat(keys)).values.cpu().numpy()
    assert np.array_equal((k >> 32).astype(np.int32), want["edge_k1"]) and np.array_equal((k & 0xffffffff).astype(np.int32), want["edge_k2"])
    eq, lab = E.equiv_replay(inp.n_rows, (k >> 32).astype(np.int32), (k & 0xffffffff).astype(np.int32), inp.a["row_exact"])
    assert np.array_equal(eq.y, want["eq_y"]) and np.array_equal(lab, want["labels"])


def test_properties_at_scale(gpu_ctx):
    """Size-independent properties where the oracle is too slow: sorted forest, one bucket per edge,
    labels == connected components of edges + clonotype_id merge, idempotent under re-run."""
    rep = E.generate(1, n_cells=300000, seed=1)
    p = E.default_params(True)
    res = gpu_ctx.join(p, rep)
    k1, k2 = res.edge_k1.astype(np.int64), res.edge_k2.astype(np.int64)
    key = k1 << 32 | k2
    assert np.all(key[1:] > key[:-1]) and np.all(k1 < k2)
    inp = rep.to_join_input()
    lens = inp.a["row_len"].reshape(-1, 2)
    assert np.array_equal(lens[k1], lens[k2])
    c3 = inp.a["row_cdr3_len"].reshape(-1, 2)
    assert np.array_equal(c3[k1], c3[k2])
    eq, lab = E.equiv_replay(inp.n_rows, res.edge_k1, res.edge_k2, inp.a["row_exact"])
    assert np.array_equal(lab, res.labels) and np.array_equal(eq.y, res.eq_y)
    assert eq.norbits() == len(np.unique(res.labels))
    assert res.stats["forest_edges"] - res.stats["two_cell_deleted"] == len(k1)
    # every accepted score obeys the reference's threshold logic (join_one.rs:710-715)
    d = res.edge_d
    assert np.all((res.edge_score <= 100000.0) | (d >= 15) | (res.edge_cd <= 100))


def test_error_behaviour(gpu_ctx):
    _, p, inp, _ = next(c for c in CASES if c[0] == "base_join")
    with pytest.raises(E.JoinError) as ei:
        gpu_ctx.join(E.default_params(True, unsupported_modes=_abi.EJB_MODE_BASIC), inp)
    assert ei.value.code == _abi.EJB_ERR_UNSUPPORTED
    bad = E.JoinInput(**{**inp.a, "row_tig_off": inp.a["row_tig_off"] + 10 ** 9})
    with pytest.raises(E.JoinError) as ei:
        gpu_ctx.join(p, bad)
    assert ei.value.code == _abi.EJB_ERR_INVALID
    # the context stays usable after an error
    assert len(gpu_ctx.join(p, inp).edge_k1) == 1


def test_stirling_range_error(gpu_ctx):
    w = cases.World()
    w.vh = cases.rand_seq(w.rng, 2400); w.VH = 2400
    w.seg_vh = w.b.seg(w.vh); w.ref_vh = w.b.ref(w.vh, 77)
    h0, l0 = w.heavy(), w.light()
    shared = list(range(10, 2380))[:1600]   # k = 2 + 2 * 1600 > 3000
    w.member(cases.mutate(h0, shared + [3]), l0, feats=False); w.member(cases.mutate(h0, shared + [7]), l0, feats=False)
    with pytest.raises(E.JoinError) as ei:
        gpu_ctx.join(E.default_params(True), w.build())
    assert ei.value.code == _abi.EJB_ERR_RANGE


def test_empty_and_ragged_inputs(gpu_ctx):
    p = E.default_params(True)
    empty = E.JoinInput()
    r = gpu_ctx.join(p, empty)
    assert len(r.edge_k1) == 0 and len(r.labels) == 0
    _, p, inp, _ = next(c for c in CASES if c[0] == "onesies_never_join")
    r = gpu_ctx.join(p, inp)
    assert len(r.edge_k1) == 0 and r.labels.tolist() == [0, 1]


@pytest.mark.parametrize("block", range(5))
def test_randomized_differential(gpu_ctx, block):
    """Random small inputs mixing rare branches of join_one (gaps with and without shifted seq_del_amino, odd bytes,
    alternative V/J references, gene-name / UTR variants, jittered feature starts, donors, barcode overlaps, parameter
    variants): CUDA result == oracle, for ASCII and for 2-bit packed contigs."""
    import fuzz
    n_edges = 0
    for seed in range(block * 60, block * 60 + 60):
        p, inp = fuzz.random_world(seed)
        want = pyoracle.run(p, inp.c_ptr(), threads=1)
        assert_same_result(gpu_ctx.join(p, inp), want, f"fuzz seed {seed}")
        assert_same_result(gpu_ctx.join(p, inp.to_packed()), want, f"fuzz seed {seed} packed")
        n_edges += len(want["edge_k1"])
    assert n_edges > 0


def _join_without_tags(ctx, p, inp):
    ctx.set_option("no_tags", 1)
    try:
        return ctx.join(p, inp)
    finally:
        ctx.set_option("no_tags", 0)


def test_reference_set_tags_do_not_change_the_result(gpu_ctx):
    """Stage 1 drops CDR3-close pairs one of whose rows is already known to fail the degradation test
    (join_one.rs:434-440) against the other's reference set.  Pure work saving: same edges, tuples and EquivRel
    as the oracle and as a run with the tags switched off."""
    rep = E.generate(1, n_cells=60000, seed=5)
    p = E.default_params(True)
    want = pyoracle.run(p, rep.c_ptr, threads=0)
    got = gpu_ctx.join(p, rep)
    assert_same_result(got, want, "tags on")
    off = _join_without_tags(gpu_ctx, p, rep)
    assert_same_result(off, want, "tags off")
    assert off.stats["stage1_ref_kills"] == 0 and off.stats["stage1_flag_skips"] == 0


@pytest.mark.parametrize("opt,value", [("thin", 0), ("thin", 96), ("thin", 128), ("quot_table", 0)])
def test_kernel_variants_do_not_change_the_result(gpu_ctx, opt, value):
    """The thin-unit stage-1 kernel (thread per column) against the tile kernel, and p1 from the constant quotient table
    against the three-quotient divisions: every variant gives the oracle's result, bit for bit."""
    rep = E.generate(1, n_cells=60000, seed=9)
    p = E.default_params(True)
    want = pyoracle.run(p, rep.c_ptr, threads=0)
    default = {"thin": 16, "quot_table": 1}[opt]
    gpu_ctx.set_option(opt, value)
    try:
        got = gpu_ctx.join(p, rep)
    finally:
        gpu_ctx.set_option(opt, default)
    assert_same_result(got, want, f"{opt}={value}")
    assert_same_result(gpu_ctx.join(p, rep.to_join_input().to_packed().to_packed_cdr3()), want, "defaults, packed input")


def test_reference_set_tags_at_scale(gpu_ctx):
    rep = E.generate(1, n_cells=300000, seed=4)
    p = E.default_params(True)
    inp = rep.to_join_input().to_packed()
    on = gpu_ctx.join(p, inp)
    off = _join_without_tags(gpu_ctx, p, inp)
    for f in ("edge_k1", "edge_k2", "edge_cd", "edge_shares", "edge_indeps", "edge_score", "labels", "eq_x", "eq_y", "eq_z"):
        assert np.array_equal(getattr(on, f), getattr(off, f)), f
    assert off.stats["stage1_ref_kills"] == 0 and off.stats["stage1_flag_skips"] == 0
    assert on.stats["stage1_ref_kills"] + on.stats["stage1_flag_skips"] > 0, "a 300k-cell heavy-tailed repertoire with 1 % indel rows must exercise the kill"
    assert on.stats["stage1_candidates"] < off.stats["stage1_candidates"]


@pytest.mark.parametrize("config,cells,world,seed", [(1, 40000, 4, 31), (1, 60000, 8, 32), (3, 30000, 3, 33)])
def test_row_range_split_of_large_sub_buckets(gpu_ctx, config, cells, world, seed):
    """The multi-GPU partitioner cuts the heaviest sub-buckets into row ranges (ejb_set_row_roles); every part is joined on
    its own (here: back to back on one GPU), the partial forests are merged (forest of forests + two-cell filter) and the
    result must be the oracle's: edges, PotentialJoin numerics, EquivRel."""
    from enclone_ranger_b200 import multi
    rep = E.generate(config, n_cells=cells, seed=seed)
    inp = rep.to_join_input()
    p = E.default_params(bool(rep.config.is_bcr))
    want = pyoracle.run(p, inp.c_ptr(), threads=0)
    parts, split = multi.partition_work(inp, world)
    assert split.any()
    got, pairs = [], 0
    try:
        for part in parts:
            sub, _ = inp.subset_rows(part.rows)
            gpu_ctx.set_row_roles(part.roles)
            res = gpu_ctx.join(p, sub)
            pairs += res.stats["pairs_scored"]
            d = {f: getattr(res, f) for f in multi.EDGE_FIELDS}
            d["edge_k1"] = part.rows[res.edge_k1].astype(np.int32)
            d["edge_k2"] = part.rows[res.edge_k2].astype(np.int32)
            got.append(d)
    finally:
        gpu_ctx.set_row_roles(None)
    assert pairs == want["pairs_scored"]
    merged = multi.merge_edge_parts(got)
    sh = merged["edge_shares"]
    min_sh = np.where(merged["edge_nrefs"] == 2, sh.min(axis=1), sh[:, 0])
    keep = multi.merge_split_edges(inp.n_rows, merged["edge_k1"], merged["edge_k2"], merged["edge_cd"], min_sh, split, multi.ncells_per_row(inp))
    assert (~keep).any(), "partial forests overlap: the merge must drop something"
    final = {f: v[keep] for f, v in merged.items()}
    assert_same_result(final, {k: want[k] for k in want if k.startswith("edge_")}, "row-range split")
    eq, lab = E.equiv_replay(inp.n_rows, final["edge_k1"], final["edge_k2"], inp.a["row_exact"])
    assert np.array_equal(eq.x, want["eq_x"]) and np.array_equal(eq.y, want["eq_y"]) and np.array_equal(eq.z, want["eq_z"])
    assert np.array_equal(lab, want["labels"])


def test_probe_live_fractions(gpu_ctx):
    """The partitioner's sample join: fractions in [0, 1] for the large sub-buckets only, and a partition built from them
    still covers every row once."""
    from enclone_ranger_b200 import multi
    rep = E.generate(1, n_cells=120000, seed=41)
    inp = rep.to_join_input()
    p = E.default_params(True)
    live = multi.probe_live_fractions(inp, gpu_ctx, p, min_rows=1024, sample=128)
    assert len(live) >= 1 and all(0.0 <= v <= 1.0 for v in live.values())
    parts, split = multi.partition_work(inp, 4, live_fraction=live)
    first = np.zeros(inp.n_rows, dtype=np.int64)
    for part in parts:
        roles = part.roles if part.roles is not None else np.zeros(len(part.rows), dtype=np.uint8)
        first[part.rows[(roles & multi.ROLE_COLUMN_ONLY) == 0]] += 1
    assert first.max() == 1


def test_input_arena_single_copy_upload(gpu_ctx):
    """ejb_set_input_arena: arrays inside the declared arena are uploaded with one copy and resolved by offset; arrays
    outside it (here: cdr3_bytes) go one by one.  Same result, fewer bytes in separate copies."""
    rep = E.generate(1, n_cells=30000, seed=51)
    inp = rep.to_join_input().to_packed()
    p = E.default_params(True)
    want = gpu_ctx.join(p, inp)
    total = sum((a.nbytes + 255) & ~255 for a in inp.a.values())
    arena = np.zeros(total + 512, dtype=np.uint8)
    off0 = (-arena.ctypes.data) & 255
    off, views = off0, {}
    for name, a in inp.a.items():
        if name == "cdr3_bytes":
            views[name] = a
            continue
        v = arena[off:off + a.nbytes].view(a.dtype).reshape(a.shape)
        v[...] = a
        views[name] = v
        off += (a.nbytes + 255) & ~255
    inp2 = E.JoinInput(**views)
    try:
        gpu_ctx.set_input_arena(arena.ctypes.data + off0, off - off0)
        got = gpu_ctx.join(p, inp2)
        again = gpu_ctx.join(p, inp)          # an input that lives outside the arena still works while it is set
    finally:
        gpu_ctx.set_input_arena(None, 0)
    for f in ("edge_k1", "edge_k2", "edge_cd", "edge_shares", "edge_indeps", "edge_score", "labels", "eq_x", "eq_y", "eq_z"):
        assert np.array_equal(getattr(got, f), getattr(want, f)), f
        assert np.array_equal(getattr(again, f), getattr(want, f)), f


def test_stirling_table_built_on_device_is_bit_exact(gpu_ctx):
    """f3: stirling2_ratio_table_double (start.rs:50-99) is built by k_sr_table on the device; every entry must carry the bits
    of the oracle's host build (same double-double operation sequence): whole rows incl. the reference's KAT rows 219 / 722,
    the last rows, and a seeded sample over the triangle."""
    assert gpu_ctx.create_ms() > 0.0
    rows = list(range(0, 40)) + [219, 722, 1500, 2047, 2048, 2049, 2999, 3000]
    for n in rows:
        got = gpu_ctx.sr_table_read(n * (n + 1) // 2, n + 1)
        want = np.array([pyoracle.sr_entry(n, k) for k in range(n + 1)])
        assert np.array_equal(got, want), f"row {n}"
    rng = np.random.default_rng(5)
    for _ in range(3000):
        n = int(rng.integers(1, 3001))
        k = int(rng.integers(0, n + 1))
        got = gpu_ctx.sr_table_read(n * (n + 1) // 2 + k, 1)[0]
        assert tuple(got) == pyoracle.sr_entry(n, k), (n, k)


@pytest.mark.parametrize("config,cells", [(0, 20000), (1, 40000), (2, 100000), (3, 20000)])
def test_packed_cdr3s(gpu_ctx, config, cells):
    """f1 intake: CDR3s supplied 2-bit packed (row_cdr32_off / cdr32_words) next to 2-bit contigs; no ASCII CDR3 byte is
    uploaded and the result is the oracle's."""
    rep = E.generate(config, n_cells=cells, seed=91)
    p = E.default_params(bool(rep.config.is_bcr))
    inp = rep.to_join_input()
    pk = inp.to_packed().to_packed_cdr3()
    assert pk.a["cdr3_bytes"].size == 0 and pk.a["cdr32_words"].size > 0
    want = pyoracle.run(p, inp.c_ptr(), threads=0)
    got = gpu_ctx.join(p, pk)
    assert_same_result(got, want, f"config {config} packed CDR3s")
    assert got.stats["h2d_bytes"] < gpu_ctx.join(p, inp.to_packed()).stats["h2d_bytes"]


def test_packed_cdr3s_micro_cases(gpu_ctx):
    for name, params, inp, expected in CASES:
        want = pyoracle.run(params, inp.c_ptr(), threads=1)
        assert_same_result(gpu_ctx.join(params, inp.to_packed().to_packed_cdr3()), want, name + " (packed CDR3)")


def test_sync_and_launch_accounting(gpu_ctx):
    """One read-back for the prepare phase, one per scheduler round, one at the end; Boruvka converges on the device."""
    rep = E.generate(1, n_cells=60000, seed=3)
    p = E.default_params(True)
    gpu_ctx.upload(p, rep.to_join_input().to_packed())
    st = gpu_ctx.run()
    assert st["host_syncs"] == st["rounds"] + st["overflow_retries"] + 2, st
    assert st["boruvka_iterations"] >= st["rounds"] - 1
    assert 0 < st["pairs_distance_computed"] <= st["pairs_evaluated"] * 1.2
    assert st["pairs_distance_full"] <= st["pairs_distance_computed"]


@pytest.mark.parametrize("config,cells,seed", [(1, 30000, 61), (3, 20000, 62), (2, 60000, 63), (0, 8000, 64)])
def test_join_one_batch_matches_oracle(gpu_ctx, config, cells, seed):
    """f2: define_mat's second use of the pair predicate (enclone_process/src/define_mat.rs:172,247).  join_one for arbitrary
    (k1, k2) lists — joined pairs in both orders, pairs inside one class (no class skip here), random pairs of one lens
    bucket, pairs across buckets, a row with itself, rows without two chains — must equal the oracle's join_one, members
    included."""
    rep = E.generate(config, n_cells=cells, seed=seed)
    inp = rep.to_join_input()
    p = E.default_params(bool(rep.config.is_bcr))
    res = gpu_ctx.join(p, inp.to_packed().to_packed_cdr3())
    rng = np.random.default_rng(seed)
    N = inp.n_rows
    lens = inp.a["row_len"].reshape(-1, 2)
    head = np.ones(N, dtype=bool)
    head[1:] = (lens[1:] != lens[:-1]).any(axis=1) | (inp.a["row_nchains"][1:] != inp.a["row_nchains"][:-1])
    bucket = np.cumsum(head) - 1
    start = np.nonzero(head)[0]
    size = np.diff(np.append(start, N))
    k1, k2 = [res.edge_k1, res.edge_k2], [res.edge_k2, res.edge_k1]           # joined pairs, both orders
    lab = res.labels
    order = np.argsort(lab, kind="stable")                                       # pairs inside one class that are not forest edges
    same = np.nonzero(lab[order][1:] == lab[order][:-1])[0][:4000]
    k1.append(order[same + 1]); k2.append(order[same])
    a = rng.integers(0, N, 20000)                                                # random pairs of one lens bucket
    b = start[bucket[a]] + rng.integers(0, 1 << 30, 20000) % size[bucket[a]]
    k1.append(a); k2.append(b)
    k1.append(rng.integers(0, N, 3000)); k2.append(rng.integers(0, N, 3000))     # arbitrary pairs (mostly different lens)
    k1.append(np.arange(0, N, max(N // 500, 1))); k2.append(np.arange(0, N, max(N // 500, 1)))   # a row with itself
    k1 = np.concatenate(k1).astype(np.int32)
    k2 = np.concatenate(k2).astype(np.int32)
    want = pyoracle.join_one_batch(p, inp.c_ptr(), k1, k2)
    got = gpu_ctx.join_one_batch(k1, k2)
    assert np.array_equal(got["passed"], want["passed"]), f"{int((got['passed'] != want['passed']).sum())} of {len(k1)} decisions differ"
    ok = want["passed"] == 1
    assert ok.sum() >= len(res.edge_k1), "every forest edge is a pair join_one accepts"
    for f in ("cd", "nrefs", "shares", "indeps", "k", "d", "n", "p1", "mult", "score", "err"):
        assert np.array_equal(got[f][ok], want[f][ok]), f
    # the reverse order of a joined pair is accepted too unless an order-dependent test says otherwise: whatever it is, equal to the oracle
    assert got["passed"][len(res.edge_k1):2 * len(res.edge_k1)].sum() > 0


def test_multi_device_context_with_sliced_arena_upload():
    """ejb_join on a multi-device context with a declared input arena: every context uploads only its slice of the arena, the
    rest arrives by GPU-to-GPU copies (the input crosses PCIe once).  Same result as the oracle."""
    rep = E.generate(1, n_cells=50000, seed=45)
    src = rep.to_join_input()
    inp = src.to_packed().to_packed_cdr3()
    p = E.default_params(True)
    want = pyoracle.run(p, src.c_ptr(), threads=0)
    total = sum((a.nbytes + 255) & ~255 for a in inp.a.values())
    arena = np.zeros(total + 512, dtype=np.uint8)
    off0 = (-arena.ctypes.data) & 255
    off, views = off0, {}
    for name, a in inp.a.items():
        v = arena[off:off + a.nbytes].view(a.dtype).reshape(a.shape)
        v[...] = a
        views[name] = v
        off += (a.nbytes + 255) & ~255
    inp2 = E.JoinInput(**views)
    multi = E.JoinContext(device=[0, 0, 0, 0], cand_cap=1 << 23)
    try:
        multi.set_input_arena(arena.ctypes.data + off0, off - off0)
        got = multi.join(p, inp2)
        assert got.stats["n_devices"] == 4
        assert got.stats["h2d_bytes"] < 1.3 * (off - off0), "the arena must cross PCIe once, not once per context"
        assert_same_result(got, want, "4-context ejb_join, sliced arena")
    finally:
        multi.close()
