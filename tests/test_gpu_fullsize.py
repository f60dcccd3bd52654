"""Parity at the sizes BASELINE.json names (north_star: "identical clonotype assignments ... on a 10M-cell synthetic BCR
repertoire").  Three kinds of evidence, all through the C ABI:

  * configs[1] (1M cells): the CPU oracle joins the WHOLE repertoire on the GPU box's host cores (about a minute) and every
    edge, PotentialJoin member, EquivRel vector and label is compared;
  * every config at full size: sha256 digests of the GPU result against tests/golden/digests.json, which
    tests/golden/make_digests.py produced offline from full-size oracle runs (join.rs:93-214, join_core.rs:23-50);
  * configs[3] / configs[4]: buckets are independent (join.rs:70-88), so the oracle is also run live on a seeded set of
    whole lens buckets INCLUDING THE LARGEST ONE and compared with the GPU's edges restricted to those buckets;
  * the stage-1 work savers (reference-set tags, donor kills) at the 8M-cell regime: identical results with them off.
"""
import json
import os
import sys

import numpy as np
import pytest

import enclone_ranger_b200 as E
from oracle import pyoracle
from util import GOLDEN, assert_same_result

sys.path.insert(0, GOLDEN)
from make_digests import ARRAYS, COUNTERS, digest_of  # noqa: E402

pytestmark = pytest.mark.gpu
DIGESTS = json.load(open(os.path.join(GOLDEN, "digests.json"))) if os.path.exists(os.path.join(GOLDEN, "digests.json")) else {}


def _check_digest(key, res, n_rows):
    if key not in DIGESTS:
        pytest.skip(f"no committed oracle digest for {key}")
    want = DIGESTS[key]
    got = digest_of(res, n_rows)
    assert got["n_rows"] == want["n_rows"] and got["n_edges"] == want["n_edges"], (key, got["n_edges"], want["n_edges"])
    assert got["edges_per_64k_rows"] == want["edges_per_64k_rows"], f"{key}: edges per 64k-row block differ"
    for f in ARRAYS:
        assert got[f] == want[f], f"{key}: sha256({f}) differs from the full-size oracle run"
    assert got["orbits"] == want["orbits"]
    for c in COUNTERS:
        assert int(res.stats[c]) == want[c], (key, c, int(res.stats[c]), want[c])


def _bucket_ids(inp):
    lens = inp.a["row_len"].reshape(-1, 2)
    nch = inp.a["row_nchains"]
    head = np.ones(inp.n_rows, dtype=bool)
    head[1:] = (lens[1:] != lens[:-1]).any(axis=1) | (nch[1:] != nch[:-1])
    return np.cumsum(head) - 1


def _sampled_buckets_vs_oracle(inp, p, res, n_extra, seed, max_rows):
    """Oracle on whole lens buckets (the largest + a seeded sample) == the GPU's edges restricted to those buckets."""
    bucket = _bucket_ids(inp)
    sizes = np.bincount(bucket)
    two = np.zeros(len(sizes), dtype=bool)
    two[bucket[inp.a["row_nchains"] == 2]] = True
    cand = np.nonzero(two & (sizes >= 2))[0]
    largest = cand[np.argmax(sizes[cand])]
    chosen, rows_left = [int(largest)], max_rows - int(sizes[largest])
    for b in np.random.default_rng(seed).permutation(cand):
        if len(chosen) > n_extra:
            break
        if b != largest and sizes[b] <= rows_left:
            chosen.append(int(b))
            rows_left -= int(sizes[b])
    rows = np.nonzero(np.isin(bucket, np.array(chosen)))[0]
    os.environ["ORACLE_P1_MEMO"] = "1"   # p1 is a pure function of (k, d, n): same bits, bounded run time on hypermutated clones
    sub = inp.take_rows(rows)            # keep the subset alive while the oracle reads it
    want = pyoracle.run(p, sub.c_ptr(), threads=0)
    sel = np.nonzero(np.isin(res.edge_k1, rows))[0]
    got = {f: np.asarray(getattr(res, f))[sel] for f in ARRAYS if f.startswith("edge_")}
    want = {k: v for k, v in want.items() if k.startswith("edge_")}
    want["edge_k1"] = rows[want["edge_k1"]].astype(np.int32)
    want["edge_k2"] = rows[want["edge_k2"]].astype(np.int32)
    assert len(want["edge_k1"]) > 0
    assert_same_result(got, want, f"{len(chosen)} whole buckets (largest {int(sizes[largest])} rows), {len(rows)} rows")
    return len(chosen), int(sizes[largest]), len(rows)


def test_config1_fullsize_matches_oracle(gpu_ctx):
    """BASELINE configs[1] at its full 1M cells: CUDA == oracle on every output, plus the committed digest."""
    rep = E.generate(1)
    p = E.default_params(True)
    inp = rep.to_join_input()
    rep.close()
    got = gpu_ctx.join(p, inp.to_packed())
    want = pyoracle.run(p, inp.c_ptr(), threads=0)
    assert_same_result(got, want, "configs[1] 1M cells")
    for c in COUNTERS:
        assert int(got.stats[c]) == int(want[c]), c
    assert np.array_equal(got.labels, want["labels"])
    _check_digest("config1", got, inp.n_rows)


@pytest.mark.parametrize("config", [0, 2])
def test_fullsize_digest(gpu_ctx, config):
    rep = E.generate(config)
    p = E.default_params(bool(rep.config.is_bcr))
    inp = rep.to_join_input()
    rep.close()
    _check_digest(f"config{config}", gpu_ctx.join(p, inp.to_packed()), inp.n_rows)


@pytest.mark.parametrize("config,n_extra,max_rows", [(3, 8, 300000), (4, 24, 600000)])
def test_fullsize_digest_and_sampled_buckets(gpu_ctx, config, n_extra, max_rows):
    """configs[3] (2M hypermutated) and configs[4] (10M): full-size digest + live oracle on whole buckets incl. the largest."""
    rep = E.generate(config)
    p = E.default_params(True)
    inp = rep.to_join_input()
    rep.close()
    res = gpu_ctx.join(p, inp.to_packed())
    _sampled_buckets_vs_oracle(inp, p, res, n_extra, seed=config, max_rows=max_rows)
    _check_digest(f"config{config}", res, inp.n_rows)


def test_stage1_work_savers_at_8m_cells(gpu_ctx):
    """The regime where stage 1's donor / reference-set kills reach ~1e8 pairs (configs[1] law at 8M cells, 4 donors):
    same edges, tuples and EquivRel with the work savers switched off (EJB_NO_TAGS)."""
    rep = E.generate(4, n_cells=8000000, seed=20261019 + 104)
    p = E.default_params(True)
    inp = rep.to_join_input().to_packed()
    rep.close()
    on = gpu_ctx.join(p, inp).detach()
    assert on.stats["stage1_donor_kills"] + on.stats["stage1_ref_kills"] + on.stats["stage1_flag_skips"] > 10 ** 7
    gpu_ctx.set_option("no_tags", 1)
    try:
        off = gpu_ctx.join(p, inp)
        assert off.stats["stage1_donor_kills"] == 0 and off.stats["stage1_ref_kills"] == 0 and off.stats["stage1_flag_skips"] == 0
        for f in ARRAYS:
            assert np.array_equal(getattr(on, f), getattr(off, f)), f
    finally:
        gpu_ctx.set_option("no_tags", 0)
