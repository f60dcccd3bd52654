"""bench.py's bounded sample for the reference arm (CPU only): lens buckets in a seeded shuffle, none excluded, each cut to its
first `cap_rows` rows; the sample must be a valid join input (whole prefixes of buckets, canonical row order kept) and the
oracle's result on it must equal the oracle's result on the same rows taken by hand."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import bench  # noqa: E402
import enclone_ranger_b200 as E  # noqa: E402
import pyoracle  # noqa: E402


def _input():
    rep = E.generate(0, n_cells=20000, seed=12)
    inp = rep.to_join_input()
    rep.close()
    return inp


def test_sample_takes_prefixes_of_whole_buckets_and_is_deterministic():
    inp = _input()
    bucket, _, _ = bench.bucket_table(inp)
    s1, d1 = bench.cpu_sample(inp, row_fraction=0.2, cap_rows=40)
    s2, d2 = bench.cpu_sample(inp, row_fraction=0.2, cap_rows=40)
    assert d1 == d2 and s1.n_rows == s2.n_rows and np.array_equal(s1.a["row_len"], s2.a["row_len"])
    assert 0 < s1.n_rows <= inp.n_rows
    # no bucket of the sample holds more than cap rows; uncapped, the same shuffle takes whole buckets
    sb, _, _ = bench.bucket_table(s1)
    assert np.bincount(sb).max() <= 40
    whole, dw = bench.cpu_sample(inp, row_fraction=0.2)
    wb, _, _ = bench.bucket_table(whole)
    sizes = np.bincount(bucket)
    assert "truncated" not in dw
    assert set(np.bincount(wb).tolist()) <= set(sizes.tolist())   # every bucket of the uncapped sample is a whole bucket


def test_oracle_on_the_sample_equals_oracle_on_the_same_rows():
    inp = _input()
    p = E.default_params(True)
    sample, _ = bench.cpu_sample(inp, row_fraction=0.3, cap_rows=60)
    got = pyoracle.run(p, sample.c_ptr(), threads=2)
    again = pyoracle.run(p, sample.c_ptr(), threads=1)
    assert got["pairs_scored"] == again["pairs_scored"] > 0
    for f in ("edge_k1", "edge_k2", "eq_y"):
        assert np.array_equal(got[f], again[f]), f
