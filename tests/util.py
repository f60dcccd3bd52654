import os

import numpy as np

import enclone_ranger_b200 as E

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TUPLE_INT = ("cd", "nrefs", "shares", "indeps", "k", "d", "n", "err")
TUPLE_F64 = ("p1", "mult", "score")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name))
    inp = E.JoinInput(**{k[3:]: z[k] for k in z.files if k.startswith("in_")})
    ora = {k[3:]: z[k] for k in z.files if k.startswith("or_")}
    return inp, ora, E.default_params(bool(z["is_bcr"][0]))


def assert_same_result(got, want, what=""):
    """got: JoinResult or dict; want: oracle dict.  Integer members bit-exact; fp64 members bit-exact too
    (both sides run the same double-double operation sequence), which implies the 1e-12 bar."""
    g = (lambda f: np.asarray(got[f])) if isinstance(got, dict) else (lambda f: np.asarray(getattr(got, f)))
    assert g("edge_k1").tolist() == want["edge_k1"].tolist(), f"{what}: edge k1"
    assert g("edge_k2").tolist() == want["edge_k2"].tolist(), f"{what}: edge k2"
    for f in TUPLE_INT:
        assert np.array_equal(g("edge_" + f), want["edge_" + f]), f"{what}: edge_{f}"
    for f in TUPLE_F64:
        a, b = g("edge_" + f), want["edge_" + f]
        rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
        assert np.all((a == b) | (rel <= 1e-12)), f"{what}: edge_{f} max rel {rel.max() if len(rel) else 0}"
        assert np.array_equal(a, b), f"{what}: edge_{f} not bit-exact"
    for f in ("labels", "eq_x", "eq_y", "eq_z"):
        if f in want and (not isinstance(got, dict) or f in got):
            assert np.array_equal(g(f), want[f]), f"{what}: {f}"
