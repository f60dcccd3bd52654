"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes front-end of oracle/liboracle.so.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from enclone_ranger_b200._abi import EjbInput, EjbParams  # noqa: E402  (struct mirrors of the shared C header only)

_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)
LIB_ORACLE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "liboracle.so")
_cache = {}


class OracleResult(C.Structure):
    _fields_ = [
        ("n_edges", C.c_int64),
        ("edge_k1", _i32p),
        ("edge_k2", _i32p),
        ("edge_cd", _i32p),
        ("edge_nrefs", _i32p),
        ("edge_shares", _i32p),
        ("edge_indeps", _i32p),
        ("edge_k", _i32p),
        ("edge_d", _i32p),
        ("edge_n", _i32p),
        ("edge_p1", _f64p),
        ("edge_mult", _f64p),
        ("edge_score", _f64p),
        ("edge_err", _u8p),
        ("n_rows", C.c_int64),
        ("eq_x", _i32p),
        ("eq_y", _i32p),
        ("eq_z", _i32p),
        ("labels", _i32p),
        ("n_buckets", C.c_int64),
        ("pairs_lens", C.c_int64),
        ("pairs_scored", C.c_int64),
        ("join_one_calls", C.c_int64),
        ("pot_found", C.c_int64),
        ("two_cell_deleted", C.c_int64),
        ("joins", C.c_int64),
        ("errors", C.c_int64),
        ("seconds_join", C.c_double),
        ("threads", C.c_int32),
        ("error", C.c_char * 256),
    ]



def load_oracle():
    """Load the CPU oracle (test infrastructure; only tests / smoke / bench baselines call this)."""
    if "oracle" not in _cache:
        if not os.path.exists(LIB_ORACLE):
            raise RuntimeError(f"{LIB_ORACLE} is missing — run `make -C oracle`")
        L = C.CDLL(LIB_ORACLE)
        L.oracle_join.argtypes = [C.POINTER(EjbParams), C.POINTER(EjbInput), C.c_int, C.POINTER(OracleResult)]
        L.oracle_join.restype = C.c_int
        L.oracle_result_free.argtypes = [C.POINTER(OracleResult)]
        L.oracle_result_free.restype = None
        L.oracle_sr_entry.argtypes = [C.c_int, C.c_int, _f64p, _f64p]
        L.oracle_sr_entry.restype = None
        L.oracle_p1.argtypes = [C.c_int64, C.c_int64, C.c_int64, _f64p]
        L.oracle_p1.restype = C.c_int
        L.oracle_dd_ops.argtypes = [C.c_double] * 4 + [_f64p]
        L.oracle_dd_ops.restype = None
        L.oracle_equiv_replay.argtypes = [C.c_int32, _i32p, _i32p, C.c_int64, _i32p, _i32p, _i32p]
        L.oracle_equiv_replay.restype = None
        _cache["oracle"] = L
    return _cache["oracle"]



def _np(ptr, n, dt):
    if n == 0:
        return np.zeros(0, dtype=dt)
    return np.ctypeslib.as_array(ptr, shape=(n,)).copy()


def run(params, c_input_ptr, threads=0):
    """Run the CPU restatement of join_exacts.  Returns a dict of numpy arrays + counters."""
    L = load_oracle()
    r = OracleResult()
    rc = L.oracle_join(C.byref(params), c_input_ptr, int(threads), C.byref(r))
    if rc != 0:
        msg = r.error.decode()
        raise RuntimeError(f"oracle_join failed ({rc}): {msg}")
    E, N = r.n_edges, r.n_rows
    out = dict(
        edge_k1=_np(r.edge_k1, E, np.int32), edge_k2=_np(r.edge_k2, E, np.int32),
        edge_cd=_np(r.edge_cd, E, np.int32), edge_nrefs=_np(r.edge_nrefs, E, np.int32),
        edge_shares=_np(r.edge_shares, 2 * E, np.int32).reshape(-1, 2),
        edge_indeps=_np(r.edge_indeps, 2 * E, np.int32).reshape(-1, 2),
        edge_k=_np(r.edge_k, E, np.int32), edge_d=_np(r.edge_d, E, np.int32), edge_n=_np(r.edge_n, E, np.int32),
        edge_p1=_np(r.edge_p1, E, np.float64), edge_mult=_np(r.edge_mult, E, np.float64),
        edge_score=_np(r.edge_score, E, np.float64), edge_err=_np(r.edge_err, E, np.uint8),
        eq_x=_np(r.eq_x, N, np.int32), eq_y=_np(r.eq_y, N, np.int32), eq_z=_np(r.eq_z, N, np.int32),
        labels=_np(r.labels, N, np.int32),
        n_buckets=r.n_buckets, pairs_lens=r.pairs_lens, pairs_scored=r.pairs_scored,
        join_one_calls=r.join_one_calls, pot_found=r.pot_found, two_cell_deleted=r.two_cell_deleted,
        joins=r.joins, errors=r.errors, seconds_join=r.seconds_join, threads=r.threads,
    )
    L.oracle_result_free(C.byref(r))
    return out


def join_one_batch(params, c_input_ptr, k1, k2):
    """join_one for arbitrary (k1, k2) pairs (define_mat's use): dict with pass + the PotentialJoin members."""
    L = load_oracle()
    k1 = np.ascontiguousarray(k1, dtype=np.int32)
    k2 = np.ascontiguousarray(k2, dtype=np.int32)
    n = len(k1)
    out = dict(passed=np.zeros(n, np.uint8), cd=np.zeros(n, np.int32), nrefs=np.zeros(n, np.int32), shares=np.zeros(2 * n, np.int32),
               indeps=np.zeros(2 * n, np.int32), k=np.zeros(n, np.int32), d=np.zeros(n, np.int32), n=np.zeros(n, np.int32),
               p1=np.zeros(n, np.float64), mult=np.zeros(n, np.float64), score=np.zeros(n, np.float64), err=np.zeros(n, np.uint8))
    ct = {np.dtype(np.uint8): C.c_uint8, np.dtype(np.int32): C.c_int32, np.dtype(np.float64): C.c_double}
    ptr = lambda a: a.ctypes.data_as(C.POINTER(ct[a.dtype]))
    L.oracle_join_one_batch.restype = C.c_int
    rc = L.oracle_join_one_batch(C.byref(params), c_input_ptr, ptr(k1), ptr(k2), C.c_int64(n), ptr(out["passed"]), ptr(out["cd"]), ptr(out["nrefs"]),
                                 ptr(out["shares"]), ptr(out["indeps"]), ptr(out["k"]), ptr(out["d"]), ptr(out["n"]), ptr(out["p1"]), ptr(out["mult"]),
                                 ptr(out["score"]), ptr(out["err"]))
    if rc != 0:
        raise RuntimeError(f"oracle_join_one_batch failed ({rc}): the reference would panic on one of the pairs")
    out["shares"] = out["shares"].reshape(-1, 2)
    out["indeps"] = out["indeps"].reshape(-1, 2)
    return out


def p1(k, d, n):
    L = load_oracle()
    v = C.c_double()
    rc = L.oracle_p1(int(k), int(d), int(n), C.byref(v))
    if rc != 0:
        raise ValueError("out of range")
    return v.value


def sr_entry(n, k):
    L = load_oracle()
    hi, lo = C.c_double(), C.c_double()
    L.oracle_sr_entry(int(n), int(k), C.byref(hi), C.byref(lo))
    return hi.value, lo.value


def dd_ops(a, b):
    L = load_oracle()
    out = (C.c_double * 8)()
    L.oracle_dd_ops(a[0], a[1], b[0], b[1], out)
    return [(out[2 * i], out[2 * i + 1]) for i in range(4)]


def equiv_replay(n, a, b):
    L = load_oracle()
    a = np.ascontiguousarray(a, dtype=np.int32)
    b = np.ascontiguousarray(b, dtype=np.int32)
    x = np.zeros(n, np.int32); y = np.zeros(n, np.int32); z = np.zeros(n, np.int32)
    p = lambda v: v.ctypes.data_as(C.POINTER(C.c_int32))
    L.oracle_equiv_replay(n, p(a), p(b), len(a), p(x), p(y), p(z))
    return x, y, z
