"""The C-ABI library: it loads without a GPU, exports every symbol include/enclone_join_b200.h declares,
the ctypes mirrors have the sizes the library was compiled with, and there is no CPU fallback."""
import ctypes as C
import os
import re

import pytest
import torch

from enclone_ranger_b200 import _abi

HEADER = os.path.join(_abi.ROOT, "include", "enclone_join_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ejb_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    L = C.CDLL(_abi.LIB_JOIN)
    names = declared_functions()
    assert {"ejb_create", "ejb_join", "ejb_result_free", "ejb_destroy", "ejb_last_error", "ejb_upload", "ejb_run", "ejb_download"} <= set(names)
    for n in names:
        assert hasattr(L, n), f"{n} is declared in the header but not exported"


def test_struct_sizes_match_the_library():
    L = _abi.load_join()
    out = (C.c_int64 * 4)()
    L.ejb_abi_sizes(out)
    assert list(out) == [C.sizeof(_abi.EjbParams), C.sizeof(_abi.EjbInput), C.sizeof(_abi.EjbStats), C.sizeof(_abi.EjbResult)]
    assert L.ejb_abi_version() == 4


def test_default_params_match_proc_args():
    # enclone_args/src/proc_args.rs:36-41,152-163
    L = _abi.load_join()
    p = _abi.EjbParams()
    L.ejb_params_default(C.byref(p), 1)
    from enclone_ranger_b200 import default_params
    q = default_params(True)
    for name, _ in _abi.EjbParams._fields_:
        assert getattr(p, name) == getattr(q, name), name
    assert (p.max_score, p.cdr3_mult, p.mult_pow, p.join_cdr3_ident, p.fwr1_cdr12_delta) == (100000.0, 5.0, 80.0, 85.0, 20.0)
    assert (p.max_cdr3_diffs, p.cdr3_normal_len, p.auto_share, p.comp_filt, p.comp_filt_bound) == (1000, 42, 15, 8, 80)
    assert (p.max_diffs, p.max_degradation, p.ref_v_trim, p.ref_j_trim) == (1000000, 2, 15, 15)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback_without_a_device():
    import enclone_ranger_b200 as E
    with pytest.raises(E.JoinError) as ei:
        E.JoinContext()
    assert ei.value.code == _abi.EJB_ERR_NO_DEVICE


def test_product_never_imports_the_oracle():
    """Only tests/, smoke() and bench.py's baseline legs may touch oracle/ — the package must not."""
    pkg = os.path.join(_abi.ROOT, "enclone_ranger_b200")
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|pyoracle|liboracle|oracle_join|#include\s+\".*oracle", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not pat.search(src), f"{f} references the oracle"


def test_host_equiv_replay_matches_oracle_equivrel():
    import numpy as np
    import enclone_ranger_b200 as E
    from oracle import pyoracle
    rng = np.random.default_rng(0)
    n = 500
    a = rng.integers(0, n, 700).astype(np.int32)
    b = rng.integers(0, n, 700).astype(np.int32)
    eq, lab = E.equiv_replay(n, a, b)
    x, y, z = pyoracle.equiv_replay(n, a, b)
    assert np.array_equal(eq.x, x) and np.array_equal(eq.y, y) and np.array_equal(eq.z, z)
    # labels = smallest member of the orbit
    for i in range(n):
        assert lab[i] == min(eq.orbit(i))
