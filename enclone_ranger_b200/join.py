"""Host-side mirror of the reference's join interface on top of the C ABI.

    reference                                   here
    ---------                                   ----
    join_exacts(to_bc, refdata, ctl,            join_exacts(ctx, params, inp) -> (EquivRel, JoinResult)
                exact_clonotypes, info,             inp    : JoinInput / Repertoire (the flattened
                &mut join_info, &mut raw_joins,              info[] / exact_clonotypes[] / to_bc / refdata)
                sr, dref) -> EquivRel               params : EjbParams (the join knobs of `ctl`)
    (enclone/src/join.rs:31-41)                     result.raw_joins == what the reference appends to raw_joins

The product path is the CUDA library only: constructing a JoinContext fails when
libenclone_join_b200.so is missing or no B200 is visible.  Nothing here imports oracle/.
"""
import ctypes as C

import numpy as np

from . import _abi
from ._abi import EjbResult, EjbStats


class JoinError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"ejb error {code}: {msg}")
        self.code = code


class EquivRel:
    """The reference's union-find as raw x/y/z vectors (equiv/src/lib.rs:34-138), read-only view
    with the accessors downstream code uses (class_id, orbit_reps, orbit, orbit_size, norbits)."""

    def __init__(self, x, y, z):
        self.x, self.y, self.z = x, y, z

    def class_id(self, a):
        return int(self.y[a])

    def orbit_size(self, a):
        return int(self.z[self.y[a]])

    def orbit_reps(self):
        return np.nonzero(self.y == np.arange(len(self.y), dtype=self.y.dtype))[0].astype(np.int32)

    def norbits(self):
        return int(len(self.orbit_reps()))

    def orbit(self, a):
        o = [int(a)]
        b = int(self.x[a])
        while b != a:
            o.append(b)
            b = int(self.x[b])
        return o


class JoinResult:
    """Edges (raw_joins order), their PotentialJoin numerics, labels and the EquivRel vectors.

    With copy=False the arrays are zero-copy views into the context's pinned result arena: valid until the next
    join()/download() on the same context (the C ABI's ownership rule); call .detach() to keep them longer."""

    def __init__(self, r, copy=True):
        E, N = r.n_edges, r.n_rows

        def arr(p, n, dt):
            if n == 0:
                return np.zeros(0, dtype=dt)
            v = np.ctypeslib.as_array(p, shape=(n,))
            return v.copy() if copy else v

        self.edge_k1 = arr(r.edge_k1, E, np.int32)
        self.edge_k2 = arr(r.edge_k2, E, np.int32)
        self.edge_cd = arr(r.edge_cd, E, np.int32)
        self.edge_nrefs = arr(r.edge_nrefs, E, np.int32)
        self.edge_shares = arr(r.edge_shares, 2 * E, np.int32).reshape(-1, 2)
        self.edge_indeps = arr(r.edge_indeps, 2 * E, np.int32).reshape(-1, 2)
        self.edge_k = arr(r.edge_k, E, np.int32)
        self.edge_d = arr(r.edge_d, E, np.int32)
        self.edge_n = arr(r.edge_n, E, np.int32)
        self.edge_p1 = arr(r.edge_p1, E, np.float64)
        self.edge_mult = arr(r.edge_mult, E, np.float64)
        self.edge_score = arr(r.edge_score, E, np.float64)
        self.edge_err = arr(r.edge_err, E, np.uint8)
        self.labels = arr(r.labels, N, np.int32)
        self.eq_x = arr(r.eq_x, N, np.int32)
        self.eq_y = arr(r.eq_y, N, np.int32)
        self.eq_z = arr(r.eq_z, N, np.int32)
        self.stats = r.stats.as_dict()

    def detach(self):
        for k, v in list(self.__dict__.items()):
            if isinstance(v, np.ndarray):
                setattr(self, k, v.copy())
        return self

    @property
    def raw_joins(self):
        return np.stack([self.edge_k1, self.edge_k2], axis=1)

    def equiv(self):
        return EquivRel(self.eq_x, self.eq_y, self.eq_z)


def _c_input(inp):
    if hasattr(inp, "c_ptr") and callable(inp.c_ptr):
        return inp.c_ptr()
    if hasattr(inp, "c_ptr"):
        return inp.c_ptr
    return inp


class JoinContext:
    """Owns one ejb_ctx.  `device` is one GPU index or a list of them (ejb_create with n_devices > 1: the library shards the
    buckets over the GPUs itself).  `rank`/`world` shard buckets across processes (one per GPU)."""

    def __init__(self, device=0, rank=0, world=1, cand_cap=0, pair_budget=0):
        self.L = _abi.load_join()
        devices = [int(d) for d in device] if isinstance(device, (list, tuple)) else [int(device)]
        self.devices = devices
        self.device = devices[0]
        self.h = C.c_void_p()
        dev = (C.c_int * len(devices))(*devices)
        rc = self.L.ejb_create(C.byref(self.h), dev, len(devices))
        if rc != 0:
            raise JoinError(rc, "ejb_create failed (no CUDA device? there is no CPU fallback)")
        self.L.ejb_set_shard.argtypes = [C.c_void_p, C.c_int, C.c_int]
        self.L.ejb_set_capacity.argtypes = [C.c_void_p, C.c_ulonglong, C.c_ulonglong]
        if world > 1:
            self._ck(self.L.ejb_set_shard(self.h, int(rank), int(world)))
        if cand_cap or pair_budget:
            self._ck(self.L.ejb_set_capacity(self.h, int(cand_cap), int(pair_budget)))

    def _ck(self, rc):
        if rc != 0:
            raise JoinError(rc, self.L.ejb_last_error(self.h).decode())

    def close(self):
        if self.h:
            self.L.ejb_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, name, value):
        """Developer / test knob (ejb_set_option): "first_unit", "growth", "s1_mode", "no_tags", "trace"."""
        self._ck(self.L.ejb_set_option(self.h, name.encode(), int(value)))

    def create_ms(self):
        """Milliseconds ejb_create took (the Stirling ratio table is built on the device there)."""
        return float(self.L.ejb_create_ms(self.h))

    def sr_table_read(self, first, count):
        """(hi, lo) pairs of entries [first, first+count) of the device-built Stirling ratio table (triangular layout)."""
        out = np.zeros(2 * int(count), dtype=np.float64)
        self._ck(self.L.ejb_sr_table_read(self.h, int(first), int(count), out.ctypes.data_as(C.POINTER(C.c_double))))
        return out.reshape(-1, 2)

    # --- the drop-in call ---
    def join(self, params, inp, copy=True):
        r = EjbResult()
        self._ck(self.L.ejb_join(self.h, C.byref(params), _c_input(inp), C.byref(r)))
        out = JoinResult(r, copy=copy)
        self.L.ejb_result_free(self.h, C.byref(r))
        return out

    # --- the same in three steps (device-resident timing) ---
    def upload(self, params, inp):
        self._ck(self.L.ejb_upload(self.h, C.byref(params), _c_input(inp)))

    def run(self):
        self._ck(self.L.ejb_run(self.h))
        s = EjbStats()
        self.L.ejb_last_stats(self.h, C.byref(s))
        return s.as_dict()

    def download(self, copy=True):
        r = EjbResult()
        self._ck(self.L.ejb_download(self.h, C.byref(r)))
        out = JoinResult(r, copy=copy)
        self.L.ejb_result_free(self.h, C.byref(r))
        return out

    def set_input_arena(self, address, nbytes):
        """Declare one (page-locked) host allocation that holds the input arrays of the following uploads: it is copied
        with a single DMA (ejb_set_input_arena).  address=None clears."""
        self._ck(self.L.ejb_set_input_arena(self.h, C.c_void_p(address) if address else None, int(nbytes) if address else 0))

    def set_row_roles(self, roles):
        """Row roles of a sub-bucket split over several contexts (ejb_set_row_roles): uint8 per row of the input the next
        runs work on (bit 0 column-only, bit 1 split sub-bucket); None clears them."""
        if roles is None:
            self._ck(self.L.ejb_set_row_roles(self.h, None, 0))
            return
        r = np.ascontiguousarray(roles, dtype=np.uint8)
        self._ck(self.L.ejb_set_row_roles(self.h, r.ctypes.data_as(C.POINTER(C.c_uint8)), r.size))

    # --- f2: the pair predicate for arbitrary pairs (define_mat) ---
    def join_one_batch(self, k1, k2):
        """join_one(k1[i], k2[i]) for every pair of rows of the input of the last join()/run() on this context
        (ejb_join_one_batch): dict with `passed` and the PotentialJoin members."""
        k1 = np.ascontiguousarray(k1, dtype=np.int32)
        k2 = np.ascontiguousarray(k2, dtype=np.int32)
        n = len(k1)
        out = dict(passed=np.zeros(n, np.uint8), cd=np.zeros(n, np.int32), nrefs=np.zeros(n, np.int32), shares=np.zeros(2 * n, np.int32),
                   indeps=np.zeros(2 * n, np.int32), k=np.zeros(n, np.int32), d=np.zeros(n, np.int32), n=np.zeros(n, np.int32),
                   p1=np.zeros(n, np.float64), mult=np.zeros(n, np.float64), score=np.zeros(n, np.float64), err=np.zeros(n, np.uint8))
        ct = {np.dtype(np.uint8): C.c_uint8, np.dtype(np.int32): C.c_int32, np.dtype(np.float64): C.c_double}
        po = _abi.EjbPairOut(**{f: out[f].ctypes.data_as(C.POINTER(ct[out[f].dtype])) for f in out})
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
        self._ck(self.L.ejb_join_one_batch(self.h, ip(k1), ip(k2), n, C.byref(po)))
        out["shares"] = out["shares"].reshape(-1, 2)
        out["indeps"] = out["indeps"].reshape(-1, 2)
        return out

    # --- sharded join: every context holds the whole input, joins its share, one context merges (multi-GPU) ---
    def set_shard(self, rank, world):
        self._ck(self.L.ejb_set_shard(self.h, int(rank), int(world)))

    def _dev_tensor(self, ptr, n, typestr):
        import torch

        class _Dev:
            def __init__(self, ptr, n):
                self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}
        return torch.as_tensor(_Dev(ptr, n), device=f"cuda:{self.device}")

    def upload_slice(self, params, inp, part, nparts):
        """Sharded upload (ejb_upload_slice): only slice `part` of the declared input arena crosses PCIe; complete the device
        arena with an all-gather on device_arena() before run()."""
        self._ck(self.L.ejb_upload_slice(self.h, C.byref(params), _c_input(inp), int(part), int(nparts)))

    def device_arena(self):
        """(uint8 CUDA tensor over the whole device arena, slice_bytes): slice i lives at [i * slice_bytes, (i + 1) * slice_bytes)."""
        p, total, sl = C.c_void_p(), C.c_uint64(), C.c_uint64()
        self._ck(self.L.ejb_device_arena(self.h, C.byref(p), C.byref(total), C.byref(sl)))
        return self._dev_tensor(p.value, int(total.value), "|u1"), int(sl.value)

    def part_forest(self):
        """Partial forest of the last sharded run(): (n, edges int64[n] = (k1 << 28) | k2 with GLOBAL row ids, tuples uint8[24 n])
        as zero-copy torch CUDA tensors (valid until the next run)."""
        import torch
        n, e, t = C.c_int64(), C.c_void_p(), C.c_void_p()
        self._ck(self.L.ejb_part_forest(self.h, C.byref(n), C.byref(e), C.byref(t)))
        if n.value == 0:
            dev = f"cuda:{self.device}"
            return 0, torch.zeros(0, dtype=torch.int64, device=dev), torch.zeros(0, dtype=torch.uint8, device=dev)
        return n.value, self._dev_tensor(e.value, n.value, "<i8"), self._dev_tensor(t.value, 32 * n.value, "|u1")

    def merge_buffers(self, total):
        """Device buffers of the merging context for the concatenated partial forests (edges int64[total], tuples uint8[24 total])."""
        import torch
        e, t = C.c_void_p(), C.c_void_p()
        self._ck(self.L.ejb_merge_buffers(self.h, int(total), C.byref(e), C.byref(t)))
        if total == 0:
            dev = f"cuda:{self.device}"
            return torch.zeros(0, dtype=torch.int64, device=dev), torch.zeros(0, dtype=torch.uint8, device=dev)
        return self._dev_tensor(e.value, int(total), "<i8"), self._dev_tensor(t.value, 32 * int(total), "|u1")

    def merge_parts(self, total):
        """Forest of the gathered partial forests + finalize on this context; download() then returns the WHOLE join."""
        self._ck(self.L.ejb_merge_parts(self.h, int(total)))
        s = EjbStats()
        self.L.ejb_last_stats(self.h, C.byref(s))
        return s.as_dict()

    def device_tuples(self):
        """(cd, nrefs, shares[2n]) of the last run()'s edges as zero-copy torch CUDA tensors."""
        import torch
        n = C.c_int64()
        self.L.ejb_device_edges.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p]
        self._ck(self.L.ejb_device_edges(self.h, C.byref(n), None, None, None))
        dev = f"cuda:{self.device}"
        if n.value == 0:
            e = torch.zeros(0, dtype=torch.int32, device=dev)
            return e, e.clone(), e.clone()
        cd, nr, sh = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._ck(self.L.ejb_device_tuples(self.h, C.byref(cd), C.byref(nr), C.byref(sh)))

        class _Dev:
            def __init__(self, ptr, n):
                self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 2}
        return (torch.as_tensor(_Dev(cd.value, n.value), device=dev), torch.as_tensor(_Dev(nr.value, n.value), device=dev),
                torch.as_tensor(_Dev(sh.value, 2 * n.value), device=dev))

    def device_edges(self):
        """(k1, k2) of the last run() as zero-copy torch CUDA tensors (valid until the next run())."""
        import torch
        n = C.c_int64()
        k1, k2 = C.c_void_p(), C.c_void_p()
        self.L.ejb_device_edges.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p]
        self._ck(self.L.ejb_device_edges(self.h, C.byref(n), C.byref(k1), C.byref(k2), None))
        if n.value == 0:
            e = torch.zeros(0, dtype=torch.int32, device=f"cuda:{self.device}")
            return e, e.clone()

        class _Dev:
            def __init__(self, ptr, n):
                self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 2}
        dev = f"cuda:{self.device}"
        return torch.as_tensor(_Dev(k1.value, n.value), device=dev), torch.as_tensor(_Dev(k2.value, n.value), device=dev)

    def p1_batch(self, k, d, n):
        k = np.ascontiguousarray(k, dtype=np.int32)
        d = np.ascontiguousarray(d, dtype=np.int32)
        n = np.ascontiguousarray(n, dtype=np.int32)
        out = np.zeros(len(k), dtype=np.float64)
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
        self._ck(self.L.ejb_p1_batch(self.h, ip(k), ip(d), ip(n), len(k), out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def measure_peaks(self):
        a, b, m = C.c_double(), C.c_double(), C.c_double()
        self._ck(self.L.ejb_measure_peaks(self.h, C.byref(a), C.byref(b), C.byref(m)))
        return {"popc_per_s": a.value, "lop3_per_s": b.value, "sm_clock_mhz": m.value}


def join_exacts(ctx, params, inp):
    """Drop-in for enclone::join::join_exacts: returns (EquivRel, JoinResult)."""
    res = ctx.join(params, inp)
    return res.equiv(), res


def forest_merge(n_rows, k1, k2):
    """Kruskal over edges sorted by (k1, k2): boolean mask of the edges that join two components (ejb_forest_merge)."""
    L = _abi.load_join()
    k1 = np.ascontiguousarray(k1, dtype=np.int32)
    k2 = np.ascontiguousarray(k2, dtype=np.int32)
    keep = np.zeros(len(k1), dtype=np.uint8)
    if len(k1):
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
        rc = L.ejb_forest_merge(int(n_rows), ip(k1), ip(k2), len(k1), keep.ctypes.data_as(C.POINTER(C.c_uint8)))
        if rc != 0:
            raise JoinError(rc, "ejb_forest_merge (edges must be sorted by (k1, k2))")
    return keep.astype(bool)


def equiv_replay(n_rows, k1, k2, row_exact=None):
    """Replay an ordered edge list into the reference's EquivRel (host code in the library)."""
    L = _abi.load_join()
    k1 = np.ascontiguousarray(k1, dtype=np.int32)
    k2 = np.ascontiguousarray(k2, dtype=np.int32)
    x = np.zeros(n_rows, np.int32); y = np.zeros(n_rows, np.int32); z = np.zeros(n_rows, np.int32)
    lab = np.zeros(n_rows, np.int32)
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    re = None
    if row_exact is not None:
        row_exact = np.ascontiguousarray(row_exact, dtype=np.int32)
        re = ip(row_exact)
    L.ejb_equiv_replay.argtypes = [C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int64,
                                   C.POINTER(C.c_int32)] + [C.POINTER(C.c_int32)] * 4
    rc = L.ejb_equiv_replay(int(n_rows), ip(k1), ip(k2), len(k1), re, ip(x), ip(y), ip(z), ip(lab))
    if rc != 0:
        raise JoinError(rc, "ejb_equiv_replay")
    return EquivRel(x, y, z), lab
