"""Host-side containers for the flattened join input (include/enclone_join_b200.h: ejb_input).

JoinInput is numpy-backed so that tests can build micro-cases by hand; it mirrors, field for field,
what the Rust shim (rust/enclone_join_b200) flattens from info[] / exact_clonotypes[] / to_bc /
refdata before calling ejb_join.
"""
import ctypes as C

import numpy as np

from ._abi import EjbInput, SynthConfig, SynthSummary, load_synth

_FIELDS = {
    # name: (dtype, count attribute or None)
    "row_nchains": np.int32, "row_len": np.int32, "row_tig_off": np.int64, "row_has_del": np.uint8,
    "row_vs": np.int32, "row_js": np.int32, "row_cdr3_off": np.int64, "row_cdr3_len": np.int32,
    "row_exact": np.int32, "row_exact_col": np.int32, "tig_bytes": np.uint8, "cdr3_bytes": np.uint8,
    "ex_chain_off": np.int64, "ex_cell_off": np.int64,
    "ch_left": np.uint8, "ch_c_ref": np.int32, "ch_v_ref": np.int32, "ch_u_ref": np.int32,
    "ch_fr1_start": np.int32, "ch_cdr1_start": np.int32, "ch_fr2_start": np.int32,
    "ch_cdr2_start": np.int32, "ch_fr3_start": np.int32, "ch_hcomp": np.int32,
    "ch_amino_off": np.int64, "ch_amino_len": np.int32, "amino_bytes": np.uint8,
    "cell_donor": np.int32, "cell_barcode": np.uint32,
    "seg_off": np.int64, "seg_bytes": np.uint8, "ref_off": np.int64, "ref_bytes": np.uint8,
    "ref_name_id": np.int32,
    # optional 2-bit packed contigs / CDR3s
    "row_tig2_off": np.int64, "tig2_words": np.uint32,
    "row_cdr32_off": np.int64, "cdr32_words": np.uint32,
}


def _ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


_CT = {np.int32: C.c_int32, np.int64: C.c_int64, np.uint8: C.c_uint8, np.uint32: C.c_uint32}


class JoinInput:
    """Owns numpy arrays for every ejb_input field; `.c` is the matching C struct."""

    def __init__(self, **arrays):
        self.a = {}
        for name, dt in _FIELDS.items():
            v = np.ascontiguousarray(np.asarray(arrays.get(name, []), dtype=dt))
            if v.size == 0:
                v = np.zeros(1, dtype=dt)[:0].copy()
            self.a[name] = v
        a = self.a
        self.n_rows = int(a["row_nchains"].size)
        self.n_exacts = max(0, int(a["ex_chain_off"].size) - 1)
        self.n_chains = int(a["ch_left"].size)
        self.n_cells = int(a["cell_barcode"].size)
        self.n_segs = max(0, int(a["seg_off"].size) - 1)
        self.n_refs = max(0, int(a["ref_off"].size) - 1)
        # keep 1-element backing stores alive for empty arrays so that pointers are non-NULL
        self._keep = {n: (v if v.size else np.zeros(1, dtype=v.dtype)) for n, v in a.items()}
        c = EjbInput()
        for name, dt in _FIELDS.items():
            setattr(c, name, _ptr(self._keep[name], _CT[dt]))
        c.n_rows, c.n_exacts, c.n_chains = self.n_rows, self.n_exacts, self.n_chains
        c.n_cells, c.n_segs, c.n_refs = self.n_cells, self.n_segs, self.n_refs
        c.n_tig2_words = int(a["tig2_words"].size)
        if a["row_tig2_off"].size == 0:
            c.row_tig2_off = None          # NULL: every contig is ASCII
        c.n_cdr32_words = int(a["cdr32_words"].size)
        if a["row_cdr32_off"].size == 0:
            c.row_cdr32_off = None         # NULL: every CDR3 is ASCII
        c.n_tig_bytes = int(a["tig_bytes"].size)
        c.n_cdr3_bytes = int(a["cdr3_bytes"].size)
        c.n_amino_bytes = int(a["amino_bytes"].size)
        self.c = c

    def c_ptr(self):
        return C.pointer(self.c)

    @staticmethod
    def from_c(cin, copy=True):
        """Build a JoinInput from an ejb_input struct (e.g. the synthetic generator's)."""
        def arr(ptr, n, dt):
            if n == 0:
                return np.zeros(0, dtype=dt)
            v = np.ctypeslib.as_array(ptr, shape=(n,))
            return v.copy() if copy else v
        N, E, CH, X, S, R = cin.n_rows, cin.n_exacts, cin.n_chains, cin.n_cells, cin.n_segs, cin.n_refs
        sizes = {
            "row_nchains": N, "row_len": 2 * N, "row_tig_off": 2 * N, "row_has_del": 2 * N, "row_vs": 2 * N,
            "row_js": 2 * N, "row_cdr3_off": 2 * N, "row_cdr3_len": 2 * N, "row_exact": N,
            "row_exact_col": 2 * N, "tig_bytes": cin.n_tig_bytes, "cdr3_bytes": cin.n_cdr3_bytes,
            "ex_chain_off": E + 1, "ex_cell_off": E + 1, "ch_left": CH, "ch_c_ref": CH, "ch_v_ref": CH,
            "ch_u_ref": CH, "ch_fr1_start": CH, "ch_cdr1_start": CH, "ch_fr2_start": CH,
            "ch_cdr2_start": CH, "ch_fr3_start": CH, "ch_hcomp": CH, "ch_amino_off": CH,
            "ch_amino_len": CH, "amino_bytes": cin.n_amino_bytes, "cell_donor": X, "cell_barcode": X,
            "seg_off": S + 1, "seg_bytes": 0, "ref_off": R + 1, "ref_bytes": 0, "ref_name_id": R,
            "row_tig2_off": 2 * N if cin.row_tig2_off else 0, "tig2_words": cin.n_tig2_words,
            "row_cdr32_off": 2 * N if cin.row_cdr32_off else 0, "cdr32_words": cin.n_cdr32_words if cin.row_cdr32_off else 0,
        }
        out = {}
        for name, dt in _FIELDS.items():
            if name in ("seg_bytes", "ref_bytes"):
                continue
            out[name] = arr(getattr(cin, name), sizes[name], dt)
        out["seg_bytes"] = arr(cin.seg_bytes, int(out["seg_off"][-1]) if S else 0, np.uint8)
        out["ref_bytes"] = arr(cin.ref_bytes, int(out["ref_off"][-1]) if R else 0, np.uint8)
        return JoinInput(**out)

    def take_rows(self, idx):
        """A new JoinInput with only rows `idx` (in that order); other tables are kept whole."""
        idx = np.asarray(idx, dtype=np.int64)
        two = np.stack([2 * idx, 2 * idx + 1], axis=1).reshape(-1)
        b = dict(self.a)
        for n in ("row_nchains", "row_exact"):
            b[n] = self.a[n][idx]
        for n in ("row_len", "row_tig_off", "row_has_del", "row_vs", "row_js", "row_cdr3_off",
                  "row_cdr3_len", "row_exact_col", "row_tig2_off", "row_cdr32_off"):
            if self.a[n].size:
                b[n] = self.a[n][two]
        return JoinInput(**b)

    def subset_rows(self, idx):
        assert self.a["row_tig2_off"].size == 0 and self.a["row_cdr32_off"].size == 0, "subset the ASCII form, then call to_packed()"
        """Rows `idx` (ascending) with the exact / chain / cell tables compacted to what those rows reference.
        Seg and ref tables are kept whole (they are small).  Returns (JoinInput, old exact id of each new exact)."""
        idx = np.asarray(idx, dtype=np.int64)
        a = self.a
        two = np.stack([2 * idx, 2 * idx + 1], axis=1).reshape(-1)
        b = dict(a)
        for n in ("row_nchains",):
            b[n] = a[n][idx]
        for n in ("row_len", "row_has_del", "row_vs", "row_js", "row_cdr3_len", "row_exact_col"):
            b[n] = a[n][two]
        # contig / CDR3 bytes: gather the referenced spans
        def gather_spans(off, ln, data, chunk=200000):
            ln = ln.astype(np.int64)
            new_off = np.concatenate([[0], np.cumsum(ln)])[:-1]
            out = []
            for c0 in range(0, len(ln), chunk):    # bounded temporaries (a per-byte index array can be tens of GB)
                l, o = ln[c0:c0 + chunk], off[c0:c0 + chunk]
                t = int(l.sum())
                if t:
                    st = np.cumsum(l) - l
                    out.append(data[np.repeat(o - st, l) + np.arange(t, dtype=np.int64)])
            return new_off.astype(np.int64), (np.concatenate(out) if out else data[:0])
        b["row_tig_off"], b["tig_bytes"] = gather_spans(a["row_tig_off"][two], b["row_len"] * (np.repeat(b["row_nchains"], 2) > np.tile([0, 1], len(idx))), a["tig_bytes"])
        b["row_cdr3_off"], b["cdr3_bytes"] = gather_spans(a["row_cdr3_off"][two], b["row_cdr3_len"] * (np.repeat(b["row_nchains"], 2) > np.tile([0, 1], len(idx))), a["cdr3_bytes"])
        # exacts
        old_ex = a["row_exact"][idx]
        used, inv = np.unique(old_ex, return_inverse=True)
        b["row_exact"] = inv.astype(np.int32)
        def compact(offs, arrays):
            ln = (offs[used + 1] - offs[used]).astype(np.int64)
            new_off = np.concatenate([[0], np.cumsum(ln)])
            total = int(new_off[-1])
            src = np.repeat(offs[used] - new_off[:-1], ln) + np.arange(total, dtype=np.int64)
            return new_off.astype(np.int64), {n: a[n][src] for n in arrays}, src
        b["ex_chain_off"], ch, _ = compact(a["ex_chain_off"], ["ch_left", "ch_c_ref", "ch_v_ref", "ch_u_ref", "ch_fr1_start", "ch_cdr1_start",
                                                                 "ch_fr2_start", "ch_cdr2_start", "ch_fr3_start", "ch_hcomp", "ch_amino_off", "ch_amino_len"])
        b.update(ch)
        b["ex_cell_off"], ce, _ = compact(a["ex_cell_off"], ["cell_donor", "cell_barcode"])
        b.update(ce)
        return JoinInput(**b), used

    def to_packed(self, chunk_rows=100000, native=True):
        """The same input with every eligible contig (pure upper-case ACGT, has_del == 0) supplied 2-bit packed
        (what the Rust shim sends from info[k].tigsp); the other contigs stay ASCII.  Works in row chunks so that the
        temporaries stay small (a per-base index array of a 10M-cell repertoire would not fit in memory)."""
        a = self.a
        if a["row_tig2_off"].size:
            return self
        n2 = 2 * self.n_rows
        if native and self.n_rows:
            try:
                return self._to_packed_native()
            except (RuntimeError, AttributeError, OSError):
                pass   # generator library missing: numpy fallback below
        valid = np.repeat(a["row_nchains"], 2) > np.tile([0, 1], self.n_rows)
        ln_all = np.where(valid, a["row_len"].astype(np.int64), 0)
        off_all = a["row_tig_off"]
        lut = np.full(256, 255, dtype=np.uint8)
        lut[ord("A")], lut[ord("C")], lut[ord("G")], lut[ord("T")] = 0, 1, 2, 3
        shifts = (2 * np.arange(16, dtype=np.uint32)).astype(np.uint32)
        tig2_off = np.full(n2, -1, dtype=np.int64)
        new_tig_off = np.full(n2, -1, dtype=np.int64)
        word_chunks, byte_chunks = [], []
        n_words, n_bytes = 0, 0
        step = 2 * int(chunk_rows)
        for c0 in range(0, n2, step):
            c1 = min(n2, c0 + step)
            ln, off = ln_all[c0:c1], off_all[c0:c1]
            total = int(ln.sum())
            if total == 0:
                continue
            start = np.cumsum(ln) - ln
            src = np.repeat(off - start, ln) + np.arange(total, dtype=np.int64)
            flat = a["tig_bytes"][src]
            codes = lut[flat]
            seg = np.repeat(np.arange(c1 - c0), ln)
            bad = np.zeros(c1 - c0, dtype=bool)
            bad[np.unique(seg[codes == 255])] = True
            elig = (ln > 0) & ~bad & (a["row_has_del"][c0:c1] == 0)
            # packed part: pad every eligible contig to a multiple of 16 bases, then fold 16 codes into one word
            nw = np.where(elig, (ln + 15) // 16, 0)
            woff = np.cumsum(nw) - nw
            tw = int(nw.sum())
            if tw:
                padded = np.zeros(tw * 16, dtype=np.uint32)
                keep = elig[seg]
                pos = (np.arange(total, dtype=np.int64) - np.repeat(start, ln))[keep]
                padded[np.repeat(woff * 16, np.where(elig, ln, 0)) + pos] = codes[keep]
                words = np.bitwise_or.reduce(padded.reshape(-1, 16) << shifts, axis=1).astype(np.uint32)
                word_chunks.append(words)
                tig2_off[c0:c1] = np.where(elig, woff + n_words, -1)
                n_words += tw
            # ASCII part
            ln_a = np.where(elig, 0, ln)
            tb = int(ln_a.sum())
            if tb:
                byte_chunks.append(flat[~elig[seg]])
                aoff = np.cumsum(ln_a) - ln_a
                new_tig_off[c0:c1] = np.where(~elig & (ln > 0), aoff + n_bytes, -1)
                n_bytes += tb
        b = dict(a)
        b["row_tig2_off"] = tig2_off
        b["tig2_words"] = np.concatenate(word_chunks) if word_chunks else np.zeros(0, np.uint32)
        b["tig_bytes"] = np.concatenate(byte_chunks) if byte_chunks else np.zeros(0, np.uint8)
        b["row_tig_off"] = new_tig_off
        return JoinInput(**b)

    def _to_packed_native(self):
        lib = load_synth()
        n2 = 2 * self.n_rows
        t2 = np.zeros(n2, dtype=np.int64)
        to = np.zeros(n2, dtype=np.int64)
        words, nbw = C.POINTER(C.c_uint32)(), C.c_int64()
        byts, nbb = C.POINTER(C.c_uint8)(), C.c_int64()
        lib.ejb_synth_pack_contigs.argtypes = [C.POINTER(EjbInput), C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.POINTER(C.c_uint32)),
                                               C.POINTER(C.c_int64), C.POINTER(C.POINTER(C.c_uint8)), C.POINTER(C.c_int64)]
        lib.ejb_synth_pack_contigs.restype = C.c_int
        lib.ejb_synth_free_array.argtypes = [C.c_void_p]
        rc = lib.ejb_synth_pack_contigs(self.c_ptr(), t2.ctypes.data_as(C.POINTER(C.c_int64)), to.ctypes.data_as(C.POINTER(C.c_int64)),
                                        C.byref(words), C.byref(nbw), C.byref(byts), C.byref(nbb))
        if rc != 0:
            raise RuntimeError("ejb_synth_pack_contigs failed")
        b = dict(self.a)
        b["row_tig2_off"] = t2
        b["row_tig_off"] = to
        b["tig2_words"] = np.ctypeslib.as_array(words, shape=(max(nbw.value, 1),))[: nbw.value].copy()
        b["tig_bytes"] = np.ctypeslib.as_array(byts, shape=(max(nbb.value, 1),))[: nbb.value].copy()
        lib.ejb_synth_free_array(words)
        lib.ejb_synth_free_array(byts)
        return JoinInput(**b)

    def to_packed_cdr3(self):
        """The same input with every pure-ACGT CDR3 supplied 2-bit packed (ejb_input.row_cdr32_off / cdr32_words): the ASCII
        cdr3_bytes shrink to the (normally zero) CDR3s that hold another byte."""
        if self.a["row_cdr32_off"].size or self.n_rows == 0:
            return self
        lib = load_synth()
        n2 = 2 * self.n_rows
        c2 = np.zeros(n2, dtype=np.int64)
        co = np.zeros(n2, dtype=np.int64)
        words, nbw = C.POINTER(C.c_uint32)(), C.c_int64()
        byts, nbb = C.POINTER(C.c_uint8)(), C.c_int64()
        lib.ejb_synth_pack_cdr3s.argtypes = [C.POINTER(EjbInput), C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.POINTER(C.c_uint32)),
                                             C.POINTER(C.c_int64), C.POINTER(C.POINTER(C.c_uint8)), C.POINTER(C.c_int64)]
        lib.ejb_synth_pack_cdr3s.restype = C.c_int
        lib.ejb_synth_free_array.argtypes = [C.c_void_p]
        rc = lib.ejb_synth_pack_cdr3s(self.c_ptr(), c2.ctypes.data_as(C.POINTER(C.c_int64)), co.ctypes.data_as(C.POINTER(C.c_int64)),
                                      C.byref(words), C.byref(nbw), C.byref(byts), C.byref(nbb))
        if rc != 0:
            raise RuntimeError("ejb_synth_pack_cdr3s failed")
        b = dict(self.a)
        b["row_cdr32_off"] = c2
        b["row_cdr3_off"] = co
        b["cdr32_words"] = np.ctypeslib.as_array(words, shape=(max(nbw.value, 1),))[: nbw.value].copy()
        b["cdr3_bytes"] = np.ctypeslib.as_array(byts, shape=(max(nbb.value, 1),))[: nbb.value].copy()
        lib.ejb_synth_free_array(words)
        lib.ejb_synth_free_array(byts)
        return JoinInput(**b)

    def nbytes(self):
        return int(sum(v.nbytes for v in self.a.values()))


class Repertoire:
    """A synthetic repertoire produced by csrc/synth.cpp (owns the native buffers)."""

    def __init__(self, handle, lib):
        self._h, self._lib = handle, lib
        self.c_ptr = lib.ejb_synth_input(handle)
        self.c = self.c_ptr.contents
        s = SynthSummary()
        lib.ejb_synth_get_summary(handle, C.byref(s))
        self.summary = s.as_dict()

    def row_clone(self):
        return np.ctypeslib.as_array(self._lib.ejb_synth_row_clone(self._h), shape=(self.c.n_rows,)).copy()

    def to_join_input(self):
        return JoinInput.from_c(self.c)

    def close(self):
        if self._h:
            self._lib.ejb_synth_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def synth_config(config=0, **overrides):
    lib = load_synth()
    cfg = SynthConfig()
    lib.ejb_synth_preset(int(config), C.byref(cfg))
    for k, v in overrides.items():
        if v is None:
            continue
        if not hasattr(cfg, k):
            raise AttributeError(k)
        setattr(cfg, k, v)
    return cfg


def generate(config=0, **overrides):
    """Generate BASELINE.json config `config` (0..4); keyword overrides e.g. n_cells=..., seed=..."""
    lib = load_synth()
    cfg = synth_config(config, **overrides)
    h = C.c_void_p()
    rc = lib.ejb_synth_generate(C.byref(cfg), C.byref(h))
    if rc != 0:
        raise RuntimeError(f"ejb_synth_generate failed: {rc}")
    rep = Repertoire(h, lib)
    rep.config = cfg
    return rep
