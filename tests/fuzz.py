"""Random small join inputs exercising rare combinations of join_one's branches (used by the GPU differential test and,
on CPU, to check that the generator itself is sane against the oracle)."""
import numpy as np

import cases
from cases import NONE, World, mutate, rand_seq
from enclone_ranger_b200 import default_params


def random_world(seed):
    rng = np.random.default_rng(seed)
    w = World(seed=1000 + seed)
    b = w.b
    is_bcr = rng.random() < 0.8
    params = dict(is_bcr=is_bcr)
    if rng.random() < 0.15: params["mix_donors"] = 1
    if rng.random() < 0.15: params["old_light"] = 1
    if rng.random() < 0.15: params["easy"] = 1
    if rng.random() < 0.15: params["max_cdr3_diffs"] = int(rng.integers(0, 6))
    if rng.random() < 0.1: params["max_diffs"] = int(rng.integers(3, 40))
    if rng.random() < 0.1: params["comp_filt_bound"] = 0
    # alternative references
    alt_v = [b.seg(mutate(w.vh, rng.choice(280, size=int(k), replace=False).tolist())) for k in (1, 2, 6)]
    alt_j = b.seg(mutate(w.jh, [40]))
    v_dup = b.ref(w.vh, 50)                      # same sequence, different gene name
    v_oth = b.ref(mutate(w.vh, [100]), 51)       # different name and sequence
    v_short = b.ref(w.vh[:250], 52)
    u_oth = b.ref(mutate(w.uh, [59]), 53)
    h0, l0 = w.heavy(), w.light()
    n_clones = int(rng.integers(1, 4))
    for c in range(n_clones):
        # a clone: trunk mutations + per-member private ones; clones share lens, CDR3s differ between clones
        trunk = rng.choice(270, size=int(rng.integers(0, 30)), replace=False).tolist()
        junk = rng.choice(np.arange(w.VH, w.VH + w.NH), size=int(rng.integers(0, 8)) if c else 0, replace=False).tolist()
        hc = mutate(mutate(h0, trunk), junk, shift=int(rng.integers(1, 4)))
        lc = mutate(l0, rng.choice(250, size=int(rng.integers(0, 6)), replace=False).tolist())
        for m in range(int(rng.integers(2, 7))):
            h = mutate(hc, rng.choice(len(h0) - 5, size=int(rng.integers(0, 7 if is_bcr else 3)), replace=False).tolist(), shift=int(rng.integers(1, 4)))
            l = mutate(lc, rng.choice(len(l0) - 5, size=int(rng.integers(0, 3)), replace=False).tolist(), shift=int(rng.integers(1, 4)))
            if is_bcr and rng.random() < 0.5:   # CDR3 differences
                h = mutate(h, rng.choice(np.arange(w.VH - 9, w.VH + w.NH + 15), size=int(rng.integers(1, 7)), replace=False).tolist(), shift=2)
            kw = {}
            r = rng.random()
            has_del = (0, 0)
            if r < 0.12:      # V deletion with '-' (sometimes with a shifted seq_del_amino)
                hb = bytearray(h); p0 = int(rng.integers(40, 250)); dl = 3 * int(rng.integers(1, 3))
                hb[p0:p0 + dl] = b"-" * dl
                if rng.random() < 0.5:
                    am = bytearray(h); sh = int(rng.integers(0, 3)); am[p0 + sh:p0 + sh + dl] = b"-" * dl
                    kw["amino"] = bytes(am)
                h = bytes(hb); has_del = (1, 0)
                if rng.random() < 0.5:
                    kw["vs"] = (b.seg(w.vh[:p0] + w.vh[p0 + dl:]), w.seg_vl)
            elif r < 0.17:    # has_del flag without marks (V insertion removed, info.rs:125)
                has_del = (1, 0)
            elif r < 0.22:    # an odd byte
                hb = bytearray(h); hb[int(rng.integers(0, len(h)))] = ord("N"); h = bytes(hb)
            elif r < 0.27:
                lb = bytearray(l); lb[int(rng.integers(0, len(l)))] = ord("n"); l = bytes(lb)
            if "vs" not in kw and rng.random() < 0.2:
                kw["vs"] = (int(rng.choice(alt_v)), w.seg_vl)
            if rng.random() < 0.08:
                kw["js"] = (alt_j, w.seg_jl)
            if rng.random() < 0.15:
                kw["v_ref"] = int(rng.choice([v_dup, v_oth, v_short]))
            if rng.random() < 0.1:
                kw["u_ref"] = int(rng.choice([u_oth, NONE])) if rng.random() < 0.7 else w.ref_uh
            if rng.random() < 0.2:
                kw["c_ref"] = int(rng.choice([w.ref_lc, NONE]))
            kw["hcomp"] = int(rng.integers(0, 30))
            kw["donor"] = int(rng.choice([0, 0, 0, 1, NONE]))
            ncells = int(rng.integers(1, 4))
            kw["barcodes"] = rng.integers(0, 40, size=ncells).tolist()   # small pool -> overlaps
            row = w.member(h, l, has_del=has_del, feats=rng.random() < 0.85, **kw)
            # per-chain feature jitter (regions of different rows no longer line up)
            if rng.random() < 0.15:
                ch = b.exacts[-1][0][0]
                for f in ("cdr1", "fr2", "cdr2", "fr3"):
                    if f in ch and rng.random() < 0.5:
                        ch[f] = int(ch[f]) + 3 * int(rng.integers(-1, 2))
                if rng.random() < 0.3:
                    ch["fr1"] = int(ch["fr1"]) + 3
            if rng.random() < 0.05:
                b.exacts[-1][0][0]["cdr1"] = NONE
    # order rows like info.par_sort(): (lens, tigs, exact)
    b.rows.sort(key=lambda r: ([len(t) for t in r["tigs"]], list(r["tigs"]), r["exact"]))
    return default_params(params.pop("is_bcr"), **params), w.build()
