"""Hand-built micro-cases, one (or more) per decision row of join_one.

Each case is a tiny flattened join input plus the edges the reference's logic must produce,
derived by hand from enclone_core/src/join_one.rs / enclone/src/join.rs.  The CPU tests check the
oracle against these expectations; the GPU tests check the CUDA path against the oracle on the
same inputs.
"""
import numpy as np

from enclone_ranger_b200 import JoinInput, default_params

NONE = -1
B = np.frombuffer(b"ACGT", dtype=np.uint8)


def rand_seq(rng, n):
    return bytes(B[rng.integers(0, 4, n)])


def mutate(seq, positions, shift=1):
    s = bytearray(seq)
    for p in positions:
        s[p] = b"ACGT"[(b"ACGT".index(s[p]) + shift) % 4]
    return bytes(s)


class Builder:
    def __init__(self):
        self.refs, self.ref_names = [], []
        self.segs = []
        self.exacts = []   # (chains, cells)
        self.rows = []

    def ref(self, seq, name_id):
        self.refs.append(seq)
        self.ref_names.append(name_id)
        return len(self.refs) - 1

    def seg(self, seq):
        self.segs.append(seq)
        return len(self.segs) - 1

    def exact(self, chains, cells):
        """chains: list of dicts(left, c_ref, v_ref, u_ref, fr1, cdr1, fr2, cdr2, fr3, hcomp, amino);
        cells: list of (donor, barcode)."""
        self.exacts.append((chains, cells))
        return len(self.exacts) - 1

    def row(self, exact, tigs, cdr3s, vs, js, cols=(0, 1), has_del=(0, 0)):
        self.rows.append(dict(exact=exact, tigs=tigs, cdr3s=cdr3s, vs=vs, js=js, cols=cols, has_del=has_del))
        return len(self.rows) - 1

    def build(self):
        a = {k: [] for k in (
            "row_nchains", "row_len", "row_tig_off", "row_has_del", "row_vs", "row_js", "row_cdr3_off",
            "row_cdr3_len", "row_exact", "row_exact_col", "ex_chain_off", "ex_cell_off", "ch_left", "ch_c_ref",
            "ch_v_ref", "ch_u_ref", "ch_fr1_start", "ch_cdr1_start", "ch_fr2_start", "ch_cdr2_start",
            "ch_fr3_start", "ch_hcomp", "ch_amino_off", "ch_amino_len", "cell_donor", "cell_barcode",
            "seg_off", "ref_off", "ref_name_id")}
        tig, cdr3, amino, segb, refb = bytearray(), bytearray(), bytearray(), bytearray(), bytearray()
        for r in self.rows:
            n = len(r["tigs"])
            a["row_nchains"].append(n)
            a["row_exact"].append(r["exact"])
            for m in range(2):
                if m < n:
                    a["row_len"].append(len(r["tigs"][m])); a["row_tig_off"].append(len(tig)); tig += r["tigs"][m]
                    a["row_has_del"].append(r["has_del"][m]); a["row_vs"].append(r["vs"][m]); a["row_js"].append(r["js"][m])
                    a["row_cdr3_off"].append(len(cdr3)); a["row_cdr3_len"].append(len(r["cdr3s"][m])); cdr3 += r["cdr3s"][m]
                    a["row_exact_col"].append(r["cols"][m])
                else:
                    for k in ("row_len", "row_tig_off", "row_has_del", "row_vs", "row_js", "row_cdr3_off", "row_cdr3_len", "row_exact_col"):
                        a[k].append(0)
        a["ex_chain_off"].append(0); a["ex_cell_off"].append(0)
        for chains, cells in self.exacts:
            for c in chains:
                a["ch_left"].append(1 if c.get("left") else 0)
                a["ch_c_ref"].append(c.get("c_ref", NONE)); a["ch_v_ref"].append(c.get("v_ref", 0)); a["ch_u_ref"].append(c.get("u_ref", NONE))
                a["ch_fr1_start"].append(c.get("fr1", 0)); a["ch_cdr1_start"].append(c.get("cdr1", NONE)); a["ch_fr2_start"].append(c.get("fr2", NONE))
                a["ch_cdr2_start"].append(c.get("cdr2", NONE)); a["ch_fr3_start"].append(c.get("fr3", NONE)); a["ch_hcomp"].append(c.get("hcomp", 0))
                am = c.get("amino")
                if am is None:
                    a["ch_amino_off"].append(NONE); a["ch_amino_len"].append(0)
                else:
                    a["ch_amino_off"].append(len(amino)); a["ch_amino_len"].append(len(am)); amino += am
            a["ex_chain_off"].append(len(a["ch_left"]))
            for d, bc in cells:
                a["cell_donor"].append(d); a["cell_barcode"].append(bc)
            a["ex_cell_off"].append(len(a["cell_donor"]))
        a["seg_off"].append(0)
        for s in self.segs:
            segb += s; a["seg_off"].append(len(segb))
        a["ref_off"].append(0)
        for s, nm in zip(self.refs, self.ref_names):
            refb += s; a["ref_off"].append(len(refb)); a["ref_name_id"].append(nm)
        a["tig_bytes"] = np.frombuffer(bytes(tig), dtype=np.uint8)
        a["cdr3_bytes"] = np.frombuffer(bytes(cdr3), dtype=np.uint8)
        a["amino_bytes"] = np.frombuffer(bytes(amino), dtype=np.uint8)
        a["seg_bytes"] = np.frombuffer(bytes(segb), dtype=np.uint8)
        a["ref_bytes"] = np.frombuffer(bytes(refb), dtype=np.uint8)
        return JoinInput(**a)


class World:
    """A tiny germline + helpers to build clone members."""

    VH, JH, VL, JL = 300, 50, 290, 40
    NH, NL = 12, 2           # junction inserts
    FR1, CDR1, FR2, CDR2, FR3 = 30, 105, 129, 180, 204

    def __init__(self, seed=1):
        self.rng = np.random.default_rng(seed)
        r = self.rng
        self.b = Builder()
        self.vh, self.jh, self.vl, self.jl = rand_seq(r, self.VH), rand_seq(r, self.JH), rand_seq(r, self.VL), rand_seq(r, self.JL)
        self.uh = rand_seq(r, 60)
        b = self.b
        self.ref_vh = b.ref(self.vh, 0); self.ref_jh = b.ref(self.jh, 1); self.ref_vl = b.ref(self.vl, 2); self.ref_jl = b.ref(self.jl, 3)
        self.ref_uh = b.ref(self.uh, 4); self.ref_kc = b.ref(rand_seq(r, 100), 5); self.ref_lc = b.ref(rand_seq(r, 100), 6)
        self.seg_vh = b.seg(self.vh); self.seg_jh = b.seg(self.jh); self.seg_vl = b.seg(self.vl); self.seg_jl = b.seg(self.jl)
        self.nh, self.nl = rand_seq(r, self.NH), rand_seq(r, self.NL)
        self.next_bc = 100

    # germline contigs
    def heavy(self):
        return self.vh + self.nh + self.jh

    def light(self):
        return self.vl + self.nl + self.jl

    def cdr3_of(self, h, l):
        # CDR3 = 9 nt of V + insert + 15 nt of J (heavy: 36 nt, light: 26 -> padded to 27 with 16 nt of J)
        ch = h[self.VH - 9: self.VH + self.NH + 15]
        cl = l[self.VL - 9: self.VL + self.NL + 16]
        return ch, cl

    def member(self, h, l, donor=0, ncells=1, c_ref=None, hcomp=12, v_ref=None, vs=None, u_ref=None, feats=True, has_del=(0, 0), amino=None,
               barcodes=None, cdr3=None, extra_chain=False, js=None):
        b = self.b
        hc = dict(left=1, c_ref=NONE, v_ref=self.ref_vh if v_ref is None else v_ref, u_ref=self.ref_uh if u_ref is None else u_ref,
                  fr1=self.FR1, hcomp=hcomp, amino=amino)
        if feats:
            hc.update(cdr1=self.CDR1, fr2=self.FR2, cdr2=self.CDR2, fr3=self.FR3)
        lc = dict(left=0, c_ref=self.ref_kc if c_ref is None else c_ref, v_ref=self.ref_vl, u_ref=NONE, fr1=0)
        chains = [hc, lc] + ([dict(left=0, c_ref=self.ref_kc, v_ref=self.ref_vl, fr1=0)] if extra_chain else [])
        if barcodes is None:
            barcodes = list(range(self.next_bc, self.next_bc + ncells))
            self.next_bc += ncells
        e = b.exact(chains, [(donor, bc) for bc in barcodes])
        c3 = self.cdr3_of(h, l) if cdr3 is None else cdr3
        vsx = (self.seg_vh, self.seg_vl) if vs is None else vs
        jsx = (self.seg_jh, self.seg_jl) if js is None else js
        return b.row(e, (h, l), c3, vsx, jsx, has_del=has_del)

    def build(self):
        return self.b.build()


# positions used for mutations (heavy contig coordinates)
SHARED16 = [5, 22, 41, 58, 77, 90, 101, 140, 150, 161, 170, 215, 230, 244, 260, 271]   # all < VH - 15, outside CDR1/2
SHARED3 = [5, 41, 90]
PRIV_A = [12, 250]
PRIV_B = [66, 222]
CDR3_H = [World.VH - 9 + i for i in (10, 12, 14, 16, 18, 20, 22, 24, 26, 28)]          # inside the insert / J part of the CDR3


def _pair(w, shared, pa, pb, cdr3_mut_b=(), **kw_b):
    h0, l0 = w.heavy(), w.light()
    ha = mutate(mutate(h0, shared), pa)
    hb = mutate(mutate(mutate(h0, shared), pb), list(cdr3_mut_b), shift=2)
    a = w.member(ha, l0)
    b = w.member(hb, l0, **kw_b)
    return a, b


def all_cases():
    """Returns a list of (name, params, JoinInput, expected_edges or None)."""
    out = []

    def add(name, w, expected, **params):
        out.append((name, default_params(params.pop("is_bcr", True), **params), w.build(), expected))

    # --- baseline: 16 shared mutations (>= auto_share: the score threshold is skipped), cd = 0
    w = World(); _pair(w, SHARED16, PRIV_A, PRIV_B); add("base_join", w, [(0, 1)])
    # row 1: exact with 4 chains never joins (join_one.rs:87)
    w = World(); h0, l0 = w.heavy(), w.light()
    w.member(mutate(h0, SHARED16 + PRIV_A), l0)
    b = w.b; e = b.exact([dict(left=1, v_ref=0, fr1=0), dict(left=0, v_ref=2), dict(left=0, v_ref=2), dict(left=0, v_ref=2)], [(0, 7)])
    b.row(e, (mutate(h0, SHARED16 + PRIV_B), l0), w.cdr3_of(h0, l0), (w.seg_vh, w.seg_vl), (w.seg_jh, w.seg_jl))
    add("row1_four_chain_exact", w, [])
    # row 2: CDR3 lengths differ (same lens)
    w = World(); h0, l0 = w.heavy(), w.light()
    w.member(mutate(h0, SHARED16 + PRIV_A), l0)
    ch, cl = w.cdr3_of(h0, l0)
    w.member(mutate(h0, SHARED16 + PRIV_B), l0, cdr3=(ch[:-3], cl))
    add("row2_cdr3_len", w, [])
    # row 4: BCR CDR3 identity: T = 36 + 27 = 63; cd = 10 -> 0.1587 > 0.15 reject; cd = 9 -> 0.1428 passes the filter
    w = World(); _pair(w, SHARED16, PRIV_A + [150 + i for i in range(1, 40, 3)], PRIV_B, cdr3_mut_b=CDR3_H[:10]); add("row4_identity_reject", w, [])
    # cd = 9 passes the identity filter (0.1428), 9 < 5 * indeps(4) = 20, shares >= auto_share skips the score; 2 cells:
    # two-cell filter 9 > 16 / 2 deletes the edge; with EASY it is kept
    w = World(); _pair(w, SHARED16, PRIV_A, PRIV_B, cdr3_mut_b=CDR3_H[:9]); add("row4_identity_pass_two_cell_deleted", w, [])
    w = World(); _pair(w, SHARED16, PRIV_A, PRIV_B, cdr3_mut_b=CDR3_H[:9]); add("row4_identity_pass_easy", w, [(0, 1)], easy=1)
    # row 5/7: TCR — full-length differences <= 5 and cd == 0
    w = World(); _pair(w, [], [12, 250], [66, 222, 100]); add("row5_tcr_5diffs_join", w, [(0, 1)], is_bcr=False)
    w = World(); _pair(w, [], [12, 250, 31], [66, 222, 100]); add("row5_tcr_6diffs_reject", w, [], is_bcr=False)
    w = World(); _pair(w, [], [12], [66], cdr3_mut_b=CDR3_H[:1]); add("row7_tcr_cd1_reject", w, [], is_bcr=False)
    w = World(); _pair(w, SHARED16, PRIV_A, PRIV_B, cdr3_mut_b=CDR3_H[:2]); add("row7_max_cdr3_diffs_1", w, [], max_cdr3_diffs=1)
    # row 8: donors
    w = World(); _pair(w, SHARED16, PRIV_A, PRIV_B, donor=1); add("row8_donor_reject", w, [])
    w = World(); _pair(w, SHARED16, PRIV_A, PRIV_B, donor=1); add("row8_mix_donors_join_err", w, [(0, 1)], mix_donors=1)
    w = World(); h0, l0 = w.heavy(), w.light()
    w.member(mutate(h0, SHARED16 + PRIV_A), l0, donor=NONE); w.member(mutate(h0, SHARED16 + PRIV_B), l0, donor=NONE)
    add("row8_no_donors_join_err", w, [(0, 1)])
    # rows 9-11: two references
    w = World(); h0, l0 = w.heavy(), w.light()
    alt = mutate(w.vh, [200]); seg_alt = w.b.seg(alt)
    w.member(mutate(h0, SHARED16 + PRIV_A), l0); w.member(mutate(h0, SHARED16 + PRIV_B), l0, vs=(seg_alt, w.seg_vl))
    add("row9_two_refs_join", w, [(0, 1)])
    w = World(); h0, l0 = w.heavy(), w.light()
    alt = mutate(w.vh, [200, 201, 202]); seg_alt = w.b.seg(alt)
    w.member(mutate(h0, SHARED16 + PRIV_A), l0); w.member(mutate(h0, SHARED16 + PRIV_B), l0, vs=(seg_alt, w.seg_vl))
    add("row11_degradation_reject", w, [])
    # row 12: barcode overlap
    w = World(); h0, l0 = w.heavy(), w.light()
    w.member(mutate(h0, SHARED16 + PRIV_A), l0, barcodes=[5, 9]); w.member(mutate(h0, SHARED16 + PRIV_B), l0, barcodes=[9, 11, 12])
    add("row12_barcode_overlap", w, [])
    # row 13: cd >= 5 * max(1, indeps): indeps = 1 (only B has a private mutation), cd = 5
    w = World(); _pair(w, SHARED16, [], [66], cdr3_mut_b=CDR3_H[:5]); add("row13_cdr3_mult_reject", w, [])
    w = World(); _pair(w, SHARED16, [], [66], cdr3_mut_b=CDR3_H[:4]); add("row13_cdr3_mult_pass", w, [(0, 1)])  # two-cell filter: 4 > 16/2 is false
    # row 14: different light-chain constant regions
    w = World(); _pair(w, SHARED16, PRIV_A, PRIV_B, cdr3_mut_b=CDR3_H[:1], c_ref=6); add("row14_cref_cd1_reject", w, [])
    w = World(); _pair(w, SHARED16, PRIV_A, PRIV_B, c_ref=6); add("row14_cref_cd0_join", w, [(0, 1)])
    w = World(); _pair(w, SHARED16, PRIV_A, PRIV_B, cdr3_mut_b=CDR3_H[:1], c_ref=6); add("row14_old_light_join", w, [(0, 1)], old_light=1)
    # rows 15-17: score.  3 shares, 4 indeps; cd = 6 on a 36-nt heavy CDR3 -> mult = 80^7, score >> 1e5
    w = World(); _pair(w, SHARED3, PRIV_A, PRIV_B, cdr3_mut_b=CDR3_H[:6], hcomp=5); add("row17_score_reject", w, [])
    w = World(); h0, l0 = w.heavy(), w.light()  # JUN_SHARE: min(hcomp) - cd = 20 - 6 >= 8 accepts
    w.member(mutate(h0, SHARED3 + PRIV_A), l0, hcomp=20, ncells=2)
    w.member(mutate(mutate(h0, SHARED3 + PRIV_B), CDR3_H[:6], 2), l0, hcomp=25)
    add("row16_jun_share_accept", w, [(0, 1)])
    w = World(); _pair(w, SHARED3, PRIV_A, PRIV_B, hcomp=5); add("row15_low_score_two_cell_kept", w, [(0, 1)])  # cd = 0
    # two-cell filter (join.rs:120-178): cd = 2 > shares/2 = 1 with one cell each -> deleted; with 3 cells -> kept
    w = World(); _pair(w, SHARED3, PRIV_A, PRIV_B, cdr3_mut_b=CDR3_H[:2]); add("two_cell_filter_deletes", w, [])
    w = World(); _pair(w, SHARED3, PRIV_A, PRIV_B, cdr3_mut_b=CDR3_H[:2]); add("two_cell_filter_easy", w, [(0, 1)], easy=1)
    w = World(); h0, l0 = w.heavy(), w.light()
    w.member(mutate(h0, SHARED3 + PRIV_A), l0, ncells=2); w.member(mutate(mutate(h0, SHARED3 + PRIV_B), CDR3_H[:2], 2), l0)
    add("two_cell_filter_three_cells_kept", w, [(0, 1)])
    # row 18: V gene names
    w = World(); vdup = w.b.ref(w.vh, 50); _pair(w, SHARED16, PRIV_A, PRIV_B, v_ref=vdup); add("row18_vname_same_seq_join", w, [(0, 1)])
    w = World(); vother = w.b.ref(mutate(w.vh, [100]), 50); _pair(w, SHARED16, PRIV_A, PRIV_B, v_ref=vother); add("row18_vname_diff_seq_reject", w, [])
    w = World(); vdup = w.b.ref(w.vh, 50); uoth = w.b.ref(mutate(w.uh, [59]), 51)
    _pair(w, SHARED16, PRIV_A, PRIV_B, v_ref=vdup, u_ref=uoth); add("row18_utr_suffix_reject", w, [])
    w = World(); vshort = w.b.ref(w.vh[:250], 50); _pair(w, SHARED16, PRIV_A, PRIV_B, v_ref=vshort); add("row18_vname_prefix_join", w, [(0, 1)])
    # row 19: FWR1 = [30,105), 1 private diff (pos 66): identity 98.67; CDR1+CDR2 = 48 nt:
    # 11 diffs -> 98.67 - 77.08 = 21.6 >= 20 reject; 9 diffs -> 98.67 - 81.25 = 17.4 joins
    cd12 = [106, 108, 110, 112, 114, 116, 181, 183, 185, 187, 189]
    w = World(); _pair(w, SHARED16, PRIV_A, PRIV_B + cd12); add("row19_fwr1_cdr12_reject", w, [])
    w = World(); _pair(w, SHARED16, PRIV_A, PRIV_B + cd12[:9]); add("row19_fwr1_cdr12_pass", w, [(0, 1)])
    w = World(); _pair(w, SHARED16, PRIV_A, PRIV_B + cd12, feats=False); add("row19_missing_features_join", w, [(0, 1)])
    # has_del rows: '-' bytes take part in every compare (info.rs:78-110)
    w = World(); h0, l0 = w.heavy(), w.light()
    hd = bytearray(mutate(h0, SHARED16 + PRIV_A)); hd[120:123] = b"---"
    hd2 = bytearray(mutate(h0, SHARED16 + PRIV_B)); hd2[120:123] = b"---"
    vdel = w.b.seg(w.vh)
    w.member(bytes(hd), l0, has_del=(1, 0), vs=(vdel, w.seg_vl)); w.member(bytes(hd2), l0, has_del=(1, 0), vs=(vdel, w.seg_vl))
    add("has_del_pair_join", w, [(0, 1)])
    w = World(); h0, l0 = w.heavy(), w.light()
    hd = bytearray(mutate(h0, SHARED16 + PRIV_A)); hd[120:123] = b"---"
    am = bytearray(mutate(h0, SHARED16 + PRIV_A)); am[121:124] = b"---"
    w.member(bytes(hd), l0, has_del=(1, 0), amino=bytes(am)); w.member(mutate(h0, SHARED16 + PRIV_B), l0)
    add("has_del_vs_plain_amino_override", w, None)
    # odd byte ('N'): byte compare differs from the 2-bit compare
    w = World(); h0, l0 = w.heavy(), w.light()
    hn = bytearray(mutate(h0, SHARED16 + PRIV_A)); hn[50] = ord("N")
    w.member(bytes(hn), l0); w.member(mutate(h0, SHARED16 + PRIV_B), l0)
    add("odd_byte_N", w, None)
    # V bail-out: the heavy V segment of row B is longer than contig + 15 (join_one.rs:371-373)
    w = World(); h0, l0 = w.heavy(), w.light()
    vlong = w.b.seg(w.vh + rand_seq(w.rng, 100))
    w.member(mutate(h0, SHARED16 + PRIV_A), l0); w.member(mutate(h0, SHARED16 + PRIV_B), l0, vs=(vlong, w.seg_vl))
    add("v_bailout_reject", w, [])
    # short junction: V and J compare ranges overlap (positions counted against both references)
    w = World(); jlong = rand_seq(w.rng, 120); seg_jlong = w.b.seg(jlong)
    hshort = w.vh[:200] + jlong[20:]   # L = 300, V range 285 + J range 105 > 300
    cd3 = (hshort[191:227], w.light()[w.VL - 9: w.VL + w.NL + 16])
    w.member(mutate(hshort, SHARED16[:12] + [12]), w.light(), cdr3=cd3, js=(seg_jlong, w.seg_jl), feats=False)
    w.member(mutate(hshort, SHARED16[:12] + [66]), w.light(), cdr3=cd3, js=(seg_jlong, w.seg_jl), feats=False)
    add("vj_overlap", w, None)
    # order dependence (join_core.rs:24-48): A~B, A~C, B~C all pass -> pot = (A,B), (A,C)
    w = World(); h0, l0 = w.heavy(), w.light()
    for priv in ([12, 250], [66, 222], [33, 133]):
        w.member(mutate(h0, SHARED16 + priv), l0)
    add("order_dependent_forest", w, [(0, 1), (0, 2)])
    # buckets are RUNS of equal lens (join.rs:70-88): an interposed row splits the run
    w = World(); h0, l0 = w.heavy(), w.light()
    w.member(mutate(h0, SHARED16 + PRIV_A), l0)
    w.member(mutate(h0, SHARED16)[:-3], l0)
    w.member(mutate(h0, SHARED16 + PRIV_B), l0)
    add("bucket_runs_split", w, [])
    # finish_join (join2.rs:69-84): two rows of one 3-chain exact end in one orbit although no join_one pair passes
    w = World(); h0, l0 = w.heavy(), w.light()
    b = w.b
    e = b.exact([dict(left=1, v_ref=w.ref_vh, fr1=0), dict(left=0, v_ref=w.ref_vl, c_ref=w.ref_kc), dict(left=0, v_ref=w.ref_vl, c_ref=w.ref_kc)], [(0, 1), (0, 2)])
    l2 = w.vl + rand_seq(w.rng, 5) + w.jl
    b.row(e, (h0, l0), w.cdr3_of(h0, l0), (w.seg_vh, w.seg_vl), (w.seg_jh, w.seg_jl), cols=(0, 1))
    b.row(e, (h0, l2), w.cdr3_of(h0, l0), (w.seg_vh, w.seg_vl), (w.seg_jh, w.seg_jl), cols=(0, 2))
    w.member(mutate(h0, SHARED16), w.light())
    add("finish_join_three_chain_merge", w, None)
    # onesie rows (lens.len() == 1) never join and form their own buckets
    w = World(); h0, l0 = w.heavy(), w.light()
    b = w.b
    for _ in range(2):
        e = b.exact([dict(left=1, v_ref=w.ref_vh, fr1=0)], [(0, w.next_bc)]); w.next_bc += 1
        b.row(e, (h0,), (w.cdr3_of(h0, l0)[0],), (w.seg_vh,), (w.seg_jh,), cols=(0,))
    add("onesies_never_join", w, [])
    return out
