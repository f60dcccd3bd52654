"""Generates the committed golden fixtures (run in the build container only; never at test time).

  hs_bcr_vdjrefs7.npz   a small human BCR repertoire whose contigs are recombined from the REAL genes of
                        the reference's bundled vdj_ann_ref/vdj_refs_7.0/human/fasta/regions.fa
                        (IGH + IGK/IGL, BASELINE.json configs[0] at fixture scale), with the oracle's result.
  synth_bcr_2k.npz / synth_tcr_20k.npz   the seeded synthetic generator at small scale + oracle result.

The .npz holds every ejb_input array plus the oracle's edges / tuples / EquivRel, so the tests can
(a) pin the oracle against regressions and (b) check the CUDA path on the GPU box, where
/root/reference does not exist.  The reference itself ships no join fixture.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import enclone_ranger_b200 as E  # noqa: E402
from cases import Builder, mutate, rand_seq  # noqa: E402
from oracle import pyoracle  # noqa: E402

FASTA = "/root/reference/vdj_ann_ref/vdj_refs_7.0/human/fasta/regions.fa"
ORACLE_KEYS = ("edge_k1", "edge_k2", "edge_cd", "edge_nrefs", "edge_shares", "edge_indeps", "edge_k", "edge_d", "edge_n",
               "edge_p1", "edge_mult", "edge_score", "edge_err", "eq_x", "eq_y", "eq_z", "labels")


def save(name, inp, params, ora):
    d = {"in_" + k: v for k, v in inp.a.items()}
    d.update({"or_" + k: ora[k] for k in ORACLE_KEYS})
    d["is_bcr"] = np.array([params.is_bcr])
    np.savez_compressed(os.path.join(HERE, name), **d)
    print(name, "rows", inp.n_rows, "edges", len(ora["edge_k1"]), "pairs_scored", ora["pairs_scored"])


def read_fasta(path):
    recs, name, seq = [], None, []
    for line in open(path):
        line = line.strip()
        if line.startswith(">"):
            if name:
                recs.append((name, "".join(seq)))
            name, seq = line[1:], []
        else:
            seq.append(line)
    recs.append((name, "".join(seq)))
    return recs


def human_bcr(seed=7, n_clones=60):
    rng = np.random.default_rng(seed)
    recs = read_fasta(FASTA)
    b = Builder()
    names = {}
    genes = {}
    for hdr, seq in recs:
        f = hdr.split("|")
        gene, region, chain = f[2], f[3], f[5]
        if chain not in ("IGH", "IGK", "IGL"):
            continue
        stripped = gene.split("*")[0]
        nid = names.setdefault((stripped, region), len(names))
        rid = b.ref(seq.upper().encode(), nid)
        genes.setdefault((chain, region), []).append((gene, rid, seq.upper().encode()))
    utr = {(c, g): rid for c in ("IGH", "IGK", "IGL") for g, rid, _ in genes.get((c, "5'UTR"), [])}
    segs = {}

    def seg(s):
        if s not in segs:
            segs[s] = b.seg(s)
        return segs[s]

    bc = 0
    for c in range(n_clones):
        hv = genes[("IGH", "L-REGION+V-REGION")][rng.integers(len(genes[("IGH", "L-REGION+V-REGION")]))]
        hj = genes[("IGH", "J-REGION")][rng.integers(len(genes[("IGH", "J-REGION")]))]
        lchain = "IGK" if rng.random() < 0.6 else "IGL"
        lv = genes[(lchain, "L-REGION+V-REGION")][rng.integers(len(genes[(lchain, "L-REGION+V-REGION")]))]
        lj = genes[(lchain, "J-REGION")][rng.integers(len(genes[(lchain, "J-REGION")]))]
        lc = genes[(lchain, "C-REGION")][rng.integers(len(genes[(lchain, "C-REGION")]))]
        vt, jt = int(rng.integers(0, 7)), int(rng.integers(0, 7))
        nh = rand_seq(rng, int(rng.integers(6, 30)))
        h0 = hv[2][: len(hv[2]) - vt] + nh + hj[2][jt:]
        l0 = lv[2][: len(lv[2]) - int(rng.integers(0, 4))] + rand_seq(rng, int(rng.integers(0, 4))) + lj[2][int(rng.integers(0, 4)):]
        ch0 = len(hv[2]) - 9
        clen = 3 * ((9 - vt + len(nh) + 24) // 3)
        cl0 = len(lv[2]) - 12
        cll = 3 * (min(len(l0) - cl0, 33) // 3)
        size = int(min(40, rng.zipf(1.6)))
        trunk = rng.choice(len(h0) - 40, size=int(rng.integers(0, 25)), replace=False).tolist()
        members = []
        for e in range(size):
            if e == 0:
                h, l = mutate(h0, trunk), l0
            else:
                ph, pl = members[rng.integers(len(members))]
                h = mutate(ph, rng.choice(len(h0), size=int(rng.integers(1, 5)), replace=False).tolist(), shift=int(rng.integers(1, 4)))
                l = mutate(pl, rng.choice(len(l0), size=int(rng.integers(0, 3)), replace=False).tolist(), shift=int(rng.integers(1, 4)))
            members.append((h, l))
            ncells = int(rng.geometric(0.7))
            fr1 = 57
            hc = dict(left=1, c_ref=-1, v_ref=hv[1], u_ref=utr.get(("IGH", hv[0]), -1), fr1=fr1, cdr1=fr1 + 75, fr2=fr1 + 99, cdr2=fr1 + 150,
                      fr3=fr1 + 174, hcomp=len(nh) // 2)
            lcd = dict(left=0, c_ref=lc[1], v_ref=lv[1], u_ref=utr.get((lchain, lv[0]), -1), fr1=0)
            ex = b.exact([hc, lcd], [(0, bc + i) for i in range(ncells)])
            bc += ncells
            b.row(ex, (h, l), (h[ch0: ch0 + clen], l[cl0: cl0 + cll]), (seg(hv[2]), seg(lv[2])), (seg(hj[2]), seg(lj[2])))
    # info.par_sort(): lens, then tigs (info.rs:374)
    b.rows.sort(key=lambda r: ([len(t) for t in r["tigs"]], list(r["tigs"]), r["exact"]))
    return b.build()


if __name__ == "__main__":
    p = E.default_params(True)
    inp = human_bcr()
    save("hs_bcr_vdjrefs7.npz", inp, p, pyoracle.run(p, inp.c_ptr(), threads=0))
    rep = E.generate(0, n_cells=2000, seed=4242)
    inp = rep.to_join_input()
    save("synth_bcr_2k.npz", inp, p, pyoracle.run(p, inp.c_ptr(), threads=0))
    p = E.default_params(False)
    rep = E.generate(2, n_cells=20000, seed=4243)
    inp = rep.to_join_input()
    save("synth_tcr_20k.npz", inp, p, pyoracle.run(p, inp.c_ptr(), threads=0))
