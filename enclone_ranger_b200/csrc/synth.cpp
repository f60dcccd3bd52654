// synth.cpp — seeded synthetic repertoire generator at the JOIN BOUNDARY.
//
// The upstream enclone stages (annotation, exact-subclonotype grouping, build_info, allele
// calling) cannot be run here (no Rust toolchain), so the BASELINE.json configs are generated
// directly as the flattened join input of include/enclone_join_b200.h — i.e. the values that
// info[] (CloneInfo, enclone_core/src/defs.rs:736-758), exact_clonotypes[] (:698-728), to_bc
// (enclone_stuff/src/start.rs:331-340) and refdata (vdj_ann/src/refx.rs:24-38) would hold when
// main_enclone_start calls join_exacts (start.rs:365-375).  Row order follows the derive(Ord)
// of CloneInfo used by info.par_sort() (enclone/src/info.rs:374): lens, then tigs, ...
//
// Germline genes are SYNTHETIC (random gene families with the segment counts and lengths of the
// human references bundled as vdj_ann_ref/vdj_refs_7.0: IGHV 64 @ 345-377 nt, IGKV 43, IGLV 38,
// IGHJ 12 @ 46-63, IGKJ 5, IGLJ 7; TRAV 46, TRBV 83, TRAJ 58, TRBJ 16) so that nothing from the
// reference tree is needed at run time.  Bench / test infrastructure, not part of the join.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/enclone_join_b200.h"
#include "synth.h"

namespace {

struct Rng {  // xoshiro256**
    uint64_t s[4];
    static uint64_t splitmix(uint64_t& x) {
        uint64_t z = (x += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    explicit Rng(uint64_t seed) {
        for (auto& v : s) v = splitmix(seed);
    }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() {
        uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    int64_t below(int64_t n) { return (int64_t)(uni() * (double)n); }  // [0, n)
    int64_t range(int64_t a, int64_t b) { return a + below(b - a + 1); }  // [a, b]
    int poisson(double lam) {
        if (lam <= 0) return 0;
        if (lam < 30) {
            double l = std::exp(-lam), p = 1.0;
            int k = 0;
            do { k++; p *= uni(); } while (p > l);
            return k - 1;
        }
        double u1 = uni(), u2 = uni();
        double g = std::sqrt(-2.0 * std::log(u1 + 1e-300)) * std::cos(6.283185307179586 * u2);
        int k = (int)std::lround(lam + std::sqrt(lam) * g);
        return k < 0 ? 0 : k;
    }
};

const char kB[4] = {'A', 'C', 'G', 'T'};

std::string rand_seq(Rng& r, int n) {
    std::string s(n, 'A');
    for (auto& c : s) c = kB[r.below(4)];
    return s;
}
std::string mutate_frac(Rng& r, const std::string& a, double frac) {
    std::string s = a;
    for (auto& c : s)
        if (r.uni() < frac) {
            char o = c;
            while (c == o) c = kB[r.below(4)];
        }
    return s;
}

struct VGene {
    std::string seq;      // L + V region
    int fr1, cdr1, fr2, cdr2, fr3;  // feature starts on V..J (== on the V ref, contig starts at V)
    int ref_id, u_ref_id; // refdata indices
    int name_id;
    int locus;            // 0 = heavy/left (IGH/TRB), 1 = kappa (or TRA), 2 = lambda
};
struct JGene {
    std::string seq;
    int motif_end;  // offset just after the W/F codon that ends the CDR3
    int ref_id;
    int locus;
    int c_ref;      // constant region paired with this J (light), EJB_NONE for heavy
};

struct Germline {
    std::vector<VGene> v[3];
    std::vector<JGene> j[3];
    std::vector<std::string> d;  // D genes (heavy only)
    std::vector<std::string> ref_seq;
    std::vector<int> ref_name;
    int n_names = 0;
    int add_ref(const std::string& s, int name_id) {
        ref_seq.push_back(s);
        ref_name.push_back(name_id);
        return (int)ref_seq.size() - 1;
    }
};

void build_v_family(Rng& r, Germline& g, int locus, int count, int nfam) {
    struct Anc { std::string lead, f1, c1, f2, c2, f3; };
    std::vector<Anc> anc(nfam);
    for (auto& a : anc) {
        a.lead = rand_seq(r, 57);
        a.f1 = rand_seq(r, 75);
        a.c1 = rand_seq(r, 24);
        a.f2 = rand_seq(r, 51);
        a.c2 = rand_seq(r, 24);
        a.f3 = rand_seq(r, 108);
    }
    for (int i = 0; i < count; i++) {
        const Anc& a = anc[i % nfam];
        const double div = 0.04 + 0.12 * r.uni();
        // leader and CDR lengths vary in whole codons: V lengths land in ~[339, 378]
        int dl = 3 * (int)r.range(-2, 2), d1 = 3 * (int)r.range(0, 2), d2 = 3 * (int)r.range(-1, 2);
        std::string lead = mutate_frac(r, a.lead, div);
        if (dl < 0) lead = lead.substr(0, lead.size() + dl); else lead += rand_seq(r, dl);
        std::string c1 = mutate_frac(r, a.c1, div * 1.5) + rand_seq(r, d1);
        std::string c2 = mutate_frac(r, a.c2, div * 1.5);
        if (d2 < 0) c2 = c2.substr(0, c2.size() + d2); else c2 += rand_seq(r, d2);
        std::string f1 = mutate_frac(r, a.f1, div), f2 = mutate_frac(r, a.f2, div), f3 = mutate_frac(r, a.f3, div);
        // conserved 3' end: Cys codon + 6 nt; the CDR3 starts at the Cys codon, 9 nt before the V end
        static const char* tails[3] = {"TGTGCGAGA", "TGTCAGCAG", "TGTCAGTCC"};
        std::string tail = tails[locus];
        if (r.uni() < 0.4) tail[5 + r.below(4)] = kB[r.below(4)];
        VGene v;
        v.seq = lead + f1 + c1 + f2 + c2 + f3 + tail;
        v.fr1 = (int)lead.size();
        v.cdr1 = v.fr1 + (int)f1.size();
        v.fr2 = v.cdr1 + (int)c1.size();
        v.cdr2 = v.fr2 + (int)f2.size();
        v.fr3 = v.cdr2 + (int)c2.size();
        v.locus = locus;
        v.name_id = g.n_names++;
        v.ref_id = g.add_ref(v.seq, v.name_id);
        v.u_ref_id = g.add_ref(rand_seq(r, (int)r.range(30, 120)), g.n_names++);
        g.v[locus].push_back(v);
        const double q = r.uni();
        if (q < 0.25) {  // a second allele: same stripped name, 1-3 nt apart
            VGene w = v;
            for (int t = (int)r.range(1, 3); t > 0; t--) {
                int p = (int)r.below((int64_t)w.seq.size() - 20);
                char o = w.seq[p];
                while (w.seq[p] == o) w.seq[p] = kB[r.below(4)];
            }
            w.ref_id = g.add_ref(w.seq, v.name_id);
            g.v[locus].push_back(w);
        } else if (q < 0.30) {  // a duplicated gene: different name, identical sequence + UTR
            VGene w = v;
            w.name_id = g.n_names++;
            w.ref_id = g.add_ref(w.seq, w.name_id);
            g.v[locus].push_back(w);
        }
    }
}

void build_j(Rng& r, Germline& g, int locus, int count, int lmin, int lmax, const char* motif) {
    for (int i = 0; i < count; i++) {
        const int len = (int)r.range(lmin, lmax);
        const int pre = (int)r.range(8, std::max(8, len - 34));  // bases before the W/F codon
        JGene j;
        j.seq = rand_seq(r, pre) + motif + rand_seq(r, len - pre - (int)strlen(motif));
        j.motif_end = pre + 3;
        j.locus = locus;
        j.ref_id = g.add_ref(j.seq, g.n_names++);
        j.c_ref = EJB_NONE;  // light-chain constant regions are attached in build_germline
        g.j[locus].push_back(j);
    }
}

Germline build_germline(bool bcr, uint64_t seed) {
    Rng r(seed ^ (bcr ? 0x1234567ull : 0x7654321ull));
    Germline g;
    if (bcr) {
        build_v_family(r, g, 0, 64, 7);
        build_v_family(r, g, 1, 43, 6);
        build_v_family(r, g, 2, 38, 10);
        build_j(r, g, 0, 12, 46, 63, "TGGGGCCAGGGA");
        build_j(r, g, 1, 5, 37, 39, "TTCGGCGGAGGG");
        build_j(r, g, 2, 7, 38, 42, "TTCGGCGGAGGG");
        for (int i = 0; i < 27; i++) g.d.push_back(rand_seq(r, (int)r.range(11, 37)));
        // constant regions: one IGKC, four IGLC
        int kc = g.add_ref(rand_seq(r, 320), g.n_names++);
        for (auto& j : g.j[1]) j.c_ref = kc;
        int lc[4];
        for (int& c : lc) c = g.add_ref(rand_seq(r, 317), g.n_names++);
        for (size_t i = 0; i < g.j[2].size(); i++) g.j[2][i].c_ref = lc[i % 4];
    } else {
        build_v_family(r, g, 0, 83, 20);  // TRBV (left)
        build_v_family(r, g, 1, 46, 20);  // TRAV
        build_j(r, g, 0, 16, 47, 53, "TTCGGCCCAGGG");
        build_j(r, g, 1, 58, 52, 66, "TTTGGAAAAGGA");
        for (int i = 0; i < 2; i++) g.d.push_back(rand_seq(r, (int)r.range(12, 16)));
        int ac = g.add_ref(rand_seq(r, 420), g.n_names++);
        for (auto& j : g.j[1]) j.c_ref = ac;
    }
    return g;
}

struct Recomb {      // one recombined chain of a clone (unmutated)
    std::string seq; // V..J
    int vi, ji, locus;
    int cdr3_start, cdr3_len;
    int hcomp;
};

Recomb recombine(Rng& r, const Germline& g, int locus, int vi, int ji, bool heavy) {
    const VGene& v = g.v[locus][vi];
    const JGene& j = g.j[locus][ji];
    Recomb c;
    c.vi = vi; c.ji = ji; c.locus = locus;
    const int vt = (int)r.range(0, 6);
    const int jt = (int)r.range(0, std::min(6, j.motif_end - 3));
    std::string ins;
    int n_ins = 0;
    if (heavy && !g.d.empty()) {
        const std::string& d = g.d[r.below((int64_t)g.d.size())];
        int a = (int)r.range(0, (int64_t)d.size() / 3), b = (int)r.range(0, (int64_t)d.size() / 3);
        std::string n1 = rand_seq(r, (int)r.range(0, 8)), n2 = rand_seq(r, (int)r.range(0, 8));
        ins = n1 + d.substr(a, d.size() - a - b) + n2;
        n_ins = (int)(n1.size() + n2.size());
    } else {
        ins = rand_seq(r, (int)r.range(0, 3));
        n_ins = (int)ins.size();
    }
    // CDR3 = Cys codon (9 nt before the germline V end) .. end of the J W/F codon; force a whole
    // number of codons by padding the insert.
    const int cdr3_start = (int)v.seq.size() - 9;
    for (;;) {
        int len = (9 - vt) + (int)ins.size() + (j.motif_end - jt);
        if (len % 3 == 0 && len >= 15) break;
        ins += kB[r.below(4)];
        n_ins++;
    }
    c.seq = v.seq.substr(0, v.seq.size() - vt) + ins + j.seq.substr(jt);
    c.cdr3_start = cdr3_start;
    c.cdr3_len = (9 - vt) + (int)ins.size() + (j.motif_end - jt);
    c.hcomp = n_ins + (int)r.range(0, 4);
    return c;
}

struct ExactRec {         // one exact subclonotype under construction
    std::string h, l, l2; // heavy, light, optional second light (3-chain)
    int32_t clone;        // clone index
    int32_t ncells;
    uint8_t nchain_kind;  // 0: H+L, 1: H+L+L2, 2: H only, 3: L only
    uint8_t del_kind;     // 0 none, 1 deletion (dashes), 2 insertion (has_del flag only)
    int16_t del_pos, del_len;
};

struct CloneRec {
    Recomb h, l, l2;
    int donor;
    bool alt_allele;      // vs of heavy is a donor allele not in refdata
    bool c_none;
    int64_t first_exact, n_exacts;
};

}  // namespace

struct ejb_synth_data {
    ejb_input in;
    ejb_synth_summary sum;
    std::vector<int32_t> row_nchains, row_len, row_vs, row_js, row_cdr3_len, row_exact, row_exact_col;
    std::vector<int64_t> row_tig_off, row_cdr3_off;
    std::vector<uint8_t> row_has_del, tig_bytes, cdr3_bytes;
    std::vector<int64_t> ex_chain_off, ex_cell_off;
    std::vector<uint8_t> ch_left;
    std::vector<int32_t> ch_c_ref, ch_v_ref, ch_u_ref, ch_fr1, ch_cdr1, ch_fr2, ch_cdr2, ch_fr3, ch_hcomp, ch_amino_len;
    std::vector<int64_t> ch_amino_off;
    std::vector<uint8_t> amino_bytes;
    std::vector<int32_t> cell_donor;
    std::vector<uint32_t> cell_barcode;
    std::vector<int64_t> seg_off, ref_off;
    std::vector<uint8_t> seg_bytes, ref_bytes;
    std::vector<int32_t> ref_name_id;
    std::vector<int32_t> row_clone;  // ground truth: clone index of each row
};

extern "C" {

void ejb_synth_preset(int config, ejb_synth_config* c) {
    memset(c, 0, sizeof(*c));
    c->seed = 20261019ull + (uint64_t)config;
    c->is_bcr = 1;
    c->n_cells = 10000;
    c->zipf_s = 1.6;
    c->max_clone = 500;
    c->shm_min = 0.0;
    c->shm_max = 0.08;
    c->n_datasets = 1;
    c->n_donors = 1;
    c->frac_has_del = 0.01;
    c->frac_three_chain = 0.02;
    c->frac_onesie = 0.03;
    c->v_skew = 0.0;
    c->cells_per_exact_p = 0.7;
    c->frac_alt_allele = 0.03;
    c->frac_contam = 0.0005;
    c->threads = 0;
    switch (config) {
        case 0: break;
        case 1: c->n_cells = 1000000; c->zipf_s = 1.4; c->max_clone = 50000; c->shm_max = 0.10; c->n_datasets = 4; c->n_donors = 2; break;
        case 2: c->n_cells = 1000000; c->is_bcr = 0; c->zipf_s = 1.8; c->max_clone = 5000; c->shm_min = 0; c->shm_max = 0;
                c->n_datasets = 4; c->n_donors = 2; c->frac_has_del = 0.0; break;
        case 3: c->n_cells = 2000000; c->zipf_s = 1.3; c->max_clone = 100000; c->shm_min = 0.08; c->shm_max = 0.20; c->v_skew = 1.0;
                c->n_datasets = 4; c->n_donors = 2; break;
        case 4: c->n_cells = 10000000; c->zipf_s = 1.4; c->max_clone = 50000; c->shm_max = 0.10; c->n_datasets = 16; c->n_donors = 4; break;
        default: break;
    }
}

static void mutate_into(Rng& r, std::string& s, int nmut, const VGene& v, int vlen_in_tig) {
    // SHM: substitutions anywhere on V..J, rate x3 inside CDR1 / CDR2
    const int L = (int)s.size();
    for (int t = 0; t < nmut; t++) {
        int p;
        for (;;) {
            p = (int)r.below(L);
            bool hot = (p >= v.cdr1 && p < v.fr2) || (p >= v.cdr2 && p < v.fr3);
            if (hot || r.uni() < (1.0 / 3.0)) break;
        }
        (void)vlen_in_tig;
        char o = s[p];
        while (s[p] == o) s[p] = kB[r.below(4)];
    }
}

int ejb_synth_generate(const ejb_synth_config* cfg, ejb_synth_data** out_data) {
    *out_data = nullptr;
    if (cfg->n_cells <= 0 || cfg->n_datasets <= 0 || cfg->n_donors <= 0 || cfg->n_datasets > 1000) return EJB_ERR_INVALID;
    const bool bcr = cfg->is_bcr != 0;
    Germline g = build_germline(bcr, 0xB200ull);
    Rng main_rng(cfg->seed);

    // ---- clone sizes: truncated Zipf ----
    std::vector<int64_t> sizes;
    {
        const int64_t M = std::max<int64_t>(1, cfg->max_clone);
        std::vector<double> cdf(M);
        double acc = 0;
        for (int64_t i = 1; i <= M; i++) {
            acc += std::pow((double)i, -cfg->zipf_s);
            cdf[i - 1] = acc;
        }
        int64_t left = cfg->n_cells;
        while (left > 0) {
            double u = main_rng.uni() * acc;
            int64_t s = (int64_t)(std::lower_bound(cdf.begin(), cdf.end(), u) - cdf.begin()) + 1;
            s = std::min(s, left);
            sizes.push_back(s);
            left -= s;
        }
    }
    const int64_t n_clones = (int64_t)sizes.size();

    // V usage weights per locus
    std::vector<double> vcdf[3];
    for (int loc = 0; loc < 3; loc++) {
        const size_t n = g.v[loc].size();
        vcdf[loc].resize(n);
        double acc = 0;
        for (size_t i = 0; i < n; i++) {
            acc += (cfg->v_skew > 0) ? std::pow((double)(i + 1), -1.6 * cfg->v_skew) : 1.0;
            vcdf[loc][i] = acc;
        }
    }
    auto pick_v = [&](Rng& r, int loc) {
        double u = r.uni() * vcdf[loc].back();
        return (int)(std::lower_bound(vcdf[loc].begin(), vcdf[loc].end(), u) - vcdf[loc].begin());
    };

    // ---- exact counts per clone (cells per exact ~ Geometric(p)) drawn sequentially ----
    std::vector<CloneRec> clones(n_clones);
    std::vector<int32_t> ex_ncells_all;
    std::vector<uint64_t> clone_seed(n_clones);
    {
        const double p = std::min(1.0, std::max(0.01, cfg->cells_per_exact_p));
        for (int64_t c = 0; c < n_clones; c++) {
            clone_seed[c] = main_rng.next();
            clones[c].first_exact = (int64_t)ex_ncells_all.size();
            int64_t left = sizes[c];
            if (!bcr) {
                // TCR: no SHM — one dominant exact plus rare sequencing-error variants
                int64_t nvar = 0;
                for (int64_t t = 0; t < left - 1; t++)
                    if (main_rng.uni() < 0.001) nvar++;
                ex_ncells_all.push_back((int32_t)(left - nvar));
                for (int64_t t = 0; t < nvar; t++) ex_ncells_all.push_back(1);
            } else {
                while (left > 0) {
                    int64_t k = 1;
                    while (k < left && main_rng.uni() > p) k++;
                    ex_ncells_all.push_back((int32_t)k);
                    left -= k;
                }
            }
            clones[c].n_exacts = (int64_t)ex_ncells_all.size() - clones[c].first_exact;
        }
    }
    const int64_t n_ex_base = (int64_t)ex_ncells_all.size();
    std::vector<ExactRec> ex(n_ex_base);

    // ---- per clone: recombination + lineage (parallel, deterministic per clone) ----
    int nthreads = cfg->threads > 0 ? cfg->threads : (int)std::thread::hardware_concurrency();
    if (nthreads <= 0) nthreads = 1;
    std::atomic<int64_t> next_clone(0);
    auto clone_worker = [&]() {
        for (;;) {
            int64_t c0 = next_clone.fetch_add(64);
            if (c0 >= n_clones) break;
            for (int64_t c = c0; c < std::min(n_clones, c0 + 64); c++) {
                Rng r(clone_seed[c]);
                CloneRec& cl = clones[c];
                const int lloc = bcr ? (r.uni() < 0.6 ? 1 : 2) : 1;
                cl.h = recombine(r, g, 0, pick_v(r, 0), (int)r.below((int64_t)g.j[0].size()), true);
                cl.l = recombine(r, g, lloc, pick_v(r, lloc), (int)r.below((int64_t)g.j[lloc].size()), false);
                const int lloc2 = bcr ? (r.uni() < 0.6 ? 1 : 2) : 1;
                cl.l2 = recombine(r, g, lloc2, pick_v(r, lloc2), (int)r.below((int64_t)g.j[lloc2].size()), false);
                cl.donor = (int)r.below(cfg->n_donors);
                cl.alt_allele = r.uni() < cfg->frac_alt_allele;
                cl.c_none = r.uni() < 0.02;
                const VGene& vh = g.v[0][cl.h.vi];
                const VGene& vl = g.v[lloc][cl.l.vi];
                const int64_t ne = cl.n_exacts;
                const double delta = cfg->shm_min + (cfg->shm_max - cfg->shm_min) * r.uni();
                const double depth = 1.0 + std::log((double)std::max<int64_t>(1, ne));
                const int Lh = (int)cl.h.seq.size(), Ll = (int)cl.l.seq.size();
                std::string rh = cl.h.seq, rl = cl.l.seq;
                if (bcr) {  // trunk mutations shared by the whole clone
                    mutate_into(r, rh, r.poisson(0.5 * delta * Lh), vh, 0);
                    mutate_into(r, rl, r.poisson(0.35 * delta * Ll), vl, 0);
                }
                const bool three = ne > 0 && r.uni() < cfg->frac_three_chain * 1.0;
                for (int64_t e = 0; e < ne; e++) {
                    ExactRec& x = ex[cl.first_exact + e];
                    x.clone = (int32_t)c;
                    x.ncells = ex_ncells_all[cl.first_exact + e];
                    x.nchain_kind = 0;
                    x.del_kind = 0;
                    x.del_pos = x.del_len = 0;
                    if (e == 0) {
                        x.h = rh;
                        x.l = rl;
                    } else {
                        const ExactRec& par = ex[cl.first_exact + r.below(e)];
                        x.h = par.h;
                        x.l = par.l;
                        if (bcr) {
                            int mh = r.poisson(0.5 * delta * Lh / depth), ml = r.poisson(0.35 * delta * Ll / depth);
                            if (mh + ml == 0) { if (r.uni() < 0.6) mh = 1; else ml = 1; }
                            mutate_into(r, x.h, mh, vh, 0);
                            mutate_into(r, x.l, ml, vl, 0);
                        } else {
                            // sequencing error variant: 1-2 substitutions outside the CDR3s
                            int nm = (int)r.range(1, 2);
                            for (int t = 0; t < nm; t++) {
                                std::string& s = (r.uni() < 0.5) ? x.h : x.l;
                                const Recomb& rc = (&s == &x.h) ? cl.h : cl.l;
                                int p;
                                do { p = (int)r.below((int64_t)s.size()); } while (p >= rc.cdr3_start && p < rc.cdr3_start + rc.cdr3_len);
                                char o = s[p];
                                while (s[p] == o) s[p] = kB[r.below(4)];
                            }
                        }
                    }
                    // rare structural variants
                    const double q = r.uni();
                    if (three && e == 0) {
                        x.nchain_kind = 1;
                        x.l2 = cl.l2.seq;
                    } else if (q < cfg->frac_onesie) {
                        x.nchain_kind = (r.uni() < 0.5) ? 2 : 3;
                    } else if (bcr && q < cfg->frac_onesie + cfg->frac_has_del) {
                        if (r.uni() < 0.75) {
                            x.del_kind = 1;
                            x.del_len = (int16_t)(3 * r.range(1, 3));
                            x.del_pos = (int16_t)r.range(vh.fr1 + 10, vh.fr3 + 60);
                        } else {
                            x.del_kind = 2;
                        }
                    }
                }
            }
        }
    };
    {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; t++) th.emplace_back(clone_worker);
        for (auto& t : th) t.join();
    }

    // ---- flatten: exacts, chains, cells ----
    ejb_synth_data* D = new ejb_synth_data();
    // refs
    D->ref_off.push_back(0);
    for (size_t i = 0; i < g.ref_seq.size(); i++) {
        D->ref_bytes.insert(D->ref_bytes.end(), g.ref_seq[i].begin(), g.ref_seq[i].end());
        D->ref_off.push_back((int64_t)D->ref_bytes.size());
        D->ref_name_id.push_back(g.ref_name[i]);
    }
    // segs: every V and J ref is a seg (index = 2*?); extra segs for alt alleles / deletions are appended
    std::unordered_map<std::string, int> seg_index;
    D->seg_off.push_back(0);
    auto add_seg = [&](const std::string& s) -> int {
        auto it = seg_index.find(s);
        if (it != seg_index.end()) return it->second;
        int id = (int)D->seg_off.size() - 1;
        D->seg_bytes.insert(D->seg_bytes.end(), s.begin(), s.end());
        D->seg_off.push_back((int64_t)D->seg_bytes.size());
        seg_index.emplace(s, id);
        return id;
    };

    // barcodes: per dataset a random permutation of the 737,280-entry whitelist
    const uint32_t WL = 737280;
    std::vector<std::vector<uint32_t>> perm(cfg->n_datasets);
    std::vector<size_t> perm_pos(cfg->n_datasets, 0);
    for (int dsi = 0; dsi < cfg->n_datasets; dsi++) {
        auto& p = perm[dsi];
        p.resize(WL);
        for (uint32_t i = 0; i < WL; i++) p[i] = i;
        Rng r(cfg->seed * 7919 + (uint64_t)dsi);
        for (uint32_t i = WL - 1; i > 0; i--) std::swap(p[i], p[(uint32_t)r.below((int64_t)i + 1)]);
    }
    std::vector<std::vector<int>> donor_ds(cfg->n_donors);
    for (int dsi = 0; dsi < cfg->n_datasets; dsi++) donor_ds[dsi % cfg->n_donors].push_back(dsi);
    for (auto& v : donor_ds)
        if (v.empty()) v.push_back(0);

    struct RowTmp {
        int32_t exact;
        int32_t nch;
        int32_t col[2];
        int32_t vs[2], js[2];
        int32_t cdr3_start[2], cdr3_len[2];
        uint8_t has_del[2];
        std::string tig[2];
        int32_t clone;
    };
    std::vector<RowTmp> rows;
    rows.reserve(n_ex_base + n_ex_base / 16);

    D->ex_chain_off.push_back(0);
    D->ex_cell_off.push_back(0);
    Rng cell_rng(cfg->seed ^ 0xCE11ull);
    for (int64_t e = 0; e < n_ex_base; e++) {
        const ExactRec& x = ex[e];
        const CloneRec& cl = clones[x.clone];
        // chains of the exact, in share order: heavy first
        struct Ch { const std::string* s; const Recomb* rc; bool left; };
        std::vector<Ch> chs;
        if (x.nchain_kind != 3) chs.push_back({&x.h, &cl.h, true});
        if (x.nchain_kind != 2) chs.push_back({&x.l, &cl.l, false});
        if (x.nchain_kind == 1) chs.push_back({&x.l2, &cl.l2, false});
        std::vector<std::string> tigs(chs.size());
        std::vector<int> segv(chs.size()), segj(chs.size());
        std::vector<uint8_t> hd(chs.size(), 0);
        for (size_t j = 0; j < chs.size(); j++) {
            const VGene& v = g.v[chs[j].rc->locus][chs[j].rc->vi];
            const JGene& jj = g.j[chs[j].rc->locus][chs[j].rc->ji];
            tigs[j] = *chs[j].s;
            std::string vseg = v.seq;
            int64_t amino_off = EJB_NONE;
            int32_t amino_len = 0;
            if (chs[j].left && cl.alt_allele) {  // donor allele substituted by sub_alts (allele.rs:566-567)
                vseg[40] = (vseg[40] == 'A') ? 'C' : 'A';
                vseg[200] = (vseg[200] == 'G') ? 'T' : 'G';
            }
            if (chs[j].left && x.del_kind == 1) {
                // V deletion: '-' inserted into tigs (info.rs:78-110), vs = ref with the deleted
                // bases removed (info.rs:199-207), seq_del_amino shifted to a codon boundary
                hd[j] = 1;
                for (int t = 0; t < x.del_len; t++) tigs[j][x.del_pos + t] = '-';
                vseg = vseg.substr(0, x.del_pos) + vseg.substr(x.del_pos + x.del_len);
                if (x.del_pos % 3 != 0) {
                    std::string am = *chs[j].s;
                    const int sh = 3 - x.del_pos % 3;
                    for (int t = 0; t < x.del_len; t++) am[x.del_pos + sh + t] = '-';
                    amino_off = (int64_t)D->amino_bytes.size();
                    amino_len = (int32_t)am.size();
                    D->amino_bytes.insert(D->amino_bytes.end(), am.begin(), am.end());
                }
            } else if (chs[j].left && x.del_kind == 2) {
                hd[j] = 1;  // V insertion removed from tigs, has_del still set (info.rs:125)
            }
            segv[j] = add_seg(vseg);
            segj[j] = add_seg(jj.seq);
            D->ch_left.push_back(chs[j].left ? 1 : 0);
            D->ch_c_ref.push_back((chs[j].left || cl.c_none) ? EJB_NONE : jj.c_ref);
            D->ch_v_ref.push_back(v.ref_id);
            D->ch_u_ref.push_back(((v.ref_id * 2654435761u) % 10 == 0) ? EJB_NONE : v.u_ref_id);
            D->ch_fr1.push_back(v.fr1);
            D->ch_cdr1.push_back(v.cdr1);
            D->ch_fr2.push_back(v.fr2);
            D->ch_cdr2.push_back(v.cdr2);
            D->ch_fr3.push_back(((v.ref_id * 40503u) % 97 == 0) ? EJB_NONE : v.fr3);
            D->ch_hcomp.push_back(chs[j].left && chs.size() == 2 ? chs[j].rc->hcomp : 0);
            D->ch_amino_off.push_back(amino_off);
            D->ch_amino_len.push_back(amino_len);
        }
        D->ex_chain_off.push_back((int64_t)D->ch_left.size());
        // cells
        for (int c = 0; c < x.ncells; c++) {
            int donor = cl.donor;
            if (cell_rng.uni() < cfg->frac_contam) donor = (int)cell_rng.below(cfg->n_donors);
            const auto& dsl = donor_ds[donor];
            int dsi = dsl[cell_rng.below((int64_t)dsl.size())];
            if (perm_pos[dsi] >= WL) {  // whitelist exhausted: take any dataset with room
                for (int t = 0; t < cfg->n_datasets; t++)
                    if (perm_pos[t] < WL) { dsi = t; break; }
            }
            uint32_t bc = perm_pos[dsi] < WL ? perm[dsi][perm_pos[dsi]++] : (uint32_t)(WL + D->cell_barcode.size());
            D->cell_donor.push_back(dsi % cfg->n_donors);
            D->cell_barcode.push_back(bc);
        }
        D->ex_cell_off.push_back((int64_t)D->cell_barcode.size());
        // rows: one per left x right pair (info.rs:277-319), else a onesie row (info.rs:334-359)
        bool placed = false;
        for (size_t i1 = 0; i1 < chs.size(); i1++)
            if (chs[i1].left)
                for (size_t i2 = 0; i2 < chs.size(); i2++)
                    if (!chs[i2].left) {
                        placed = true;
                        RowTmp rt;
                        rt.exact = (int32_t)e;
                        rt.nch = 2;
                        rt.clone = x.clone;
                        const size_t idx[2] = {i1, i2};
                        for (int m = 0; m < 2; m++) {
                            rt.col[m] = (int32_t)idx[m];
                            rt.vs[m] = segv[idx[m]];
                            rt.js[m] = segj[idx[m]];
                            rt.has_del[m] = hd[idx[m]];
                            rt.tig[m] = tigs[idx[m]];
                            rt.cdr3_start[m] = chs[idx[m]].rc->cdr3_start;
                            rt.cdr3_len[m] = chs[idx[m]].rc->cdr3_len;
                        }
                        rows.push_back(std::move(rt));
                    }
        if (!placed && chs.size() == 1) {
            RowTmp rt;
            rt.exact = (int32_t)e;
            rt.nch = 1;
            rt.clone = x.clone;
            rt.col[0] = 0; rt.col[1] = 0;
            rt.vs[0] = segv[0]; rt.js[0] = segj[0]; rt.vs[1] = rt.js[1] = 0;
            rt.has_del[0] = hd[0]; rt.has_del[1] = 0;
            rt.tig[0] = tigs[0];
            rt.cdr3_start[0] = chs[0].rc->cdr3_start; rt.cdr3_len[0] = chs[0].rc->cdr3_len;
            rt.cdr3_start[1] = rt.cdr3_len[1] = 0;
            rows.push_back(std::move(rt));
        }
    }
    ex.clear();
    ex.shrink_to_fit();

    // ---- sort rows like info.par_sort(): lens (as a vector), then tigs, then clonotype_id ----
    const int64_t N = (int64_t)rows.size();
    std::vector<int64_t> order(N);
    for (int64_t i = 0; i < N; i++) order[i] = i;
    auto lens_key = [&](int64_t i) -> uint64_t {
        const RowTmp& r = rows[i];
        // [a] sorts before [a, b]; encode the second length + 1 (0 = absent)
        return ((uint64_t)r.tig[0].size() << 32) | (uint64_t)(r.nch == 2 ? r.tig[1].size() + 1 : 0);
    };
    {
        std::vector<uint64_t> keys(N);
        for (int64_t i = 0; i < N; i++) keys[i] = lens_key(i);
        std::sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return keys[a] < keys[b] || (keys[a] == keys[b] && a < b); });
        // sort inside each lens group in parallel
        std::vector<std::pair<int64_t, int64_t>> groups;
        for (int64_t i = 0; i < N;) {
            int64_t j = i;
            while (j < N && keys[order[j]] == keys[order[i]]) j++;
            groups.emplace_back(i, j);
            i = j;
        }
        std::sort(groups.begin(), groups.end(), [](auto& a, auto& b) { return (a.second - a.first) > (b.second - b.first); });
        std::atomic<size_t> gi(0);
        auto sorter = [&]() {
            for (;;) {
                size_t q = gi.fetch_add(1);
                if (q >= groups.size()) break;
                std::sort(order.begin() + groups[q].first, order.begin() + groups[q].second, [&](int64_t a, int64_t b) {
                    const RowTmp &x = rows[a], &y = rows[b];
                    int c = x.tig[0].compare(y.tig[0]);
                    if (c) return c < 0;
                    if (x.nch == 2) {
                        c = x.tig[1].compare(y.tig[1]);
                        if (c) return c < 0;
                    }
                    if (x.has_del[0] != y.has_del[0]) return x.has_del[0] < y.has_del[0];
                    if (x.exact != y.exact) return x.exact < y.exact;
                    return a < b;
                });
            }
        };
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; t++) th.emplace_back(sorter);
        for (auto& t : th) t.join();
    }

    // ---- emit rows ----
    D->row_nchains.resize(N);
    D->row_exact.resize(N);
    D->row_clone.resize(N);
    D->row_len.assign(2 * N, 0);
    D->row_tig_off.assign(2 * N, 0);
    D->row_has_del.assign(2 * N, 0);
    D->row_vs.assign(2 * N, 0);
    D->row_js.assign(2 * N, 0);
    D->row_cdr3_off.assign(2 * N, 0);
    D->row_cdr3_len.assign(2 * N, 0);
    D->row_exact_col.assign(2 * N, 0);
    {
        size_t tb = 0, cb = 0;
        for (auto& r : rows) {
            tb += r.tig[0].size() + (r.nch == 2 ? r.tig[1].size() : 0);
            cb += r.cdr3_len[0] + r.cdr3_len[1];
        }
        D->tig_bytes.reserve(tb + 16);
        D->cdr3_bytes.reserve(cb + 16);
    }
    for (int64_t k = 0; k < N; k++) {
        RowTmp& r = rows[order[k]];
        D->row_nchains[k] = r.nch;
        D->row_exact[k] = r.exact;
        D->row_clone[k] = r.clone;
        for (int m = 0; m < r.nch; m++) {
            D->row_len[2 * k + m] = (int32_t)r.tig[m].size();
            D->row_tig_off[2 * k + m] = (int64_t)D->tig_bytes.size();
            D->tig_bytes.insert(D->tig_bytes.end(), r.tig[m].begin(), r.tig[m].end());
            D->row_has_del[2 * k + m] = r.has_del[m];
            D->row_vs[2 * k + m] = r.vs[m];
            D->row_js[2 * k + m] = r.js[m];
            D->row_exact_col[2 * k + m] = r.col[m];
            // cdr3_dna comes from the unedited seq (x.cdr3_dna), never contains '-'
            D->row_cdr3_off[2 * k + m] = (int64_t)D->cdr3_bytes.size();
            D->row_cdr3_len[2 * k + m] = r.cdr3_len[m];
            for (int t = 0; t < r.cdr3_len[m]; t++) {
                char c = r.tig[m][r.cdr3_start[m] + t];
                D->cdr3_bytes.push_back((uint8_t)(c == '-' ? 'A' : c));
            }
            std::string().swap(r.tig[m]);
        }
    }
    rows.clear();

    // ---- wire up the ejb_input view ----
    ejb_input& in = D->in;
    memset(&in, 0, sizeof(in));
    in.n_rows = N;
    in.row_nchains = D->row_nchains.data();
    in.row_len = D->row_len.data();
    in.row_tig_off = D->row_tig_off.data();
    in.row_has_del = D->row_has_del.data();
    in.row_vs = D->row_vs.data();
    in.row_js = D->row_js.data();
    in.row_cdr3_off = D->row_cdr3_off.data();
    in.row_cdr3_len = D->row_cdr3_len.data();
    in.row_exact = D->row_exact.data();
    in.row_exact_col = D->row_exact_col.data();
    in.tig_bytes = D->tig_bytes.data();
    in.n_tig_bytes = (int64_t)D->tig_bytes.size();
    in.cdr3_bytes = D->cdr3_bytes.data();
    in.n_cdr3_bytes = (int64_t)D->cdr3_bytes.size();
    in.n_exacts = n_ex_base;
    in.ex_chain_off = D->ex_chain_off.data();
    in.ex_cell_off = D->ex_cell_off.data();
    in.n_chains = (int64_t)D->ch_left.size();
    in.ch_left = D->ch_left.data();
    in.ch_c_ref = D->ch_c_ref.data();
    in.ch_v_ref = D->ch_v_ref.data();
    in.ch_u_ref = D->ch_u_ref.data();
    in.ch_fr1_start = D->ch_fr1.data();
    in.ch_cdr1_start = D->ch_cdr1.data();
    in.ch_fr2_start = D->ch_fr2.data();
    in.ch_cdr2_start = D->ch_cdr2.data();
    in.ch_fr3_start = D->ch_fr3.data();
    in.ch_hcomp = D->ch_hcomp.data();
    in.ch_amino_off = D->ch_amino_off.data();
    in.ch_amino_len = D->ch_amino_len.data();
    D->amino_bytes.push_back(0);
    in.amino_bytes = D->amino_bytes.data();
    in.n_amino_bytes = (int64_t)D->amino_bytes.size() - 1;
    in.n_cells = (int64_t)D->cell_barcode.size();
    in.cell_donor = D->cell_donor.data();
    in.cell_barcode = D->cell_barcode.data();
    in.n_segs = (int64_t)D->seg_off.size() - 1;
    in.seg_off = D->seg_off.data();
    in.seg_bytes = D->seg_bytes.data();
    in.n_refs = (int64_t)D->ref_off.size() - 1;
    in.ref_off = D->ref_off.data();
    in.ref_bytes = D->ref_bytes.data();
    in.ref_name_id = D->ref_name_id.data();

    D->sum.n_cells = in.n_cells;
    D->sum.n_clones = n_clones;
    D->sum.n_exacts = n_ex_base;
    D->sum.n_rows = N;
    D->sum.max_clone_cells = *std::max_element(sizes.begin(), sizes.end());
    int64_t mx = 0;
    for (auto& c : clones) mx = std::max(mx, c.n_exacts);
    D->sum.max_clone_exacts = mx;
    D->sum.tig_bytes = in.n_tig_bytes;
    *out_data = D;
    return EJB_OK;
}

const ejb_input* ejb_synth_input(const ejb_synth_data* d) { return &d->in; }
const int32_t* ejb_synth_row_clone(const ejb_synth_data* d) { return d->row_clone.data(); }
void ejb_synth_get_summary(const ejb_synth_data* d, ejb_synth_summary* s) { *s = d->sum; }
void ejb_synth_free(ejb_synth_data* d) { delete d; }

}  // extern "C"

// ------------------------------------------------------------------------------------------------------------
// Host helper: 2-bit pack every eligible contig of a flattened input (pure upper-case ACGT, has_del == 0) the way
// include/enclone_join_b200.h expects it (16 bases per uint32, base i at bits 2*(i%16), A,C,G,T = 0..3); other
// contigs stay ASCII.  Returns malloc'ed arrays (free with ejb_synth_free_array).
// ------------------------------------------------------------------------------------------------------------
extern "C" int ejb_synth_pack_contigs(const ejb_input* in, int64_t* row_tig2_off /* [2*n_rows] out */, int64_t* new_tig_off /* [2*n_rows] out */,
                                      uint32_t** words_out, int64_t* n_words_out, uint8_t** bytes_out, int64_t* n_bytes_out) {
    const int64_t n2 = 2 * in->n_rows;
    static int8_t lut[256];
    static bool lut_init = false;
    if (!lut_init) {
        for (int i = 0; i < 256; i++) lut[i] = -1;
        lut[(int)'A'] = 0; lut[(int)'C'] = 1; lut[(int)'G'] = 2; lut[(int)'T'] = 3;
        lut_init = true;
    }
    // pass 1: eligibility and sizes
    std::vector<char> elig((size_t)n2, 0);
    int64_t nw = 0, nb = 0;
    for (int64_t i = 0; i < n2; i++) {
        row_tig2_off[i] = -1;
        new_tig_off[i] = -1;
        if ((i & 1) >= in->row_nchains[i >> 1]) continue;
        const int64_t L = in->row_len[i];
        if (L <= 0) continue;
        const uint8_t* t = in->tig_bytes + in->row_tig_off[i];
        bool ok = in->row_has_del[i] == 0;
        for (int64_t p = 0; ok && p < L; p++) ok = lut[t[p]] >= 0;
        elig[(size_t)i] = ok;
        if (ok) { row_tig2_off[i] = nw; nw += (L + 15) / 16; } else { new_tig_off[i] = nb; nb += L; }
    }
    uint32_t* words = (uint32_t*)malloc(std::max<int64_t>(nw, 1) * 4);
    uint8_t* bytes = (uint8_t*)malloc(std::max<int64_t>(nb, 1));
    if (!words || !bytes) { free(words); free(bytes); return EJB_ERR_OOM; }
    for (int64_t i = 0; i < n2; i++) {
        if ((i & 1) >= in->row_nchains[i >> 1]) continue;
        const int64_t L = in->row_len[i];
        if (L <= 0) continue;
        const uint8_t* t = in->tig_bytes + in->row_tig_off[i];
        if (elig[(size_t)i]) {
            uint32_t* w = words + row_tig2_off[i];
            for (int64_t q = 0; q < (L + 15) / 16; q++) {
                uint32_t v = 0;
                const int64_t m = std::min<int64_t>(16, L - 16 * q);
                for (int64_t j = 0; j < m; j++) v |= (uint32_t)lut[t[16 * q + j]] << (2 * j);
                w[q] = v;
            }
        } else {
            memcpy(bytes + new_tig_off[i], t, (size_t)L);
        }
    }
    *words_out = words; *n_words_out = nw; *bytes_out = bytes; *n_bytes_out = nb;
    return EJB_OK;
}
// Same for the CDR3s (info[k].cdr3s[m] -> 2-bit words, ejb_input.row_cdr32_off / cdr32_words): a CDR3 with a byte outside
// upper-case ACGT stays ASCII (offset -1 in row_cdr32_off, its bytes are re-packed into bytes_out).
extern "C" int ejb_synth_pack_cdr3s(const ejb_input* in, int64_t* row_cdr32_off /* [2*n_rows] out */, int64_t* new_cdr3_off /* [2*n_rows] out */,
                                    uint32_t** words_out, int64_t* n_words_out, uint8_t** bytes_out, int64_t* n_bytes_out) {
    const int64_t n2 = 2 * in->n_rows;
    int8_t lut[256];
    for (int i = 0; i < 256; i++) lut[i] = -1;
    lut[(int)'A'] = 0; lut[(int)'C'] = 1; lut[(int)'G'] = 2; lut[(int)'T'] = 3;
    std::vector<char> elig((size_t)n2, 0);
    int64_t nw = 0, nb = 0;
    for (int64_t i = 0; i < n2; i++) {
        row_cdr32_off[i] = -1;
        new_cdr3_off[i] = -1;
        if ((i & 1) >= in->row_nchains[i >> 1]) continue;
        const int64_t L = in->row_cdr3_len[i];
        const uint8_t* t = in->cdr3_bytes + in->row_cdr3_off[i];
        bool ok = true;
        for (int64_t p = 0; ok && p < L; p++) ok = lut[t[p]] >= 0;
        elig[(size_t)i] = ok;
        if (ok) { row_cdr32_off[i] = nw; nw += (L + 15) / 16; } else { new_cdr3_off[i] = nb; nb += L; }
    }
    uint32_t* words = (uint32_t*)malloc(std::max<int64_t>(nw, 1) * 4);
    uint8_t* bytes = (uint8_t*)malloc(std::max<int64_t>(nb, 1));
    if (!words || !bytes) { free(words); free(bytes); return EJB_ERR_OOM; }
    for (int64_t i = 0; i < n2; i++) {
        if ((i & 1) >= in->row_nchains[i >> 1]) continue;
        const int64_t L = in->row_cdr3_len[i];
        const uint8_t* t = in->cdr3_bytes + in->row_cdr3_off[i];
        if (elig[(size_t)i]) {
            uint32_t* w = words + row_cdr32_off[i];
            for (int64_t q = 0; q < (L + 15) / 16; q++) {
                uint32_t v = 0;
                const int64_t m = std::min<int64_t>(16, L - 16 * q);
                for (int64_t j = 0; j < m; j++) v |= (uint32_t)lut[t[16 * q + j]] << (2 * j);
                w[q] = v;
            }
        } else if (L > 0) {
            memcpy(bytes + new_cdr3_off[i], t, (size_t)L);
        }
    }
    *words_out = words; *n_words_out = nw; *bytes_out = bytes; *n_bytes_out = nb;
    return EJB_OK;
}
extern "C" void ejb_synth_free_array(void* p) { free(p); }
