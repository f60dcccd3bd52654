// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/dd.hpp header).  Never linked into or called by
// the product library; used by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.
//
// A CPU restatement, in C++17, of the reference's clonotype join on the flattened input of
// include/enclone_join_b200.h.  The real reference is pure Rust and cannot be built here (no
// rustc/cargo, no vendored crates), so this follows the source line by line:
//
//   EquivRel                       equiv/src/lib.rs:34-138
//   stirling2_ratio_table_double   enclone_stuff/src/start.rs:50-99
//   p_at_most_m_distinct_..._double enclone_core/src/join_one.rs:22-43
//   join_one                       enclone_core/src/join_one.rs:66-111, 214-567, 708-887
//                                  (alternative modes :113-212 and SUPER_COMP_FILT :568-705 are
//                                   rejected, they are unreachable from enclone_ranger)
//   join_core                      enclone/src/join_core.rs:11-51
//   join_exacts                    enclone/src/join.rs:68-214, 582-589 (logging :216-579 skipped,
//                                   it is off under the default quiet=true)
//   finish_join                    enclone/src/join2.rs:36-87
//   helpers                        vector_utils/src/lib.rs:38-47,115-132,134-150,352-354;
//                                  stats_utils/src/lib.rs:116-121
//
// PARITY: pinned at f64 level by the Stirling KATs (stirling_numbers/src/lib.rs:261-315) and
// hand-built micro-cases per decision row; UNPINNED at double-double bit level (qd source
// absent) and for debruijn's 2-bit semantics (restated: A,C,G,T -> 0..3, anything else -> 0).
// The reference holds no test or fixture of join_exacts / join_one / EquivRel.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../include/enclone_join_b200.h"
#include "dd.hpp"

namespace oracle {

// ---------------------------------------------------------------------------------------------
// EquivRel (equiv/src/lib.rs:34-138)
// ---------------------------------------------------------------------------------------------
struct EquivRel {
    std::vector<int32_t> x, y, z;  // next in orbit, class id, orbit size
    explicit EquivRel(int32_t n) : x(n), y(n), z(n) {
        for (int32_t i = 0; i < n; i++) {
            x[i] = i;
            y[i] = i;
            z[i] = 1;
        }
    }
    int32_t orbit_size(int32_t a) const { return z[y[a]]; }
    int32_t class_id(int32_t a) const { return y[a]; }
    void join(int32_t a, int32_t b) {
        int32_t ax = a, bx = b;
        if (y[ax] != y[bx]) {
            if (orbit_size(ax) < orbit_size(bx)) std::swap(ax, bx);  // always move the smaller orbit
            int32_t new_size = orbit_size(ax) + orbit_size(bx);
            std::swap(x[ax], x[bx]);
            int32_t n = x[ax];
            for (;;) {
                if (y[n] == y[ax]) break;
                y[n] = y[ax];
                n = x[n];
            }
            z[y[bx]] = new_size;
        }
    }
    void orbit_reps(std::vector<int32_t>& reps) const {
        reps.clear();
        for (size_t i = 0; i < x.size(); i++)
            if ((int32_t)i == y[i]) reps.push_back((int32_t)i);
    }
    void orbit(int32_t a, std::vector<int32_t>& o) const {
        o.clear();
        o.push_back(a);
        int32_t b = a;
        for (;;) {
            b = x[b];
            if (b == a) break;
            o.push_back(b);
        }
    }
};

// ---------------------------------------------------------------------------------------------
// Stirling ratio table (start.rs:50-99): s[n][k] = S(n,k) * k! / k^n, in double-double.
// Stored triangular: row n starts at n(n+1)/2.
// ---------------------------------------------------------------------------------------------
struct SrTable {
    int n_max = -1;
    std::vector<DD> s;
    const DD& at(int n, int k) const { return s[(size_t)n * (n + 1) / 2 + k]; }
    DD& at(int n, int k) { return s[(size_t)n * (n + 1) / 2 + k]; }
};

static void build_sr(SrTable& t, int n_max) {
    t.n_max = n_max;
    t.s.assign((size_t)(n_max + 1) * (n_max + 2) / 2, DD(0.0));
    const DD zero(0.0), one(1.0);
    t.at(0, 0) = one;
    std::vector<DD> z;
    std::vector<DD> n2n1(n_max + 1, DD(0.0));
    for (int n = 2; n <= n_max; n++) n2n1[n] = dd_div(DD((double)(n - 2)), DD((double)(n - 1)));
    std::vector<DD> k1k(n_max, DD(0.0));
    for (int k = 1; k < n_max; k++) k1k[k] = dd_div(DD((double)(k - 1)), DD((double)k));
    std::vector<DD> njn(n_max + 1, DD(0.0));
    for (int n = 1; n <= n_max; n++) {
        DD p = one;
        for (int j = 1; j <= n; j++) p = dd_mul(p, dd_div(DD((double)j), DD((double)n)));
        njn[n] = p;
    }
    for (int n = 1; n <= n_max; n++) {
        t.at(n, 0) = zero;
        for (int k = 1; k + 1 < n; k++) z[k - 1] = dd_mul(z[k - 1], k1k[k]);
        if (n >= 2) z.push_back(dd_powi(n2n1[n], n - 1));
        for (int k = 1; k < n; k++) {
            DD x = z[k - 1];  // ((k-1)/k)^(n-1)
            t.at(n, k) = dd_add(t.at(n - 1, k), dd_mul(t.at(n - 1, k - 1), x));
        }
        t.at(n, n) = njn[n];
    }
}

static SrTable g_sr;
static std::once_flag g_sr_once;
static const SrTable& sr_table() {
    std::call_once(g_sr_once, [] { build_sr(g_sr, 3000); });  // start.rs:344
    return g_sr;
}

// join_one.rs:22-43.  Returns false where the reference would panic (index out of range).
static bool p_at_most_m_distinct(int64_t m, int64_t x, int64_t n, const SrTable& sr, double& out) {
    if (x > sr.n_max || x < 0 || m < 0) return false;
    DD p(1.0);
    for (int64_t u = m + 1; u <= x; u++) {
        DD z = sr.at((int)x, (int)u);
        for (int64_t i = 0; i < x; i++) z = dd_mul(z, dd_div(DD((double)u), DD((double)n)));
        for (int64_t v = 1; v <= u; v++) {
            if (v > n) return false;  // usize underflow of (n - v) + 1: Rust evaluates n - v first
            z = dd_mul(z, dd_div(DD((double)(n - v + 1)), DD((double)(u - v + 1))));
        }
        p = dd_sub(p, z);
    }
    if (dd_lt(p, DD(0.0))) p = DD(0.0);
    out = p.hi;
    return true;
}

// Digest generation only (tests/golden/make_digests.py sets ORACLE_P1_MEMO=1): p1 is a pure function of (k, d, n), so a
// per-thread memo returns the very same bits while sparing the full-size runs the O(d (k + u)) dd loop per call.  The
// timed CPU baselines never set it: the reference evaluates p1 afresh for every pair (join_one.rs:499-502).
static bool p1_maybe_memo(int64_t k, int64_t d, int64_t n, const SrTable& sr, double& out) {
    static const bool memo = getenv("ORACLE_P1_MEMO") != nullptr;
    if (!memo) return p_at_most_m_distinct(k - d, k, n, sr, out);
    thread_local std::unordered_map<uint64_t, double> cache;
    if (k < 0 || k > sr.n_max || d < 0 || d > k || n < 0 || n >= (1ll << 32)) return p_at_most_m_distinct(k - d, k, n, sr, out);
    const uint64_t key = ((uint64_t)n << 24) | ((uint64_t)k << 12) | (uint64_t)d;
    auto it = cache.find(key);
    if (it != cache.end()) {
        out = it->second;
        return true;
    }
    if (!p_at_most_m_distinct(k - d, k, n, sr, out)) return false;
    cache.emplace(key, out);
    return true;
}

// ---------------------------------------------------------------------------------------------
// Views over the flattened input.
// ---------------------------------------------------------------------------------------------
struct In {
    const ejb_input* in;
    const ejb_params* p;
    std::vector<int32_t> seg_canon;  // content-canonical seg id (DnaString == compares content)
    int nchains(int64_t k) const { return in->row_nchains[k]; }
    int len(int64_t k, int m) const { return in->row_len[2 * k + m]; }
    const uint8_t* tig(int64_t k, int m) const { return in->tig_bytes + in->row_tig_off[2 * k + m]; }
    bool has_del(int64_t k, int m) const { return in->row_has_del[2 * k + m] != 0; }
    const uint8_t* cdr3(int64_t k, int m) const { return in->cdr3_bytes + in->row_cdr3_off[2 * k + m]; }
    int cdr3_len(int64_t k, int m) const { return in->row_cdr3_len[2 * k + m]; }
    const uint8_t* seg(int s) const { return in->seg_bytes + in->seg_off[s]; }
    int64_t seg_len(int s) const { return in->seg_off[s + 1] - in->seg_off[s]; }
    const uint8_t* ref(int r) const { return in->ref_bytes + in->ref_off[r]; }
    int64_t ref_len(int r) const { return in->ref_off[r + 1] - in->ref_off[r]; }
    int64_t ex_nchains(int e) const { return in->ex_chain_off[e + 1] - in->ex_chain_off[e]; }
    int64_t ex_ncells(int e) const { return in->ex_cell_off[e + 1] - in->ex_cell_off[e]; }
    int64_t chain(int e, int j) const { return in->ex_chain_off[e] + j; }
};

// debruijn: base_to_bits — A,C,G,T (either case) -> 0..3, anything else -> 0.
static inline int base2bit(uint8_t b) {
    switch (b) {
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return 0;
    }
}
static const uint8_t kBits2Base[4] = {'A', 'C', 'G', 'T'};

struct Pot {  // numeric members of PotentialJoin (defs.rs:870-891)
    int64_t k1, k2;
    int nrefs;
    int64_t cd;
    int64_t shares[2], indeps[2];
    double score, p1, mult;
    bool err;
    int64_t k, d, n;
};

enum JoinOneStatus { J_FALSE = 0, J_TRUE = 1, J_PANIC = 2 };

// ---------------------------------------------------------------------------------------------
// join_one (join_one.rs:66-887), default algorithm only.
// ---------------------------------------------------------------------------------------------
static JoinOneStatus join_one(int64_t k1, int64_t k2, const In& I, const SrTable& sr,
                              std::vector<Pot>& pot, std::string& why) {
    const ejb_input& in = *I.in;
    const ejb_params& ctl = *I.p;
    const bool is_bcr = ctl.is_bcr != 0;

    // :83-96
    const int clono1 = in.row_exact[k1], clono2 = in.row_exact[k2];
    const int64_t chains1 = I.ex_nchains(clono1), chains2 = I.ex_nchains(clono2);
    if (!(chains1 >= 2 && chains1 <= 3) || !(chains2 >= 2 && chains2 <= 3)) return J_FALSE;
    const int nv1 = I.nchains(k1), nv2 = I.nchains(k2);  // info.vs.len() == info.lens.len()
    if (nv1 == 1 || nv2 == 4) return J_FALSE;
    if (nv1 > 2) return J_FALSE;

    // :101-109 CDR3s must have the same lengths
    if (nv1 != nv2) return J_FALSE;  // x1.len() != x2.len()
    for (int i = 0; i < nv1; i++)
        if (I.cdr3_len(k1, i) != I.cdr3_len(k2, i)) return J_FALSE;

    // :216-234 identity filter on CDR3s for BCR
    if (is_bcr) {
        int64_t cd = 0, total = 0;
        for (int z = 0; z < 2; z++) {
            const int n = I.cdr3_len(k1, z);
            if (n != I.cdr3_len(k2, z)) return J_FALSE;
            const uint8_t *a = I.cdr3(k1, z), *b = I.cdr3(k2, z);
            for (int m = 0; m < n; m++)
                if (a[m] != b[m]) cd++;
            total += n;
        }
        if ((double)cd / (double)total > 1.0 - ctl.join_cdr3_ident / 100.0) return J_FALSE;
    }

    // :238-267 number of differences; by default applied only to TCR
    if (!is_bcr || ctl.max_diffs < 1000000) {
        int64_t diffs = 0;
        for (int x = 0; x < nv1; x++) {
            const int n = I.len(k1, x);
            if (I.len(k2, x) != n) {
                why = "rows of one bucket with different lens";
                return J_PANIC;
            }
            const uint8_t *a = I.tig(k1, x), *b = I.tig(k2, x);
            if (!I.has_del(k1, x) && !I.has_del(k2, x)) {
                // ndiffs on tigsp (2-bit packing of the unedited seq == tigs when !has_del)
                for (int j = 0; j < n; j++)
                    if (base2bit(a[j]) != base2bit(b[j])) diffs++;
            } else {
                for (int j = 0; j < n; j++)
                    if (a[j] != b[j]) diffs++;
            }
        }
        if (diffs > ctl.max_diffs) return J_FALSE;
        if (!is_bcr && diffs > 5) return J_FALSE;
    }

    // :271-282 junction diffs
    int64_t cd = 0, hcd = 0;
    for (int l = 0; l < nv1; l++) {
        const int n = I.cdr3_len(k1, l);
        const uint8_t *a = I.cdr3(k1, l), *b = I.cdr3(k2, l);
        for (int m = 0; m < n; m++)
            if (a[m] != b[m]) {
                if (l == 0) hcd++;
                cd++;
            }
    }
    (void)hcd;

    // :286-290 cap CDR3 diffs for TCR or as requested
    if ((ctl.max_cdr3_diffs < 1000 || !is_bcr) && (cd > ctl.max_cdr3_diffs || (!is_bcr && cd > 0)))
        return J_FALSE;

    // :301-323 donors
    std::vector<int64_t> donors1, donors2;
    const int ex1 = clono1, ex2 = clono2;  // clonotype_index == clonotype_id
    for (int64_t j = in.ex_cell_off[ex1]; j < in.ex_cell_off[ex1 + 1]; j++)
        if (in.cell_donor[j] != EJB_NONE) donors1.push_back(in.cell_donor[j]);
    for (int64_t j = in.ex_cell_off[ex2]; j < in.ex_cell_off[ex2 + 1]; j++)
        if (in.cell_donor[j] != EJB_NONE) donors2.push_back(in.cell_donor[j]);
    auto unique_sort = [](auto& v) {
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end()), v.end());
    };
    unique_sort(donors1);
    unique_sort(donors2);
    if (!ctl.mix_donors && !donors1.empty() && !donors2.empty() && donors1 != donors2) return J_FALSE;
    const bool err = donors1 != donors2 || donors1.size() != 1 || donors2.size() != 1;

    // :329-334 one or two references
    int nrefs = 1;
    for (int m = 0; m < 2; m++) {
        if (I.seg_canon[in.row_vs[2 * k1 + m]] != I.seg_canon[in.row_vs[2 * k2 + m]] ||
            I.seg_canon[in.row_js[2 * k1 + m]] != I.seg_canon[in.row_js[2 * k2 + m]])
            nrefs = 2;
    }
    int64_t shares[2] = {0, 0}, indeps[2] = {0, 0};
    int64_t total[2][2] = {{0, 0}, {0, 0}};
    // :343-429
    for (int u = 0; u < nrefs; u++) {
        const int64_t k = (u == 0) ? k1 : k2;
        const int nchains = I.nchains(k1);
        for (int m = 0; m < nchains; m++) {
            const uint8_t *tig1 = I.tig(k1, m), *tig2 = I.tig(k2, m);
            const int64_t tl1 = I.len(k1, m), tl2 = I.len(k2, m);
            for (int si = 0; si < 2; si++) {
                const int sg = (si == 0) ? in.row_vs[2 * k + m] : in.row_js[2 * k + m];
                const uint8_t* seg = I.seg(sg);
                const int64_t seg_len = I.seg_len(sg);
                const int64_t ref_trim = (si == 1) ? ctl.ref_j_trim : ctl.ref_v_trim;
                if (seg_len < ref_trim) {
                    why = "seg.len() - ref_trim underflows (join_one.rs:365)";
                    return J_PANIC;
                }
                for (int64_t p = 0; p < seg_len - ref_trim; p++) {
                    uint8_t t1, t2, r;
                    if (si == 0) {
                        if (p >= tl1 || p >= tl2) return J_FALSE;  // "ugly bailout" :371-373
                        t1 = tig1[p];
                        t2 = tig2[p];
                        r = kBits2Base[base2bit(seg[p])];
                    } else {
                        if (p + 1 > tl1 || p + 1 > tl2) {
                            why = "tig shorter than J range (join_one.rs:388-389 index panics)";
                            return J_PANIC;
                        }
                        t1 = tig1[tl1 - p - 1];
                        t2 = tig2[tl2 - p - 1];
                        r = kBits2Base[base2bit(seg[seg_len - p - 1])];
                    }
                    if (t1 == t2 && t1 != r) {
                        shares[u]++;
                    } else if ((t1 == r && t2 != r) || (t2 == r && t1 != r)) {
                        indeps[u]++;
                    } else if (t1 != r && t2 != r) {
                        indeps[u] += 2;
                    }
                    if (t1 != r) total[u][0]++;
                    if (t2 != r) total[u][1]++;
                }
            }
        }
    }

    // :434-440
    if (nrefs == 2) {
        for (int m = 0; m < 2; m++) {
            const int64_t a = total[0][m], b = total[1][m];
            const int64_t ad = (a <= b) ? b - a : a - b;
            if (ad > ctl.max_degradation) return J_FALSE;
        }
    }

    // :444-447
    int64_t min_shares = shares[0], min_indeps = indeps[0];
    if (nrefs == 2) {
        min_shares = std::min(shares[0], shares[1]);
        min_indeps = std::min(indeps[0], indeps[1]);
    }

    // :451-470 reject if barcode overlap.  bcs = union over info.origin (the exact's datasets) of
    // to_bc[(origin, clonotype_id)] = all barcodes of the exact's cells (start.rs:331-340).
    {
        std::vector<uint32_t> bcs1(in.cell_barcode + in.ex_cell_off[ex1], in.cell_barcode + in.ex_cell_off[ex1 + 1]);
        std::vector<uint32_t> bcs2(in.cell_barcode + in.ex_cell_off[ex2], in.cell_barcode + in.ex_cell_off[ex2 + 1]);
        unique_sort(bcs1);
        unique_sort(bcs2);
        size_t i = 0, j = 0;
        while (i < bcs1.size() && j < bcs2.size()) {  // meet()
            if (bcs1[i] < bcs2[j]) i++;
            else if (bcs2[j] < bcs1[i]) j++;
            else return J_FALSE;
        }
    }

    // :474-476 concentration of SHM in the junction regions
    if ((double)cd >= ctl.cdr3_mult * (double)std::max<int64_t>(1, min_indeps)) return J_FALSE;

    // :481-493 different light chain constant regions, unless cd == 0
    if (!ctl.old_light) {
        for (int i = 0; i < nv1; i++) {
            const int64_t c1 = I.chain(ex1, in.row_exact_col[2 * k1 + i]);
            const int64_t c2 = I.chain(ex2, in.row_exact_col[2 * k2 + i]);
            if (!in.ch_left[c1] && in.ch_c_ref[c1] != EJB_NONE && in.ch_c_ref[c2] != EJB_NONE &&
                in.ch_c_ref[c1] != in.ch_c_ref[c2] && cd > 0)
                return J_FALSE;
        }
    }

    // :499-503
    const int64_t n = 3 * ((int64_t)I.len(k1, 0) + (int64_t)I.len(k1, 1));
    const int64_t k = min_indeps + 2 * min_shares;
    const int64_t d = min_shares;
    double p1;
    if (!p1_maybe_memo(k, d, n, sr, p1)) {
        why = "Stirling table index out of range (k > 3000, join_one.rs:30)";
        return J_PANIC;
    }

    // :509-540 (old_mult rejected)
    double mult;
    {
        int64_t cd1 = 0, cd2 = 0;
        const int64_t n1 = I.cdr3_len(k1, 0), n2 = I.cdr3_len(k1, 1);
        const uint8_t *a0 = I.cdr3(k1, 0), *b0 = I.cdr3(k2, 0);
        for (int64_t m = 0; m < n1; m++)
            if (a0[m] != b0[m]) cd1++;
        const uint8_t *a1 = I.cdr3(k1, 1), *b1 = I.cdr3(k2, 1);
        for (int64_t m = 0; m < n2; m++)
            if (a1[m] != b1[m]) cd2++;
        const int64_t cdx = ctl.cdr3_normal_len;
        mult = std::pow(ctl.mult_pow, (double)cdx * (double)cd1 / (double)n1);
        mult *= std::pow(ctl.mult_pow, (double)cdx * (double)cd2 / (double)n2);
    }

    // :544
    const double score = p1 * mult;

    // :548-567 JUN_SHARE
    bool accept = false;
    const int64_t c1_0 = I.chain(ex1, 0), c1_1 = I.chain(ex1, 1);
    if (ctl.comp_filt < 1000000 && score > ctl.max_score && min_shares < ctl.auto_share &&
        (ctl.comp_filt_bound == 0 || min_indeps <= ctl.comp_filt_bound) && chains1 == 2 &&
        chains2 == 2 && in.ch_left[c1_0] != in.ch_left[c1_1]) {
        const int64_t h1 = I.chain(ex1, in.row_exact_col[2 * k1 + 0]);
        const int64_t h2 = I.chain(ex2, in.row_exact_col[2 * k2 + 0]);
        const int64_t comp = std::min(in.ch_hcomp[h1], in.ch_hcomp[h2]);
        if (comp - cd >= ctl.comp_filt) accept = true;
        // else: SUPER_COMP_FILT branch (:568-705) needs super_comp_filt > 0 — rejected mode
    }

    // :710-715 threshold on score
    if (!accept && score > ctl.max_score && min_shares < ctl.auto_share) return J_FALSE;

    // :722-758 V gene names
    for (int i = 0; i < nv1; i++) {
        const int64_t x1 = I.chain(ex1, in.row_exact_col[2 * k1 + i]);
        const int64_t x2 = I.chain(ex2, in.row_exact_col[2 * k2 + i]);
        const int v1 = in.ch_v_ref[x1], v2 = in.ch_v_ref[x2];
        if (in.ref_name_id[v1] != in.ref_name_id[v2]) {
            const uint8_t *y1 = I.ref(v1), *y2 = I.ref(v2);
            const int64_t l1 = I.ref_len(v1), l2 = I.ref_len(v2);
            const int64_t nn = std::min(l1, l2);  // equal lengths: whole compare == prefix compare
            for (int64_t m = 0; m < nn; m++)
                if (base2bit(y1[m]) != base2bit(y2[m])) return J_FALSE;
            const int u1 = in.ch_u_ref[x1], u2 = in.ch_u_ref[x2];
            if (u1 != EJB_NONE && u2 != EJB_NONE) {
                const uint8_t *w1 = I.ref(u1), *w2 = I.ref(u2);
                const int64_t m1 = I.ref_len(u1), m2 = I.ref_len(u2);
                const int64_t mm = std::min(m1, m2);
                for (int64_t m = 0; m < mm; m++)
                    if (base2bit(w1[m1 - 1 - m]) != base2bit(w2[m2 - 1 - m])) return J_FALSE;
            }
        }
    }

    // :766-854 FWR1 identity minus CDR1+2 identity, heavy chain
    {
        const int nchains = I.nchains(k1);
        int64_t fwr1_len = 0, cdr1_len = 0, cdr2_len = 0;
        int64_t fwr1_diffs = 0, cdr1_diffs = 0, cdr2_diffs = 0;
        for (int m = 0; m < nchains; m++) {
            const int64_t x1 = I.chain(ex1, in.row_exact_col[2 * k1 + m]);
            const int64_t x2 = I.chain(ex2, in.row_exact_col[2 * k2 + m]);
            if (!in.ch_left[x1]) continue;
            // seq_del_amino of each chain (falls back to the row's tig)
            const uint8_t *s1, *s2;
            int64_t sl1, sl2;
            if (in.ch_amino_off[x1] != EJB_NONE) {
                s1 = in.amino_bytes + in.ch_amino_off[x1];
                sl1 = in.ch_amino_len[x1];
            } else {
                s1 = I.tig(k1, m);
                sl1 = I.len(k1, m);
            }
            if (in.ch_amino_off[x2] != EJB_NONE) {
                s2 = in.amino_bytes + in.ch_amino_off[x2];
                sl2 = in.ch_amino_len[x2];
            } else {
                s2 = I.tig(k2, m);
                sl2 = I.len(k2, m);
            }
            auto region = [&](int64_t a1, int64_t b1, int64_t a2, int64_t b2, int64_t& rlen,
                              int64_t& rdiffs) -> bool {
                // len = stop1 - start1 (usize); equal lengths required, else the region is skipped
                const int64_t len = b1 - a1;
                if (b2 - a2 == len) {
                    int64_t diffs = 0;
                    for (int64_t p = 0; p < len; p++) {
                        if (p + a1 >= sl1 || p + a2 >= sl2) return false;  // index panic
                        if (s1[p + a1] != s2[p + a2]) diffs++;
                    }
                    rlen = len;
                    rdiffs = diffs;
                }
                return true;
            };
            if (in.ch_cdr1_start[x1] != EJB_NONE && in.ch_cdr1_start[x2] != EJB_NONE) {
                const int64_t a1 = in.ch_fr1_start[x1], b1 = in.ch_cdr1_start[x1];
                const int64_t a2 = in.ch_fr1_start[x2], b2 = in.ch_cdr1_start[x2];
                if (b1 < a1 || b2 < a2) {
                    why = "fr1_stop - fr1_start underflows (join_one.rs:778-779)";
                    return J_PANIC;
                }
                if (!region(a1, b1, a2, b2, fwr1_len, fwr1_diffs)) {
                    why = "seq_del_amino index out of range";
                    return J_PANIC;
                }
            }
            if (in.ch_cdr1_start[x1] != EJB_NONE && in.ch_fr2_start[x1] != EJB_NONE &&
                in.ch_cdr1_start[x2] != EJB_NONE && in.ch_fr2_start[x2] != EJB_NONE) {
                const int64_t a1 = in.ch_cdr1_start[x1], b1 = in.ch_fr2_start[x1];
                const int64_t a2 = in.ch_cdr1_start[x2], b2 = in.ch_fr2_start[x2];
                if (a1 <= b1 && a2 <= b2) {
                    if (!region(a1, b1, a2, b2, cdr1_len, cdr1_diffs)) {
                        why = "seq_del_amino index out of range";
                        return J_PANIC;
                    }
                }
            }
            if (in.ch_cdr2_start[x1] != EJB_NONE && in.ch_fr3_start[x1] != EJB_NONE &&
                in.ch_cdr2_start[x2] != EJB_NONE && in.ch_fr3_start[x2] != EJB_NONE) {
                const int64_t a1 = in.ch_cdr2_start[x1], b1 = in.ch_fr3_start[x1];
                const int64_t a2 = in.ch_cdr2_start[x2], b2 = in.ch_fr3_start[x2];
                if (a1 <= b1 && a2 <= b2) {
                    if (!region(a1, b1, a2, b2, cdr2_len, cdr2_diffs)) {
                        why = "seq_del_amino index out of range";
                        return J_PANIC;
                    }
                }
            }
        }
        if (fwr1_len > 0 && cdr1_len > 0 && cdr2_len > 0) {
            int64_t len = fwr1_len, diffs = fwr1_diffs;
            const double fwr1_identity = 100.0 * (double)(len - diffs) / (double)len;
            len = cdr1_len + cdr2_len;
            diffs = cdr1_diffs + cdr2_diffs;
            const double cdr12_identity = 100.0 * (double)(len - diffs) / (double)len;
            if (fwr1_identity - cdr12_identity >= ctl.fwr1_cdr12_delta) return J_FALSE;
        }
    }

    // :860-886
    Pot pj;
    pj.k1 = k1;
    pj.k2 = k2;
    pj.nrefs = nrefs;
    pj.cd = cd;
    pj.shares[0] = shares[0];
    pj.shares[1] = (nrefs == 2) ? shares[1] : 0;
    pj.indeps[0] = indeps[0];
    pj.indeps[1] = (nrefs == 2) ? indeps[1] : 0;
    pj.score = score;
    pj.err = err;
    pj.p1 = p1;
    pj.mult = mult;
    pj.k = k;
    pj.d = d;
    pj.n = n;
    pot.push_back(pj);
    return J_TRUE;
}

// ---------------------------------------------------------------------------------------------
// Per-bucket work: join_core (join_core.rs:11-51) + the closure of join.rs:93-214.
// ---------------------------------------------------------------------------------------------
struct BucketResult {
    int64_t i = 0, j = 0;
    int64_t joins = 0, errors = 0;
    int64_t join_one_calls = 0, pot_found = 0, two_cell_deleted = 0;
    std::vector<Pot> join_list;  // the PotentialJoins of r.join_list, in order
    bool panicked = false;
    std::string why;
};

static void run_bucket(BucketResult& r, const In& I, const SrTable& sr) {
    const int64_t i = r.i, j = r.j;
    std::vector<Pot> pot;

    // join_core
    {
        EquivRel eq((int32_t)(j - i));
        for (int64_t k1 = i; k1 < j; k1++) {
            for (int64_t k2 = k1 + 1; k2 < j; k2++) {
                if (eq.class_id((int32_t)(k1 - i)) == eq.class_id((int32_t)(k2 - i))) continue;
                r.join_one_calls++;
                JoinOneStatus st = join_one(k1, k2, I, sr, pot, r.why);
                if (st == J_PANIC) {
                    r.panicked = true;
                    return;
                }
                if (st == J_TRUE) eq.join((int32_t)(k1 - i), (int32_t)(k2 - i));
            }
        }
    }
    r.pot_found = (int64_t)pot.size();

    // join.rs:120-178 — two passes of the "only two cells" filter
    const ejb_input& in = *I.in;
    for (int pass = 0; pass < 2; pass++) {
        EquivRel eq((int32_t)(j - i));
        for (const Pot& p : pot) eq.join((int32_t)(p.k1 - i), (int32_t)(p.k2 - i));
        std::vector<std::vector<size_t>> to_pot(j - i);
        for (size_t pj = 0; pj < pot.size(); pj++) to_pot[pot[pj].k1 - i].push_back(pj);
        std::vector<char> to_delete(pot.size(), 0);
        std::vector<int32_t> reps, x;
        eq.orbit_reps(reps);
        for (int32_t rep : reps) {
            eq.orbit(rep, x);
            int64_t ncells = 0;
            for (int32_t t : x) {
                const int64_t k = (int64_t)t + i;
                ncells += I.ex_ncells(in.row_exact[k]);
            }
            if (ncells == 2 && x.size() == 2) {
                const int64_t ka = (int64_t)x[0] + i, kb = (int64_t)x[1] + i;
                const int64_t k = std::min(ka, kb);
                for (size_t pj : to_pot[k - i]) {
                    const int64_t cd = pot[pj].cd;
                    int64_t min_shares = pot[pj].shares[0];
                    if (pot[pj].nrefs == 2) min_shares = std::min(min_shares, pot[pj].shares[1]);
                    if (cd > min_shares / 2 && !I.p->easy) to_delete[pj] = 1;  // isize division
                }
            }
        }
        // erase_if
        size_t count = 0;
        for (size_t q = 0; q < pot.size(); q++)
            if (!to_delete[q]) {
                if (q != count) std::swap(pot[q], pot[count]);
                count++;
            }
        r.two_cell_deleted += (int64_t)(pot.size() - count);
        pot.resize(count);
    }

    // join.rs:182-214 — analyse potential joins
    {
        EquivRel eq((int32_t)(j - i));
        for (const Pot& pj : pot) {
            if (eq.class_id((int32_t)(pj.k1 - i)) == eq.class_id((int32_t)(pj.k2 - i))) continue;
            r.join_list.push_back(pj);
            r.joins++;
            if (pj.err) r.errors++;
            eq.join((int32_t)(pj.k1 - i), (int32_t)(pj.k2 - i));
        }
    }
}

}  // namespace oracle

// ---------------------------------------------------------------------------------------------
// C interface (ctypes / tests / bench CPU baseline)
// ---------------------------------------------------------------------------------------------
extern "C" {

typedef struct oracle_result {
    int64_t n_edges;
    int32_t *edge_k1, *edge_k2, *edge_cd, *edge_nrefs, *edge_shares, *edge_indeps;
    int32_t *edge_k, *edge_d, *edge_n;
    double *edge_p1, *edge_mult, *edge_score;
    uint8_t* edge_err;
    int64_t n_rows;
    int32_t *eq_x, *eq_y, *eq_z;
    int32_t* labels;  // min row index of each orbit
    int64_t n_buckets, pairs_lens, pairs_scored, join_one_calls, pot_found, two_cell_deleted;
    int64_t joins, errors;
    double seconds_join;  // bucket loop + finish_join (excludes the Stirling table build)
    int32_t threads;
    char error[256];
} oracle_result;

int oracle_join(const ejb_params* p, const ejb_input* in, int nthreads, oracle_result* out);
void oracle_result_free(oracle_result* r);
void oracle_sr_entry(int n, int k, double* hi, double* lo);
int oracle_p1(int64_t k, int64_t d, int64_t n, double* out);
void oracle_dd_ops(double ahi, double alo, double bhi, double blo, double* out8);
void oracle_equiv_replay(int32_t n, const int32_t* a, const int32_t* b, int64_t m, int32_t* x,
                         int32_t* y, int32_t* z);
}

template <class T>
static T* dup_vec(const std::vector<T>& v) {
    T* p = (T*)malloc(std::max<size_t>(1, v.size()) * sizeof(T));
    if (!v.empty()) memcpy(p, v.data(), v.size() * sizeof(T));
    return p;
}

static void make_in(oracle::In& I, const ejb_params* p, const ejb_input* in) {
    using namespace oracle;
    I.in = in;
    I.p = p;
    // canonical seg ids by content
    std::vector<std::pair<std::string, int>> v;
    for (int64_t s = 0; s < in->n_segs; s++) {
        std::string str((const char*)in->seg_bytes + in->seg_off[s], (size_t)(in->seg_off[s + 1] - in->seg_off[s]));
        for (auto& c : str) c = (char)kBits2Base[base2bit((uint8_t)c)];
        v.emplace_back(std::move(str), (int)s);
    }
    std::sort(v.begin(), v.end());
    I.seg_canon.assign(in->n_segs, 0);
    int id = -1;
    for (size_t q = 0; q < v.size(); q++) {
        if (q == 0 || v[q].first != v[q - 1].first) id++;
        I.seg_canon[v[q].second] = id;
    }
}

// join_one (join_one.rs:66-887) for arbitrary pairs, as define_mat calls it (enclone_process/src/define_mat.rs:172,247): the
// checker of ejb_join_one_batch.  Outputs are zero where join_one returns false.
extern "C" int oracle_join_one_batch(const ejb_params* p, const ejb_input* in, const int32_t* k1, const int32_t* k2, int64_t n, uint8_t* pass,
                                     int32_t* cd, int32_t* nrefs, int32_t* shares, int32_t* indeps, int32_t* k, int32_t* d, int32_t* nn, double* p1,
                                     double* mult, double* score, uint8_t* err) {
    using namespace oracle;
    const SrTable& sr = sr_table();
    In I;
    make_in(I, p, in);
    for (int64_t i = 0; i < n; i++) {
        std::vector<Pot> pot;
        std::string why;
        pass[i] = 0;
        cd[i] = nrefs[i] = shares[2 * i] = shares[2 * i + 1] = indeps[2 * i] = indeps[2 * i + 1] = k[i] = d[i] = nn[i] = 0;
        p1[i] = mult[i] = score[i] = 0.0;
        err[i] = 0;
        // the caller's filter (define_mat.rs:169, 244): equal lens
        bool same = in->row_nchains[k1[i]] == in->row_nchains[k2[i]];
        for (int m = 0; same && m < in->row_nchains[k1[i]] && m < 2; m++) same = in->row_len[2 * (int64_t)k1[i] + m] == in->row_len[2 * (int64_t)k2[i] + m];
        if (!same) continue;
        const JoinOneStatus st = join_one(k1[i], k2[i], I, sr, pot, why);
        if (st == J_PANIC) return EJB_ERR_INVALID;
        if (st != J_TRUE || pot.empty()) continue;
        const Pot& q = pot.back();
        pass[i] = 1;
        cd[i] = (int32_t)q.cd;
        nrefs[i] = q.nrefs;
        shares[2 * i] = (int32_t)q.shares[0]; shares[2 * i + 1] = (int32_t)q.shares[1];
        indeps[2 * i] = (int32_t)q.indeps[0]; indeps[2 * i + 1] = (int32_t)q.indeps[1];
        k[i] = (int32_t)q.k; d[i] = (int32_t)q.d; nn[i] = (int32_t)q.n;
        p1[i] = q.p1; mult[i] = q.mult; score[i] = q.score;
        err[i] = q.err ? 1 : 0;
    }
    return EJB_OK;
}

int oracle_join(const ejb_params* p, const ejb_input* in, int nthreads, oracle_result* out) {
    using namespace oracle;
    memset(out, 0, sizeof(*out));
    if (p->unsupported_modes != 0) {
        snprintf(out->error, sizeof(out->error), "unsupported join mode 0x%x", p->unsupported_modes);
        return EJB_ERR_UNSUPPORTED;
    }
    const SrTable& sr = sr_table();
    In I;
    make_in(I, p, in);
    const int64_t N = in->n_rows;

    // join.rs:68-88 — buckets = runs of equal lens
    std::vector<BucketResult> results;
    {
        int64_t i = 0;
        while (i < N) {
            int64_t j = i + 1;
            while (j < N) {
                bool same = in->row_nchains[j] == in->row_nchains[i];
                for (int m = 0; same && m < in->row_nchains[i]; m++)
                    if (in->row_len[2 * j + m] != in->row_len[2 * i + m]) same = false;
                if (!same) break;
                j++;
            }
            BucketResult r;
            r.i = i;
            r.j = j;
            results.push_back(std::move(r));
            i = j;
        }
    }
    out->n_buckets = (int64_t)results.size();
    for (auto& r : results) {
        const int64_t nb = r.j - r.i;
        if (in->row_nchains[r.i] == 2) {
            out->pairs_lens += nb * (nb - 1) / 2;
            // metric unit: pairs with equal per-chain CDR3 lengths (SURVEY §8 d)
            std::vector<std::pair<int, int>> cl;
            for (int64_t k = r.i; k < r.j; k++) cl.emplace_back(in->row_cdr3_len[2 * k], in->row_cdr3_len[2 * k + 1]);
            std::sort(cl.begin(), cl.end());
            size_t a = 0;
            while (a < cl.size()) {
                size_t b = a;
                while (b < cl.size() && cl[b] == cl[a]) b++;
                out->pairs_scored += (int64_t)(b - a) * (int64_t)(b - a - 1) / 2;
                a = b;
            }
        }
    }

    if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
    if (nthreads <= 0) nthreads = 1;
    out->threads = nthreads;

    auto t0 = std::chrono::steady_clock::now();
    // join.rs:582 — one task per bucket; largest first so the dynamic schedule balances like rayon
    std::vector<size_t> order(results.size());
    for (size_t q = 0; q < order.size(); q++) order[q] = q;
    std::sort(order.begin(), order.end(), [&](size_t a, size_t b) {
        return (results[a].j - results[a].i) > (results[b].j - results[b].i);
    });
    std::atomic<size_t> next(0);
    auto worker = [&]() {
        for (;;) {
            size_t q = next.fetch_add(1);
            if (q >= order.size()) break;
            run_bucket(results[order[q]], I, sr);
        }
    };
    if (nthreads == 1) {
        worker();
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; t++) th.emplace_back(worker);
        for (auto& t : th) t.join();
    }
    for (auto& r : results)
        if (r.panicked) {
            snprintf(out->error, sizeof(out->error), "reference would panic: %s", r.why.c_str());
            return (r.why.find("Stirling") != std::string::npos) ? EJB_ERR_RANGE : EJB_ERR_INVALID;
        }

    // join.rs:584-588 raw_joins; join2.rs:45-54
    std::vector<int32_t> k1v, k2v, cdv, nrv, shv, inv, kv, dv, nv;
    std::vector<double> p1v, mv, sv;
    std::vector<uint8_t> ev;
    EquivRel eq((int32_t)N);
    for (auto& r : results) {
        out->joins += r.joins;
        out->errors += r.errors;
        out->join_one_calls += r.join_one_calls;
        out->pot_found += r.pot_found;
        out->two_cell_deleted += r.two_cell_deleted;
        for (const Pot& pj : r.join_list) {
            k1v.push_back((int32_t)pj.k1);
            k2v.push_back((int32_t)pj.k2);
            cdv.push_back((int32_t)pj.cd);
            nrv.push_back(pj.nrefs);
            shv.push_back((int32_t)pj.shares[0]);
            shv.push_back((int32_t)pj.shares[1]);
            inv.push_back((int32_t)pj.indeps[0]);
            inv.push_back((int32_t)pj.indeps[1]);
            kv.push_back((int32_t)pj.k);
            dv.push_back((int32_t)pj.d);
            nv.push_back((int32_t)pj.n);
            p1v.push_back(pj.p1);
            mv.push_back(pj.mult);
            sv.push_back(pj.score);
            ev.push_back(pj.err ? 1 : 0);
            eq.join((int32_t)pj.k1, (int32_t)pj.k2);
        }
    }
    // join2.rs:69-84 — join orbits that cross subclones of a clone
    {
        std::vector<std::pair<int64_t, int32_t>> ox;
        ox.reserve(N);
        for (int64_t i = 0; i < N; i++) ox.emplace_back(in->row_exact[i], eq.class_id((int32_t)i));
        std::sort(ox.begin(), ox.end());
        size_t i = 0;
        while (i < ox.size()) {
            size_t j = i + 1;
            while (j < ox.size() && ox[j].first == ox[i].first) j++;  // next_diff1_2
            for (size_t k = i; k + 1 < j; k++) eq.join(ox[k].second, ox[k + 1].second);
            i = j;
        }
    }
    out->seconds_join = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

    out->n_edges = (int64_t)k1v.size();
    out->edge_k1 = dup_vec(k1v);
    out->edge_k2 = dup_vec(k2v);
    out->edge_cd = dup_vec(cdv);
    out->edge_nrefs = dup_vec(nrv);
    out->edge_shares = dup_vec(shv);
    out->edge_indeps = dup_vec(inv);
    out->edge_k = dup_vec(kv);
    out->edge_d = dup_vec(dv);
    out->edge_n = dup_vec(nv);
    out->edge_p1 = dup_vec(p1v);
    out->edge_mult = dup_vec(mv);
    out->edge_score = dup_vec(sv);
    out->edge_err = dup_vec(ev);
    out->n_rows = N;
    out->eq_x = dup_vec(eq.x);
    out->eq_y = dup_vec(eq.y);
    out->eq_z = dup_vec(eq.z);
    {
        std::vector<int32_t> lab(N), mn(N, INT32_MAX);
        for (int64_t i = 0; i < N; i++) mn[eq.y[i]] = std::min(mn[eq.y[i]], (int32_t)i);
        for (int64_t i = 0; i < N; i++) lab[i] = mn[eq.y[i]];
        out->labels = dup_vec(lab);
    }
    return EJB_OK;
}

void oracle_result_free(oracle_result* r) {
    free(r->edge_k1); free(r->edge_k2); free(r->edge_cd); free(r->edge_nrefs);
    free(r->edge_shares); free(r->edge_indeps); free(r->edge_k); free(r->edge_d); free(r->edge_n);
    free(r->edge_p1); free(r->edge_mult); free(r->edge_score); free(r->edge_err);
    free(r->eq_x); free(r->eq_y); free(r->eq_z); free(r->labels);
    memset(r, 0, sizeof(*r));
}

void oracle_sr_entry(int n, int k, double* hi, double* lo) {
    const oracle::SrTable& sr = oracle::sr_table();
    *hi = sr.at(n, k).hi;
    *lo = sr.at(n, k).lo;
}

int oracle_p1(int64_t k, int64_t d, int64_t n, double* out) {
    return oracle::p_at_most_m_distinct(k - d, k, n, oracle::sr_table(), *out) ? 0 : EJB_ERR_RANGE;
}

// a+b, a-b, a*b, a/b as (hi, lo) pairs — for the mpmath cross-check of the dd restatement
void oracle_dd_ops(double ahi, double alo, double bhi, double blo, double* out8) {
    using namespace oracle;
    DD a(ahi, alo), b(bhi, blo);
    DD r = dd_add(a, b);
    out8[0] = r.hi; out8[1] = r.lo;
    r = dd_sub(a, b);
    out8[2] = r.hi; out8[3] = r.lo;
    r = dd_mul(a, b);
    out8[4] = r.hi; out8[5] = r.lo;
    r = dd_div(a, b);
    out8[6] = r.hi; out8[7] = r.lo;
}

void oracle_equiv_replay(int32_t n, const int32_t* a, const int32_t* b, int64_t m, int32_t* x,
                         int32_t* y, int32_t* z) {
    oracle::EquivRel eq(n);
    for (int64_t i = 0; i < m; i++) eq.join(a[i], b[i]);
    memcpy(x, eq.x.data(), (size_t)n * 4);
    memcpy(y, eq.y.data(), (size_t)n * 4);
    memcpy(z, eq.z.data(), (size_t)n * 4);
}
