// ejb_types.cuh — shared device/host types and helpers of the sm_100a clonotype join.
//
// Kernel map (reference function each one replaces):
//   ejb_pack.cuh    k_heads / k_keys / k_subheads / k_subinfo  bucket scan of join_exacts, enclone/src/join.rs:68-88
//                   k_pack_rows  the "bucketer": 2-bit planes of contigs + their V/J references, compare
//                                masks, reference-difference bitmaps (the per-ROW half of the per-base
//                                compares of enclone_core/src/join_one.rs:343-429)
//   ejb_stage1.cuh  k_stage1<W1> CDR3 Hamming pre-filter, join_one.rs:216-234, 271-290, + the same-class
//                                skip of enclone/src/join_core.rs:32
//   ejb_stage2.cuh  k_stage2a / k_stage2_slow  shares / indeps / diffs + integer tests,
//                                join_one.rs:238-267, 301-440, 474-493, region diffs of :766-854
//                   k_stage2b    barcode meet, score, JUN_SHARE, V names, FWR1/CDR12, join_one.rs:451-470,
//                                499-567, 710-854
//   ejb_p1.cuh      k_p1_*       p_at_most_m_distinct_in_sample_of_x_from_n_double, join_one.rs:22-43
//   ejb_forest.cuh  k_bor_*      lexicographic spanning forest == pot of join_core.rs:23-50
//                   k_two_cell / k_exact_union / k_labels   join.rs:120-178, join2.rs:69-84
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dd.cuh"

namespace ejb {

constexpr int TILE = 128;          // rows / cols of one stage-1 tile
constexpr int S1_THREADS = 256;
#ifndef S1_STRIP_N
#define S1_STRIP_N 16
#endif
constexpr int S1_STRIP = S1_STRIP_N;   // column tiles per CTA (row fragment stays in registers)
constexpr int MAXW1 = 6;           // concatenated CDR3 (heavy+light) up to 192 nt on the fast path
constexpr int MAXW4 = 8;           // contig up to 1024 nt on the stage-2 fast path
constexpr int SR_NMAX = 3000;      // stirling2_ratio_table_double(3000), start.rs:344

// ---- per-row record `aux`: AUX_INTS int32 per packed row position ----
enum AuxField {
    AX_FLAGS = 0, AX_VSH, AX_VSL, AX_JSH, AX_JSL, AX_TOT, AX_EXACT, AX_NCELLS,          // uint4 #0, #1
    AX_DONLO, AX_DONHI, AX_CREFH, AX_CREFL, AX_HCOMP, AX_VREFH, AX_VREFL, AX_BC0,       // uint4 #2, #3
    AX_VNAMEH, AX_VNAMEL, AX_FR1, AX_CDR1, AX_FR2, AX_CDR2, AX_FR3, AX_CHH,             // uint4 #4, #5
    AX_CHL, AX_RGN0, AX_RGN1, AX_RGN2,                                                  // uint4 #6: FWR1 / CDR1 / CDR2 ranges of chain 0,
                                                                                        // lo | hi << 16, empty (0) unless B_REG_OK
    AUX_INTS
};
static_assert(AUX_INTS == 28, "aux record is 7 uint4");

enum RowFlags : uint32_t {
    F_ODDBYTE_H = 1u, F_ODDBYTE_L = 2u,   // contig has a byte outside ACGT ('-' is fine when has_del is set)
    F_CDR3_IRREG = 4u,                    // a CDR3 byte outside ACGT
    F_SLOWGEOM = 8u,                      // V/J ranges overlap, or the V range exceeds the contig
    F_PANIC = 16u,                        // the reference would panic on this row (J range > contig ...)
    F_HASDEL_H = 32u, F_HASDEL_L = 64u,
    F_SLOWMASK = F_ODDBYTE_H | F_ODDBYTE_L | F_CDR3_IRREG | F_SLOWGEOM | F_PANIC,
    // facts about the row's exact subclonotype
    B_NCH_OK = 1u << 16,      // share.len() in 2..=3           join_one.rs:87
    B_NCH2 = 1u << 17,        // share.len() == 2               join_one.rs:554
    B_JUN_LEFT = 1u << 18,    // share[0].left != share[1].left join_one.rs:556
    B_LEFT_H = 1u << 19,      // left flag of the row's chain 0 / 1
    B_LEFT_L = 1u << 20,
    B_AMINO_TIG_H = 1u << 21, // seq_del_amino of chain 0 is the row's contig
    B_REG_OK = 1u << 22       // chain 0's FWR1/CDR1/CDR2 ranges are well formed and inside the contig: the region planes are valid
};

struct SubBucket {
    int64_t j0;       // first position in the dense (sorted) order
    int64_t q0;       // first packed row position (multiple of 4: sub-buckets are padded to 16 bytes)
    int64_t s1_off;   // word offset of this sub-bucket's CDR3 planes in cdr3w
    int64_t rs_off;   // uint4 offset of this sub-bucket's row records in rowstore
    int32_t n;        // rows
    int32_t npad;     // n rounded up to a multiple of 4 (16-byte TMA granules)
    int32_t ch, cl;   // CDR3 nt lengths
    int32_t Lh, Ll;   // contig lengths
    int32_t W1;       // words per plane of the concatenated CDR3; 0 => longer than the fast path
    int32_t W4h, W4l; // uint4 per plane of the heavy / light contig; 0 => longer than the fast path
    int32_t gap;      // bit m set: chain m of some row has_del => a gap plane is stored for that chain
    int32_t rec_u4;   // uint4 per row record
    int32_t maxcd;    // stage-1 threshold on the CDR3 distance (superset of join_one.rs:231,286-290)
    int32_t blk0;     // first global 128-row block id
    int32_t n3;       // 3 * (Lh + Ll)  (join_one.rs:499)
    int32_t owner;    // rank that owns this sub-bucket (multi-GPU sharding; -1: split over several ranks by row ranges)
    int32_t r_beg;    // rows [r_beg, n_act) act as the FIRST row of a pair on this context: [0, n) when it owns the whole
    int32_t n_act;    // sub-bucket, a sub-range when the sub-bucket is split by row ranges over several contexts (in-library
                      // sharding, or the caller's ejb_set_row_roles), empty when another shard owns it (then it is not packed)
    int32_t raw;      // part of a split sub-bucket: its forest edges are returned without the two-cell filter
    int32_t rot;      // the concatenated CDR3 is stored ROTATED left by `rot` positions (Hamming distance does not care): plane
                      // word 0 then holds the 32 positions centred in the heavy junction — the most diverse ones — instead of the
                      // germline-encoded V end, which makes stage 1's first-word test reject unrelated pairs after one popc
    int32_t pad2;
};
constexpr uint8_t ROLE_COLUMN_ONLY = 1, ROLE_SPLIT = 2;
// Row record layout (384 B for contigs up to 512 nt).  Per chain m the record is PLANE-MAJOR with a fixed plane order; W4
// (uint4 per plane) is 4 for contigs up to 512 nt and 8 up to 1024 nt, so that the common case has a compile-time stride:
//     chain m, plane pl, window g (128 positions)  ->  uint4 index  off_m + pl * W4_m + g
//   PL_TLO, PL_THI : 2-bit code of the contig base ((byte >> 1) & 3: A,C,T,G = 0,1,2,3; '-' carries the gap bit)
//   PL_D           : cmp & (contig base != own reference base), cmp = positions join_one compares
//                    (p < |V|-ref_v_trim  or  p >= L-(|J|-ref_j_trim), join_one.rs:365-391)
//   [gap]          : '-' positions, last plane, only when SubBucket.gap has bit m
// What only the rare two-reference candidates need (the OTHER row's V / J reference bases in contig coordinates and its
// compare mask) is not stored per row: stage 2a slices it out of the per-segment planes (segpl, ref_word() below).  The
// FWR1 / CDR1 / CDR2 masks of chain 0 are rebuilt from three packed ranges in the aux record (AX_RGN*).
enum Plane { PL_TLO = 0, PL_THI = 1, PL_D = 2, PL_GAP = 3 };
__host__ __device__ __forceinline__ int rec_planes(int m, int gapbit) { (void)m; return 3 + (gapbit ? 1 : 0); }
__host__ __device__ __forceinline__ int gap_plane(int m) { (void)m; return PL_GAP; }
__device__ __forceinline__ uint32_t base_code(uint32_t byte) { return (byte >> 1) & 3u; }   // A,C,T,G -> 0,1,2,3
__device__ __forceinline__ bool is_acgt_fast(uint32_t b) { return (b & 0xe0u) == 0x40u && ((0x0010008au >> (b & 31u)) & 1u); }

// Per-segment reference planes (internal base code, 2 planes x SEGPL_WORDS words, position 0 = first base of the
// segment, zero beyond its end), indexed by the CANONICAL seg id, plus the segment lengths: the bucketer slices / shifts
// them into contig coordinates instead of touching segment bytes, stage 2a does the same for two-reference candidates.
constexpr int SEGPL_WORDS = 16;   // segments up to 512 nt (longer ones send their rows to the byte-exact path)
// bits [0, n) set (n may be <= 0 or >= 32)
__device__ __forceinline__ uint32_t mask_lt(long long n) { return n <= 0 ? 0u : (n >= 32 ? 0xffffffffu : ((1u << (int)n) - 1u)); }
// same for a 32-bit n: PTX shr.u32 clamps shift amounts above 31, so all-ones >> (32 - clamp(n, 0, 32)) needs no branches
__device__ __forceinline__ uint32_t mask_lt32(int n) {
    uint32_t r;
    const int sh = 32 - min(max(n, 0), 32);
    asm("shr.u32 %0, %1, %2;" : "=r"(r) : "r"(0xffffffffu), "r"(sh));
    return r;
}
// 32 bits of a SEGPL_WORDS-word plane starting at bit index `idx` (bits outside the plane read as 0)
__device__ __forceinline__ uint32_t plane_bits(const uint32_t* __restrict__ pl, long long idx) {
    const long long wi = idx >> 5;   // floor
    const int sh = (int)(idx & 31);
    const uint32_t a = (wi >= 0 && wi < SEGPL_WORDS) ? pl[wi] : 0u;
    const uint32_t b = (wi + 1 >= 0 && wi + 1 < SEGPL_WORDS) ? pl[wi + 1] : 0u;
    return sh ? ((a >> sh) | (b << (32 - sh))) : a;
}
// One row-chain's V / J reference in CONTIG coordinates (V from the start, J aligned to the end, join_one.rs:365-391).
// Only valid for rows without F_PANIC / F_SLOWGEOM: 0 <= vr, 0 <= jr, vr + jr <= L, both segments <= 32 * SEGPL_WORDS.
struct RefGeom {
    uint32_t vo, jo;   // word offsets of the V / J lo planes in segpl; the hi planes follow at + SEGPL_WORDS
    int vr, jr;        // compared lengths: |V| - ref_v_trim, |J| - ref_j_trim
    int jshift;        // J index of contig position p = p + jshift  (|J| - L)
    int L;
};
__device__ __forceinline__ RefGeom ref_geom(const int32_t* __restrict__ seglen, int sv, int sj, int L, long long ref_v_trim, long long ref_j_trim) {
    RefGeom G;
    G.vo = (uint32_t)sv * 2u * SEGPL_WORDS;
    G.jo = (uint32_t)sj * 2u * SEGPL_WORDS;
    const long long vlen = seglen[sv], jlen = seglen[sj];
    G.vr = (int)max(min(vlen - ref_v_trim, 1ll << 30), -(1ll << 30));
    G.jr = (int)max(min(jlen - ref_j_trim, 1ll << 30), -(1ll << 30));
    G.jshift = (int)(jlen - L);
    G.L = L;
    return G;
}
// word wi (positions 32 wi .. 32 wi + 31): reference planes and compare mask
__device__ __forceinline__ void ref_word(const uint32_t* __restrict__ segpl, const RefGeom& G, int wi, uint32_t& rlo, uint32_t& rhi, uint32_t& cm) {
    const uint32_t inl = mask_lt32(G.L - 32 * wi);
    const uint32_t cmv = mask_lt32(G.vr - 32 * wi) & inl;
    const uint32_t cmj = inl & ~mask_lt32((G.L - G.jr) - 32 * wi) & ~cmv;   // p >= L - jr; the V range has priority
    const long long jidx = (long long)32 * wi + G.jshift;
    const uint32_t* vp = segpl + G.vo;
    const uint32_t* jp = segpl + G.jo;
    const uint32_t vlo = wi < SEGPL_WORDS ? vp[wi] : 0u, vhi = wi < SEGPL_WORDS ? vp[SEGPL_WORDS + wi] : 0u;
    rlo = (vlo & cmv) | (plane_bits(jp, jidx) & cmj);
    rhi = (vhi & cmv) | (plane_bits(jp + SEGPL_WORDS, jidx) & cmj);
    cm = cmv | cmj;
}

// Device view of the raw (uploaded) input, same meaning as ejb_input.
struct DevIn {
    int64_t n_rows, n_exacts, n_chains, n_cells, n_segs, n_refs;
    int64_t n_tig_bytes, n_cdr3_bytes, n_amino_bytes, n_seg_bytes, n_ref_bytes;
    const int32_t *row_nchains, *row_len, *row_vs, *row_js, *row_cdr3_len, *row_exact, *row_exact_col;
    const int64_t *row_tig_off, *row_cdr3_off;
    const uint8_t *row_has_del, *tig_bytes, *cdr3_bytes;
    const int64_t *ex_chain_off, *ex_cell_off;
    const uint8_t* ch_left;
    const int32_t *ch_c_ref, *ch_v_ref, *ch_u_ref, *ch_fr1, *ch_cdr1, *ch_fr2, *ch_cdr2, *ch_fr3, *ch_hcomp, *ch_amino_len;
    const int64_t* ch_amino_off;
    const uint8_t* amino_bytes;
    const int32_t* cell_donor;
    const uint32_t* cell_barcode;
    const int64_t* seg_off;
    const uint8_t* seg_bytes;
    const int32_t* seg_canon;  // content-canonical seg id (host-built)
    const int64_t* ref_off;
    const uint8_t* ref_bytes;
    const int32_t* ref_name_id;
    const int64_t* row_tig2_off;   // optional 2-bit packed contigs (NULL: all ASCII)
    const uint32_t* tig2_words;
    int64_t n_tig2_words;
    const int64_t* row_cdr32_off;  // optional 2-bit packed CDR3s (NULL: all ASCII)
    const uint32_t* cdr32_words;
    int64_t n_cdr32_words;
};

// Flat join knobs on the device.
struct DevParams {
    int is_bcr, mix_donors, easy, old_light;
    double max_score, cdr3_mult, cdr3_ident_thr /* 1.0 - join_cdr3_ident/100.0 */, fwr1_cdr12_delta;
    long long max_cdr3_diffs, auto_share, comp_filt, comp_filt_bound, max_diffs, max_degradation, ref_v_trim, ref_j_trim;
};

struct Survivor {       // a candidate that passed the integer tests of stage 2
    uint32_t q1, q2;
    uint16_t cd1, cd2;
    int32_t sh0, in0, sh1, in1;
    uint8_t nrefs, err, ok, regfast;   // regfast: fd/c1d/c2d below are valid
    uint16_t fd, c1d, c2d, pad;        // FWR1 / CDR1 / CDR2 nucleotide differences (join_one.rs:766-854)
    uint32_t slot;                     // output slot (pair-batch mode, ejb_join_one_batch)
};

struct EdgeTuple {      // numeric members of PotentialJoin (defs.rs:870-891); pair-batch mode output
    int32_t cd, nrefs, sh0, sh1, in0, in1, k, d, n, err, pass, pad;
    double p1, mult, score;
};
struct PassTuple {      // what a pass edge keeps of its PotentialJoin: the rest (k, d, n, mult, score) is rebuilt from it.  32 bytes.
    uint32_t cd12;      // cd heavy | cd light << 16
    int32_t sh0, in0, sh1, in1;
    uint32_t nrefs_err; // nrefs | err << 8 | PT_SPLIT
    double p1;          // the value stage 2b accepted the edge with (travels with the edge: a merging rank never re-evaluates it)
};
static_assert(sizeof(PassTuple) == 32, "PassTuple is part of the partial-forest exchange format");
constexpr uint32_t PT_SPLIT = 1u << 16;   // the edge comes from a PART of a sub-bucket that was split over several ranks by row ranges

struct Counters {       // device-side counters, one instance per context
    // ---- zeroed at the start of every scheduler round (one memset over [n_cand, round_end)) ----
    unsigned long long n_cand, n_slow, n_surv, n_pass, n_newkeys, overflow, n_surv_total, s2_work;
    unsigned long long any_cut;       // a unit of this round was truncated after stage 1 (k_truncate)
    unsigned long long p1_touched, p1_desc;   // p1 term groups touched / terms listed for evaluation in this round (ejb_p1.cuh)
    unsigned long long alive_it[64];  // Boruvka: edges still joining two components, per iteration of this round
    unsigned long long round_end;     // marker only
    // ---- per run ----
    unsigned long long pairs_eval, n_forest;
    unsigned long long n_two_cell, n_errflag_range, n_errflag_panic, n_final;
    unsigned long long n_multi;       // rows whose exact subclonotype owns two or more rows (finish_join's clonotype_id merge, join2.rs:69-84)
    unsigned long long p1_total;      // keys ever inserted into the p1 cache since it was last cleared
    unsigned long long p1_full;       // a probe sequence of the p1 cache ran through the whole table (must not happen: the host grows it)
    unsigned long long p1_groups, p1_pool_used;   // (n, k) term groups and term-pool entries handed out since the p1 structures were reset
    unsigned long long p1_terms, p1_direct;       // terms evaluated; keys summed by the direct path (a term structure was full)
    unsigned long long s1_first;      // stage 1: lane-pairs whose first CDR3 word went through XOR + popc
    unsigned long long s1_full;       // stage 1: lane-pairs whose remaining CDR3 words were evaluated too
    unsigned long long bor_iters;     // Boruvka iterations over all rounds
    unsigned long long n_s1_kill;     // stage-1 hits dropped: known dead against the other's reference set, or different donor sets
    unsigned long long n_s1_donor;    // ... of which for the donor sets
    unsigned long long n_s1_flag;     // stage-1 pairs skipped on the 4-bit row flags alone, before any distance (not necessarily CDR3-close)
    unsigned long long n_two_ref;     // stage-2a candidates with two reference sets
    unsigned long long n_memo_dead;   // ... of which rejected on memoised totals (phase 0)
    unsigned long long n_degr_dead;   // ... of which rejected by the degradation test on freshly computed totals
    unsigned long long n_rej[8];      // rejections of the integer tests by first failing test (trace): 0 CDR3 identity,
                                      // 1 max_diffs, 2 max_cdr3_diffs, 3 donors, 4 cd >= cdr3_mult * indeps, 5 constant regions, 6 k range
};

// Reference-set tags (stage-1 kill of pairs that are known to fail join_one.rs:434-440).
//   tagq[q]  = canonical (vs_h, vs_l, js_h, js_l) ids of row q, 16 bits each; TAG_SLOW for rows the plane path cannot score
//   deadq[q] = tag of a FOREIGN reference set R with |tot_q - T(q, R)| > max_degradation (written by stage 2a), else DEAD_NONE
// Requires every canonical seg id < 65535 (otherwise the tags are switched off).
constexpr unsigned long long DEAD_NONE = 0xffffffffffffffffull;
constexpr unsigned long long TAG_SLOW = 0xfffffffffffffffeull;
__host__ __device__ __forceinline__ unsigned long long make_ref_tag(int vsh, int vsl, int jsh, int jsl) {
    return (unsigned long long)(uint32_t)vsh | ((unsigned long long)(uint32_t)vsl << 16) | ((unsigned long long)(uint32_t)jsh << 32) |
           ((unsigned long long)(uint32_t)jsl << 48);
}

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int base2bit(uint8_t b) {  // debruijn base_to_bits: non-ACGT -> 0
    switch (b) {
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return 0;
    }
}
__device__ __forceinline__ bool is_acgt(uint8_t b) { return b == 'A' || b == 'C' || b == 'G' || b == 'T'; }
__device__ __forceinline__ uint8_t bit2base(int x) { return (uint8_t)("ACGT"[x]); }

// A contig that is either ASCII bytes or 2-bit packed words (16 bases per word, A,C,G,T = 0..3); the byte-exact
// paths read it through this view.
struct TigRef {
    const uint8_t* b;
    const uint32_t* w;
    __device__ __forceinline__ uint8_t operator[](long long p) const {
        if (b) return b[p];
        const uint32_t x = w[p >> 4];
        return (uint8_t)("ACGT"[(x >> (2 * (int)(p & 15))) & 3u]);
    }
};
__device__ __forceinline__ TigRef tig_ref(const DevIn& in, int64_t k, int m) {
    TigRef r;
    r.b = nullptr;
    r.w = nullptr;
    const int64_t o2 = in.row_tig2_off ? in.row_tig2_off[2 * k + m] : -1;
    if (o2 >= 0) r.w = in.tig2_words + o2;
    else r.b = in.tig_bytes + in.row_tig_off[2 * k + m];
    return r;
}

__device__ __forceinline__ TigRef cdr3_ref(const DevIn& in, int64_t k, int m) {   // info[k].cdr3s[m], ASCII or 2-bit packed
    TigRef r;
    r.b = nullptr;
    r.w = nullptr;
    const int64_t o2 = in.row_cdr32_off ? in.row_cdr32_off[2 * k + m] : -1;
    if (o2 >= 0) r.w = in.cdr32_words + o2;
    else r.b = in.cdr3_bytes + in.row_cdr3_off[2 * k + m];
    return r;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t phase) {
    uint32_t ok;
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(phase)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    while (!mbar_try_wait(bar, phase)) {
    }
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

}  // namespace ejb
