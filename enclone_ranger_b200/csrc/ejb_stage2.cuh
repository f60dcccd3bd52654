// ejb_stage2.cuh — the exact join_one decision for stage-1 candidates.
//
//   k_stage2a      8 lanes per candidate over the packed planes: CDR3 distances, shared / independent
//                  mutation counts against one or two references, full-length differences, FWR1/CDR1/CDR2
//                  differences, then the integer tests of join_one.rs:216-323, 434-440, 474-493.
//   k_stage2_slow  byte-exact warp-per-candidate restatement for the rare rows the planes cannot
//                  represent (bytes outside ACGT-, overlapping V/J ranges, over-long contigs).
//   k_stage2b      one thread per survivor: barcode meet, p1 * mult score, JUN_SHARE, score threshold,
//                  V gene names, FWR1 - CDR12 identity (join_one.rs:451-470, 499-567, 710-854).
#pragma once
#include "ejb_p1.cuh"
#include "ejb_types.cuh"

namespace ejb {

struct S2Ctx {
    DevIn in;
    DevParams P;
    const SubBucket* subs;
    const int32_t* aux;          // AUX_INTS per packed row
    const int32_t* rowq;         // canonical row of a packed position
    const int32_t* subq;         // sub-bucket of a packed position
    const int32_t* qcut;         // per sub-bucket: candidates whose first row is at a packed position >= qcut are not part of this round
    unsigned int* sub_pass;      // per sub-bucket: pass edges found in this round
    int32_t* memo;               // per packed row, 8 ints: {state, vs_h, vs_l, js_h, js_l, T, -, -} — see memo_lookup()
    const unsigned long long* tagq;   // reference-set tag per packed row (nullptr: tags off), see ejb_types.cuh
    unsigned long long* deadq;        // per packed row: tag of a foreign reference set the row is dead against (read by stage 1)
    const uint32_t* cdr3w;
    const uint4* rowstore;
    const uint32_t* segpl;       // per canonical segment: reference planes (two-reference candidates slice them, ref_word())
    const int32_t* seglen;
    const uint64_t* bc_sorted;   // (exact << 32 | barcode), sorted: each exact's barcodes ascending
    const double* powtab;        // mult_pow ^ (cdr3_normal_len * cd / n), indexed pow_off[n] + cd
    const int32_t* pow_off;
    P1Cache p1;
    Counters* ctr;
    Survivor* surv;
    unsigned long long surv_cap;
    uint2* slow;                 // candidates routed to the byte-exact slow path
    uint32_t* slow_slot;
    unsigned long long slow_cap;
    int tuple_mode;              // 0: filter -> pass edges (+ PassTuple); 1: pair-batch mode (ejb_join_one_batch): an EdgeTuple per input pair
    int trace;                   // count the rejections of the integer tests by reason (developer aid, EJB_TRACE)
};

struct RegionDiffs {
    int regfast, fd, c1d, c2d;
};

constexpr uint32_t SURV_INVALID = 0xffffffffu;   // Survivor.q1 of an unused slot
constexpr int SURV_CHUNK = 64;                    // slots a warp reserves per atomicAdd

// Integer tests shared by the fast and the slow path, executed by one lane per candidate.
// Returns true and fills `s` (+ requests its p1) unless join_one would already return false.
__device__ __forceinline__ bool s2_integer_tests(const S2Ctx& X, uint32_t q1, uint32_t q2, uint32_t slot, const SubBucket& sb, int cd1, int cd2,
                                                 long long diffs, int nrefs, int sh0, int in0, int sh1, int in1, RegionDiffs rg, bool pre_ok,
                                                 bool force_emit, Survivor& s) {
    const DevParams& P = X.P;
    const int32_t* a1 = X.aux + (size_t)q1 * AUX_INTS;
    const int32_t* a2 = X.aux + (size_t)q2 * AUX_INTS;
    const uint32_t f1 = (uint32_t)a1[AX_FLAGS];
    const int cd = cd1 + cd2;
    bool ok = pre_ok;
    int why = -1;   // first failing test (trace only)
    // join_one.rs:216-234
    if (P.is_bcr) {
        const int total = sb.ch + sb.cl;
        if ((double)cd / (double)total > P.cdr3_ident_thr) { ok = false; if (why < 0) why = 0; }
    }
    // :238-267
    if (!P.is_bcr || P.max_diffs < 1000000) {
        if (diffs > P.max_diffs) { ok = false; if (why < 0) why = 1; }
        if (!P.is_bcr && diffs > 5) { ok = false; if (why < 0) why = 1; }
    }
    // :286-290
    if ((P.max_cdr3_diffs < 1000 || !P.is_bcr) && (cd > P.max_cdr3_diffs || (!P.is_bcr && cd > 0))) { ok = false; if (why < 0) why = 2; }
    // :301-323
    const unsigned long long m1 = ((unsigned long long)(uint32_t)a1[AX_DONHI] << 32) | (uint32_t)a1[AX_DONLO];
    const unsigned long long m2 = ((unsigned long long)(uint32_t)a2[AX_DONHI] << 32) | (uint32_t)a2[AX_DONLO];
    if (!P.mix_donors && m1 != 0 && m2 != 0 && m1 != m2) { ok = false; if (why < 0) why = 3; }
    const bool err = (m1 != m2) || __popcll(m1) != 1 || __popcll(m2) != 1;
    // :444-447
    const int min_sh = (nrefs == 2) ? min(sh0, sh1) : sh0;
    const int min_in = (nrefs == 2) ? min(in0, in1) : in0;
    // :474-476
    if ((double)cd >= P.cdr3_mult * (double)max(1, min_in)) { ok = false; if (why < 0) why = 4; }
    // :481-493
    if (!P.old_light && cd > 0) {
        if (!(f1 & B_LEFT_H) && a1[AX_CREFH] != -1 && a2[AX_CREFH] != -1 && a1[AX_CREFH] != a2[AX_CREFH]) { ok = false; if (why < 0) why = 5; }
        if (!(f1 & B_LEFT_L) && a1[AX_CREFL] != -1 && a2[AX_CREFL] != -1 && a1[AX_CREFL] != a2[AX_CREFL]) { ok = false; if (why < 0) why = 5; }
    }
    if (X.trace && pre_ok && why >= 0 && !X.tuple_mode) atomicAdd(&X.ctr->n_rej[why], 1ull);
    if (!ok && !force_emit) return false;
    const int k = min_in + 2 * min_sh;
    if (k > SR_NMAX) {
        if (ok) atomicAdd(&X.ctr->n_errflag_range, 1ull);  // the reference panics on sr[x][u]
        if (!force_emit) return false;
    }
    s.q1 = q1;
    s.q2 = q2;
    s.cd1 = (uint16_t)cd1;
    s.cd2 = (uint16_t)cd2;
    s.sh0 = sh0;
    s.in0 = in0;
    s.sh1 = sh1;
    s.in1 = in1;
    s.nrefs = (uint8_t)nrefs;
    s.err = err ? 1 : 0;
    s.ok = ok ? 1 : 0;
    s.regfast = (uint8_t)rg.regfast;
    s.fd = (uint16_t)rg.fd;
    s.c1d = (uint16_t)rg.c1d;
    s.c2d = (uint16_t)rg.c2d;
    s.pad = 0;
    s.slot = slot;
    if (k <= SR_NMAX) p1_request(X.p1, sb.n3, k, min_sh, X.ctr);
    return true;
}

// Warp-private slot reservation: one atomicAdd per SURV_CHUNK survivors instead of one per survivor
// (a single-address atomic per candidate serialises the whole kernel in L2).
struct SurvWriter {
    unsigned long long pos, end;   // warp-uniform
    unsigned long long written;
    __device__ __forceinline__ void init() { pos = end = 0; written = 0; }
    __device__ __forceinline__ void retire(const S2Ctx& X, int lane) {   // mark the unused tail of the current chunk
        for (unsigned long long i = pos + lane; i < end && i < X.surv_cap; i += 32) X.surv[i].q1 = SURV_INVALID;
        pos = end;
    }
    __device__ __forceinline__ void push(const S2Ctx& X, int lane, bool emit, const Survivor& s) {   // all 32 lanes must call
        const unsigned int m = __ballot_sync(0xffffffffu, emit);
        if (!m) return;
        const int cnt = __popc(m);
        if (pos + cnt > end) {
            retire(X, lane);
            unsigned long long b = 0;
            if (lane == 0) b = atomicAdd(&X.ctr->n_surv, (unsigned long long)SURV_CHUNK);
            b = __shfl_sync(0xffffffffu, b, 0);
            pos = b;
            end = b + SURV_CHUNK;
            if (end > X.surv_cap && lane == 0) X.ctr->overflow = 1ull;
        }
        if (emit) {
            const unsigned long long o = pos + __popc(m & ((1u << lane) - 1u));
            if (o < X.surv_cap) X.surv[o] = s;
        }
        pos += cnt;
        written += cnt;
    }
    __device__ __forceinline__ void finish(const S2Ctx& X, int lane) {
        retire(X, lane);
        if (lane == 0 && written) atomicAdd(&X.ctr->n_surv_total, written);
    }
};

__device__ __forceinline__ void ld4(const uint4* p, uint32_t* o) {
    const uint4 v = __ldg(p);
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}

// Write-once per-row memo of T(row, refset) = number of compared positions where the row's contigs differ from ANOTHER
// row's V/J references (on that reference set's compare ranges).  It depends only on the row and the foreign
// reference set (all rows of a sub-bucket have equal contig lengths), and it is all the degradation test
// (join_one.rs:434-440) needs, so a row that keeps meeting the same foreign reference set — an indel row against its
// whole clone — is rejected without touching the planes again.  state: 0 empty, 1 being written, 2 ready (immutable).
__device__ __forceinline__ bool memo_lookup(const int32_t* memo, uint32_t q, int vsh, int vsl, int jsh, int jsl, int& T) {
    const int32_t* m = memo + (size_t)q * 8;
    if (*reinterpret_cast<const volatile int32_t*>(m) != 2) return false;
    __threadfence();
    if (m[1] != vsh || m[2] != vsl || m[3] != jsh || m[4] != jsl) return false;
    T = m[5];
    return true;
}
// `dead` = the totals fail the degradation test for row q against that reference set: the winner of the slot also publishes
// the fact to stage 1 (deadq[q] = tag of the foreign reference set; written once, consistent with the memo).
__device__ __forceinline__ void memo_store(int32_t* memo, uint32_t q, int vsh, int vsl, int jsh, int jsl, int T, bool dead,
                                           unsigned long long* deadq, const unsigned long long* tagq, uint32_t q_other) {
    int32_t* m = memo + (size_t)q * 8;
    if (atomicCAS(m, 0, 1) != 0) return;   // somebody else owns the slot
    m[1] = vsh; m[2] = vsl; m[3] = jsh; m[4] = jsl; m[5] = T;
    __threadfence();
    *reinterpret_cast<volatile int32_t*>(m) = 2;
    if (dead && deadq) deadq[q] = tagq[q_other];
}

// Per-candidate values the owner lane derives once and hands to the group working on it (by shuffle).
struct CandInfo {
    uint32_t meta;        // W1 | W4h<<3 | W4l<<7 | gap<<11 | fast<<13 | nrefs1<<14 | ch<<16
    uint32_t q1, q2;      // packed row positions (the group reads region ranges / segment ids from their aux records)
    uint32_t rec1, rec2;  // uint4 index of the two row records in rowstore
    uint32_t cw1, cw2;    // word index of the two rows in plane 0 / word 0 of cdr3w
    uint32_t npad;        // row stride of the CDR3 planes
    uint32_t L2;          // contig lengths: heavy | light << 16 (both <= 1024 on this path)
    uint32_t rc;          // CDR3 rotation | light CDR3 length << 16
    int tot1, tot2;
};
__device__ __forceinline__ CandInfo shfl_cand(const CandInfo& c, int src) {
    CandInfo r;
    r.meta = __shfl_sync(0xffffffffu, c.meta, src);
    r.q1 = __shfl_sync(0xffffffffu, c.q1, src);
    r.q2 = __shfl_sync(0xffffffffu, c.q2, src);
    r.rec1 = __shfl_sync(0xffffffffu, c.rec1, src);
    r.rec2 = __shfl_sync(0xffffffffu, c.rec2, src);
    r.cw1 = __shfl_sync(0xffffffffu, c.cw1, src);
    r.cw2 = __shfl_sync(0xffffffffu, c.cw2, src);
    r.npad = __shfl_sync(0xffffffffu, c.npad, src);
    r.L2 = __shfl_sync(0xffffffffu, c.L2, src);
    r.rc = __shfl_sync(0xffffffffu, c.rc, src);
    r.tot1 = __shfl_sync(0xffffffffu, c.tot1, src);
    r.tot2 = __shfl_sync(0xffffffffu, c.tot2, src);
    return r;
}
// reference geometries of chain M of the two rows of a two-reference candidate
template <int M>
__device__ __forceinline__ void chain_geoms(const S2Ctx& X, uint32_t q1, uint32_t q2, int L, RefGeom& Ga, RefGeom& Gb) {
    const int32_t* a1 = X.aux + (size_t)q1 * AUX_INTS;
    const int32_t* a2 = X.aux + (size_t)q2 * AUX_INTS;
    Ga = ref_geom(X.seglen, a1[M ? AX_VSL : AX_VSH], a1[M ? AX_JSL : AX_JSH], L, X.P.ref_v_trim, X.P.ref_j_trim);
    Gb = ref_geom(X.seglen, a2[M ? AX_VSL : AX_VSH], a2[M ? AX_JSL : AX_JSH], L, X.P.ref_v_trim, X.P.ref_j_trim);
}

// Plane math of ONE candidate by a group of G lanes (G = 4 or 8); lane g handles window g (128 positions) of
// every plane.  Results are packed sums (every field stays < 65536 after the group reduction):
//   r[0] = cd1 | cd2<<16      CDR3 differences heavy | light
//   r[1] = A | B<<16          A = popc(d1 & d2), B = popc(d1 & d2 & e); e = bases differ (incl. gaps)
//   r[2] = diffs | fd<<16     full-length differences | FWR1 differences
//   r[3] = c1d | c2d<<16      CDR1 | CDR2 differences                       (regions of row 1, AX_RGN*)
//   two references only: r[4] = A0|B0, r[5] = O0|T01, r[6] = A1|B1, r[7] = O1|T10
//     x21 = (t2 != r1) on cmp1, x12 = (t1 != r2) on cmp2;  A0 = popc(d1&x21) ... (join_one.rs:343-429)
//     r1 / cmp1 = row 1's V / J reference bases and compare mask in contig coordinates, sliced from the segment planes
// WS = plane stride in uint4 when known at compile time (4), 0 = runtime stride `st`
template <int WS, int M>
__device__ __forceinline__ void s2_chain(const S2Ctx& X, uint32_t q1, uint32_t q2, int L, const uint4* __restrict__ pa, const uint4* __restrict__ pb, int st,
                                         int g, bool gapbit, bool nrefs1, uint32_t& pk1, uint32_t& pk2, uint32_t& pk3, uint32_t& n0, uint32_t& n1,
                                         uint32_t& n2, uint32_t& n3) {
    const int S = WS ? WS : st;
    uint32_t al[4], ah[4], ad[4], ag[4] = {0, 0, 0, 0}, bl[4], bh[4], bd[4], bg[4] = {0, 0, 0, 0}, e[4];
    ld4(pa + PL_TLO * S, al); ld4(pa + PL_THI * S, ah); ld4(pa + PL_D * S, ad);
    ld4(pb + PL_TLO * S, bl); ld4(pb + PL_THI * S, bh); ld4(pb + PL_D * S, bd);
    if (gapbit) { ld4(pa + PL_GAP * S, ag); ld4(pb + PL_GAP * S, bg); }
    int A = 0, B = 0, df = 0;
#pragma unroll
    for (int w = 0; w < 4; w++) {
        e[w] = (al[w] ^ bl[w]) | (ah[w] ^ bh[w]) | (ag[w] ^ bg[w]);
        const uint32_t both = ad[w] & bd[w];
        A += __popc(both);
        B += __popc(both & e[w]);
        df += __popc(e[w]);
    }
    pk1 += (uint32_t)A | ((uint32_t)B << 16);
    pk2 += (uint32_t)df;
    if (M == 0) {   // FWR1 / CDR1 / CDR2 of row 1: pairwise disjoint ranges [lo, hi), all empty unless B_REG_OK
        const int32_t* rgn = X.aux + (size_t)q1 * AUX_INTS + AX_RGN0;
        const int r0 = rgn[0], r1 = rgn[1], r2 = rgn[2];
        int fd = 0, c1 = 0, c2 = 0;
#pragma unroll
        for (int w = 0; w < 4; w++) {
            const int p0 = 32 * (4 * g + w);
            fd += __popc(e[w] & mask_lt32((r0 >> 16) - p0) & ~mask_lt32((r0 & 0xffff) - p0));
            c1 += __popc(e[w] & mask_lt32((r1 >> 16) - p0) & ~mask_lt32((r1 & 0xffff) - p0));
            c2 += __popc(e[w] & mask_lt32((r2 >> 16) - p0) & ~mask_lt32((r2 & 0xffff) - p0));
        }
        pk2 += (uint32_t)fd << 16;
        pk3 += (uint32_t)c1 | ((uint32_t)c2 << 16);
    }
    if (!nrefs1) {
        RefGeom Ga, Gb;
        chain_geoms<M>(X, q1, q2, L, Ga, Gb);
        int A0 = 0, B0 = 0, O0 = 0, T01 = 0, A1 = 0, B1 = 0, O1 = 0, T10 = 0;
#pragma unroll
        for (int w = 0; w < 4; w++) {
            uint32_t arl, arh, ac, brl, brh, bc;
            ref_word(X.segpl, Ga, 4 * g + w, arl, arh, ac);
            ref_word(X.segpl, Gb, 4 * g + w, brl, brh, bc);
            const uint32_t x21 = ((bl[w] ^ arl) | (bh[w] ^ arh) | bg[w]) & ac;  // t2 != r1 on k1's ranges
            const uint32_t x12 = ((al[w] ^ brl) | (ah[w] ^ brh) | ag[w]) & bc;  // t1 != r2 on k2's ranges
            const uint32_t d1 = ad[w], d2 = bd[w];
            A0 += __popc(d1 & x21); B0 += __popc(d1 & x21 & e[w]); O0 += __popc(d1 ^ x21); T01 += __popc(x21);
            A1 += __popc(x12 & d2); B1 += __popc(x12 & d2 & e[w]); O1 += __popc(x12 ^ d2); T10 += __popc(x12);
        }
        n0 += (uint32_t)A0 | ((uint32_t)B0 << 16);
        n1 += (uint32_t)O0 | ((uint32_t)T01 << 16);
        n2 += (uint32_t)A1 | ((uint32_t)B1 << 16);
        n3 += (uint32_t)O1 | ((uint32_t)T10 << 16);
    }
}

// T01 | T10<<16 of one chain window: the two totals the degradation test needs (join_one.rs:434-440)
template <int WS, int M>
__device__ __forceinline__ uint32_t s2_chain_degr(const S2Ctx& X, uint32_t q1, uint32_t q2, int L, const uint4* __restrict__ pa, const uint4* __restrict__ pb,
                                                  int st, int g, bool gapbit) {
    const int S = WS ? WS : st;
    RefGeom Ga, Gb;
    chain_geoms<M>(X, q1, q2, L, Ga, Gb);
    uint32_t al[4], ah[4], ag[4] = {0, 0, 0, 0}, bl[4], bh[4], bg[4] = {0, 0, 0, 0};
    ld4(pa + PL_TLO * S, al); ld4(pa + PL_THI * S, ah); ld4(pb + PL_TLO * S, bl); ld4(pb + PL_THI * S, bh);
    if (gapbit) { ld4(pa + PL_GAP * S, ag); ld4(pb + PL_GAP * S, bg); }
    int T01 = 0, T10 = 0;
#pragma unroll
    for (int w = 0; w < 4; w++) {
        uint32_t arl, arh, ac, brl, brh, bc;
        ref_word(X.segpl, Ga, 4 * g + w, arl, arh, ac);
        ref_word(X.segpl, Gb, 4 * g + w, brl, brh, bc);
        T01 += __popc(((bl[w] ^ arl) | (bh[w] ^ arh) | bg[w]) & ac);
        T10 += __popc(((al[w] ^ brl) | (ah[w] ^ brh) | ag[w]) & bc);
    }
    return (uint32_t)T01 | ((uint32_t)T10 << 16);
}

template <int G>
__device__ __forceinline__ void s2_planes(const S2Ctx& X, const CandInfo& ci, int g, unsigned int gmask, uint32_t* r) {
    const int Lh = (int)(ci.L2 & 0xffff), Ll = (int)(ci.L2 >> 16);
    const int W1 = ci.meta & 7, W4h = (ci.meta >> 3) & 15, W4l = (ci.meta >> 7) & 15, gap = (ci.meta >> 11) & 3, ch = ci.meta >> 16;
    bool active = (ci.meta >> 13) & 1;
    const bool nrefs1 = (ci.meta >> 14) & 1;
    constexpr int WS = (G == 4) ? 4 : 0;   // the 4-lane path only runs when both chains have W4 == 4
    const int off1 = W4h * rec_planes(0, gap & 1);
    const uint4* pa0 = X.rowstore + ci.rec1 + g;
    const uint4* pb0 = X.rowstore + ci.rec2 + g;
    const uint4* pa1 = pa0 + off1;
    const uint4* pb1 = pb0 + off1;
    uint32_t pk0 = 0, pk1 = 0, pk2 = 0, pk3 = 0, n0 = 0, n1 = 0, n2 = 0, n3 = 0;
    // Two references: nearly every such candidate dies on the degradation test (join_one.rs:434-440), which only
    // needs T01 = popc(t2 != r1 on cmp1) and T10 = popc(t1 != r2 on cmp2) — evaluate that first and stop early.
    if (active && !nrefs1 && !X.tuple_mode) {
        uint32_t tt = 0;
        if (g < W4h) tt += s2_chain_degr<WS, 0>(X, ci.q1, ci.q2, Lh, pa0, pb0, W4h, g, gap & 1);
        if (g < W4l) tt += s2_chain_degr<WS, 1>(X, ci.q1, ci.q2, Ll, pa1, pb1, W4l, g, gap & 2);
#pragma unroll
        for (int o = 1; o < G; o <<= 1) tt += __shfl_xor_sync(gmask, tt, o);
        const int T01 = (int)(tt & 0xffff), T10 = (int)(tt >> 16);
        if (abs(ci.tot1 - T10) > X.P.max_degradation || abs(T01 - ci.tot2) > X.P.max_degradation) {
            active = false;          // rejected: phase 2 repeats the test on these totals and drops the candidate
            n1 = (uint32_t)T01 << 16;
            n3 = (uint32_t)T10 << 16;
        }
    }
    if (active) {
        // CDR3 distances heavy | light.  The stored concatenation is rotated left by `rot`: stored positions [0, ch - rot) and
        // [T - rot, T) are heavy, the ones in between light (bits >= T are zero in every plane).
        const int rot = (int)(ci.rc & 0xffff), T = ch + (int)(ci.rc >> 16);
        for (int w = g; w < W1; w += G) {
            const uint32_t* b = X.cdr3w;
            const uint32_t x = (b[ci.cw1 + w * ci.npad] ^ b[ci.cw2 + w * ci.npad]) | (b[ci.cw1 + (W1 + w) * ci.npad] ^ b[ci.cw2 + (W1 + w) * ci.npad]);
            const int lo_bit = w * 32;
            const uint32_t hm = mask_lt32(ch - rot - lo_bit) | (mask_lt32(T - lo_bit) & ~mask_lt32(T - rot - lo_bit));
            pk0 += (uint32_t)__popc(x & hm) | ((uint32_t)__popc(x & ~hm) << 16);
        }
        if (g < W4h) s2_chain<WS, 0>(X, ci.q1, ci.q2, Lh, pa0, pb0, W4h, g, gap & 1, nrefs1, pk1, pk2, pk3, n0, n1, n2, n3);
        if (g < W4l) s2_chain<WS, 1>(X, ci.q1, ci.q2, Ll, pa1, pb1, W4l, g, gap & 2, nrefs1, pk1, pk2, pk3, n0, n1, n2, n3);
    }
#pragma unroll
    for (int o = 1; o < G; o <<= 1) {
        pk0 += __shfl_xor_sync(0xffffffffu, pk0, o);
        pk1 += __shfl_xor_sync(0xffffffffu, pk1, o);
        pk2 += __shfl_xor_sync(0xffffffffu, pk2, o);
        pk3 += __shfl_xor_sync(0xffffffffu, pk3, o);
    }
    if (active && !nrefs1) {  // group-uniform branch: shuffle inside the group only
#pragma unroll
        for (int o = 1; o < G; o <<= 1) {
            n0 += __shfl_xor_sync(gmask, n0, o);
            n1 += __shfl_xor_sync(gmask, n1, o);
            n2 += __shfl_xor_sync(gmask, n2, o);
            n3 += __shfl_xor_sync(gmask, n3, o);
        }
    }
    r[0] = pk0; r[1] = pk1; r[2] = pk2; r[3] = pk3; r[4] = n0; r[5] = n1; r[6] = n2; r[7] = n3;
}

// Stage 2a.  A warp works on batches of 32 consecutive candidates:
//   phase 1: the plane math, G lanes per candidate (G = 4 when both contigs fit 4 uint4 = 512 nt, else 8),
//            packed results to shared memory;
//   phase 2: one lane per candidate runs the scalar integer tests and emits the survivor.
// A CTA takes CONTIGUOUS chunks of the candidate list: consecutive candidates come from one stage-1 tile
// (<= 128 + 64 distinct rows), so their row records are re-read from L1 instead of L2.
#ifndef S2A_MINB
#define S2A_MINB 2   // 128 registers, no spills: measured 6 % faster than 3 (80 registers, 500 B of spills); 4 is 30 % slower
#endif
__global__ void __launch_bounds__(256, S2A_MINB) k_stage2a(S2Ctx X, const uint2* __restrict__ cand, const uint32_t* __restrict__ cand_slot,
                                                    const unsigned long long* __restrict__ n_cand_p, unsigned long long cand_cap) {
    __shared__ unsigned long long s_chunk;
    __shared__ uint32_t s_res[8][32][9];   // [warp][candidate][8 packed results]; padded row: conflict-free column reads
    if (X.ctr->overflow) return;   // the round cannot be committed: the host splits it and retries
    const unsigned long long n_cand = min(*n_cand_p, cand_cap);
    // contiguous chunks keep the rows of one stage-1 tile in L1; a short list (the later rounds) is cut into smaller chunks so
    // that it still spreads over every SM instead of keeping a few CTAs busy for eight batches each
    const unsigned long long S2A_CHUNK = max(256ull, min(2048ull, (n_cand / gridDim.x + 255ull) & ~255ull));
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5, nwb = blockDim.x >> 5;
    SurvWriter wr;
    wr.init();
    unsigned int n_two = 0, n_memo = 0, n_degr = 0;   // diagnostics, per lane; reduced once at the end
    for (;;) {
      __syncthreads();
      if (threadIdx.x == 0) s_chunk = atomicAdd(&X.ctr->s2_work, S2A_CHUNK);
      __syncthreads();
      const unsigned long long c0 = s_chunk;
      if (c0 >= n_cand) break;
      const unsigned long long c1 = min(c0 + S2A_CHUNK, n_cand);
      for (unsigned long long base = c0 + (unsigned long long)wib * 32; base < c1; base += (unsigned long long)nwb * 32) {
        // ---- my own candidate (phase 2 owner) ----
        const unsigned long long ci = base + lane;
        bool valid = ci < c1;
        uint2 pr = make_uint2(0, 0);
        if (valid) pr = cand[ci];
        const uint32_t q1 = pr.x, q2 = pr.y;
        const int sid = valid ? X.subq[q1] : 0;
        if (valid && !X.tuple_mode && (int)q1 >= X.qcut[sid]) valid = false;   // row truncated from this round's unit after stage 1
        // a batch whose candidates were all cut away (the rows behind the first dense block of a unit) costs nothing more
        if (!__any_sync(0xffffffffu, valid)) continue;
        const SubBucket sb = X.subs[sid];
        const int32_t* a1 = X.aux + (size_t)(valid ? q1 : 0) * AUX_INTS;
        const int32_t* a2 = X.aux + (size_t)(valid ? q2 : 0) * AUX_INTS;
        uint4 a10 = make_uint4(0, 0, 0, 0), a20 = make_uint4(0, 0, 0, 0);
        if (valid) {
            a10 = __ldg(reinterpret_cast<const uint4*>(a1));
            a20 = __ldg(reinterpret_cast<const uint4*>(a2));
        }
        const uint32_t f1 = a10.x, f2 = a20.x;
        const bool nrefs1 = valid && (a10.y == a20.y && a10.z == a20.z && a10.w == a20.w && a1[AX_JSL] == a2[AX_JSL]);
        const bool fast = valid && ((f1 | f2) & F_SLOWMASK) == 0 && sb.W1 > 0 && sb.W4h > 0 && sb.W4l > 0;
        const bool wide = fast && (sb.W4h > 4 || sb.W4l > 4);
        const int tot1 = valid ? a1[AX_TOT] : 0, tot2 = valid ? a2[AX_TOT] : 0;
        // phase 0: two references — try the memoised totals first
        bool memo_dead = false;
        if (fast && !nrefs1 && !X.tuple_mode) {
            int T;
            if (memo_lookup(X.memo, q1, (int)a20.y, (int)a20.z, (int)a20.w, a2[AX_JSL], T) && abs(tot1 - T) > X.P.max_degradation) memo_dead = true;        // T10
            else if (memo_lookup(X.memo, q2, (int)a10.y, (int)a10.z, (int)a10.w, a1[AX_JSL], T) && abs(T - tot2) > X.P.max_degradation) memo_dead = true;   // T01
        }
        CandInfo me;
        me.meta = (uint32_t)sb.W1 | ((uint32_t)sb.W4h << 3) | ((uint32_t)sb.W4l << 7) | ((uint32_t)sb.gap << 11) | ((uint32_t)(fast && !memo_dead) << 13) |
                  ((uint32_t)nrefs1 << 14) | ((uint32_t)sb.ch << 16);
        me.q1 = q1;
        me.q2 = q2;
        me.L2 = (uint32_t)(sb.Lh & 0xffff) | ((uint32_t)(sb.Ll & 0xffff) << 16);
        me.rc = (uint32_t)(sb.rot & 0xffff) | ((uint32_t)(sb.cl & 0xffff) << 16);
        me.rec1 = (uint32_t)(sb.rs_off + (int64_t)(q1 - sb.q0) * sb.rec_u4);
        me.rec2 = (uint32_t)(sb.rs_off + (int64_t)(q2 - sb.q0) * sb.rec_u4);
        me.cw1 = (uint32_t)(sb.s1_off + (q1 - sb.q0));
        me.cw2 = (uint32_t)(sb.s1_off + (q2 - sb.q0));
        me.npad = (uint32_t)sb.npad;
        me.tot1 = tot1;
        me.tot2 = tot2;
        // The column row of a candidate is cold more often than not (a row is a column of only a few candidates): every owner
        // lane asks L2 for the lines of its two records now, so that the plane loads of phase 1 — eight candidates at a time —
        // find them there instead of paying the DRAM latency one batch after the other.
        if (fast && !memo_dead) {
            const char* r1p = reinterpret_cast<const char*>(X.rowstore + me.rec1);
            const char* r2p = reinterpret_cast<const char*>(X.rowstore + me.rec2);
            const int rec_bytes = sb.rec_u4 * 16;
            for (int o = 0; o < rec_bytes; o += 128) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(r2p + o));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(r1p + o));
            }
        }
        // ---- phase 1 ----
        if (!__any_sync(0xffffffffu, wide)) {
#pragma unroll 1
            for (int it = 0; it < 4; it++) {
                const int src = it * 8 + (lane >> 2);           // candidate (lane index of its owner) this group works on
                const CandInfo gc = shfl_cand(me, src);
                uint32_t r[8];
                s2_planes<4>(X, gc, lane & 3, 0xfu << (lane & 28), r);
                if ((lane & 3) == 0) {
#pragma unroll
                    for (int i = 0; i < 8; i++) s_res[wib][src][i] = r[i];
                }
            }
        } else {
#pragma unroll 1
            for (int it = 0; it < 8; it++) {
                const int src = it * 4 + (lane >> 3);
                const CandInfo gc = shfl_cand(me, src);
                uint32_t r[8];
                s2_planes<8>(X, gc, lane & 7, 0xffu << (lane & 24), r);
                if ((lane & 7) == 0) {
#pragma unroll
                    for (int i = 0; i < 8; i++) s_res[wib][src][i] = r[i];
                }
            }
        }
        __syncwarp();
        // ---- phase 2: one lane per candidate ----
        bool emit = false;
        Survivor sv;
        if (valid) {
            const uint32_t slot = cand_slot ? cand_slot[ci] : (uint32_t)ci;
            if (fast && !nrefs1 && !X.tuple_mode) n_two++;
            if (fast && memo_dead) {
                // rejected by the degradation test on memoised totals
                n_memo++;
            } else if (fast) {
                const uint32_t* r = s_res[wib][lane];
                const int cd1 = (int)(r[0] & 0xffff), cd2 = (int)(r[0] >> 16);
                const int diffs = (int)(r[2] & 0xffff);
                // region differences are usable when both rows: chain 0 is the only left chain, seq_del_amino is the
                // contig, well-formed ranges, and identical feature starts
                RegionDiffs rg;
                const uint32_t need = B_LEFT_H | B_AMINO_TIG_H | B_REG_OK;
                const uint4 a14 = __ldg(reinterpret_cast<const uint4*>(a1) + 4), a15 = __ldg(reinterpret_cast<const uint4*>(a1) + 5);
                const uint4 a24 = __ldg(reinterpret_cast<const uint4*>(a2) + 4), a25 = __ldg(reinterpret_cast<const uint4*>(a2) + 5);
                const bool same = a14.z == a24.z && a14.w == a24.w && a15.x == a25.x && a15.y == a25.y && a15.z == a25.z;
                rg.regfast = (same && (f1 & (need | B_LEFT_L)) == need && (f2 & (need | B_LEFT_L)) == need) ? 1 : 0;
                rg.fd = (int)(r[2] >> 16);
                rg.c1d = (int)(r[3] & 0xffff);
                rg.c2d = (int)(r[3] >> 16);
                if (nrefs1) {
                    const int A = (int)(r[1] & 0xffff), B = (int)(r[1] >> 16);
                    emit = s2_integer_tests(X, q1, q2, slot, sb, cd1, cd2, diffs, 1, A - B, tot1 + tot2 - 2 * A + 2 * B, 0, 0, rg, true, X.tuple_mode != 0, sv);
                } else {
                    const int A0 = (int)(r[4] & 0xffff), B0 = (int)(r[4] >> 16), O0 = (int)(r[5] & 0xffff), T01 = (int)(r[5] >> 16);
                    const int A1 = (int)(r[6] & 0xffff), B1 = (int)(r[6] >> 16), O1 = (int)(r[7] & 0xffff), T10 = (int)(r[7] >> 16);
                    // join_one.rs:434-440: total[0] = (tot1, T01), total[1] = (T10, tot2)
                    const bool dead1 = abs(tot1 - T10) > X.P.max_degradation;   // q1 fails against q2's reference set, whoever carries it
                    const bool dead2 = abs(T01 - tot2) > X.P.max_degradation;
                    if (!X.tuple_mode) {
                        memo_store(X.memo, q1, (int)a20.y, (int)a20.z, (int)a20.w, a2[AX_JSL], T10, dead1, X.deadq, X.tagq, q2);
                        memo_store(X.memo, q2, (int)a10.y, (int)a10.z, (int)a10.w, a1[AX_JSL], T01, dead2, X.deadq, X.tagq, q1);
                    }
                    const bool ok = !dead1 && !dead2;
                    if (!ok && !X.tuple_mode) n_degr++;
                    if (ok || X.tuple_mode)
                        emit = s2_integer_tests(X, q1, q2, slot, sb, cd1, cd2, diffs, 2, A0 - B0, O0 + 2 * B0, A1 - B1, O1 + 2 * B1, rg, ok, X.tuple_mode != 0, sv);
                }
            } else {
                const unsigned long long si = atomicAdd(&X.ctr->n_slow, 1ull);
                if (si < X.slow_cap) {
                    X.slow[si] = pr;
                    X.slow_slot[si] = slot;
                } else {
                    X.ctr->overflow = 1ull;
                }
            }
        }
        __syncwarp();   // s_res is rewritten by the next batch
        wr.push(X, lane, emit, sv);
      }
    }
    wr.finish(X, lane);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n_two += __shfl_xor_sync(0xffffffffu, n_two, o);
        n_memo += __shfl_xor_sync(0xffffffffu, n_memo, o);
        n_degr += __shfl_xor_sync(0xffffffffu, n_degr, o);
    }
    if (lane == 0 && n_two) {
        atomicAdd(&X.ctr->n_two_ref, (unsigned long long)n_two);
        if (n_memo) atomicAdd(&X.ctr->n_memo_dead, (unsigned long long)n_memo);
        if (n_degr) atomicAdd(&X.ctr->n_degr_dead, (unsigned long long)n_degr);
    }
}

// Slow path: one warp per candidate, literal byte-level restatement of join_one.rs:216-440
// ('N' and other odd bytes, overlapping V/J ranges, the V bail-out, over-long contigs).
__global__ void __launch_bounds__(256) k_stage2_slow(S2Ctx X, const unsigned long long* __restrict__ n_slow_p) {
    if (X.ctr->overflow) return;
    const unsigned long long n_slow = min(*n_slow_p, X.slow_cap);
    const int lane = threadIdx.x & 31;
    const unsigned long long warp0 = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned long long nwarps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    const DevIn& in = X.in;
    for (unsigned long long ci = warp0; ci < n_slow; ci += nwarps) {
        const uint2 pr = X.slow[ci];
        const uint32_t q1 = pr.x, q2 = pr.y, slot = X.slow_slot[ci];
        const SubBucket sb = X.subs[X.subq[q1]];
        const int64_t k1 = X.rowq[q1], k2 = X.rowq[q2];
        const int32_t* a1 = X.aux + (size_t)q1 * AUX_INTS;
        const int32_t* a2 = X.aux + (size_t)q2 * AUX_INTS;
        const uint32_t f1 = (uint32_t)a1[AX_FLAGS], f2 = (uint32_t)a2[AX_FLAGS];
        if ((f1 | f2) & F_PANIC) {
            if (lane == 0) atomicAdd(&X.ctr->n_errflag_panic, 1ull);
            if (!X.tuple_mode) continue;
        }
        // CDR3 distances on bytes
        int cd1 = 0, cd2 = 0;
        {
            const TigRef a = cdr3_ref(in, k1, 0), b = cdr3_ref(in, k2, 0);
            for (int p = lane; p < sb.ch; p += 32) cd1 += (a[p] != b[p]);
            const TigRef a2 = cdr3_ref(in, k1, 1), b2 = cdr3_ref(in, k2, 1);
            for (int p = lane; p < sb.cl; p += 32) cd2 += (a2[p] != b2[p]);
        }
        // full-length differences (join_one.rs:238-260)
        int diffs = 0;
        for (int m = 0; m < 2; m++) {
            const int L = m ? sb.Ll : sb.Lh;
            const TigRef a = tig_ref(in, k1, m), b = tig_ref(in, k2, m);
            const bool hd = in.row_has_del[2 * k1 + m] || in.row_has_del[2 * k2 + m];
            if (!hd) {
                for (int p = lane; p < L; p += 32) diffs += (base2bit(a[p]) != base2bit(b[p]));
            } else {
                for (int p = lane; p < L; p += 32) diffs += (a[p] != b[p]);
            }
        }
        const int nrefs = (a1[AX_VSH] == a2[AX_VSH] && a1[AX_VSL] == a2[AX_VSL] && a1[AX_JSH] == a2[AX_JSH] && a1[AX_JSL] == a2[AX_JSL]) ? 1 : 2;
        int sh[2] = {0, 0}, ind[2] = {0, 0}, tot[2][2] = {{0, 0}, {0, 0}};
        bool bail = false;
        for (int u = 0; u < nrefs; u++) {
            const int64_t k = u ? k2 : k1;
            for (int m = 0; m < 2; m++) {
                const int L = m ? sb.Ll : sb.Lh;
                const TigRef t1 = tig_ref(in, k1, m), t2 = tig_ref(in, k2, m);
                for (int si = 0; si < 2; si++) {
                    const int sg = si ? in.row_js[2 * k + m] : in.row_vs[2 * k + m];
                    const uint8_t* seg = in.seg_bytes + in.seg_off[sg];
                    const long long slen = in.seg_off[sg + 1] - in.seg_off[sg];
                    long long rng = slen - (si ? X.P.ref_j_trim : X.P.ref_v_trim);
                    if (rng < 0) rng = 0;  // flagged F_PANIC
                    if (si == 0 && rng > L) {
                        bail = true;  // p reaches tig.len(): return false (join_one.rs:371-373)
                        rng = L;
                    }
                    if (si == 1 && rng > L) rng = L;  // flagged F_PANIC
                    for (long long p = lane; p < rng; p += 32) {
                        uint8_t a, b, r;
                        if (si == 0) {
                            a = t1[p];
                            b = t2[p];
                            r = bit2base(base2bit(seg[p]));
                        } else {
                            a = t1[L - p - 1];
                            b = t2[L - p - 1];
                            r = bit2base(base2bit(seg[slen - p - 1]));
                        }
                        if (a == b && a != r) sh[u]++;
                        else if ((a == r && b != r) || (b == r && a != r)) ind[u]++;
                        else if (a != r && b != r) ind[u] += 2;
                        if (a != r) tot[u][0]++;
                        if (b != r) tot[u][1]++;
                    }
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            cd1 += __shfl_xor_sync(0xffffffffu, cd1, o);
            cd2 += __shfl_xor_sync(0xffffffffu, cd2, o);
            diffs += __shfl_xor_sync(0xffffffffu, diffs, o);
#pragma unroll
            for (int u = 0; u < 2; u++) {
                sh[u] += __shfl_xor_sync(0xffffffffu, sh[u], o);
                ind[u] += __shfl_xor_sync(0xffffffffu, ind[u], o);
                tot[u][0] += __shfl_xor_sync(0xffffffffu, tot[u][0], o);
                tot[u][1] += __shfl_xor_sync(0xffffffffu, tot[u][1], o);
            }
        }
        if (lane == 0) {
            bool ok = !bail && !((f1 | f2) & F_PANIC);
            if (nrefs == 2) {  // join_one.rs:434-440
                for (int m = 0; m < 2; m++) {
                    const long long a = tot[0][m], b = tot[1][m];
                    if ((a <= b ? b - a : a - b) > X.P.max_degradation) ok = false;
                }
            }
            RegionDiffs rg = {0, 0, 0, 0};
            Survivor sv;
            if ((ok || X.tuple_mode) &&
                s2_integer_tests(X, q1, q2, slot, sb, cd1, cd2, diffs, nrefs, sh[0], ind[0], sh[1], ind[1], rg, ok, X.tuple_mode != 0, sv)) {
                const unsigned long long si = atomicAdd(&X.ctr->n_surv, 1ull);   // rare path: one atomic per survivor is fine
                if (si < X.surv_cap) X.surv[si] = sv; else X.ctr->overflow = 1ull;
                atomicAdd(&X.ctr->n_surv_total, 1ull);
            }
        }
    }
}

// sorted-list intersection of the two exacts' barcodes (vector_utils meet, join_one.rs:451-470)
__device__ __forceinline__ bool bc_meet(const S2Ctx& X, int e1, int e2) {
    int64_t i = X.in.ex_cell_off[e1], ie = X.in.ex_cell_off[e1 + 1];
    int64_t j = X.in.ex_cell_off[e2], je = X.in.ex_cell_off[e2 + 1];
    while (i < ie && j < je) {
        const uint32_t a = (uint32_t)X.bc_sorted[i], b = (uint32_t)X.bc_sorted[j];
        if (a < b) i++;
        else if (b < a) j++;
        else return true;
    }
    return false;
}

// Stage 2b: one thread per survivor.  Emits pass edges (filter mode) or EdgeTuples (tuple mode).
#ifndef S2B_MINB
#define S2B_MINB 4   // 64 registers: measured 0.25 ms faster per join than the unconstrained 128
#endif
__global__ void __launch_bounds__(256, S2B_MINB) k_stage2b(S2Ctx X, const unsigned long long* __restrict__ n_surv_p, unsigned long long* __restrict__ pass_w,
                                                 PassTuple* __restrict__ pass_t, unsigned long long pass_cap, EdgeTuple* __restrict__ tuples) {
    if (X.ctr->overflow) return;
    const unsigned long long n_surv = min(*n_surv_p, X.surv_cap);
    const DevIn& in = X.in;
    const DevParams& P = X.P;
    const int lane = threadIdx.x & 31;
    for (unsigned long long sbase = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) - lane; sbase < n_surv;
         sbase += (unsigned long long)gridDim.x * blockDim.x) {
      const unsigned long long si = sbase + lane;
      bool pass_edge = false;
      unsigned long long pass_val = 0;
      PassTuple pass_tup;
      int pass_sub = -1;
      Survivor s;
      s.q1 = SURV_INVALID;
      if (si < n_surv) s = X.surv[si];
      if (s.q1 != SURV_INVALID) {
        const uint32_t q1 = s.q1, q2 = s.q2;
        const SubBucket sb = X.subs[X.subq[q1]];
        const int32_t* a1 = X.aux + (size_t)q1 * AUX_INTS;
        const int32_t* a2 = X.aux + (size_t)q2 * AUX_INTS;
        const uint32_t f1 = (uint32_t)a1[AX_FLAGS], f2 = (uint32_t)a2[AX_FLAGS];
        bool ok = s.ok != 0;
        const int cd = (int)s.cd1 + (int)s.cd2;
        const int min_sh = (s.nrefs == 2) ? min(s.sh0, s.sh1) : s.sh0;
        const int min_in = (s.nrefs == 2) ? min(s.in0, s.in1) : s.in0;
        // :83-89 exact subclonotypes with 2 or 3 chains only
        if (!(f1 & B_NCH_OK) || !(f2 & B_NCH_OK)) ok = false;
        // :451-470 barcode overlap
        if (ok || X.tuple_mode) {
            if (a1[AX_NCELLS] == 1 && a2[AX_NCELLS] == 1) {
                if (a1[AX_BC0] == a2[AX_BC0]) ok = false;
            } else if (bc_meet(X, a1[AX_EXACT], a2[AX_EXACT])) {
                ok = false;
            }
        }
        // :499-544
        const int k = min_in + 2 * min_sh, d = min_sh;
        double p1 = 0.0;
        if (k <= SR_NMAX) p1 = p1_lookup(X.p1, sb.n3, k, d); else ok = false;
        const double mult = __dmul_rn(X.powtab[X.pow_off[sb.ch] + s.cd1], X.powtab[X.pow_off[sb.cl] + s.cd2]);
        const double score = __dmul_rn(p1, mult);
        // :548-567 JUN_SHARE
        bool accept = false;
        if (P.comp_filt < 1000000 && score > P.max_score && min_sh < P.auto_share && (P.comp_filt_bound == 0 || min_in <= P.comp_filt_bound) &&
            (f1 & B_NCH2) && (f2 & B_NCH2) && (f1 & B_JUN_LEFT)) {
            const long long comp = min(a1[AX_HCOMP], a2[AX_HCOMP]);
            if (comp - cd >= P.comp_filt) accept = true;
        }
        // :710-715
        if (!accept && score > P.max_score && min_sh < P.auto_share) ok = false;
        // :722-758 V gene names (rare: names differ)
        if (ok && (a1[AX_VNAMEH] != a2[AX_VNAMEH] || a1[AX_VNAMEL] != a2[AX_VNAMEL])) {
            for (int i = 0; i < 2 && ok; i++) {
                if (a1[AX_VNAMEH + i] == a2[AX_VNAMEH + i]) continue;
                const int va = a1[AX_VREFH + i], vb = a2[AX_VREFH + i];
                const uint8_t *y1 = in.ref_bytes + in.ref_off[va], *y2 = in.ref_bytes + in.ref_off[vb];
                const long long l1 = in.ref_off[va + 1] - in.ref_off[va], l2 = in.ref_off[vb + 1] - in.ref_off[vb];
                const long long nn = min(l1, l2);
                for (long long m = 0; m < nn; m++)
                    if (base2bit(y1[m]) != base2bit(y2[m])) {
                        ok = false;
                        break;
                    }
                const int x1 = a1[AX_CHH + i], x2 = a2[AX_CHH + i];
                const int ua = in.ch_u_ref[x1], ub = in.ch_u_ref[x2];
                if (ok && ua != -1 && ub != -1) {
                    const uint8_t *w1 = in.ref_bytes + in.ref_off[ua], *w2 = in.ref_bytes + in.ref_off[ub];
                    const long long m1 = in.ref_off[ua + 1] - in.ref_off[ua], m2 = in.ref_off[ub + 1] - in.ref_off[ub];
                    const long long mm = min(m1, m2);
                    for (long long m = 0; m < mm; m++)
                        if (base2bit(w1[m1 - 1 - m]) != base2bit(w2[m2 - 1 - m])) {
                            ok = false;
                            break;
                        }
                }
            }
        }
        // :766-854 FWR1 identity minus CDR1+2 identity on the left chain
        if (ok) {
            long long fwr1_len = 0, cdr1_len = 0, cdr2_len = 0, fwr1_diffs = 0, cdr1_diffs = 0, cdr2_diffs = 0;
            bool panic = false;
            if (s.regfast) {
                // both rows: chain 0 is the only left chain, identical feature starts, seq_del_amino == contig
                const int fr1 = a1[AX_FR1], cdr1 = a1[AX_CDR1], fr2 = a1[AX_FR2], cdr2 = a1[AX_CDR2], fr3 = a1[AX_FR3];
                if (cdr1 != -1) { fwr1_len = cdr1 - fr1; fwr1_diffs = s.fd; }
                if (cdr1 != -1 && fr2 != -1 && cdr1 <= fr2) { cdr1_len = fr2 - cdr1; cdr1_diffs = s.c1d; }
                if (cdr2 != -1 && fr3 != -1 && cdr2 <= fr3) { cdr2_len = fr3 - cdr2; cdr2_diffs = s.c2d; }
            } else {
                const int64_t k1 = X.rowq[q1], k2 = X.rowq[q2];
                for (int m = 0; m < 2; m++) {
                    const int x1 = a1[AX_CHH + m], x2 = a2[AX_CHH + m];
                    if (!in.ch_left[x1]) continue;
                    TigRef s1, s2;
                    long long sl1, sl2;
                    if (in.ch_amino_off[x1] != -1) { s1.b = in.amino_bytes + in.ch_amino_off[x1]; s1.w = nullptr; sl1 = in.ch_amino_len[x1]; }
                    else { s1 = tig_ref(in, k1, m); sl1 = m ? sb.Ll : sb.Lh; }
                    if (in.ch_amino_off[x2] != -1) { s2.b = in.amino_bytes + in.ch_amino_off[x2]; s2.w = nullptr; sl2 = in.ch_amino_len[x2]; }
                    else { s2 = tig_ref(in, k2, m); sl2 = m ? sb.Ll : sb.Lh; }
                    auto region = [&](long long r1a, long long r1b, long long r2a, long long r2b, long long& rlen, long long& rdiffs) {
                        const long long len = r1b - r1a;
                        if (r2b - r2a == len) {
                            if (len > 0 && (r1a + len > sl1 || r2a + len > sl2 || r1a < 0 || r2a < 0)) { panic = true; return; }
                            long long df = 0;
                            for (long long p = 0; p < len; p++) df += (s1[p + r1a] != s2[p + r2a]);
                            rlen = len;
                            rdiffs = df;
                        }
                    };
                    const int f1a = in.ch_fr1[x1], c1a = in.ch_cdr1[x1], f2a = in.ch_fr2[x1], c2a = in.ch_cdr2[x1], f3a = in.ch_fr3[x1];
                    const int f1b = in.ch_fr1[x2], c1b = in.ch_cdr1[x2], f2b = in.ch_fr2[x2], c2b = in.ch_cdr2[x2], f3b = in.ch_fr3[x2];
                    if (c1a != -1 && c1b != -1) {
                        if (c1a < f1a || c1b < f1b) panic = true;  // usize underflow in the reference
                        else region(f1a, c1a, f1b, c1b, fwr1_len, fwr1_diffs);
                    }
                    if (c1a != -1 && f2a != -1 && c1b != -1 && f2b != -1 && c1a <= f2a && c1b <= f2b) region(c1a, f2a, c1b, f2b, cdr1_len, cdr1_diffs);
                    if (c2a != -1 && f3a != -1 && c2b != -1 && f3b != -1 && c2a <= f3a && c2b <= f3b) region(c2a, f3a, c2b, f3b, cdr2_len, cdr2_diffs);
                }
            }
            if (panic) {
                atomicAdd(&X.ctr->n_errflag_panic, 1ull);
                ok = false;
            } else if (fwr1_len > 0 && cdr1_len > 0 && cdr2_len > 0) {
                const double fwr1_identity = 100.0 * (double)(fwr1_len - fwr1_diffs) / (double)fwr1_len;
                const long long len = cdr1_len + cdr2_len, df = cdr1_diffs + cdr2_diffs;
                const double cdr12_identity = 100.0 * (double)(len - df) / (double)len;
                if (fwr1_identity - cdr12_identity >= P.fwr1_cdr12_delta) ok = false;
            }
        }
        if (X.tuple_mode) {
            EdgeTuple t;
            t.cd = cd;
            t.nrefs = s.nrefs;
            t.sh0 = s.sh0;
            t.sh1 = (s.nrefs == 2) ? s.sh1 : 0;
            t.in0 = s.in0;
            t.in1 = (s.nrefs == 2) ? s.in1 : 0;
            t.k = k;
            t.d = d;
            t.n = sb.n3;
            t.err = s.err;
            t.pass = ok ? 1 : 0;
            t.pad = 0;
            t.p1 = p1;
            t.mult = mult;
            t.score = score;
            tuples[s.slot] = t;
        } else if (ok) {
            pass_edge = true;
            pass_val = ((unsigned long long)(uint32_t)X.rowq[q1] << 28) | (unsigned long long)(uint32_t)X.rowq[q2];
            pass_sub = X.subq[q1];
            pass_tup.cd12 = (uint32_t)s.cd1 | ((uint32_t)s.cd2 << 16);
            pass_tup.sh0 = s.sh0;
            pass_tup.in0 = s.in0;
            pass_tup.sh1 = (s.nrefs == 2) ? s.sh1 : 0;
            pass_tup.in1 = (s.nrefs == 2) ? s.in1 : 0;
            pass_tup.p1 = p1;
            pass_tup.nrefs_err = (uint32_t)s.nrefs | ((uint32_t)s.err << 8) | (sb.owner < 0 ? PT_SPLIT : 0u);   // owner < 0: sub-bucket split over ranks
        }
      }
      // warp-aggregated append: one atomic per warp
      const unsigned int m = __ballot_sync(0xffffffffu, pass_edge);
      if (m) {
          unsigned long long b = 0;
          if (lane == 0) b = atomicAdd(&X.ctr->n_pass, (unsigned long long)__popc(m));
          b = __shfl_sync(0xffffffffu, b, 0);
          if (pass_edge) {
              const unsigned long long o = b + __popc(m & ((1u << lane) - 1u));
              if (o < pass_cap) {
                  pass_w[o] = pass_val;
                  pass_t[o] = pass_tup;
              } else {
                  X.ctr->overflow = 1ull;
              }
              // per-sub-bucket pass counter, aggregated over the lanes of the warp that hit the same sub-bucket
              const unsigned int same = __match_any_sync(m, pass_sub);
              if ((same & ((1u << lane) - 1u)) == 0) atomicAdd(&X.sub_pass[pass_sub], (unsigned int)__popc(same));
          }
      }
    }
}

}  // namespace ejb
