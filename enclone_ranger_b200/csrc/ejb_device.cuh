// ejb_device.cuh — sm_100a device code of the clonotype join.
//
// Kernel map (reference function each one replaces):
//   k_heads / k_keys / k_subheads / k_subinfo   bucket scan of join_exacts   enclone/src/join.rs:68-88
//   k_pack_rows                                 the "bucketer": 2-bit planes + reference-difference
//                                               bitmaps (restates the per-base compares of
//                                               enclone_core/src/join_one.rs:343-429 per ROW)
//   k_stage1<W1>                                CDR3 Hamming pre-filter       join_one.rs:216-234, 271-290
//                                               + same-class skip            enclone/src/join_core.rs:32
//   k_stage2a / k_stage2_slow                   shares/indeps/diffs + integer tests
//                                                                            join_one.rs:238-267, 301-440, 474-493
//   k_p1                                        p_at_most_m_distinct_..._double join_one.rs:22-43
//   k_stage2b                                   barcode meet, score, JUN_SHARE, V names, FWR1/CDR12
//                                                                            join_one.rs:451-470, 499-567, 710-854
//   k_bor_*                                     lexicographic spanning forest == pot of join_core
//                                                                            join_core.rs:23-50
//   k_two_cell / k_exact_union / k_labels       join.rs:120-178, join2.rs:69-84
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dd.cuh"

namespace ejb {

constexpr int TILE = 128;          // rows / cols of one stage-1 tile
constexpr int S1_THREADS = 256;
constexpr int S1_STRIP = 4;        // column tiles per CTA (row fragment stays in registers)
constexpr int MAXW1 = 6;           // concatenated CDR3 (heavy+light) up to 192 nt on the fast path
constexpr int MAXW4 = 8;           // contig up to 1024 nt on the stage-2 fast path
constexpr int SR_NMAX = 3000;      // stirling2_ratio_table_double(3000), start.rs:344

enum RowFlags : uint32_t {
    F_NONACGT_H = 1u, F_NONACGT_L = 2u,   // tig has a byte outside ACGT (incl. '-')
    F_CDR3_IRREG = 4u,                    // a CDR3 byte outside ACGT
    F_SLOWGEOM = 8u,                      // V/J ranges overlap, or the V range exceeds the contig
    F_PANIC = 16u,                        // the reference would panic on this row (J range > contig ...)
    F_HASDEL_H = 32u, F_HASDEL_L = 64u,
    F_SLOWMASK = F_NONACGT_H | F_NONACGT_L | F_CDR3_IRREG | F_SLOWGEOM | F_PANIC
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
    int32_t rec_u4;   // uint4 per row record
    int32_t maxcd;    // stage-1 threshold on the CDR3 distance (superset of join_one.rs:231,286-290)
    int32_t blk0;     // first global 128-row block id
    int32_t n3;       // 3 * (Lh + Ll)  (join_one.rs:499)
    int32_t owner;    // rank that owns this sub-bucket (multi-GPU sharding)
};

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
};

// Flat join knobs on the device.
struct DevParams {
    int is_bcr, mix_donors, easy, old_light;
    double max_score, cdr3_mult, cdr3_ident_thr /* 1.0 - join_cdr3_ident/100.0 */, fwr1_cdr12_delta;
    long long max_cdr3_diffs, auto_share, comp_filt, comp_filt_bound, max_diffs, max_degradation, ref_v_trim, ref_j_trim;
};

// Packed-row (position q) metadata, SoA.
struct RowMeta {
    int32_t* row;       // canonical row k
    int32_t* sub;       // sub-bucket id
    uint32_t* flags;
    int2* vs;           // canonical seg ids (heavy, light)
    int2* js;
    int32_t* tot;       // popc of the reference-difference bitmap (V and J ranges, both chains)
    int32_t* exact;
    int2* chain;        // chain ids of (heavy, light)
};

struct Survivor {       // a candidate that passed the integer tests of stage 2
    uint32_t q1, q2;
    uint16_t cd1, cd2;
    int32_t sh0, in0, sh1, in1;
    uint8_t nrefs, err;
    uint16_t pad;
    uint32_t slot;      // output slot (tuple mode)
};

struct EdgeTuple {      // numeric members of PotentialJoin (defs.rs:870-891)
    int32_t cd, nrefs, sh0, sh1, in0, in1, k, d, n, err, pass, pad;
    double p1, mult, score;
};

struct Counters {       // device-side counters, one instance per context
    unsigned long long n_cand, n_slow, n_surv, n_pass, n_newkeys, pairs_eval, n_forest, n_alive;
    unsigned long long n_two_cell, n_errflag_range, n_errflag_panic, overflow, n_cand_total, n_surv_total, n_pass_total;
    unsigned long long p1_unique;
};

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

// ------------------------------------------------------------------------------------------------
// Bucket scan (join.rs:68-88): head flag = lens vector differs from the previous row's.
// Also validates per-row offsets / ids (error bits OR-ed into *err).
// ------------------------------------------------------------------------------------------------
__global__ void k_heads(DevIn in, int32_t* head, unsigned int* err) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= in.n_rows) return;
    const int nc = in.row_nchains[k];
    unsigned int e = 0;
    if (nc < 1 || nc > 4) e |= 1;
    if (in.row_exact[k] < 0 || in.row_exact[k] >= in.n_exacts) e |= 2;
    for (int m = 0; m < 2 && m < nc; m++) {
        const int64_t i = 2 * k + m;
        const int L = in.row_len[i];
        if (L < 0 || in.row_tig_off[i] < 0 || in.row_tig_off[i] + L > in.n_tig_bytes) e |= 4;
        const int c = in.row_cdr3_len[i];
        if (c < 0 || c > 65535 || in.row_cdr3_off[i] < 0 || in.row_cdr3_off[i] + c > in.n_cdr3_bytes) e |= 8;
        if (in.row_vs[i] < 0 || in.row_vs[i] >= in.n_segs || in.row_js[i] < 0 || in.row_js[i] >= in.n_segs) e |= 16;
        if (e == 0) {
            const int ex = in.row_exact[k];
            const int64_t nch = in.ex_chain_off[ex + 1] - in.ex_chain_off[ex];
            if (in.row_exact_col[i] < 0 || in.row_exact_col[i] >= nch) e |= 32;
        }
    }
    if (e) atomicOr(err, e);
    int h = 1;
    if (k > 0) {
        h = 0;
        if (in.row_nchains[k - 1] != nc) h = 1;
        for (int m = 0; m < 2 && m < nc && !h; m++)
            if (in.row_len[2 * k + m] != in.row_len[2 * (k - 1) + m]) h = 1;
        // lens vectors longer than 2 (improper rows) never join; compare only what is stored
    }
    head[k] = h;
}

// validates exact / chain / cell tables
__global__ void k_validate_tables(DevIn in, unsigned int* err) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int e = 0;
    if (i < in.n_exacts) {
        if (in.ex_chain_off[i] < 0 || in.ex_chain_off[i + 1] < in.ex_chain_off[i] || in.ex_chain_off[i + 1] > in.n_chains) e |= 64;
        if (in.ex_cell_off[i] < 0 || in.ex_cell_off[i + 1] < in.ex_cell_off[i] || in.ex_cell_off[i + 1] > in.n_cells) e |= 128;
    }
    if (i < in.n_chains) {
        if (in.ch_v_ref[i] < 0 || in.ch_v_ref[i] >= in.n_refs) e |= 256;
        if (in.ch_u_ref[i] < -1 || in.ch_u_ref[i] >= in.n_refs) e |= 256;
        if (in.ch_amino_off[i] != -1 && (in.ch_amino_off[i] < 0 || in.ch_amino_len[i] < 0 || in.ch_amino_off[i] + in.ch_amino_len[i] > in.n_amino_bytes)) e |= 512;
    }
    if (i < in.n_cells) {
        if (in.cell_donor[i] < -1 || in.cell_donor[i] >= 64) e |= 1024;  // donor sets are 64-bit masks
    }
    if (i < in.n_segs) {
        if (in.seg_off[i] < 0 || in.seg_off[i + 1] < in.seg_off[i] || in.seg_off[i + 1] > in.n_seg_bytes) e |= 2048;
    }
    if (i < in.n_refs) {
        if (in.ref_off[i] < 0 || in.ref_off[i + 1] < in.ref_off[i] || in.ref_off[i + 1] > in.n_ref_bytes) e |= 4096;
    }
    if (e) atomicOr(err, e);
}

// sort key of a row: (bucket id, heavy CDR3 len, light CDR3 len); rows that can never join
// (lens.len() != 2, join_one.rs:83-96) get the maximal key and fall off the end.
__global__ void k_keys(DevIn in, const int32_t* bucket_incl, uint64_t* keys, uint32_t* vals, unsigned long long* n_active) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t key = ~0ull;
    bool act = false;
    if (k < in.n_rows && in.row_nchains[k] == 2) {
        const uint64_t b = (uint64_t)(bucket_incl[k] - 1);
        key = (b << 32) | ((uint64_t)in.row_cdr3_len[2 * k] << 16) | (uint64_t)in.row_cdr3_len[2 * k + 1];
        act = true;
    }
    const unsigned int m = __ballot_sync(0xffffffffu, act);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_active, (unsigned long long)__popc(m));
    if (k < in.n_rows) {
        keys[k] = key;
        vals[k] = (uint32_t)k;
    }
}

__global__ void k_subheads(const uint64_t* keys, int64_t n_active, int32_t* head) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_active) return;
    head[q] = (q == 0 || keys[q] != keys[q - 1]) ? 1 : 0;
}

struct SubInfo {
    int64_t q0;
    uint64_t key;
    int32_t Lh, Ll;
};
__global__ void k_subinfo(DevIn in, const uint64_t* keys, const uint32_t* row_of, const int32_t* head, const int32_t* sub_incl,
                          int64_t n_active, SubInfo* out) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_active || !head[q]) return;
    const int s = sub_incl[q] - 1;
    const int64_t k = row_of[q];
    SubInfo si;
    si.q0 = q;
    si.key = keys[q];
    si.Lh = in.row_len[2 * k];
    si.Ll = in.row_len[2 * k + 1];
    out[s] = si;
}

// ------------------------------------------------------------------------------------------------
// The bucketer: one warp per packed row.  Writes
//   * CDR3 planes (heavy ++ light concatenated), layout [plane*W1 + w][npad] per sub-bucket
//   * the row record  [lo_h | hi_h | d_h | lo_l | hi_l | d_l], each W4 uint4
//     d = 1 where the contig base differs from its own V (positions < |V|-ref_v_trim) or
//     J (last |J|-ref_j_trim positions) reference base: the per-row half of join_one.rs:365-425
//   * row metadata
// ------------------------------------------------------------------------------------------------
__global__ void k_pack_rows(DevIn in, DevParams P, const SubBucket* subs, const int32_t* sub_incl, const uint32_t* row_of,
                            int64_t n_active, uint32_t* cdr3w, uint4* rowstore, RowMeta M) {
    const int lane = threadIdx.x & 31;
    const int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (j >= n_active) return;
    const int s = sub_incl[j] - 1;
    const SubBucket sb = subs[s];
    const int64_t k = row_of[j];
    const int64_t li = j - sb.j0;
    const int64_t q = sb.q0 + li;
    uint32_t flags = 0;
    int tot = 0;
    int2 vs, js;
    vs.x = in.seg_canon[in.row_vs[2 * k]];
    vs.y = in.seg_canon[in.row_vs[2 * k + 1]];
    js.x = in.seg_canon[in.row_js[2 * k]];
    js.y = in.seg_canon[in.row_js[2 * k + 1]];

    // ---- CDR3 planes ----
    {
        const uint8_t* c0 = in.cdr3_bytes + in.row_cdr3_off[2 * k];
        const uint8_t* c1 = in.cdr3_bytes + in.row_cdr3_off[2 * k + 1];
        const int T = sb.ch + sb.cl;
        const int nw = (T + 31) / 32;
        bool irr = false;
        for (int w = 0; w < nw; w++) {
            const int p = w * 32 + lane;
            int code = 0;
            if (p < T) {
                const uint8_t b = (p < sb.ch) ? c0[p] : c1[p - sb.ch];
                if (!is_acgt(b)) irr = true;
                code = base2bit(b);
            }
            const uint32_t lo = __ballot_sync(0xffffffffu, code & 1);
            const uint32_t hi = __ballot_sync(0xffffffffu, code & 2);
            if (sb.W1 > 0 && lane == 0) {
                cdr3w[sb.s1_off + (int64_t)(w)*sb.npad + li] = lo;
                cdr3w[sb.s1_off + (int64_t)(sb.W1 + w) * sb.npad + li] = hi;
            }
        }
        if (__any_sync(0xffffffffu, irr)) flags |= F_CDR3_IRREG;
    }

    // ---- contig planes + reference-difference bitmap ----
    uint4* rec = rowstore + sb.rs_off + li * (int64_t)sb.rec_u4;
    uint32_t* rec32 = reinterpret_cast<uint32_t*>(rec);
    int base_u4 = 0;
    for (int m = 0; m < 2; m++) {
        const int L = (m == 0) ? sb.Lh : sb.Ll;
        const int W4 = (m == 0) ? sb.W4h : sb.W4l;
        const uint8_t* tig = in.tig_bytes + in.row_tig_off[2 * k + m];
        const int sv = in.row_vs[2 * k + m], sj = in.row_js[2 * k + m];
        const uint8_t* vseg = in.seg_bytes + in.seg_off[sv];
        const uint8_t* jseg = in.seg_bytes + in.seg_off[sj];
        const long long vlen = in.seg_off[sv + 1] - in.seg_off[sv];
        const long long jlen = in.seg_off[sj + 1] - in.seg_off[sj];
        const long long vr = vlen - P.ref_v_trim, jr = jlen - P.ref_j_trim;  // compared ranges
        if (vr < 0 || jr < 0 || jr > L) flags |= F_PANIC;   // usize underflow / index panic in the reference
        if (vr > L) flags |= F_SLOWGEOM;                    // "ugly bailout" join_one.rs:371-373
        if (vr + jr > L) flags |= F_SLOWGEOM;               // V and J ranges overlap: positions counted twice
        if (in.row_has_del[2 * k + m]) flags |= (m == 0) ? F_HASDEL_H : F_HASDEL_L;
        bool irr = false;
        const int nw = (L + 31) / 32;
        for (int w = 0; w < 4 * W4; w++) {
            uint32_t lo = 0, hi = 0, dm = 0;
            if (w < nw) {
                const int p = w * 32 + lane;
                int code = 0, dbit = 0;
                if (p < L) {
                    const uint8_t b = tig[p];
                    if (!is_acgt(b)) irr = true;
                    code = base2bit(b);
                    if (p < vr) {
                        dbit = (b != bit2base(base2bit(vseg[p])));
                    } else if (p >= L - jr) {
                        // position counted from the contig end: tig[L-1-pp] vs J[|J|-1-pp]
                        const long long pp = (long long)L - 1 - p;
                        dbit = (b != bit2base(base2bit(jseg[jlen - 1 - pp])));
                    }
                }
                lo = __ballot_sync(0xffffffffu, code & 1);
                hi = __ballot_sync(0xffffffffu, code & 2);
                dm = __ballot_sync(0xffffffffu, dbit);
                tot += __popc(dm);
            }
            if (lane == 0) {
                rec32[(base_u4)*4 + w] = lo;
                rec32[(base_u4 + W4) * 4 + w] = hi;
                rec32[(base_u4 + 2 * W4) * 4 + w] = dm;
            }
        }
        if (W4 == 0) {  // contig longer than the fast path: still need the irregular flag
            for (int p = lane; p < L; p += 32)
                if (!is_acgt(tig[p])) irr = true;
        }
        if (__any_sync(0xffffffffu, irr)) flags |= (m == 0) ? F_NONACGT_H : F_NONACGT_L;
        base_u4 += 3 * W4;
    }
    if (lane == 0) {
        M.row[q] = (int32_t)k;
        M.sub[q] = s;
        M.flags[q] = flags;
        M.vs[q] = vs;
        M.js[q] = js;
        M.tot[q] = tot;
        const int ex = in.row_exact[k];
        M.exact[q] = ex;
        int2 c;
        c.x = (int)(in.ex_chain_off[ex] + in.row_exact_col[2 * k]);
        c.y = (int)(in.ex_chain_off[ex] + in.row_exact_col[2 * k + 1]);
        M.chain[q] = c;
    }
}

// per exact: donor set as a 64-bit mask (join_one.rs:301-315) and sort keys for the barcode lists
__global__ void k_exact_prep(DevIn in, unsigned long long* donor_mask, uint64_t* bc_keys) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= in.n_exacts) return;
    unsigned long long m = 0;
    for (int64_t j = in.ex_cell_off[e]; j < in.ex_cell_off[e + 1]; j++) {
        const int d = in.cell_donor[j];
        if (d >= 0) m |= 1ull << d;
        bc_keys[j] = ((uint64_t)e << 32) | (uint64_t)in.cell_barcode[j];
    }
    donor_mask[e] = m;
}

// ------------------------------------------------------------------------------------------------
// Stage 1: all pairs of a sub-bucket, tiled 128 x 128.  One CTA = one row block x a strip of up to
// S1_STRIP column blocks.  Row/column CDR3 planes + class labels are staged in shared memory by TMA
// bulk copies (double-buffered columns); each thread keeps an 8-row fragment in registers and walks
// 8 columns per tile.  Distance = popc of (lo^lo')|(hi^hi') per 32 bases, words pre-reduced with a
// carry-save adder so that 3 words cost 2 popc.  A pair becomes a candidate iff
// cd <= maxcd  &&  class labels differ (join_core.rs:32 — a pair already connected by
// lexicographically smaller joins can never be in pot).
// ------------------------------------------------------------------------------------------------
struct RowTask {       // one row block of one sub-bucket
    int32_t sub;
    int32_t rb;
    int64_t cta0;      // first CTA of this task inside the round
};

template <int W1>
__device__ __forceinline__ int cdr3_dist(const uint32_t* rl, const uint32_t* rh, const uint32_t* cl, const uint32_t* chh) {
    uint32_t x[W1];
#pragma unroll
    for (int w = 0; w < W1; w++) x[w] = (rl[w] ^ cl[w]) | (rh[w] ^ chh[w]);
    if (W1 == 1) return __popc(x[0]);
    if (W1 == 2) return __popc(x[0]) + __popc(x[1]);
    uint32_t s = x[0] ^ x[1] ^ x[2];
    uint32_t c = (x[0] & x[1]) | (x[2] & (x[0] ^ x[1]));
    if (W1 == 3) return __popc(s) + 2 * __popc(c);
    if (W1 == 4) return __popc(s) + __popc(x[3]) + 2 * __popc(c);
    if (W1 == 5) {
        uint32_t s2 = s ^ x[3] ^ x[4];
        uint32_t c2 = (s & x[3]) | (x[4] & (s ^ x[3]));
        return __popc(s2) + 2 * (__popc(c) + __popc(c2));
    }
    // W1 == 6
    uint32_t s3 = x[3] ^ x[4] ^ x[5];
    uint32_t c3 = (x[3] & x[4]) | (x[5] & (x[3] ^ x[4]));
    return __popc(s) + __popc(s3) + 2 * (__popc(c) + __popc(c3));
}

template <int W1>
__global__ void __launch_bounds__(S1_THREADS) k_stage1(const SubBucket* __restrict__ subs, const RowTask* __restrict__ tasks, int n_tasks,
                                                       const uint32_t* __restrict__ cdr3w, const int32_t* __restrict__ labq,
                                                       const int32_t* __restrict__ blk_lab, uint2* __restrict__ cand,
                                                       unsigned long long cand_cap, Counters* __restrict__ ctr) {
    constexpr int NP = 2 * W1;           // plane-words per row
    constexpr int SLOT = (NP + 1) * TILE; // words per staged tile (planes + labels)
    __shared__ __align__(16) uint32_t s_row[SLOT];
    __shared__ __align__(16) uint32_t s_col[2][SLOT];
    __shared__ __align__(8) uint64_t s_bar[3];
    __shared__ uint16_t s_hits[TILE * TILE / 2];
    __shared__ unsigned int s_nhits;
    __shared__ unsigned long long s_base;
    __shared__ int s_cbs[S1_STRIP];
    __shared__ int s_ncb;

    // ---- locate the task: binary search on cta0 ----
    const int64_t cta = blockIdx.x;
    int lo = 0, hi = n_tasks - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (tasks[mid].cta0 <= cta) lo = mid; else hi = mid - 1;
    }
    const RowTask t = tasks[lo];
    const SubBucket sb = subs[t.sub];
    if (sb.W1 != W1) return;  // launched once per W1; other widths are handled by their own launch
    const int nb = (sb.n + TILE - 1) / TILE;
    const int rb = t.rb;
    const int cb_first = rb + (int)(cta - t.cta0) * S1_STRIP;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;

    // class-skip at block level: both blocks uniformly in one class
    if (tid == 0) {
        int n = 0;
        const int lr = blk_lab[sb.blk0 + rb];
        for (int c = cb_first; c < nb && c < cb_first + S1_STRIP; c++) {
            const int lc = blk_lab[sb.blk0 + c];
            if (lr >= 0 && lr == lc) continue;
            s_cbs[n++] = c;
        }
        s_ncb = n;
        s_nhits = 0;
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        mbar_init(&s_bar[2], 1);
        mbar_fence_init();
    }
    __syncthreads();
    const int ncb = s_ncb;
    if (ncb == 0) return;

    const uint32_t* base = cdr3w + sb.s1_off;
    auto issue = [&](uint32_t* dst, int blk, uint64_t* bar) {
        // rows [blk*128, min(n, blk*128+128)) of every plane-word + labels, 16-byte granules
        const int r0 = blk * TILE;
        const int cnt = min(TILE, sb.npad - r0);  // npad is a multiple of 4
        const uint32_t bytes = (uint32_t)cnt * 4u;
        mbar_expect_tx(bar, bytes * (NP + 1));
#pragma unroll
        for (int pw = 0; pw < NP; pw++) tma_load_1d(dst + pw * TILE, base + (int64_t)pw * sb.npad + r0, bytes, bar);
        tma_load_1d(dst + NP * TILE, labq + sb.q0 + r0, bytes, bar);   // q0 is a multiple of 4
    };
    if (tid == 0) {
        issue(s_row, rb, &s_bar[2]);
        issue(s_col[0], s_cbs[0], &s_bar[0]);
        if (ncb > 1) issue(s_col[1], s_cbs[1], &s_bar[1]);
    }

    // ---- row fragment -> registers ----
    mbar_wait(&s_bar[2], 0);
    uint32_t rl[8][W1], rh[8][W1];
    int rlab[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int r = i * 16 + ty;
#pragma unroll
        for (int w = 0; w < W1; w++) {
            rl[i][w] = s_row[w * TILE + r];
            rh[i][w] = s_row[(W1 + w) * TILE + r];
        }
        rlab[i] = (int)s_row[NP * TILE + r];
    }
    const int row_lim = sb.n - rb * TILE;  // valid local rows are < row_lim
    const int maxcd = sb.maxcd;
    unsigned long long evaluated = 0;

    for (int it = 0; it < ncb; it++) {
        const int cb = s_cbs[it];
        const int buf = it & 1;
        mbar_wait(&s_bar[buf], (it >> 1) & 1);
        const uint32_t* sc = s_col[buf];
        const int col_lim = sb.n - cb * TILE;
        const bool diag = (cb == rb);
        const bool full = !diag && row_lim >= TILE && col_lim >= TILE;
        // two halves of 64 columns: the hit list holds at most 128 x 64 entries
        for (int half = 0; half < 2; half++) {
            if (full) {
#pragma unroll 2
                for (int j = half * 4; j < half * 4 + 4; j++) {
                    const int c = j * 16 + tx;
                    uint32_t cl[W1], chh[W1];
#pragma unroll
                    for (int w = 0; w < W1; w++) {
                        cl[w] = sc[w * TILE + c];
                        chh[w] = sc[(W1 + w) * TILE + c];
                    }
                    const int clab = (int)sc[NP * TILE + c];
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const int cd = cdr3_dist<W1>(rl[i], rh[i], cl, chh);
                        if (cd <= maxcd && rlab[i] != clab) {
                            const unsigned int h = atomicAdd(&s_nhits, 1u);
                            s_hits[h] = (uint16_t)(((i * 16 + ty) << 7) | c);
                        }
                    }
                }
            } else {
                for (int j = half * 4; j < half * 4 + 4; j++) {
                    const int c = j * 16 + tx;
                    uint32_t cl[W1], chh[W1];
#pragma unroll
                    for (int w = 0; w < W1; w++) {
                        cl[w] = sc[w * TILE + c];
                        chh[w] = sc[(W1 + w) * TILE + c];
                    }
                    const int clab = (int)sc[NP * TILE + c];
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const int r = i * 16 + ty;
                        const bool valid = r < row_lim && c < col_lim && (!diag || r < c);
                        const int cd = cdr3_dist<W1>(rl[i], rh[i], cl, chh);
                        if (valid && cd <= maxcd && rlab[i] != clab) {
                            const unsigned int h = atomicAdd(&s_nhits, 1u);
                            s_hits[h] = (uint16_t)((r << 7) | c);
                        }
                    }
                }
            }
            __syncthreads();  // all hits of this half recorded (and, after half 1, s_col[buf] drained)
            const unsigned int nh = s_nhits;
            if (tid == 0) {
                if (half == 1) {
                    // prefetch the tile after next into the buffer just drained
                    if (it + 2 < ncb) issue(s_col[buf], s_cbs[it + 2], &s_bar[buf]);
                    const int rr = min(TILE, row_lim), cc = min(TILE, col_lim);
                    evaluated += diag ? (unsigned long long)rr * (rr - 1) / 2 : (unsigned long long)rr * cc;
                }
                if (nh) s_base = atomicAdd(&ctr->n_cand, (unsigned long long)nh);
            }
            if (nh) {
                __syncthreads();
                const unsigned long long b = s_base;
                const uint32_t q_r0 = (uint32_t)(sb.q0 + (int64_t)rb * TILE), q_c0 = (uint32_t)(sb.q0 + (int64_t)cb * TILE);
                for (unsigned int h = tid; h < nh; h += S1_THREADS) {
                    const unsigned long long o = b + h;
                    if (o < cand_cap) {
                        const uint16_t v = s_hits[h];
                        cand[o] = make_uint2(q_r0 + (v >> 7), q_c0 + (v & 127));
                    } else {
                        ctr->overflow = 1ull;
                    }
                }
                __syncthreads();
                if (tid == 0) s_nhits = 0;
                __syncthreads();
            }
        }
    }
    if (tid == 0) atomicAdd(&ctr->pairs_eval, evaluated);
}

// ------------------------------------------------------------------------------------------------
// p1 cache: open-addressing hash (n, k, d) -> f64, filled by k_p1.
// ------------------------------------------------------------------------------------------------
struct P1Cache {
    unsigned long long* keys;  // ~0 = empty
    double* vals;
    uint32_t* newslots;        // slots inserted since the last k_p1
    uint32_t cap_mask;
};
__device__ __forceinline__ unsigned long long p1_key(int n, int k, int d) {
    return ((unsigned long long)(uint32_t)n << 24) | ((unsigned long long)(uint32_t)k << 12) | (unsigned long long)(uint32_t)d;
}
__device__ __forceinline__ uint32_t p1_hash(unsigned long long key) {
    key ^= key >> 33;
    key *= 0xff51afd7ed558ccdull;
    key ^= key >> 33;
    key *= 0xc4ceb9fe1a85ec53ull;
    key ^= key >> 33;
    return (uint32_t)key;
}
__device__ __forceinline__ void p1_request(const P1Cache& c, int n, int k, int d, Counters* ctr) {
    const unsigned long long key = p1_key(n, k, d);
    uint32_t h = p1_hash(key) & c.cap_mask;
    for (;;) {
        const unsigned long long cur = c.keys[h];
        if (cur == key) return;
        if (cur == ~0ull) {
            const unsigned long long old = atomicCAS(&c.keys[h], ~0ull, key);
            if (old == ~0ull) {
                const unsigned long long i = atomicAdd(&ctr->n_newkeys, 1ull);
                c.newslots[i] = h;
                return;
            }
            if (old == key) return;
        }
        h = (h + 1) & c.cap_mask;
    }
}
__device__ __forceinline__ double p1_lookup(const P1Cache& c, int n, int k, int d) {
    const unsigned long long key = p1_key(n, k, d);
    uint32_t h = p1_hash(key) & c.cap_mask;
    for (;;) {
        const unsigned long long cur = c.keys[h];
        if (cur == key) return c.vals[h];
        if (cur == ~0ull) return __longlong_as_double(0x7ff8000000000000ll);  // NaN: must not happen
        h = (h + 1) & c.cap_mask;
    }
}

// One warp per new key: p = 1 - sum_{u=m+1..x} sr[x][u] * (u/n)^x * prod_{v=1..u} (n-v+1)/(u-v+1),
// same operation order as join_one.rs:28-42 (terms are independent; the subtraction runs in
// increasing u so that the double-double rounding sequence is the reference's).
__device__ __forceinline__ double p1_eval_warp(const dd_t* __restrict__ sr, int k, int d, int n, int lane) {
    const int x = k, m = k - d;
    dd_t p = dd_make(1.0);
    const dd_t* row = sr + (size_t)x * (x + 1) / 2;
    for (int u0 = m + 1; u0 <= x; u0 += 32) {
        const int u = u0 + lane;
        dd_t z = dd_make(0.0);
        if (u <= x) {
            z = row[u];
            const dd_t f = dd_div(dd_make((double)u), dd_make((double)n));
            for (int i = 0; i < x; i++) z = dd_mul(z, f);
            for (int v = 1; v <= u; v++) z = dd_mul(z, dd_div(dd_make((double)(n - v + 1)), dd_make((double)(u - v + 1))));
        }
        const int cnt = min(32, x - u0 + 1);
        for (int i = 0; i < cnt; i++) {
            dd_t zi;
            zi.hi = __shfl_sync(0xffffffffu, z.hi, i);
            zi.lo = __shfl_sync(0xffffffffu, z.lo, i);
            p = dd_sub(p, zi);
        }
    }
    if (dd_lt(p, dd_make(0.0))) p = dd_make(0.0);
    return p.hi;
}
__global__ void k_p1(P1Cache c, const dd_t* __restrict__ sr, const unsigned long long* n_new) {
    const int lane = threadIdx.x & 31;
    const unsigned long long warp0 = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned long long nwarps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    const unsigned long long n = *n_new;
    for (unsigned long long wi = warp0; wi < n; wi += nwarps) {
        const uint32_t slot = c.newslots[wi];
        const unsigned long long key = c.keys[slot];
        const int d = (int)(key & 0xfff), k = (int)((key >> 12) & 0xfff), nn = (int)(key >> 24);
        const double v = p1_eval_warp(sr, k, d, nn, lane);
        if (lane == 0) c.vals[slot] = v;
    }
}

// p1 for explicit triples (parity tests): one warp per triple, same code path as k_p1
__global__ void k_p1_batch(const dd_t* __restrict__ sr, const int32_t* kk, const int32_t* ddv, const int32_t* nn, int64_t count, double* out) {
    const int lane = threadIdx.x & 31;
    const int64_t wi = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (wi >= count) return;
    const double v = p1_eval_warp(sr, kk[wi], ddv[wi], nn[wi], lane);
    if (lane == 0) out[wi] = v;
}

// ------------------------------------------------------------------------------------------------
// Stage 2 context shared by the kernels below.
// ------------------------------------------------------------------------------------------------
struct S2Ctx {
    DevIn in;
    DevParams P;
    const SubBucket* subs;
    RowMeta M;
    const uint32_t* cdr3w;
    const uint4* rowstore;
    const unsigned long long* donor_mask;
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
    int tuple_mode;              // 0: filter -> pass edges; 1: write an EdgeTuple per input pair
};

// Integer tests shared by the fast and the slow path, executed by one lane per candidate.
// Returns false when join_one would return false; otherwise appends a Survivor + requests p1.
__device__ __forceinline__ bool s2_integer_tests(const S2Ctx& X, uint32_t q1, uint32_t q2, uint32_t slot, const SubBucket& sb, int cd1, int cd2,
                                                 long long diffs, int nrefs, int sh0, int in0, int sh1, int in1, bool pre_ok, bool force_emit) {
    const DevParams& P = X.P;
    const int cd = cd1 + cd2;
    bool ok = pre_ok;
    // join_one.rs:216-234
    if (P.is_bcr) {
        const int total = sb.ch + sb.cl;
        if ((double)cd / (double)total > P.cdr3_ident_thr) ok = false;
    }
    // :238-267
    if (!P.is_bcr || P.max_diffs < 1000000) {
        if (diffs > P.max_diffs) ok = false;
        if (!P.is_bcr && diffs > 5) ok = false;
    }
    // :286-290
    if ((P.max_cdr3_diffs < 1000 || !P.is_bcr) && (cd > P.max_cdr3_diffs || (!P.is_bcr && cd > 0))) ok = false;
    // :301-323
    const int e1 = X.M.exact[q1], e2 = X.M.exact[q2];
    const unsigned long long m1 = X.donor_mask[e1], m2 = X.donor_mask[e2];
    if (!P.mix_donors && m1 != 0 && m2 != 0 && m1 != m2) ok = false;
    const bool err = (m1 != m2) || __popcll(m1) != 1 || __popcll(m2) != 1;
    // :444-447
    const int min_sh = (nrefs == 2) ? min(sh0, sh1) : sh0;
    const int min_in = (nrefs == 2) ? min(in0, in1) : in0;
    // :474-476
    if ((double)cd >= P.cdr3_mult * (double)max(1, min_in)) ok = false;
    // :481-493
    if (!P.old_light) {
        const int2 c1 = X.M.chain[q1], c2 = X.M.chain[q2];
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const int a = i ? c1.y : c1.x, b = i ? c2.y : c2.x;
            if (!X.in.ch_left[a] && X.in.ch_c_ref[a] != -1 && X.in.ch_c_ref[b] != -1 && X.in.ch_c_ref[a] != X.in.ch_c_ref[b] && cd > 0) ok = false;
        }
    }
    if (!ok && !force_emit) return false;
    const int k = min_in + 2 * min_sh;
    if (k > SR_NMAX) {
        if (ok) atomicAdd(&X.ctr->n_errflag_range, 1ull);  // the reference panics on sr[x][u]
        if (!force_emit) return false;
    }
    const unsigned long long si = atomicAdd(&X.ctr->n_surv, 1ull);
    if (si >= X.surv_cap) {
        X.ctr->overflow = 1ull;
        return false;
    }
    Survivor s;
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
    s.pad = ok ? 1 : 0;   // integer tests passed
    s.slot = slot;
    X.surv[si] = s;
    if (k <= SR_NMAX) p1_request(X.p1, sb.n3, k, min_sh, X.ctr);
    return true;
}

// Fast path: 8 lanes per candidate.  Lane g handles uint4 g of every plane (128 positions).
//   e = (lo1^lo2)|(hi1^hi2)   bases differ            (t1 != t2)
//   A = popc(d1 & d2)          both differ from the reference
//   B = popc(d1 & d2 & e)      ... and from each other
//   shares = A - B;  indeps = tot1 + tot2 - 2A + 2B   (join_one.rs:402-419)
__global__ void __launch_bounds__(256) k_stage2a(S2Ctx X, const uint2* __restrict__ cand, const uint32_t* __restrict__ cand_slot,
                                                 const unsigned long long* __restrict__ n_cand_p, unsigned long long cand_cap) {
    const unsigned long long n_cand = min(*n_cand_p, cand_cap);
    const int lane = threadIdx.x & 31;
    const int g = lane & 7, gid = lane >> 3;
    const unsigned long long warp0 = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned long long nwarps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    for (unsigned long long base = warp0 * 4; base < n_cand; base += nwarps * 4) {
        const unsigned long long ci = base + gid;
        const bool valid = ci < n_cand;
        uint2 pr = make_uint2(0, 0);
        if (valid) pr = cand[ci];
        const uint32_t q1 = pr.x, q2 = pr.y;
        const uint32_t slot = (valid && cand_slot) ? cand_slot[ci] : (uint32_t)ci;
        const SubBucket sb = X.subs[X.M.sub[q1]];
        const uint32_t f1 = X.M.flags[q1], f2 = X.M.flags[q2];
        const int2 v1 = X.M.vs[q1], v2 = X.M.vs[q2], j1 = X.M.js[q1], j2 = X.M.js[q2];
        const bool nrefs1 = (v1.x == v2.x && v1.y == v2.y && j1.x == j2.x && j1.y == j2.y);
        const bool fast = nrefs1 && ((f1 | f2) & F_SLOWMASK) == 0 && sb.W1 > 0 && sb.W4h > 0 && sb.W4l > 0;
        int cd1 = 0, cd2 = 0, A = 0, B = 0, diffs = 0;
        if (valid && fast) {
            // CDR3 distances, split at the heavy/light boundary of the concatenated planes
            if (g < sb.W1) {
                const uint32_t* b = X.cdr3w + sb.s1_off;
                const int64_t i1 = q1 - sb.q0, i2 = q2 - sb.q0;
                const uint32_t x = (b[(int64_t)g * sb.npad + i1] ^ b[(int64_t)g * sb.npad + i2]) |
                                   (b[(int64_t)(sb.W1 + g) * sb.npad + i1] ^ b[(int64_t)(sb.W1 + g) * sb.npad + i2]);
                const int lo_bit = g * 32;
                uint32_t hm;  // bits of this word that belong to the heavy CDR3
                if (sb.ch >= lo_bit + 32) hm = 0xffffffffu;
                else if (sb.ch <= lo_bit) hm = 0u;
                else hm = (1u << (sb.ch - lo_bit)) - 1u;
                cd1 = __popc(x & hm);
                cd2 = __popc(x & ~hm);
            }
            const uint4* r1 = X.rowstore + sb.rs_off + (int64_t)(q1 - sb.q0) * sb.rec_u4;
            const uint4* r2 = X.rowstore + sb.rs_off + (int64_t)(q2 - sb.q0) * sb.rec_u4;
#pragma unroll
            for (int m = 0; m < 2; m++) {
                const int W4 = m ? sb.W4l : sb.W4h;
                const int off = m ? 3 * sb.W4h : 0;
                if (g < W4) {
                    const uint4 a_lo = r1[off + g], a_hi = r1[off + W4 + g], a_d = r1[off + 2 * W4 + g];
                    const uint4 b_lo = r2[off + g], b_hi = r2[off + W4 + g], b_d = r2[off + 2 * W4 + g];
                    const uint32_t al[4] = {a_lo.x, a_lo.y, a_lo.z, a_lo.w}, ah[4] = {a_hi.x, a_hi.y, a_hi.z, a_hi.w};
                    const uint32_t ad[4] = {a_d.x, a_d.y, a_d.z, a_d.w};
                    const uint32_t bl[4] = {b_lo.x, b_lo.y, b_lo.z, b_lo.w}, bh[4] = {b_hi.x, b_hi.y, b_hi.z, b_hi.w};
                    const uint32_t bd[4] = {b_d.x, b_d.y, b_d.z, b_d.w};
#pragma unroll
                    for (int w = 0; w < 4; w++) {
                        const uint32_t e = (al[w] ^ bl[w]) | (ah[w] ^ bh[w]);
                        const uint32_t both = ad[w] & bd[w];
                        A += __popc(both);
                        B += __popc(both & e);
                        diffs += __popc(e);
                    }
                }
            }
        }
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            cd1 += __shfl_xor_sync(0xffffffffu, cd1, o);
            cd2 += __shfl_xor_sync(0xffffffffu, cd2, o);
            A += __shfl_xor_sync(0xffffffffu, A, o);
            B += __shfl_xor_sync(0xffffffffu, B, o);
            diffs += __shfl_xor_sync(0xffffffffu, diffs, o);
        }
        if (valid && g == 0) {
            if (fast) {
                const int shares = A - B;
                const int indeps = X.M.tot[q1] + X.M.tot[q2] - 2 * A + 2 * B;
                s2_integer_tests(X, q1, q2, slot, sb, cd1, cd2, diffs, 1, shares, indeps, 0, 0, true, X.tuple_mode != 0);
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
    }
}

// Slow path: one warp per candidate, literal byte-level restatement of join_one.rs:216-440
// (two references, '-' / non-ACGT bytes, overlapping V/J ranges, the V bail-out).
__global__ void __launch_bounds__(256) k_stage2_slow(S2Ctx X, const unsigned long long* __restrict__ n_slow_p) {
    const unsigned long long n_slow = min(*n_slow_p, X.slow_cap);
    const int lane = threadIdx.x & 31;
    const unsigned long long warp0 = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned long long nwarps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    const DevIn& in = X.in;
    for (unsigned long long ci = warp0; ci < n_slow; ci += nwarps) {
        const uint2 pr = X.slow[ci];
        const uint32_t q1 = pr.x, q2 = pr.y, slot = X.slow_slot[ci];
        const SubBucket sb = X.subs[X.M.sub[q1]];
        const int64_t k1 = X.M.row[q1], k2 = X.M.row[q2];
        const uint32_t f1 = X.M.flags[q1], f2 = X.M.flags[q2];
        if ((f1 | f2) & F_PANIC) {
            if (lane == 0) atomicAdd(&X.ctr->n_errflag_panic, 1ull);
            if (!X.tuple_mode) continue;
        }
        // CDR3 distances on bytes
        int cd1 = 0, cd2 = 0;
        {
            const uint8_t *a = in.cdr3_bytes + in.row_cdr3_off[2 * k1], *b = in.cdr3_bytes + in.row_cdr3_off[2 * k2];
            for (int p = lane; p < sb.ch; p += 32) cd1 += (a[p] != b[p]);
            a = in.cdr3_bytes + in.row_cdr3_off[2 * k1 + 1];
            b = in.cdr3_bytes + in.row_cdr3_off[2 * k2 + 1];
            for (int p = lane; p < sb.cl; p += 32) cd2 += (a[p] != b[p]);
        }
        // full-length differences (join_one.rs:238-260)
        int diffs = 0;
        for (int m = 0; m < 2; m++) {
            const int L = m ? sb.Ll : sb.Lh;
            const uint8_t *a = in.tig_bytes + in.row_tig_off[2 * k1 + m], *b = in.tig_bytes + in.row_tig_off[2 * k2 + m];
            const bool hd = in.row_has_del[2 * k1 + m] || in.row_has_del[2 * k2 + m];
            if (!hd) {
                for (int p = lane; p < L; p += 32) diffs += (base2bit(a[p]) != base2bit(b[p]));
            } else {
                for (int p = lane; p < L; p += 32) diffs += (a[p] != b[p]);
            }
        }
        const int2 v1 = X.M.vs[q1], v2 = X.M.vs[q2], j1 = X.M.js[q1], j2 = X.M.js[q2];
        const int nrefs = (v1.x == v2.x && v1.y == v2.y && j1.x == j2.x && j1.y == j2.y) ? 1 : 2;
        int sh[2] = {0, 0}, ind[2] = {0, 0}, tot[2][2] = {{0, 0}, {0, 0}};
        bool bail = false;
        for (int u = 0; u < nrefs; u++) {
            const int64_t k = u ? k2 : k1;
            for (int m = 0; m < 2; m++) {
                const int L = m ? sb.Ll : sb.Lh;
                const uint8_t *t1 = in.tig_bytes + in.row_tig_off[2 * k1 + m], *t2 = in.tig_bytes + in.row_tig_off[2 * k2 + m];
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
            if (ok || X.tuple_mode)
                s2_integer_tests(X, q1, q2, slot, sb, cd1, cd2, diffs, nrefs, sh[0], ind[0], sh[1], ind[1], ok, X.tuple_mode != 0);
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

// Stage 2b: one thread per survivor — barcode overlap, p1 * mult score, JUN_SHARE, score threshold,
// V gene names, FWR1 vs CDR1+2 identity.  Emits pass edges (filter mode) or EdgeTuples (tuple mode).
__global__ void __launch_bounds__(256) k_stage2b(S2Ctx X, const unsigned long long* __restrict__ n_surv_p, unsigned long long* __restrict__ pass_w,
                                                 unsigned long long pass_cap, EdgeTuple* __restrict__ tuples) {
    const unsigned long long n_surv = min(*n_surv_p, X.surv_cap);
    const DevIn& in = X.in;
    const DevParams& P = X.P;
    for (unsigned long long si = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; si < n_surv;
         si += (unsigned long long)gridDim.x * blockDim.x) {
        const Survivor s = X.surv[si];
        const uint32_t q1 = s.q1, q2 = s.q2;
        const SubBucket sb = X.subs[X.M.sub[q1]];
        const int e1 = X.M.exact[q1], e2 = X.M.exact[q2];
        const int64_t k1 = X.M.row[q1], k2 = X.M.row[q2];
        bool ok = s.pad != 0;
        const int cd = (int)s.cd1 + (int)s.cd2;
        const int min_sh = (s.nrefs == 2) ? min(s.sh0, s.sh1) : s.sh0;
        const int min_in = (s.nrefs == 2) ? min(s.in0, s.in1) : s.in0;
        // :451-470
        if (ok || X.tuple_mode) {
            if (bc_meet(X, e1, e2)) ok = false;
        }
        // :499-544
        const int k = min_in + 2 * min_sh, d = min_sh;
        double p1 = 0.0;
        if (k <= SR_NMAX) p1 = p1_lookup(X.p1, sb.n3, k, d); else ok = false;
        const double mult = __dmul_rn(X.powtab[X.pow_off[sb.ch] + s.cd1], X.powtab[X.pow_off[sb.cl] + s.cd2]);
        const double score = __dmul_rn(p1, mult);
        // :548-567
        bool accept = false;
        const int64_t c10 = in.ex_chain_off[e1], c20 = in.ex_chain_off[e2];
        const int64_t nch1 = in.ex_chain_off[e1 + 1] - c10, nch2 = in.ex_chain_off[e2 + 1] - c20;
        if (!(nch1 >= 2 && nch1 <= 3) || !(nch2 >= 2 && nch2 <= 3)) ok = false;  // :87
        const int2 ch1 = X.M.chain[q1], ch2 = X.M.chain[q2];
        if (P.comp_filt < 1000000 && score > P.max_score && min_sh < P.auto_share && (P.comp_filt_bound == 0 || min_in <= P.comp_filt_bound) &&
            nch1 == 2 && nch2 == 2 && in.ch_left[c10] != in.ch_left[c10 + 1]) {
            const long long comp = min(in.ch_hcomp[ch1.x], in.ch_hcomp[ch2.x]);
            if (comp - cd >= P.comp_filt) accept = true;
        }
        // :710-715
        if (!accept && score > P.max_score && min_sh < P.auto_share) ok = false;
        // :722-758 V gene names
        if (ok) {
            for (int i = 0; i < 2 && ok; i++) {
                const int x1 = i ? ch1.y : ch1.x, x2 = i ? ch2.y : ch2.x;
                const int va = in.ch_v_ref[x1], vb = in.ch_v_ref[x2];
                if (in.ref_name_id[va] != in.ref_name_id[vb]) {
                    const uint8_t *y1 = in.ref_bytes + in.ref_off[va], *y2 = in.ref_bytes + in.ref_off[vb];
                    const long long l1 = in.ref_off[va + 1] - in.ref_off[va], l2 = in.ref_off[vb + 1] - in.ref_off[vb];
                    const long long nn = min(l1, l2);
                    for (long long m = 0; m < nn; m++)
                        if (base2bit(y1[m]) != base2bit(y2[m])) {
                            ok = false;
                            break;
                        }
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
        }
        // :766-854 FWR1 identity minus CDR1+2 identity on the left chain
        if (ok) {
            long long fwr1_len = 0, cdr1_len = 0, cdr2_len = 0, fwr1_diffs = 0, cdr1_diffs = 0, cdr2_diffs = 0;
            bool panic = false;
            for (int m = 0; m < 2; m++) {
                const int x1 = m ? ch1.y : ch1.x, x2 = m ? ch2.y : ch2.x;
                if (!in.ch_left[x1]) continue;
                const uint8_t *s1, *s2;
                long long sl1, sl2;
                if (in.ch_amino_off[x1] != -1) { s1 = in.amino_bytes + in.ch_amino_off[x1]; sl1 = in.ch_amino_len[x1]; }
                else { s1 = in.tig_bytes + in.row_tig_off[2 * k1 + m]; sl1 = m ? sb.Ll : sb.Lh; }
                if (in.ch_amino_off[x2] != -1) { s2 = in.amino_bytes + in.ch_amino_off[x2]; sl2 = in.ch_amino_len[x2]; }
                else { s2 = in.tig_bytes + in.row_tig_off[2 * k2 + m]; sl2 = m ? sb.Ll : sb.Lh; }
                auto region = [&](long long a1, long long b1, long long a2, long long b2, long long& rlen, long long& rdiffs) {
                    const long long len = b1 - a1;
                    if (b2 - a2 == len) {
                        if (len > 0 && (a1 + len > sl1 || a2 + len > sl2 || a1 < 0 || a2 < 0)) { panic = true; return; }
                        long long df = 0;
                        for (long long p = 0; p < len; p++) df += (s1[p + a1] != s2[p + a2]);
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
            const unsigned long long o = atomicAdd(&X.ctr->n_pass, 1ull);
            if (o < pass_cap) pass_w[o] = ((unsigned long long)(uint32_t)k1 << 28) | (unsigned long long)(uint32_t)k2;
            else X.ctr->overflow = 1ull;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Lexicographic spanning forest by Borůvka rounds on edge weight w = (k1 << 28) | k2.
// pot of join_core (join_core.rs:23-50) is Kruskal's forest of the pass graph under this order;
// with distinct weights the minimum spanning forest is unique, and every component's lightest
// outgoing edge belongs to it.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int uf_find_ro(const int32_t* parent, int x) {
    int p = parent[x];
    while (p != x) {
        x = p;
        p = parent[x];
    }
    return x;
}
// lock-free union; the smaller index becomes the root, so root == smallest row of the component
__device__ __forceinline__ void uf_union(int32_t* parent, int a, int b) {
    for (;;) {
        a = uf_find_ro(parent, a);
        b = uf_find_ro(parent, b);
        if (a == b) return;
        if (a > b) {
            const int t = a;
            a = b;
            b = t;
        }
        if (atomicCAS(&parent[b], b, a) == b) return;
    }
}

constexpr int BOR_TAG_SHIFT = 56;

__global__ void k_bor_min(const unsigned long long* __restrict__ w, const unsigned long long* __restrict__ n_p, unsigned long long cap,
                          int32_t* __restrict__ parent, unsigned long long* __restrict__ best, int iter, int2* __restrict__ roots,
                          Counters* ctr) {
    const unsigned long long n = min(*n_p, cap);
    const unsigned long long tag = (unsigned long long)(63 - iter) << BOR_TAG_SHIFT;
    unsigned long long alive = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long e = w[i];
        const int a = (int)(e >> 28), b = (int)(e & 0xfffffffull);
        const int ra = uf_find_ro(parent, a), rb = uf_find_ro(parent, b);
        // pointer jumping: no hook runs concurrently with this kernel, a root is a valid parent
        if (parent[a] != ra) parent[a] = ra;
        if (parent[b] != rb) parent[b] = rb;
        roots[i] = make_int2(ra, rb);
        if (ra != rb) {
            atomicMin(&best[ra], tag | e);
            atomicMin(&best[rb], tag | e);
            alive++;
        }
    }
    // warp-aggregate the alive count
    for (int o = 16; o > 0; o >>= 1) alive += __shfl_xor_sync(0xffffffffu, alive, o);
    if ((threadIdx.x & 31) == 0 && alive) atomicAdd(&ctr->n_alive, alive);
}

__global__ void k_bor_hook(const unsigned long long* __restrict__ w, const unsigned long long* __restrict__ n_p, unsigned long long cap,
                           int32_t* __restrict__ parent, const unsigned long long* __restrict__ best, int iter, const int2* __restrict__ roots,
                           unsigned long long* __restrict__ forest, unsigned long long forest_cap, Counters* ctr) {
    const unsigned long long n = min(*n_p, cap);
    const unsigned long long tag = (unsigned long long)(63 - iter) << BOR_TAG_SHIFT;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        const int2 r = roots[i];
        if (r.x == r.y) continue;
        const unsigned long long e = w[i], te = tag | e;
        if (best[r.x] == te || best[r.y] == te) {
            const unsigned long long o = atomicAdd(&ctr->n_forest, 1ull);
            if (o < forest_cap) forest[o] = e; else ctr->overflow = 1ull;
            uf_union(parent, (int)(e >> 28), (int)(e & 0xfffffffull));
        }
    }
}

// after the hooks of one iteration: re-point every old root straight at its new root
__global__ void k_bor_compress(const unsigned long long* __restrict__ n_p, unsigned long long cap, int32_t* __restrict__ parent,
                               const int2* __restrict__ roots) {
    const unsigned long long n = min(*n_p, cap);
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        const int2 r = roots[i];
        if (r.x == r.y) continue;
        const int nx = uf_find_ro(parent, r.x), ny = uf_find_ro(parent, r.y);
        if (parent[r.x] != nx) parent[r.x] = nx;
        if (parent[r.y] != ny) parent[r.y] = ny;
    }
}

// after a round: class label of every packed row, and per 128-block uniform labels
__global__ void k_flatten(int32_t* __restrict__ parent, const int32_t* __restrict__ rowq, int64_t n_pos, int32_t* __restrict__ labq) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_pos) return;
    const int k = rowq[q];
    if (k < 0) return;  // padding position
    const int r = uf_find_ro(parent, k);
    if (parent[k] != r) parent[k] = r;
    labq[q] = r;
}
__global__ void k_block_labels(const SubBucket* __restrict__ subs, int n_subs, const int32_t* __restrict__ labq, int32_t* __restrict__ blk_lab,
                               const int32_t* __restrict__ blk_sub, int n_blocks) {
    // one warp per 128-row block
    const int lane = threadIdx.x & 31;
    const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b >= n_blocks) return;
    const SubBucket sb = subs[blk_sub[b]];
    const int rb = (int)(b - sb.blk0);
    const int r0 = rb * TILE, r1 = min(sb.n, r0 + TILE);
    const int first = labq[sb.q0 + r0];
    bool same = true;
    for (int r = r0 + lane; r < r1; r += 32)
        if (labq[sb.q0 + r] != first) same = false;
    same = __all_sync(0xffffffffu, same);
    if (lane == 0) blk_lab[b] = same ? first : -1;
    (void)n_subs;
}

__global__ void k_iota(int32_t* p, int64_t n, int32_t off = 0) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (int32_t)i + off;
}
__global__ void k_fill_u64(unsigned long long* p, int64_t n, unsigned long long v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ------------------------------------------------------------------------------------------------
// Finalize
// ------------------------------------------------------------------------------------------------
// forest edges -> (q1, q2) pairs for the tuple pass, and degree counts for the two-cell filter
__global__ void k_forest_pairs(const unsigned long long* __restrict__ forest, int64_t n, const int32_t* __restrict__ q_of_row, uint2* __restrict__ pairs,
                               int32_t* __restrict__ deg) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long e = forest[i];
    const int a = (int)(e >> 28), b = (int)(e & 0xfffffffull);
    pairs[i] = make_uint2((uint32_t)q_of_row[a], (uint32_t)q_of_row[b]);
    atomicAdd(&deg[a], 1);
    atomicAdd(&deg[b], 1);
}
__global__ void k_q_of_row(const int32_t* __restrict__ rowq, int64_t n_pos, int32_t* __restrict__ q_of_row) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q < n_pos && rowq[q] >= 0) q_of_row[rowq[q]] = (int32_t)q;
}
// join.rs:120-178: an orbit of exactly two rows holding two cells in total loses its edge when
// cd > min(shares) / 2 (unless EASY).  keep[i] = 0 for deleted edges; parent[] is un-merged.
__global__ void k_two_cell(DevIn in, DevParams P, const unsigned long long* __restrict__ forest, int64_t n, const int32_t* __restrict__ deg,
                           const EdgeTuple* __restrict__ tuples, int32_t* __restrict__ parent, uint8_t* __restrict__ keep, Counters* ctr) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long e = forest[i];
    const int a = (int)(e >> 28), b = (int)(e & 0xfffffffull);
    int kp = 1;
    if (deg[a] == 1 && deg[b] == 1) {
        const int ea = in.row_exact[a], eb = in.row_exact[b];
        const long long ncells = (in.ex_cell_off[ea + 1] - in.ex_cell_off[ea]) + (in.ex_cell_off[eb + 1] - in.ex_cell_off[eb]);
        if (ncells == 2) {
            const EdgeTuple t = tuples[i];
            const int min_sh = (t.nrefs == 2) ? min(t.sh0, t.sh1) : t.sh0;
            if (t.cd > min_sh / 2 && !P.easy) {
                kp = 0;
                parent[a] = a;
                parent[b] = b;
                atomicAdd(&ctr->n_two_cell, 1ull);
            }
        }
    }
    keep[i] = (uint8_t)kp;
}
// join2.rs:69-84: rows of one exact subclonotype end up in one orbit
__global__ void k_exact_first(DevIn in, int32_t* __restrict__ first_row) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < in.n_rows) atomicMin(&first_row[in.row_exact[k]], (int32_t)k);
}
__global__ void k_exact_union(DevIn in, const int32_t* __restrict__ first_row, int32_t* __restrict__ parent) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= in.n_rows) return;
    const int f = first_row[in.row_exact[k]];
    if (f != (int)k) uf_union(parent, f, (int)k);
}
__global__ void k_labels(const int32_t* __restrict__ parent, int64_t n, int32_t* __restrict__ labels) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) labels[k] = uf_find_ro(parent, (int)k);
}
__global__ void k_fill_i32(int32_t* p, int64_t n, int32_t v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
// compact kept forest edges into sort keys (weight) + source index
__global__ void k_final_keys(const unsigned long long* __restrict__ forest, const uint8_t* __restrict__ keep, int64_t n, unsigned long long* __restrict__ keys,
                             uint32_t* __restrict__ idx, unsigned long long* __restrict__ n_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !keep[i]) return;
    const unsigned long long o = atomicAdd(n_out, 1ull);
    keys[o] = forest[i];
    idx[o] = (uint32_t)i;
}
struct FinalOut {
    int32_t *k1, *k2, *cd, *nrefs, *shares, *indeps, *k, *d, *n;
    double *p1, *mult, *score;
    uint8_t* err;
};
__global__ void k_final_gather(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ idx, int64_t n, const EdgeTuple* __restrict__ tuples,
                               FinalOut o, Counters* ctr) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long e = keys[i];
    const EdgeTuple t = tuples[idx[i]];
    o.k1[i] = (int)(e >> 28);
    o.k2[i] = (int)(e & 0xfffffffull);
    o.cd[i] = t.cd;
    o.nrefs[i] = t.nrefs;
    o.shares[2 * i] = t.sh0;
    o.shares[2 * i + 1] = t.sh1;
    o.indeps[2 * i] = t.in0;
    o.indeps[2 * i + 1] = t.in1;
    o.k[i] = t.k;
    o.d[i] = t.d;
    o.n[i] = t.n;
    o.p1[i] = t.p1;
    o.mult[i] = t.mult;
    o.score[i] = t.score;
    o.err[i] = (uint8_t)t.err;
    (void)ctr;
}

// ------------------------------------------------------------------------------------------------
// Integer-pipe peak microbenchmarks (roofline denominators, SURVEY §8 d)
// ------------------------------------------------------------------------------------------------
__global__ void k_peak_popc(uint32_t* out, int iters) {
    uint32_t a = threadIdx.x * 2654435761u + 1u, b = blockIdx.x * 40503u + 7u, c = a ^ 0x9e3779b9u, d = b + 0x7f4a7c15u;
    uint32_t s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0, s5 = 0, s6 = 0, s7 = 0;
    for (int i = 0; i < iters; i++) {
        // 8 independent popc per iteration; operands evolve with cheap adds on another pipe
        s0 += __popc(a); s1 += __popc(b); s2 += __popc(c); s3 += __popc(d);
        s4 += __popc(a ^ 0x55555555u); s5 += __popc(b ^ 0x33333333u); s6 += __popc(c ^ 0x0f0f0f0fu); s7 += __popc(d ^ 0x00ff00ffu);
        a += 0x01010101u; b += 0x10101010u; c += 0x11111111u; d += 0x00010001u;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s0 + s1 + s2 + s3 + s4 + s5 + s6 + s7;
}
__global__ void k_peak_lop3(uint32_t* out, int iters) {
    uint32_t a = threadIdx.x * 2654435761u + 1u, b = blockIdx.x * 40503u + 7u, c = a ^ 0x9e3779b9u, d = b + 0x7f4a7c15u;
    uint32_t e = a + 3, f = b + 5, g = c + 7, h = d + 11;
    for (int i = 0; i < iters; i++) {
        // 8 independent 3-input logic ops per iteration
        a = (a ^ b) | (c & i); b = (b & c) ^ (d | i); c = (c | d) ^ (a & i); d = (d ^ a) & (b | i);
        e = (e ^ f) | (g & i); f = (f & g) ^ (h | i); g = (g | h) ^ (e & i); h = (h ^ e) & (f | i);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
}

}  // namespace ejb
