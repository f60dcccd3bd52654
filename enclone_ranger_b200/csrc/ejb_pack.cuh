// ejb_pack.cuh — bucket scan (join.rs:68-88) and the bucketer that packs rows into HBM.
#pragma once
#include "ejb_types.cuh"

namespace ejb {

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
        const int64_t o2 = in.row_tig2_off ? in.row_tig2_off[i] : -1;
        if (o2 >= 0) {   // 2-bit packed contig: pure ACGT by contract, never a has_del row
            if (L < 0 || o2 + (L + 15) / 16 > in.n_tig2_words || in.row_has_del[i]) e |= 8192;
        } else if (!in.row_tig_off || L < 0 || in.row_tig_off[i] < 0 || in.row_tig_off[i] + L > in.n_tig_bytes) e |= 4;
        const int c = in.row_cdr3_len[i];
        const int64_t c2 = in.row_cdr32_off ? in.row_cdr32_off[i] : -1;
        if (c2 >= 0) {   // 2-bit packed CDR3: pure ACGT by contract
            if (c < 0 || c > 65535 || c2 + (c + 15) / 16 > in.n_cdr32_words) e |= 8;
        } else if (!in.row_cdr3_off || c < 0 || c > 65535 || in.row_cdr3_off[i] < 0 || in.row_cdr3_off[i] + c > in.n_cdr3_bytes) e |= 8;
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

// validates exact / chain / cell / seg / ref tables
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

// first row of every lens bucket (runs of equal lens are contiguous in row order): what the host needs to replay the
// EquivRel of independent buckets in parallel
__global__ void k_bucket_starts(const int32_t* __restrict__ head, const int32_t* __restrict__ bucket_incl, int64_t n, int32_t* __restrict__ starts) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n && head[k]) starts[bucket_incl[k] - 1] = (int32_t)k;
}

// sort key of a row: (bucket id, heavy CDR3 len, light CDR3 len); rows that can never join
// (lens.len() != 2, join_one.rs:83-96) get the maximal key and fall off the end.
// Run header on the device: what the host reads back ONCE after the bucket scan and the sub-bucket table are built.
struct RunHdr {
    unsigned int err, big_exacts;   // validation mask; an exact subclonotype with more than BC_SMALL cells exists
    unsigned long long n_active;   // rows with two chains (the only ones that can join)
    unsigned long long n_subs;     // sub-buckets
    unsigned long long n_buckets;  // runs of equal lens
};
__global__ void k_keys(DevIn in, const int32_t* bucket_incl, uint64_t* keys, uint32_t* vals, RunHdr* hdr) {
    unsigned long long* n_active = &hdr->n_active;
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k == in.n_rows - 1) hdr->n_buckets = (unsigned long long)bucket_incl[k];
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

// launched over all n rows: the number of active rows is only known on the device
__global__ void k_subheads(const uint64_t* keys, const RunHdr* hdr, int64_t n, int32_t* head) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    head[q] = (q < (int64_t)hdr->n_active && (q == 0 || keys[q] != keys[q - 1])) ? 1 : 0;
}

struct SubInfo {
    int64_t j0;
    uint64_t key;
    int32_t Lh, Ll;
    int32_t gap, pad;   // gap: OR over the sub-bucket's rows of has_del (bit 0 heavy, bit 1 light)
};
__global__ void k_subinfo(DevIn in, const uint64_t* keys, const uint32_t* row_of, const int32_t* head, const int32_t* sub_incl,
                          RunHdr* hdr, SubInfo* out) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n_active = (int64_t)hdr->n_active;
    if (j >= n_active) return;
    if (j == n_active - 1) hdr->n_subs = (unsigned long long)sub_incl[j];
    const int s = sub_incl[j] - 1;
    const int64_t k = row_of[j];
    if (head[j]) {
        out[s].j0 = j;
        out[s].key = keys[j];
        out[s].Lh = in.row_len[2 * k];
        out[s].Ll = in.row_len[2 * k + 1];
    }
    const int g = (in.row_has_del[2 * k] ? 1 : 0) | (in.row_has_del[2 * k + 1] ? 2 : 0);
    if (g) atomicOr(&out[s].gap, g);
}

// Row roles (ejb_set_row_roles): per sub-bucket the number of rows that act as first rows, the split flag, and a check that
// column-only rows form a suffix of the sub-bucket (bit 16384 of the validation mask otherwise).
struct SubRole {
    int32_t n_act, raw;
};
__global__ void k_subroles(const uint8_t* __restrict__ roles, const uint32_t* __restrict__ row_of, const int32_t* __restrict__ head,
                           const int32_t* __restrict__ sub_incl, const RunHdr* __restrict__ hdr, SubRole* __restrict__ out, unsigned int* __restrict__ err) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= (int64_t)hdr->n_active) return;
    const int s = sub_incl[j] - 1;
    const uint8_t r = roles[row_of[j]];
    if (!(r & ROLE_COLUMN_ONLY)) atomicAdd(&out[s].n_act, 1);
    if (r & ROLE_SPLIT) atomicOr(&out[s].raw, 1);
    if (j > 0 && !head[j] && !(r & ROLE_COLUMN_ONLY) && (roles[row_of[j - 1]] & ROLE_COLUMN_ONLY)) atomicOr(err, 16384u);
}

// Dominant reference-set tag and donor set of every sub-bucket: the most frequent value among five evenly spaced rows (a
// wrong guess only makes the flag shortcut of stage 1 fire less often; the exact 64-bit tests behind it do not depend on it).
struct SubDom {
    unsigned long long tag, don;
};
__global__ void k_sub_dominant(const SubBucket* __restrict__ subs, int n_subs, const unsigned long long* __restrict__ tagq,
                               const unsigned long long* __restrict__ donq, SubDom* __restrict__ out) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_subs) return;
    const SubBucket sb = subs[s];
    if (sb.n_act <= sb.r_beg) {   // not packed on this context (another shard owns it)
        out[s].tag = DEAD_NONE;
        out[s].don = 0ull;
        return;
    }
    unsigned long long t[5], d[5];
    for (int i = 0; i < 5; i++) {
        const int64_t q = sb.q0 + (int64_t)(((int64_t)sb.n * (2 * i + 1)) / 10);
        t[i] = tagq[q];
        d[i] = donq[q];
    }
    int bt = 0, bd = 0, ct = 0, cd = 0;
    for (int i = 0; i < 5; i++) {
        int a = 0, b = 0;
        for (int j = 0; j < 5; j++) {
            a += t[j] == t[i];
            b += d[j] == d[i];
        }
        if (a > ct && t[i] != TAG_SLOW) { ct = a; bt = i; }
        if (b > cd && d[i] != 0) { cd = b; bd = i; }
    }
    out[s].tag = ct ? t[bt] : DEAD_NONE;   // DEAD_NONE never equals a tag
    out[s].don = cd ? d[bd] : 0ull;
}
// flagq: see RF_* in ejb_stage1.cuh; refreshed before every stage-1 launch (deadq grows while stage 2a learns)
__global__ void k_row_flags(const int32_t* __restrict__ subq, const int32_t* __restrict__ rowq, int64_t n_pos, const SubDom* __restrict__ dom,
                            const unsigned long long* __restrict__ tagq, const unsigned long long* __restrict__ deadq,
                            const unsigned long long* __restrict__ donq, uint32_t* __restrict__ flagq) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_pos) return;
    uint32_t f = 0;
    if (rowq[q] >= 0) {
        const SubDom d = dom[subq[q]];
        if (deadq[q] == d.tag) f |= 1u;
        if (tagq[q] == d.tag) f |= 2u;
        const unsigned long long m = donq[q];
        if (d.don != 0 && m != 0) f |= (m == d.don) ? 8u : 4u;
    }
    flagq[q] = f;
}

// per exact: donor set as a 64-bit mask (join_one.rs:301-315) and the exact's barcode list in ascending order (the sorted lists
// the barcode meet of join_one.rs:451-470 walks).  Nearly every exact has a handful of cells: the thread sorts up to
// BC_SMALL of them itself; a larger exact raises hdr->big_exacts and the host falls back to one global radix sort of the
// (exact << 32 | barcode) keys, which this kernel writes as well.
constexpr int BC_SMALL = 32;
__global__ void k_exact_prep(DevIn in, unsigned long long* donor_mask, uint64_t* bc_keys, uint64_t* bc_sorted, RunHdr* hdr) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= in.n_exacts) return;
    unsigned long long m = 0;
    const int64_t j0 = in.ex_cell_off[e], j1 = in.ex_cell_off[e + 1];
    const int64_t n = j1 - j0;
    uint32_t loc[BC_SMALL];
    for (int64_t j = j0; j < j1; j++) {
        const int d = in.cell_donor[j];
        if (d >= 0) m |= 1ull << d;
        const uint32_t bc = in.cell_barcode[j];
        bc_keys[j] = ((uint64_t)e << 32) | (uint64_t)bc;
        if (n <= BC_SMALL) {   // insertion into the sorted prefix
            int i = (int)(j - j0);
            while (i > 0 && loc[i - 1] > bc) {
                loc[i] = loc[i - 1];
                i--;
            }
            loc[i] = bc;
        }
    }
    donor_mask[e] = m;
    if (n <= BC_SMALL) {
        for (int i = 0; i < (int)n; i++) bc_sorted[j0 + i] = ((uint64_t)e << 32) | (uint64_t)loc[i];
    } else {
        hdr->big_exacts = 1u;
    }
}

// Per-segment reference planes + lengths at the canonical seg id (segments with equal content write equal values).
__global__ void k_seg_planes(DevIn in, uint32_t* __restrict__ segpl, int32_t* __restrict__ seglen) {
    const int lane = threadIdx.x & 31;
    const int64_t sg = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (sg >= in.n_segs) return;
    const uint8_t* b = in.seg_bytes + in.seg_off[sg];
    const long long len = in.seg_off[sg + 1] - in.seg_off[sg];
    const int64_t cs = in.seg_canon[sg];
    if (lane == 0) seglen[cs] = (int32_t)min(len, (long long)0x7fffffff);
    for (int w = 0; w < SEGPL_WORDS; w++) {
        const long long p = (long long)w * 32 + lane;
        const uint32_t code = p < len ? base_code(b[p]) : 0u;
        const uint32_t lo = __ballot_sync(0xffffffffu, code & 1), hi = __ballot_sync(0xffffffffu, code & 2);
        if (lane == 0) {
            segpl[(cs * 2 + 0) * SEGPL_WORDS + w] = lo;
            segpl[(cs * 2 + 1) * SEGPL_WORDS + w] = hi;
        }
    }
}
// even bits of x compressed into the low 16 bits
__device__ __forceinline__ uint32_t even_bits(uint32_t x) {
    x &= 0x55555555u;
    x = (x | (x >> 1)) & 0x33333333u;
    x = (x | (x >> 2)) & 0x0f0f0f0fu;
    x = (x | (x >> 4)) & 0x00ff00ffu;
    x = (x | (x >> 8)) & 0x0000ffffu;
    return x;
}
// ------------------------------------------------------------------------------------------------
// The bucketer: one warp per packed row.  Writes
//   * CDR3 planes (heavy ++ light concatenated), layout [plane*W1 + w][npad] per sub-bucket
//   * the row record (layout in ejb_types.cuh): contig planes, reference-difference bitmap d, optional gap plane
//   * the aux record (AUX_INTS ints) and rowq / subq, reference-set tag and donor set
// ------------------------------------------------------------------------------------------------
#ifndef PACK_MINB
#define PACK_MINB 8   // latency-bound: occupancy wins over spills
#endif
// pack_off / pack_sub: the sub-buckets this context packs (all of them, or the ones its shard owns) and the prefix sums of
// their row counts; warp w handles row w - pack_off[i] of sub-bucket pack_sub[i].
// The generic path (ASCII contigs or CDR3s, '-', contigs over 512 nt, segments over 512 nt): one warp per row of the list that
// k_pack_rows8 left (slow_rows == nullptr: every row, in pack order).
__global__ void __launch_bounds__(256, PACK_MINB) k_pack_rows(DevIn in, DevParams P, const SubBucket* __restrict__ subs, const int64_t* __restrict__ pack_off,
                            const int32_t* __restrict__ pack_sub, int n_pack, const uint32_t* __restrict__ row_of, const unsigned long long* __restrict__ donor_mask,
                            const uint32_t* __restrict__ segpl, const int32_t* __restrict__ seglen, uint32_t* __restrict__ cdr3w, uint4* __restrict__ rowstore,
                            int32_t* __restrict__ aux, int32_t* __restrict__ rowq, int32_t* __restrict__ subq,
                            unsigned long long* __restrict__ tagq, unsigned long long* __restrict__ donq, const uint32_t* __restrict__ slow_rows,
                            const unsigned long long* __restrict__ n_slow_rows) {
    const int lane = threadIdx.x & 31;
    const int64_t n_list = slow_rows ? (int64_t)*n_slow_rows : pack_off[n_pack];
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    __shared__ __align__(16) int32_t s_aux[8][AUX_INTS];
    for (int64_t wl = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; wl < n_list; wl += nwarps) {
    const int64_t w = slow_rows ? (int64_t)slow_rows[wl] : wl;
    int plo = 0, phi = n_pack - 1;
    while (plo < phi) {   // last i with pack_off[i] <= w
        const int mid = (plo + phi + 1) >> 1;
        if (pack_off[mid] <= w) plo = mid; else phi = mid - 1;
    }
    const int s = pack_sub[plo];
    const SubBucket sb = subs[s];
    const int64_t li = w - pack_off[plo];
    const int64_t k = row_of[sb.j0 + li];
    const int64_t q = sb.q0 + li;
    uint32_t flags = 0;
    int tot = 0;

    // ---- CDR3 planes (ASCII bytes, or 2-bit packed words: the role tigsp plays for the contigs) ----
    {
        const int T = sb.ch + sb.cl;
        const int nw = (T + 31) / 32;
        const int64_t o0 = in.row_cdr32_off ? in.row_cdr32_off[2 * k] : -1, o1 = in.row_cdr32_off ? in.row_cdr32_off[2 * k + 1] : -1;
        const uint8_t* c0 = in.cdr3_bytes + (o0 >= 0 ? 0 : in.row_cdr3_off[2 * k]);
        const uint8_t* c1 = in.cdr3_bytes + (o1 >= 0 ? 0 : in.row_cdr3_off[2 * k + 1]);
        const uint32_t* w0 = in.cdr32_words + (o0 >= 0 ? o0 : 0);
        const uint32_t* w1 = in.cdr32_words + (o1 >= 0 ? o1 : 0);
        bool irr = false;
        for (int w = 0; w < nw; w++) {
            const int p = w * 32 + lane;
            int code = 0;
            if (p < T) {
                int op = p + sb.rot;                 // stored position p holds original position (p + rot) mod T
                if (op >= T) op -= T;
                const bool heavy = op < sb.ch;
                const int pp = heavy ? op : op - sb.ch;
                if ((heavy ? o0 : o1) >= 0) {
                    const uint32_t c2 = ((heavy ? w0 : w1)[pp >> 4] >> (2 * (pp & 15))) & 3u;   // A,C,G,T = 0..3
                    code = (int)(c2 ^ (c2 >> 1));                                               // -> internal A,C,T,G = 0,1,2,3
                } else {
                    const uint8_t b = (heavy ? c0 : c1)[pp];
                    if (!is_acgt_fast(b)) irr = true;
                    code = (int)base_code(b);
                }
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

    // ---- contig planes and the difference bitmap against the row's own V / J reference ----
    uint32_t* rec32 = reinterpret_cast<uint32_t*>(rowstore + sb.rs_off + li * (int64_t)sb.rec_u4);
    int base_u4 = 0;
    // FWR1 / CDR1 / CDR2 ranges of chain 0 (only meaningful when that chain is the left chain; stage 2 checks)
    int rF0 = 0, rF1 = 0, rC10 = 0, rC11 = 0, rC20 = 0, rC21 = 0;
    {
        const int ex0 = in.row_exact[k];
        const int chh = (int)(in.ex_chain_off[ex0] + in.row_exact_col[2 * k]);
        const int fr1 = in.ch_fr1[chh], cdr1 = in.ch_cdr1[chh], fr2 = in.ch_fr2[chh], cdr2 = in.ch_cdr2[chh], fr3 = in.ch_fr3[chh];
        const int L = sb.Lh;
        const bool has_f = cdr1 != -1;
        const bool has_c1 = cdr1 != -1 && fr2 != -1 && cdr1 <= fr2;
        const bool has_c2 = cdr2 != -1 && fr3 != -1 && cdr2 <= fr3;
        const bool okf = !has_f || (fr1 >= 0 && fr1 <= cdr1 && cdr1 <= L);
        const bool okc1 = !has_c1 || (cdr1 >= 0 && fr2 <= L);
        const bool okc2 = !has_c2 || (cdr2 >= 0 && fr3 <= L);
        // the three ranges must also be pairwise disjoint to be counted from one pass over the difference words;
        // the region masks hold 16-bit bounds (contigs on this path are <= 1024 nt)
        bool disj = true;
        if (has_f && has_c2 && !(cdr1 <= cdr2 || fr3 <= fr1)) disj = false;
        if (has_c1 && has_c2 && !(fr2 <= cdr2 || fr3 <= cdr1)) disj = false;
        if (okf && okc1 && okc2 && disj && L <= 128 * MAXW4) {
            flags |= B_REG_OK;
            if (has_f) { rF0 = fr1; rF1 = cdr1; }
            if (has_c1) { rC10 = cdr1; rC11 = fr2; }
            if (has_c2) { rC20 = cdr2; rC21 = fr3; }
        }
    }
    // ---- fast path: both contigs 2-bit packed, <= 512 nt, segments <= 512 nt.  Lanes 0-15 build the 16 words of every
    //      plane of chain 0, lanes 16-31 those of chain 1, by de-interleaving the packed words: no per-base work. ----
    bool fastrow = sb.W4h == 4 && sb.W4l == 4 && in.row_tig2_off != nullptr;
    bool longseg = false;
    for (int m = 0; m < 2; m++) {
        const int sv = in.seg_canon[in.row_vs[2 * k + m]], sj = in.seg_canon[in.row_js[2 * k + m]];
        if (seglen[sv] > 32 * SEGPL_WORDS || seglen[sj] > 32 * SEGPL_WORDS) longseg = true;
        if (fastrow && in.row_tig2_off[2 * k + m] < 0) fastrow = false;
    }
    if (longseg) {   // stage 2a cannot slice this row's references out of the segment planes: byte-exact path
        fastrow = false;
        flags |= F_SLOWGEOM;
    }
    if (fastrow) {
        const int m = lane >> 4, w = lane & 15;
        const int L = m ? sb.Ll : sb.Lh;
        const int gapbit = (sb.gap >> m) & 1;
        const RefGeom G = ref_geom(seglen, in.seg_canon[in.row_vs[2 * k + m]], in.seg_canon[in.row_js[2 * k + m]], L, P.ref_v_trim, P.ref_j_trim);
        uint32_t fl = 0;
        if (G.vr < 0 || G.jr < 0 || G.jr > L) fl |= F_PANIC;
        if (G.vr > L || G.vr + G.jr > L) fl |= F_SLOWGEOM;
        // contig planes from the packed words (A,C,G,T = 0..3 -> internal code lo' = lo ^ hi, hi' = hi)
        const uint32_t* tw = in.tig2_words + in.row_tig2_off[2 * k + m];
        const int nw16 = (L + 15) >> 4;
        const uint32_t x0 = (2 * w < nw16) ? tw[2 * w] : 0u, x1 = (2 * w + 1 < nw16) ? tw[2 * w + 1] : 0u;
        const uint32_t inl = mask_lt32(L - 32 * w);
        uint32_t lo = even_bits(x0) | (even_bits(x1) << 16), hi = even_bits(x0 >> 1) | (even_bits(x1 >> 1) << 16);
        lo = (lo ^ hi) & inl;
        hi &= inl;
        uint32_t rlo, rhi, cm;
        ref_word(segpl, G, w, rlo, rhi, cm);
        const uint32_t dm = (fl & (F_PANIC | F_SLOWGEOM)) ? 0u : (cm & ((lo ^ rlo) | (hi ^ rhi)));
        uint32_t* r = rec32 + (m ? rec_planes(0, sb.gap & 1) * sb.W4h * 4 : 0) + w;
        const int st = 16;   // 4 * W4
        r[PL_TLO * st] = lo;
        r[PL_THI * st] = hi;
        r[PL_D * st] = dm;
        if (gapbit) r[PL_GAP * st] = 0u;
        int t = __popc(dm);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            t += __shfl_xor_sync(0xffffffffu, t, o);
            fl |= __shfl_xor_sync(0xffffffffu, fl, o);
        }
        tot = t;
        flags |= fl;
    }
    if (!fastrow)
    for (int m = 0; m < 2; m++) {
        const int L = (m == 0) ? sb.Lh : sb.Ll;
        const int W4 = (m == 0) ? sb.W4h : sb.W4l;
        const int gapbit = (sb.gap >> m) & 1;
        const bool hasdel = in.row_has_del[2 * k + m] != 0;
        const int64_t off2 = in.row_tig2_off ? in.row_tig2_off[2 * k + m] : -1;
        const bool packed = off2 >= 0;
        const uint8_t* tig = in.tig_bytes + (packed ? 0 : in.row_tig_off[2 * k + m]);
        const uint32_t* tig2 = in.tig2_words + (packed ? off2 : 0);
        const int sv = in.row_vs[2 * k + m], sj = in.row_js[2 * k + m];
        const uint8_t* vseg = in.seg_bytes + in.seg_off[sv];
        const uint8_t* jseg = in.seg_bytes + in.seg_off[sj];
        const long long vlen = in.seg_off[sv + 1] - in.seg_off[sv];
        const long long jlen = in.seg_off[sj + 1] - in.seg_off[sj];
        const long long vr = vlen - P.ref_v_trim, jr = jlen - P.ref_j_trim;  // compared ranges
        if (vr < 0 || jr < 0 || jr > L) flags |= F_PANIC;   // usize underflow / index panic in the reference
        if (vr > L) flags |= F_SLOWGEOM;                    // "ugly bailout" join_one.rs:371-373
        if (vr + jr > L) flags |= F_SLOWGEOM;               // V and J ranges overlap: positions counted twice
        if (hasdel) flags |= (m == 0) ? F_HASDEL_H : F_HASDEL_L;
        bool odd = false;
        const int nw = (L + 31) / 32;
        // lane w keeps word w of every plane (4 * W4 <= 32 words) and writes them at the end
        uint32_t k_lo = 0, k_hi = 0, k_d = 0, k_gap = 0;
        const int nwords = min(nw, 4 * W4);
        const long long jshift = (long long)jlen - L;   // J index of contig position p: p + jshift  (tig[L-1-pp] vs J[|J|-1-pp], join_one.rs:388-391)
#pragma unroll 4
        for (int w = 0; w < nwords; w++) {
            const int p = w * 32 + lane;
            uint32_t b = 'A', rbyte = 0, code = 0;
            bool inr = false;
            if (p < L) {
                if (packed) {
                    const uint32_t c2 = (tig2[p >> 4] >> (2 * (p & 15))) & 3u;   // A,C,G,T = 0..3
                    code = c2 ^ (c2 >> 1);                                      // -> internal A,C,T,G = 0,1,2,3
                } else {
                    b = tig[p];
                    code = base_code(b);
                }
                if (p < vr) { rbyte = vseg[p]; inr = true; }
                else if (p >= L - jr && p + jshift >= 0) { rbyte = jseg[p + jshift]; inr = true; }
            }
            const bool gapb = (b == '-') && hasdel;
            if (p < L && !gapb && !is_acgt_fast(b)) odd = true;
            const uint32_t rcode = base_code(rbyte);
            // seg bytes are canonical upper-case ACGT (host); '-' and odd bytes never equal them
            const bool dbit = inr && (packed ? (code != rcode) : (b != rbyte));
            const uint32_t lo = __ballot_sync(0xffffffffu, code & 1), hi = __ballot_sync(0xffffffffu, code & 2);
            const uint32_t dm = __ballot_sync(0xffffffffu, dbit), gp = __ballot_sync(0xffffffffu, gapb);
            tot += __popc(dm);
            if (lane == w) {
                k_lo = lo; k_hi = hi; k_d = dm; k_gap = gp;
            }
        }
        if (lane < 4 * W4) {
            // plane-major record: one coalesced store per plane
            uint32_t* r = rec32 + base_u4 * 4 + lane;
            const int st = 4 * W4;
            r[PL_TLO * st] = k_lo;
            r[PL_THI * st] = k_hi;
            r[PL_D * st] = k_d;
            if (gapbit) r[PL_GAP * st] = k_gap;
        }
        if (W4 == 0 && !packed) {  // contig longer than the fast path: the byte-exact path handles it
            for (int p = lane; p < L; p += 32)
                if (!is_acgt_fast(tig[p]) && !(tig[p] == '-' && hasdel)) odd = true;
        }
        if (__any_sync(0xffffffffu, odd)) flags |= (m == 0) ? F_ODDBYTE_H : F_ODDBYTE_L;
        base_u4 += rec_planes(m, gapbit) * W4;
    }
    int32_t* const a = s_aux[(threadIdx.x >> 5) & 7];
    __syncwarp();
    if (lane == 0) {
        rowq[q] = (int32_t)k;
        subq[q] = s;
        const int ex = in.row_exact[k];
        const int64_t c0 = in.ex_chain_off[ex], nch = in.ex_chain_off[ex + 1] - c0;
        const int ch = (int)(c0 + in.row_exact_col[2 * k]), cl = (int)(c0 + in.row_exact_col[2 * k + 1]);
        if (nch >= 2 && nch <= 3) flags |= B_NCH_OK;
        if (nch == 2) flags |= B_NCH2;
        if (nch >= 2 && in.ch_left[c0] != in.ch_left[c0 + 1]) flags |= B_JUN_LEFT;
        if (in.ch_left[ch]) flags |= B_LEFT_H;
        if (in.ch_left[cl]) flags |= B_LEFT_L;
        if (in.ch_amino_off[ch] == -1) flags |= B_AMINO_TIG_H;
        const int64_t x0 = in.ex_cell_off[ex], ncells = in.ex_cell_off[ex + 1] - x0;
        const unsigned long long dmask = donor_mask[ex];
        a[AX_FLAGS] = (int32_t)flags;
        a[AX_VSH] = in.seg_canon[in.row_vs[2 * k]];
        a[AX_VSL] = in.seg_canon[in.row_vs[2 * k + 1]];
        a[AX_JSH] = in.seg_canon[in.row_js[2 * k]];
        a[AX_JSL] = in.seg_canon[in.row_js[2 * k + 1]];
        a[AX_TOT] = tot;
        a[AX_EXACT] = ex;
        a[AX_NCELLS] = (int32_t)ncells;
        a[AX_DONLO] = (int32_t)(uint32_t)(dmask & 0xffffffffull);
        a[AX_DONHI] = (int32_t)(uint32_t)(dmask >> 32);
        a[AX_CREFH] = in.ch_c_ref[ch];
        a[AX_CREFL] = in.ch_c_ref[cl];
        a[AX_HCOMP] = in.ch_hcomp[ch];
        a[AX_VREFH] = in.ch_v_ref[ch];
        a[AX_VREFL] = in.ch_v_ref[cl];
        a[AX_BC0] = ncells > 0 ? (int32_t)in.cell_barcode[x0] : -1;
        a[AX_VNAMEH] = in.ref_name_id[in.ch_v_ref[ch]];
        a[AX_VNAMEL] = in.ref_name_id[in.ch_v_ref[cl]];
        a[AX_FR1] = in.ch_fr1[ch];
        a[AX_CDR1] = in.ch_cdr1[ch];
        a[AX_FR2] = in.ch_fr2[ch];
        a[AX_CDR2] = in.ch_cdr2[ch];
        a[AX_FR3] = in.ch_fr3[ch];
        a[AX_CHH] = ch;
        a[AX_CHL] = cl;
        a[AX_RGN0] = rF0 | (rF1 << 16);
        a[AX_RGN1] = rC10 | (rC11 << 16);
        a[AX_RGN2] = rC20 | (rC21 << 16);
        if (tagq) {
            tagq[q] = (flags & F_SLOWMASK) ? TAG_SLOW : make_ref_tag(a[AX_VSH], a[AX_VSL], a[AX_JSH], a[AX_JSL]);
            donq[q] = P.mix_donors ? 0ull : dmask;   // join_one.rs:301-323 only rejects when donors are not mixed
        }
    }
    __syncwarp();
    if (lane < AUX_INTS / 4) reinterpret_cast<uint4*>(aux + q * AUX_INTS)[lane] = reinterpret_cast<const uint4*>(a)[lane];
    }
}

// ------------------------------------------------------------------------------------------------
// The bucketer's fast path: EIGHT lanes per row, four rows per warp.  Rows whose contigs and CDR3s arrive 2-bit packed
// (<= 512 nt, segments <= 512 nt) need no per-base work: lane g of a row's octet builds words g and g + 8 of every plane of
// both chains by bit de-interleaving, the CDR3 planes four positions at a time (octet-wide redux.or), and lane 0 of the octet
// gathers the aux record.  The per-row scalar work (sub-bucket lookup, geometry, feature ranges) is shared by 8 lanes instead
// of 32: 4x fewer warp instructions per row than the warp-per-row kernel.  Rows that do not qualify go to `slow_rows`.
// ------------------------------------------------------------------------------------------------
#ifndef PACK8_MINB
#define PACK8_MINB 4
#endif
__global__ void __launch_bounds__(256, PACK8_MINB) k_pack_rows8(DevIn in, DevParams P, const SubBucket* __restrict__ subs, const int64_t* __restrict__ pack_off,
                            const int32_t* __restrict__ pack_sub, int n_pack, const uint32_t* __restrict__ row_of, const unsigned long long* __restrict__ donor_mask,
                            const uint32_t* __restrict__ segpl, const int32_t* __restrict__ seglen, uint32_t* __restrict__ cdr3w, uint4* __restrict__ rowstore,
                            int32_t* __restrict__ aux, int32_t* __restrict__ rowq, int32_t* __restrict__ subq,
                            unsigned long long* __restrict__ tagq, unsigned long long* __restrict__ donq, uint32_t* __restrict__ slow_rows,
                            unsigned long long* __restrict__ n_slow_rows) {
    const int lane = threadIdx.x & 31, g = lane & 7, oct = lane >> 3;
    const unsigned int omask = 0xffu << (oct * 8);
    const int64_t w = ((((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) << 2) + oct;
    if (w >= pack_off[n_pack]) return;   // whole octets leave together: the collectives below name octets only
    int plo = 0, phi = n_pack - 1;
    while (plo < phi) {   // last i with pack_off[i] <= w
        const int mid = (plo + phi + 1) >> 1;
        if (pack_off[mid] <= w) plo = mid; else phi = mid - 1;
    }
    const int s = pack_sub[plo];
    const SubBucket sb = subs[s];
    const int64_t li = w - pack_off[plo];
    const int64_t k = row_of[sb.j0 + li];
    const int64_t q = sb.q0 + li;
    const int64_t o0 = in.row_cdr32_off ? in.row_cdr32_off[2 * k] : -1, o1 = in.row_cdr32_off ? in.row_cdr32_off[2 * k + 1] : -1;
    int sv[2], sj[2];
    bool fast = sb.W4h == 4 && sb.W4l == 4 && in.row_tig2_off != nullptr && o0 >= 0 && o1 >= 0;
#pragma unroll
    for (int m = 0; m < 2; m++) {
        sv[m] = in.seg_canon[in.row_vs[2 * k + m]];
        sj[m] = in.seg_canon[in.row_js[2 * k + m]];
        if (seglen[sv[m]] > 32 * SEGPL_WORDS || seglen[sj[m]] > 32 * SEGPL_WORDS) fast = false;
        if (fast && in.row_tig2_off[2 * k + m] < 0) fast = false;
    }
    if (!fast) {
        if (g == 0) slow_rows[atomicAdd(n_slow_rows, 1ull)] = (uint32_t)w;
        return;
    }
    uint32_t flags = 0;

    // ---- CDR3 planes from the 2-bit words: lane g owns positions 4g .. 4g+3 of every 32-position word ----
    {
        const int T = sb.ch + sb.cl;
        const int nw = (T + 31) / 32;
        const uint32_t* w0 = in.cdr32_words + o0;
        const uint32_t* w1 = in.cdr32_words + o1;
        for (int wd = 0; wd < nw; wd++) {
            uint32_t lo4 = 0, hi4 = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int p = wd * 32 + g * 4 + j;
                if (p < T) {
                    int op = p + sb.rot;   // stored position p holds original position (p + rot) mod T
                    if (op >= T) op -= T;
                    const bool heavy = op < sb.ch;
                    const int pp = heavy ? op : op - sb.ch;
                    const uint32_t c2 = ((heavy ? w0 : w1)[pp >> 4] >> (2 * (pp & 15))) & 3u;   // A,C,G,T = 0..3
                    const uint32_t code = c2 ^ (c2 >> 1);                                        // -> internal A,C,T,G = 0,1,2,3
                    lo4 |= (code & 1u) << j;
                    hi4 |= (code >> 1) << j;
                }
            }
            const uint32_t lo = __reduce_or_sync(omask, lo4 << (4 * g)), hi = __reduce_or_sync(omask, hi4 << (4 * g));
            if (sb.W1 > 0 && g == 0) {
                cdr3w[sb.s1_off + (int64_t)wd * sb.npad + li] = lo;
                cdr3w[sb.s1_off + (int64_t)(sb.W1 + wd) * sb.npad + li] = hi;
            }
        }
    }

    // ---- contig planes and the difference bitmap against the row's own V / J reference: words g and g + 8 of both chains ----
    uint32_t* rec32 = reinterpret_cast<uint32_t*>(rowstore + sb.rs_off + li * (int64_t)sb.rec_u4);
    int t = 0;
#pragma unroll
    for (int m = 0; m < 2; m++) {
        const int L = m ? sb.Ll : sb.Lh;
        const int gapbit = (sb.gap >> m) & 1;
        const RefGeom G = ref_geom(seglen, sv[m], sj[m], L, P.ref_v_trim, P.ref_j_trim);
        uint32_t fl = 0;
        if (G.vr < 0 || G.jr < 0 || G.jr > L) fl |= F_PANIC;
        if (G.vr > L || G.vr + G.jr > L) fl |= F_SLOWGEOM;
        flags |= fl;
        // contig planes from the packed words (A,C,G,T = 0..3 -> internal code lo' = lo ^ hi, hi' = hi)
        const uint32_t* tw = in.tig2_words + in.row_tig2_off[2 * k + m];
        const int nw16 = (L + 15) >> 4;
        uint32_t* const rbase = rec32 + (m ? rec_planes(0, sb.gap & 1) * sb.W4h * 4 : 0);
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int wi = g + 8 * h;
            const uint32_t x0 = (2 * wi < nw16) ? tw[2 * wi] : 0u, x1 = (2 * wi + 1 < nw16) ? tw[2 * wi + 1] : 0u;
            const uint32_t inl = mask_lt32(L - 32 * wi);
            uint32_t lo = even_bits(x0) | (even_bits(x1) << 16), hi = even_bits(x0 >> 1) | (even_bits(x1 >> 1) << 16);
            lo = (lo ^ hi) & inl;
            hi &= inl;
            uint32_t rlo, rhi, cm;
            ref_word(segpl, G, wi, rlo, rhi, cm);
            const uint32_t dm = fl ? 0u : (cm & ((lo ^ rlo) | (hi ^ rhi)));
            uint32_t* r = rbase + wi;
            const int st = 16;   // 4 * W4
            r[PL_TLO * st] = lo;
            r[PL_THI * st] = hi;
            r[PL_D * st] = dm;
            if (gapbit) r[PL_GAP * st] = 0u;
            t += __popc(dm);
        }
    }
    const int tot = (int)__reduce_add_sync(omask, (unsigned int)t);

    // ---- aux record: lane 0 of the octet gathers, the octet stores ----
    __shared__ __align__(16) int32_t s_aux[8][4][AUX_INTS];
    int32_t* const a = s_aux[(threadIdx.x >> 5) & 7][oct];
    if (g == 0) {
        const int ex = in.row_exact[k];
        const int64_t c0 = in.ex_chain_off[ex], nch = in.ex_chain_off[ex + 1] - c0;
        const int ch = (int)(c0 + in.row_exact_col[2 * k]), cl = (int)(c0 + in.row_exact_col[2 * k + 1]);
        // FWR1 / CDR1 / CDR2 ranges of chain 0 (only meaningful when that chain is the left chain; stage 2 checks)
        int rF0 = 0, rF1 = 0, rC10 = 0, rC11 = 0, rC20 = 0, rC21 = 0;
        const int fr1 = in.ch_fr1[ch], cdr1 = in.ch_cdr1[ch], fr2 = in.ch_fr2[ch], cdr2 = in.ch_cdr2[ch], fr3 = in.ch_fr3[ch];
        {
            const int L = sb.Lh;
            const bool has_f = cdr1 != -1;
            const bool has_c1 = cdr1 != -1 && fr2 != -1 && cdr1 <= fr2;
            const bool has_c2 = cdr2 != -1 && fr3 != -1 && cdr2 <= fr3;
            const bool okf = !has_f || (fr1 >= 0 && fr1 <= cdr1 && cdr1 <= L);
            const bool okc1 = !has_c1 || (cdr1 >= 0 && fr2 <= L);
            const bool okc2 = !has_c2 || (cdr2 >= 0 && fr3 <= L);
            bool disj = true;
            if (has_f && has_c2 && !(cdr1 <= cdr2 || fr3 <= fr1)) disj = false;
            if (has_c1 && has_c2 && !(fr2 <= cdr2 || fr3 <= cdr1)) disj = false;
            if (okf && okc1 && okc2 && disj && L <= 128 * MAXW4) {
                flags |= B_REG_OK;
                if (has_f) { rF0 = fr1; rF1 = cdr1; }
                if (has_c1) { rC10 = cdr1; rC11 = fr2; }
                if (has_c2) { rC20 = cdr2; rC21 = fr3; }
            }
        }
        rowq[q] = (int32_t)k;
        subq[q] = s;
        if (nch >= 2 && nch <= 3) flags |= B_NCH_OK;
        if (nch == 2) flags |= B_NCH2;
        if (nch >= 2 && in.ch_left[c0] != in.ch_left[c0 + 1]) flags |= B_JUN_LEFT;
        if (in.ch_left[ch]) flags |= B_LEFT_H;
        if (in.ch_left[cl]) flags |= B_LEFT_L;
        if (in.ch_amino_off[ch] == -1) flags |= B_AMINO_TIG_H;
        const int64_t x0 = in.ex_cell_off[ex], ncells = in.ex_cell_off[ex + 1] - x0;
        const unsigned long long dmask = donor_mask[ex];
        a[AX_FLAGS] = (int32_t)flags;
        a[AX_VSH] = sv[0];
        a[AX_VSL] = sv[1];
        a[AX_JSH] = sj[0];
        a[AX_JSL] = sj[1];
        a[AX_TOT] = tot;
        a[AX_EXACT] = ex;
        a[AX_NCELLS] = (int32_t)ncells;
        a[AX_DONLO] = (int32_t)(uint32_t)(dmask & 0xffffffffull);
        a[AX_DONHI] = (int32_t)(uint32_t)(dmask >> 32);
        a[AX_CREFH] = in.ch_c_ref[ch];
        a[AX_CREFL] = in.ch_c_ref[cl];
        a[AX_HCOMP] = in.ch_hcomp[ch];
        a[AX_VREFH] = in.ch_v_ref[ch];
        a[AX_VREFL] = in.ch_v_ref[cl];
        a[AX_BC0] = ncells > 0 ? (int32_t)in.cell_barcode[x0] : -1;
        a[AX_VNAMEH] = in.ref_name_id[in.ch_v_ref[ch]];
        a[AX_VNAMEL] = in.ref_name_id[in.ch_v_ref[cl]];
        a[AX_FR1] = fr1;
        a[AX_CDR1] = cdr1;
        a[AX_FR2] = fr2;
        a[AX_CDR2] = cdr2;
        a[AX_FR3] = fr3;
        a[AX_CHH] = ch;
        a[AX_CHL] = cl;
        a[AX_RGN0] = rF0 | (rF1 << 16);
        a[AX_RGN1] = rC10 | (rC11 << 16);
        a[AX_RGN2] = rC20 | (rC21 << 16);
        if (tagq) {
            tagq[q] = (flags & F_SLOWMASK) ? TAG_SLOW : make_ref_tag(sv[0], sv[1], sj[0], sj[1]);
            donq[q] = P.mix_donors ? 0ull : dmask;   // join_one.rs:301-323 only rejects when donors are not mixed
        }
    }
    __syncwarp(omask);
    if (g < AUX_INTS / 4) reinterpret_cast<uint4*>(aux + q * AUX_INTS)[g] = reinterpret_cast<const uint4*>(a)[g];
}

}  // namespace ejb
