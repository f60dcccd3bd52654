// ejb_stage1.cuh — CDR3 Hamming pre-filter over all pairs of a sub-bucket (TMA-staged tiles).
#pragma once
#include "ejb_types.cuh"

namespace ejb {

// ------------------------------------------------------------------------------------------------
// Stage 1: all pairs of a sub-bucket, tiled 128 x 128.  One CTA = one row block x a strip of up to
// S1_STRIP column blocks.  Row/column CDR3 planes + class labels are staged in shared memory by TMA
// bulk copies (double-buffered columns); each thread keeps an 8-row fragment in registers and walks
// 8 columns per tile.  Distance = popc of (lo^lo')|(hi^hi') per 32 bases, words pre-reduced with a
// carry-save adder so that 3 words cost 2 popc.  A pair becomes a candidate iff
// cd <= maxcd  &&  class labels differ (join_core.rs:32 — a pair already connected by
// lexicographically smaller joins can never be in pot).
// ------------------------------------------------------------------------------------------------
struct RowTask {       // (part of) one row block of one sub-bucket, against every column block to its right
    int32_t sub;
    int32_t rb;
    int64_t cta0;      // first CTA of this task inside the launch
    int32_t row_lo, row_hi;   // local rows [row_lo, row_hi) of the block belong to this round's unit
    int32_t tid, pad;         // task id inside the round (per-task candidate counters)
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

// MODE bits (compile time):
//   S1_RUNI   one warp vote per column when the thread's 8 rows and the column carry one class label
//   S1_TAGS   per-row tags staged next to the labels; after the pair loop of a half tile the hit list is compacted in place,
//             dropping hits that join_one is known to reject without looking at the sequences:
//             * one row is known (from an earlier round's stage 2a) to fail the degradation test of join_one.rs:434-440
//               against the other row's reference set — the decision stage 2a would take from its memo;
//             * the two exact subclonotypes come from different donor sets (join_one.rs:301-323; donq is 0 for rows without a
//               donor and for every row under MIX_DONORS).  That test precedes everything that can panic in join_one.
constexpr int S1_RUNI = 1, S1_TAGS = 2;

// Row flags (flagq): relation of a row to the dominant reference set / donor set of its sub-bucket.
constexpr uint32_t RF_DEAD_VS_DOM = 1u;   // deadq[q] == dominant tag: the row fails the degradation test against the dominant reference set
constexpr uint32_t RF_HAS_DOM = 2u;       // tagq[q]  == dominant tag
constexpr uint32_t RF_FOREIGN_DON = 4u;   // donor set non-empty and different from the dominant (non-empty) donor set
constexpr uint32_t RF_DOM_DON = 8u;       // donor set == dominant (non-empty) donor set
// pair rejected by the flags alone?  (dead-vs-dominant x has-dominant, foreign-donor x dominant-donor, either way round)
__device__ __forceinline__ bool flag_kill(uint32_t fr, uint32_t fc, unsigned int& donor_kills, unsigned int& flag_kills) {
    const uint32_t g = ((fr & 5u) << 1) | ((fr & 10u) >> 1);
    const uint32_t k = g & fc;
    if (!k) return false;
    flag_kills++;
    if (!(k & 3u)) donor_kills++;
    return true;
}

// S1_NOVOTE: no per-row warp vote before the distance (the column-level vote S1_RUNI already skips uniform 16 x 16 groups)
#ifndef S1_NOVOTE
#define S1_NOVOTE 0
#endif
#ifndef S1_MINB2
#define S1_MINB2 4
#endif
#ifndef S1_MINB3
#define S1_MINB3 3
#endif
template <int W1, int MODE>
__global__ void __launch_bounds__(S1_THREADS, (W1 <= 2) ? S1_MINB2 : ((W1 <= 3) ? S1_MINB3 : 2)) k_stage1(const SubBucket* __restrict__ subs, const RowTask* __restrict__ tasks, int n_tasks,
                                                       const uint32_t* __restrict__ cdr3w, const int32_t* __restrict__ labq,
                                                       const int32_t* __restrict__ blk_lab, uint2* __restrict__ cand,
                                                       unsigned long long cand_cap, Counters* __restrict__ ctr,
                                                       unsigned long long* __restrict__ unit_cand,
                                                       const unsigned long long* __restrict__ tagq, const unsigned long long* __restrict__ deadq,
                                                       const unsigned long long* __restrict__ donq, const uint32_t* __restrict__ flagq) {
    constexpr bool RUNI = (MODE & S1_RUNI) != 0, TAGS = (MODE & S1_TAGS) != 0;
    constexpr int NP = 2 * W1;            // plane-words per row
    constexpr int TAG0 = (NP + 1) * TILE; // word offset of the staged reference-set tags (uint64 x TILE), then the dead tags
    constexpr int FLAG0 = (NP + 7) * TILE; // word offset of the staged flag words
    constexpr int SLOT = (NP + 1 + (TAGS ? 7 : 0)) * TILE;   // words per staged tile (planes + labels [+ tags + dead tags + donor sets + flags])
    __shared__ __align__(16) uint32_t s_row[SLOT];
    __shared__ __align__(16) uint32_t s_col[2][SLOT];
    __shared__ __align__(8) uint64_t s_bar[3];
    __shared__ uint16_t s_hits[TILE * TILE / 2];
    __shared__ unsigned int s_nhits, s_nkeep;
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
        mbar_expect_tx(bar, bytes * (NP + 1 + (TAGS ? 7 : 0)));
#pragma unroll
        for (int pw = 0; pw < NP; pw++) tma_load_1d(dst + pw * TILE, base + (int64_t)pw * sb.npad + r0, bytes, bar);
        tma_load_1d(dst + NP * TILE, labq + sb.q0 + r0, bytes, bar);   // q0 is a multiple of 4
        if (TAGS) {
            tma_load_1d(dst + TAG0, tagq + sb.q0 + r0, bytes * 2, bar);
            tma_load_1d(dst + TAG0 + 2 * TILE, deadq + sb.q0 + r0, bytes * 2, bar);
            tma_load_1d(dst + TAG0 + 4 * TILE, donq + sb.q0 + r0, bytes * 2, bar);
            tma_load_1d(dst + FLAG0, flagq + sb.q0 + r0, bytes, bar);
        }
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
    // common class label of this thread's 8 rows (-2: mixed): a column whose label equals it needs no per-row test
    int runi = rlab[0];
    if (RUNI) {
#pragma unroll
        for (int i = 1; i < 8; i++)
            if (rlab[i] != runi) runi = -2;
    }
    const int row_lo = t.row_lo;
    const int row_lim = min(sb.n - rb * TILE, t.row_hi);  // valid local rows are row_lo <= r < row_lim
    const int maxcd = sb.maxcd;
    unsigned long long evaluated = 0, killed = 0;   // thread 0 only
    unsigned int donor_kills = 0, flag_kills = 0;    // per thread

    for (int it = 0; it < ncb; it++) {
        const int cb = s_cbs[it];
        const int buf = it & 1;
        mbar_wait(&s_bar[buf], (it >> 1) & 1);
        const uint32_t* sc = s_col[buf];
        const int col_lim = sb.n - cb * TILE;
        const bool diag = (cb == rb);
        const bool full = !diag && row_lo == 0 && row_lim >= TILE && col_lim >= TILE;
        // two halves of 64 columns: the hit list holds at most 128 x 64 entries
        for (int half = 0; half < 2; half++) {
            if (full) {
#pragma unroll 2
                for (int j = half * 4; j < half * 4 + 4; j++) {
                    const int c = j * 16 + tx;
                    const int clab = (int)sc[NP * TILE + c];
                    if (RUNI && __all_sync(0xffffffffu, clab == runi)) continue;   // 16 rows x 16 columns of one class: one vote instead of eight
                    uint32_t cl[W1], chh[W1];
#pragma unroll
                    for (int w = 0; w < W1; w++) {
                        cl[w] = sc[w * TILE + c];
                        chh[w] = sc[(W1 + w) * TILE + c];
                    }
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        // class test first (join_core.rs:32): a warp whose 32 pairs are all already connected skips the
                        // distance entirely — the common case inside a clone once its first rows have been processed
                        const bool need = rlab[i] != clab;
                        if (S1_NOVOTE || __any_sync(0xffffffffu, need)) {
                            const int cd = cdr3_dist<W1>(rl[i], rh[i], cl, chh);
                            if (need && cd <= maxcd) {
                                if (TAGS && flag_kill(s_row[FLAG0 + i * 16 + ty], sc[FLAG0 + c], donor_kills, flag_kills)) continue;
                                const unsigned int h = atomicAdd(&s_nhits, 1u);
                                s_hits[h] = (uint16_t)(((i * 16 + ty) << 7) | c);
                            }
                        }
                    }
                }
            } else {
                for (int j = half * 4; j < half * 4 + 4; j++) {
                    const int c = j * 16 + tx;
                    const int clab = (int)sc[NP * TILE + c];
                    if (RUNI && __all_sync(0xffffffffu, clab == runi)) continue;
                    uint32_t cl[W1], chh[W1];
#pragma unroll
                    for (int w = 0; w < W1; w++) {
                        cl[w] = sc[w * TILE + c];
                        chh[w] = sc[(W1 + w) * TILE + c];
                    }
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const int r = i * 16 + ty;
                        const bool need = r >= row_lo && r < row_lim && c < col_lim && (!diag || r < c) && rlab[i] != clab;
                        if (S1_NOVOTE || __any_sync(0xffffffffu, need)) {
                            const int cd = cdr3_dist<W1>(rl[i], rh[i], cl, chh);
                            if (need && cd <= maxcd) {
                                if (TAGS && flag_kill(s_row[FLAG0 + r], sc[FLAG0 + c], donor_kills, flag_kills)) continue;
                                const unsigned int h = atomicAdd(&s_nhits, 1u);
                                s_hits[h] = (uint16_t)((r << 7) | c);
                            }
                        }
                    }
                }
            }
            __syncthreads();  // all hits of this half recorded (and, after half 1, s_col[buf] drained of plane reads)
            if (TAGS) {
                const unsigned int nh0 = s_nhits;
                if (nh0) {   // block-uniform
                    const unsigned long long* rt = reinterpret_cast<const unsigned long long*>(s_row + TAG0);
                    const unsigned long long* ct = reinterpret_cast<const unsigned long long*>(sc + TAG0);
                    if (tid == 0) s_nkeep = 0;
                    __syncthreads();
                    for (unsigned int b0 = 0; b0 < nh0; b0 += S1_THREADS) {
                        const unsigned int h = b0 + tid;
                        bool keep = false;
                        uint16_t v = 0;
                        if (h < nh0) {
                            v = s_hits[h];
                            const int r = v >> 7, c = v & 127;
                            const unsigned long long dr = rt[2 * TILE + r], dc = ct[2 * TILE + c];
                            const bool donors = dr != 0 && dc != 0 && dr != dc;
                            keep = !(donors || rt[TILE + r] == ct[c] || ct[TILE + c] == rt[r]);
                            donor_kills += donors;
                        }
                        __syncthreads();   // entries [b0, b0 + 256) are in registers; kept ones land below b0 + 256
                        if (keep) s_hits[atomicAdd(&s_nkeep, 1u)] = v;
                    }
                    __syncthreads();
                    if (tid == 0) {
                        killed += nh0 - s_nkeep;
                        s_nhits = s_nkeep;
                    }
                    __syncthreads();
                }
            }
            const unsigned int nh = s_nhits;
            if (tid == 0) {
                if (half == 1) {
                    // prefetch the tile after next into the buffer just drained
                    if (it + 2 < ncb) issue(s_col[buf], s_cbs[it + 2], &s_bar[buf]);
                    const long long hi = min(TILE, row_lim), lo = row_lo, cc = min(TILE, col_lim), nr = hi - lo;
                    if (nr > 0) {
                        if (!diag) evaluated += (unsigned long long)(nr * cc);
                        else {   // pairs (r, c) with lo <= r < hi, r < c < cc
                            const long long h2 = min(hi, cc);
                            const long long n2 = max(h2 - lo, 0ll);
                            evaluated += (unsigned long long)(n2 * (cc - 1) - (lo + h2 - 1) * n2 / 2);
                        }
                    }
                }
                if (nh) {
                    s_base = atomicAdd(&ctr->n_cand, (unsigned long long)nh);
                    atomicAdd(&unit_cand[t.tid], (unsigned long long)nh);
                }
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
    if (tid == 0) {
        atomicAdd(&ctr->pairs_eval, evaluated);
        if (killed) atomicAdd(&ctr->n_s1_kill, killed);
    }
    if (TAGS) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            donor_kills += __shfl_xor_sync(0xffffffffu, donor_kills, o);
            flag_kills += __shfl_xor_sync(0xffffffffu, flag_kills, o);
        }
        if ((tid & 31) == 0 && donor_kills) atomicAdd(&ctr->n_s1_donor, (unsigned long long)donor_kills);
        if ((tid & 31) == 0 && flag_kills) atomicAdd(&ctr->n_s1_kill, (unsigned long long)flag_kills);   // n_s1_kill counts every kind
    }
}


}  // namespace ejb
