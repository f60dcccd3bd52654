// ejb_stage1.cuh — CDR3 Hamming pre-filter over all pairs of a sub-bucket (TMA-staged tiles, warp-private hit queues).
#pragma once
#include "ejb_types.cuh"

namespace ejb {

// ------------------------------------------------------------------------------------------------
// Stage 1: all pairs of a sub-bucket, tiled 128 x 128.  One CTA = one row block x a strip of up to
// S1_STRIP column blocks.  Row/column CDR3 planes, class labels and the per-row tags are staged in
// shared memory by TMA bulk copies through a ring of S1_NBUF column buffers guarded by full / empty
// mbarriers: the eight warps of a CTA never meet at a CTA-wide barrier inside the tile loop, a warp
// whose 16 x 16 groups are already one class runs ahead of the ones that score.
// Each thread keeps an 8-row fragment (8 CONSECUTIVE rows; a warp covers a band of 16 consecutive rows, the granularity at
// which candidates are counted and units are cut) in registers and walks 8 columns per tile.  Per column:
//   1. one warp vote skips 16 rows x 16 columns that carry one class label (join_core.rs:32: a pair
//      already connected by lexicographically smaller joins can never be in pot);
//   2. the FIRST 32 positions of the eight pairs go through XOR + popc; a pair can only be CDR3-close
//      if that partial distance is already <= maxcd, so one more warp vote ends the column for
//      unrelated sequences (typically ~24 of 32 positions differ) after ONE popc per pair;
//   3. otherwise the remaining words are added, hits are tested against the per-row tags and pushed
//      to the warp's private queue in shared memory (ballot compaction, no atomics); a queue is
//      flushed to the global candidate list with one atomicAdd per ~100 candidates.
// A pair becomes a candidate iff cd <= maxcd && class labels differ && not rejected by the tags.
// ------------------------------------------------------------------------------------------------
struct RowTask {       // (part of) one row block of one sub-bucket, against every column block to its right
    int32_t sub;
    int32_t rb;
    int32_t cta0;      // first CTA of this task inside the launch
    int32_t row_lo, row_hi;   // local rows [row_lo, row_hi) of the block belong to this round's unit
    int32_t tid;              // task id inside the round (per-task candidate counters)
    // the unit's truncation rule (k_truncate), for the early return of blocks that are certain to be cut away:
    int32_t allow;            // candidate allowance of the unit (saturated)
    float rho_ref;            // candidates per row of the sub-bucket's last committed row range (< 0: unknown)
};
// k_truncate cuts a unit at the end of the first 16-row band where the candidates counted so far exceed `allow`, or — once a
// band's density has jumped above 8 x rho_ref + 16 (a clone that is not connected yet) — where the candidates since the jump
// exceed `strict`.  Replaying that rule on the bands of the PREVIOUS block alone (no carry from earlier blocks, counts as they
// stand: both can only be smaller than what k_truncate will see) gives a sufficient condition for "this block will be cut away":
// its candidates would be thrown away and its rows re-run next round, so its CTAs return at once.
__device__ __forceinline__ bool block_is_cut(const RowTask& prev, const unsigned long long* band_cnt /* 8, volatile reads */, const RowTask& t, double strict) {
    double cum = 0.0, aj = 0.0;
    bool jumped = t.rho_ref < 0.0f;
    for (int b = 0; b < 8; b++) {
        const int rows = min(prev.row_hi, 16 * b + 16) - max(prev.row_lo, 16 * b);
        if (rows <= 0) continue;
        const double cnt = (double)*reinterpret_cast<const volatile unsigned long long*>(&band_cnt[b]);
        if (!jumped && cnt / (double)rows > 8.0 * (double)t.rho_ref + 16.0) jumped = true;
        cum += cnt;
        if (jumped) aj += cnt;
        if (cum > (double)t.allow || (jumped && aj > strict)) return true;
    }
    return false;
}

// distance contribution of the words after the first
template <int W1>
__device__ __forceinline__ int cdr3_rest(const uint32_t* rl, const uint32_t* rh, const uint32_t* cl, const uint32_t* chh) {
    if constexpr (W1 == 1) {
        return 0;
    } else {
        uint32_t x[W1];
#pragma unroll
        for (int w = 1; w < W1; w++) x[w] = (rl[w] ^ cl[w]) | (rh[w] ^ chh[w]);
        if constexpr (W1 == 2) {
            return __popc(x[1]);
        } else if constexpr (W1 == 3) {
            return __popc(x[1]) + __popc(x[2]);
        } else {
            // carry-save adder over three words: 2 popc instead of 3
            const uint32_t s = x[1] ^ x[2] ^ x[3];
            const uint32_t c = (x[1] & x[2]) | (x[3] & (x[1] ^ x[2]));
            int d = __popc(s) + 2 * __popc(c);
            if constexpr (W1 >= 5) d += __popc(x[4]);
            if constexpr (W1 >= 6) d += __popc(x[5]);
            return d;
        }
    }
}
// popc instructions per lane of cdr3_rest (for the executed-popc statistics)
__host__ __device__ constexpr int cdr3_rest_popc(int w1) { return w1 <= 1 ? 0 : (w1 == 2 ? 1 : (w1 <= 4 ? 2 : (w1 == 5 ? 3 : 4))); }

// MODE bits (compile time):
//   S1_RUNI   one warp vote per column when the thread's 8 rows and the column carry one class label
//   S1_TAGS   per-row tags staged next to the labels; a hit is dropped when join_one is known to reject it without looking at
//             the sequences:
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
// A pair is rejected by the flags alone when: dead-vs-dominant x has-dominant, or foreign-donor x dominant-donor, either way round.

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

#ifndef S1_MINB2
#define S1_MINB2 3   // 85 registers: the 8-row fragment (32 words + 8 labels at W1 = 2) stays in registers; 4 CTAs (64 registers) spill it
#endif
#ifndef S1_MINB3
#define S1_MINB3 3
#endif
constexpr int S1_QCAP = 128;              // entries of a warp's private hit queue
constexpr int S1_WARPS = S1_THREADS / 32;
template <int W1>
struct S1Cfg {
    static constexpr int NBUF = (W1 <= 4) ? 3 : 2;   // column buffers in the TMA ring (static shared memory stays below 48 KB)
};

template <int W1, int MODE>
__global__ void __launch_bounds__(S1_THREADS, (W1 <= 2) ? S1_MINB2 : ((W1 <= 3) ? S1_MINB3 : 2)) k_stage1(const SubBucket* __restrict__ subs, const RowTask* __restrict__ tasks, int n_tasks,
                                                       const uint32_t* __restrict__ cdr3w, const int32_t* __restrict__ labq,
                                                       const int32_t* __restrict__ blk_lab, uint2* __restrict__ cand,
                                                       unsigned long long cand_cap, Counters* __restrict__ ctr,
                                                       unsigned long long* __restrict__ unit_cand,
                                                       const unsigned long long* __restrict__ tagq, const unsigned long long* __restrict__ deadq,
                                                       const unsigned long long* __restrict__ donq, const uint32_t* __restrict__ flagq, int strip,
                                                       uint8_t* __restrict__ task_skip, double strict, const int32_t* __restrict__ cta_task) {
    constexpr bool RUNI = (MODE & S1_RUNI) != 0, TAGS = (MODE & S1_TAGS) != 0;
    constexpr int NP = 2 * W1;            // plane-words per row
    constexpr int TAG0 = (NP + 1) * TILE; // word offset of the staged reference-set tags (uint64 x TILE), then the dead tags, then the donor sets
    constexpr int FLAG0 = (NP + 7) * TILE; // word offset of the staged flag words
    constexpr int SLOT = (NP + 1 + (TAGS ? 7 : 0)) * TILE;   // words per staged tile (planes + labels [+ tags + dead tags + donor sets + flags])
    constexpr int NBUF = S1Cfg<W1>::NBUF;
    __shared__ __align__(16) uint32_t s_row[SLOT];
    __shared__ __align__(16) uint32_t s_col[NBUF][SLOT];
    __shared__ __align__(8) uint64_t s_full[NBUF], s_empty[NBUF], s_rowbar;
    __shared__ __align__(8) uint2 s_q[S1_WARPS][S1_QCAP];
    __shared__ int s_cbs[S1_STRIP];
    __shared__ int s_ncb;

    // ---- locate the task: one load from the CTA -> task map of this launch (k_expand_tasks) ----
    const int64_t cta = blockIdx.x;
    const int lo = cta_task[cta];
    (void)n_tasks;
    const RowTask t = tasks[lo];
    const SubBucket sb = subs[t.sub];
    if (sb.W1 != W1) return;  // launched once per W1; other widths are handled by their own launch
    const int nb = (sb.n + TILE - 1) / TILE;
    const int rb = t.rb;
    const int cb_first = rb + (int)(cta - t.cta0) * strip;   // strip: column tiles per CTA of this launch (host: 4 .. S1_STRIP)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx = tid & 15, ty = tid >> 4;

    // class-skip at block level: both blocks uniformly in one class.  Warp 0 looks at the (up to 16) column blocks of the strip
    // in parallel and compacts the ones that need work.
    if (warp == 0) {
        // behind a dense block the rest of the unit is cut away after stage 1 (block_is_cut); one thread decides for the CTA.
        // strict <= 0 switches the early return off.
        bool skip = false;
        if (lane == 0 && strict > 0.0 && lo > 0 && tasks[lo - 1].sub == t.sub && tasks[lo - 1].rb == rb - 1) {
            const RowTask prev = tasks[lo - 1];
            skip = *reinterpret_cast<volatile uint8_t*>(&task_skip[prev.tid]) != 0 ||
                   block_is_cut(prev, unit_cand + (size_t)prev.tid * S1_WARPS, t, strict);
            if (skip) task_skip[t.tid] = 1;
        }
        skip = __shfl_sync(0xffffffffu, skip ? 1 : 0, 0) != 0;
        const int lr = blk_lab[sb.blk0 + rb];
        const int cbl = cb_first + lane;
        bool work = false;
        if (!skip && lane < strip && cbl < nb) {
            const int lc = blk_lab[sb.blk0 + cbl];
            work = !(lr >= 0 && lr == lc);
        }
        const unsigned int wm = __ballot_sync(0xffffffffu, work);
        if (work) s_cbs[__popc(wm & ((1u << lane) - 1u))] = cbl;
        if (lane == 0) {
            s_ncb = __popc(wm);
            for (int b = 0; b < NBUF; b++) {
                mbar_init(&s_full[b], 1);
                mbar_init(&s_empty[b], S1_WARPS);
            }
            mbar_init(&s_rowbar, 1);
            mbar_fence_init();
        }
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
        issue(s_row, rb, &s_rowbar);
        for (int b = 0; b < NBUF && b < ncb; b++) issue(s_col[b], s_cbs[b], &s_full[b]);
    }

    // ---- row fragment -> registers ----
    mbar_wait(&s_rowbar, 0);
    uint32_t rl[8][W1], rh[8][W1];
    int rlab[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int r = ty * 8 + i;
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
        runi = __shfl_sync(0xffffffffu, runi, lane);   // identity shuffle: keeps the value in a register (ptxas otherwise re-derives it
                                                       // from the eight labels for every column, 14 instructions instead of none)
    }
    // row flags of the thread's eight rows as four byte masks (bit i = row i): a column's flag nibble selects, with four
    // operations, the rows whose pair with that column join_one rejects without looking at the sequences (flag_kill)
    uint32_t r_dead8 = 0, r_dom8 = 0, r_fd8 = 0, r_dd8 = 0;
    if (TAGS) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t f = s_row[FLAG0 + ty * 8 + i];
            r_dead8 |= ((f >> 0) & 1u) << i;
            r_dom8 |= ((f >> 1) & 1u) << i;
            r_fd8 |= ((f >> 2) & 1u) << i;
            r_dd8 |= ((f >> 3) & 1u) << i;
        }
    }
    const int row_lo = t.row_lo;
    const int row_lim = min(sb.n - rb * TILE, t.row_hi);  // valid local rows are row_lo <= r < row_lim
    // a warp none of whose 16 rows belongs to the unit (the first rounds process 2 .. 32 rows of a block) has nothing to do in any
    // column tile: it only releases the buffers
    bool warp_rows = false;
#pragma unroll
    for (int i = 0; i < 8; i++) warp_rows |= (ty * 8 + i >= row_lo) & (ty * 8 + i < row_lim);
    warp_rows = __any_sync(0xffffffffu, warp_rows);
    const int maxcd = sb.maxcd;
    const uint32_t q_r0 = (uint32_t)(sb.q0 + (int64_t)rb * TILE);
    unsigned long long evaluated = 0;                 // thread 0 only
    unsigned int donor_kills = 0, tag_kills = 0;      // per thread
    unsigned int n_first = 0, n_full = 0;             // per lane: first-word / remaining-word evaluations
    unsigned int n_flag = 0;                          // per lane: pairs of computed columns rejected by the row flags alone
    unsigned int wq = 0;                              // warp-uniform: entries in this warp's queue
    unsigned long long w_total = 0;                   // warp-uniform: candidates this warp emitted
    uint2* const myq = s_q[warp];
    const unsigned int lt_mask = (1u << lane) - 1u;

    auto flush = [&]() {   // all lanes
        __syncwarp();
        unsigned long long b = 0;
        if (lane == 0) b = atomicAdd(&ctr->n_cand, (unsigned long long)wq);
        b = __shfl_sync(0xffffffffu, b, 0);
        for (unsigned int h = lane; h < wq; h += 32) {
            const unsigned long long o = b + h;
            if (o < cand_cap) cand[o] = myq[h]; else ctr->overflow = 1ull;
        }
        w_total += wq;
        wq = 0;
        __syncwarp();
    };

    const unsigned long long* const rt = reinterpret_cast<const unsigned long long*>(s_row + TAG0);

    for (int it = 0; it < ncb; it++) {
        const int buf = it % NBUF;
        // producer: the buffer of tile it-1 is refilled with tile it-1+NBUF once every warp has released it
        if (tid == 0 && it >= 1 && it - 1 + NBUF < ncb) {
            const int pb = (it - 1) % NBUF;
            mbar_wait(&s_empty[pb], ((it - 1) / NBUF) & 1);
            issue(s_col[pb], s_cbs[it - 1 + NBUF], &s_full[pb]);
        }
        __syncwarp();
        const int cb = s_cbs[it];
        mbar_wait(&s_full[buf], (it / NBUF) & 1);
        const uint32_t* sc = s_col[buf];
        const unsigned long long* const ct = reinterpret_cast<const unsigned long long*>(sc + TAG0);
        const int col_lim = sb.n - cb * TILE;
        const bool diag = (cb == rb);
        const bool full = !diag && row_lo == 0 && row_lim >= TILE && col_lim >= TILE;
        const uint32_t q_c0 = (uint32_t)(sb.q0 + (int64_t)cb * TILE);
        const uint32_t* pc = sc + tx;   // this thread's column of the staged tile, advanced by 16 per step
#pragma unroll 1
        for (int j = warp_rows ? 0 : 8; j < 8; j++, pc += 16) {
            const int c = j * 16 + tx;
            const int clab = (int)pc[NP * TILE];
            if (RUNI && __all_sync(0xffffffffu, clab == runi)) continue;   // 16 rows x 16 columns of one class: one vote instead of eight
            // first word of the eight pairs (the 32 most diverse CDR3 positions, SubBucket::rot): is any pair still a possible candidate?
            const uint32_t cl0 = pc[0], ch0 = pc[W1 * TILE];
            uint32_t kill8 = 0;   // bit i: the pair (row i, this column) is rejected by the row flags (reference set / donor set)
            if (TAGS) {
                const uint32_t fc = pc[FLAG0];
                kill8 = ((fc & RF_HAS_DOM) ? r_dead8 : 0u) | ((fc & RF_DEAD_VS_DOM) ? r_dom8 : 0u) | ((fc & RF_DOM_DON) ? r_fd8 : 0u) |
                        ((fc & RF_FOREIGN_DON) ? r_dd8 : 0u);
                n_flag += __popc(kill8);
            }
            bool any = false;
            if (full) {
#pragma unroll
                for (int i = 0; i < 8; i++)
                    any |= (__popc((rl[i][0] ^ cl0) | (rh[i][0] ^ ch0)) <= maxcd) & (rlab[i] != clab) & !((kill8 >> i) & 1u);
            } else {
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int r = ty * 8 + i;
                    any |= (__popc((rl[i][0] ^ cl0) | (rh[i][0] ^ ch0)) <= maxcd) & (rlab[i] != clab) & !((kill8 >> i) & 1u) & (r >= row_lo) & (r < row_lim) &
                           (c < col_lim) & (!diag | (r < c));
                }
            }
            n_first += 8;
            if (!__any_sync(0xffffffffu, any)) continue;
            // rare: some pair of this warp passed the first word — full distances for the rows that still have a live pair
            uint32_t cl[W1], chh[W1];
#pragma unroll
            for (int w = 0; w < W1; w++) {
                cl[w] = pc[w * TILE];
                chh[w] = pc[(W1 + w) * TILE];
            }
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int r = ty * 8 + i;
                const int p0 = __popc((rl[i][0] ^ cl[0]) | (rh[i][0] ^ chh[0]));
                bool hit = p0 <= maxcd && rlab[i] != clab && !((kill8 >> i) & 1u);
                if (!full) hit = hit && r >= row_lo && r < row_lim && c < col_lim && (!diag || r < c);
                if (!__any_sync(0xffffffffu, hit)) continue;
                if (W1 > 1) {
                    n_full++;
                    hit = hit && p0 + cdr3_rest<W1>(rl[i], rh[i], cl, chh) <= maxcd;
                }
                if (TAGS && hit) {   // the exact 64-bit tests behind the flags (rows off the dominant reference / donor set)
                    const unsigned long long dr = rt[2 * TILE + r], dc = ct[2 * TILE + c];
                    const bool donors = dr != 0 && dc != 0 && dr != dc;
                    if (donors || rt[TILE + r] == ct[c] || ct[TILE + c] == rt[r]) {
                        hit = false;
                        tag_kills++;
                        donor_kills += donors;
                    }
                }
                const unsigned int m = __ballot_sync(0xffffffffu, hit);
                if (m) {
                    if (hit) myq[wq + __popc(m & lt_mask)] = make_uint2(q_r0 + r, q_c0 + c);
                    wq += __popc(m);
                    if (wq > S1_QCAP - 32) flush();
                }
            }
        }
        // this warp is done with the column buffer
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[buf]);
        if (tid == 0) {
            const long long hi2 = min(TILE, row_lim), lo2 = row_lo, cc = min(TILE, col_lim), nr = hi2 - lo2;
            if (nr > 0) {
                if (!diag) evaluated += (unsigned long long)(nr * cc);
                else {   // pairs (r, c) with lo <= r < hi, r < c < cc
                    const long long h2 = min(hi2, cc);
                    const long long n2 = max(h2 - lo2, 0ll);
                    evaluated += (unsigned long long)(n2 * (cc - 1) - (lo2 + h2 - 1) * n2 / 2);
                }
            }
        }
    }
    if (wq) flush();
    if (lane == 0 && w_total) atomicAdd(&unit_cand[(size_t)t.tid * S1_WARPS + warp], w_total);   // per 16-row band
    if (tid == 0) atomicAdd(&ctr->pairs_eval, evaluated);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n_first += __shfl_xor_sync(0xffffffffu, n_first, o);
        n_full += __shfl_xor_sync(0xffffffffu, n_full, o);
        if (TAGS) {
            n_flag += __shfl_xor_sync(0xffffffffu, n_flag, o);
            donor_kills += __shfl_xor_sync(0xffffffffu, donor_kills, o);
            tag_kills += __shfl_xor_sync(0xffffffffu, tag_kills, o);
        }
    }
    if (lane == 0) {
        atomicAdd(&ctr->s1_first, (unsigned long long)n_first);
        if (n_full) atomicAdd(&ctr->s1_full, (unsigned long long)n_full);
        if (TAGS && n_flag) atomicAdd(&ctr->n_s1_flag, (unsigned long long)n_flag);
        if (TAGS && donor_kills) atomicAdd(&ctr->n_s1_donor, (unsigned long long)donor_kills);
        if (TAGS && tag_kills) atomicAdd(&ctr->n_s1_kill, (unsigned long long)tag_kills);   // n_s1_kill counts every kind
    }
}

// ------------------------------------------------------------------------------------------------
// Stage 1 for THIN units: at most `thin_rows` rows (the first units of every sub-bucket: 2 rows, then 12, then 72) against
// every column to their right.  k_stage1 stages 128-column tiles for 128 rows, so a thin task pays the whole tile pipeline
// for a fraction of a tile of pairs; here the roles are swapped: the rows (planes, labels, flags, tags) sit in shared memory,
// a THREAD owns a column (coalesced loads of the plane-major CDR3 words, labels and flags) and walks the rows.  Same
// predicates, in the same order, as k_stage1: first word, class labels, row flags, remaining words, the exact 64-bit tag tests.
// One CTA = `strip` column blocks of the task (the CTA -> task map of k_expand_tasks is shared with k_stage1).
// ------------------------------------------------------------------------------------------------
constexpr int S1_THIN_ROWS = TILE;   // a thin task never exceeds its row block
template <int W1, bool TAGS>
__global__ void __launch_bounds__(S1_THREADS) k_stage1_thin(const SubBucket* __restrict__ subs, const RowTask* __restrict__ tasks, const uint32_t* __restrict__ cdr3w,
                                                            const int32_t* __restrict__ labq, uint2* __restrict__ cand, unsigned long long cand_cap,
                                                            Counters* __restrict__ ctr, unsigned long long* __restrict__ unit_cand,
                                                            const unsigned long long* __restrict__ tagq, const unsigned long long* __restrict__ deadq,
                                                            const unsigned long long* __restrict__ donq, const uint32_t* __restrict__ flagq, int strip,
                                                            const int32_t* __restrict__ cta_task) {
    __shared__ uint32_t s_rl[S1_THIN_ROWS][W1], s_rh[S1_THIN_ROWS][W1];
    __shared__ int s_lab[S1_THIN_ROWS];
    __shared__ uint32_t s_flag[S1_THIN_ROWS];
    __shared__ unsigned long long s_tag[S1_THIN_ROWS], s_dead[S1_THIN_ROWS], s_don[S1_THIN_ROWS];
    __shared__ __align__(8) uint2 s_q[S1_WARPS][S1_QCAP];
    const int64_t cta = blockIdx.x;
    const RowTask t = tasks[cta_task[cta]];
    const SubBucket sb = subs[t.sub];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int R0 = t.rb * TILE + t.row_lo;                                  // first row (sub-bucket local)
    const int nr = min(min(t.row_hi, sb.n - t.rb * TILE) - t.row_lo, S1_THIN_ROWS);   // rows of the task
    if (nr <= 0) return;
    const uint32_t* base = cdr3w + sb.s1_off;
    if (tid < nr) {
        const int r = R0 + tid;
#pragma unroll
        for (int w = 0; w < W1; w++) {
            s_rl[tid][w] = base[(int64_t)w * sb.npad + r];
            s_rh[tid][w] = base[(int64_t)(W1 + w) * sb.npad + r];
        }
        s_lab[tid] = labq[sb.q0 + r];
        if (TAGS) {
            s_flag[tid] = flagq[sb.q0 + r];
            s_tag[tid] = tagq[sb.q0 + r];
            s_dead[tid] = deadq[sb.q0 + r];
            s_don[tid] = donq[sb.q0 + r];
        }
    }
    __syncthreads();
    const int maxcd = sb.maxcd;
    const int c_beg = max((t.rb + (int)(cta - t.cta0) * strip) * TILE, R0 + 1);
    const int c_end = min(sb.n, (t.rb + (int)(cta - t.cta0) * strip + strip) * TILE);
    uint2* const myq = s_q[warp];
    const unsigned int lt_mask = (1u << lane) - 1u;
    unsigned int wq = 0;
    __shared__ unsigned int s_band[S1_WARPS];   // candidates per 16-row band of the row block
    if (tid < S1_WARPS) s_band[tid] = 0u;
    __syncthreads();
    unsigned int n_first = 0, n_full = 0, n_flag = 0, donor_kills = 0, tag_kills = 0;
    auto flush = [&]() {   // all lanes
        __syncwarp();
        unsigned long long b = 0;
        if (lane == 0) b = atomicAdd(&ctr->n_cand, (unsigned long long)wq);
        b = __shfl_sync(0xffffffffu, b, 0);
        for (unsigned int h = lane; h < wq; h += 32) {
            const unsigned long long o = b + h;
            if (o < cand_cap) cand[o] = myq[h]; else ctr->overflow = 1ull;
        }
        wq = 0;
        __syncwarp();
    };
    // whole warps walk the column range so that the ballots below are warp-wide
    for (int c0 = c_beg - (c_beg & 31) + warp * 32; c0 < c_end; c0 += S1_THREADS) {
        const int c = c0 + lane;
        const bool cv = c >= c_beg && c < c_end;
        uint32_t cl[W1], chh[W1];
        int clab = 0;
        uint32_t fc = 0;
        if (cv) {
#pragma unroll
            for (int w = 0; w < W1; w++) {
                cl[w] = base[(int64_t)w * sb.npad + c];
                chh[w] = base[(int64_t)(W1 + w) * sb.npad + c];
            }
            clab = labq[sb.q0 + c];
            if (TAGS) fc = flagq[sb.q0 + c];
        } else {
#pragma unroll
            for (int w = 0; w < W1; w++) cl[w] = chh[w] = 0u;
        }
        for (int i = 0; i < nr; i++) {
            const int r = R0 + i;
            bool hit = cv && c > r;
            if (hit) n_first++;
            bool fk = false;
            if (TAGS) {
                const uint32_t fr = s_flag[i];
                fk = ((fc & RF_HAS_DOM) && (fr & RF_DEAD_VS_DOM)) || ((fc & RF_DEAD_VS_DOM) && (fr & RF_HAS_DOM)) || ((fc & RF_DOM_DON) && (fr & RF_FOREIGN_DON)) ||
                     ((fc & RF_FOREIGN_DON) && (fr & RF_DOM_DON));
                if (hit && fk) n_flag++;
            }
            const int p0 = __popc((s_rl[i][0] ^ cl[0]) | (s_rh[i][0] ^ chh[0]));
            hit = hit && p0 <= maxcd && s_lab[i] != clab && !fk;
            if (!__any_sync(0xffffffffu, hit)) continue;
            if (W1 > 1 && hit) {
                n_full++;
                hit = p0 + cdr3_rest<W1>(s_rl[i], s_rh[i], cl, chh) <= maxcd;
            }
            if (TAGS && hit) {   // the exact 64-bit tests behind the flags
                const unsigned long long dr = s_don[i], dc = donq[sb.q0 + c];
                const bool donors = dr != 0 && dc != 0 && dr != dc;
                if (donors || s_dead[i] == tagq[sb.q0 + c] || deadq[sb.q0 + c] == s_tag[i]) {
                    hit = false;
                    tag_kills++;
                    donor_kills += donors;
                }
            }
            const unsigned int m = __ballot_sync(0xffffffffu, hit);
            if (m) {
                if (hit) myq[wq + __popc(m & lt_mask)] = make_uint2((uint32_t)(sb.q0 + r), (uint32_t)(sb.q0 + c));
                wq += __popc(m);
                if (lane == 0) atomicAdd(&s_band[(t.row_lo + i) >> 4], (unsigned int)__popc(m));
                if (wq > S1_QCAP - 32) flush();
            }
        }
    }
    if (wq) flush();
    __syncthreads();
    if (tid < S1_WARPS && s_band[tid]) atomicAdd(&unit_cand[(size_t)t.tid * S1_WARPS + tid], (unsigned long long)s_band[tid]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n_first += __shfl_xor_sync(0xffffffffu, n_first, o);
        n_full += __shfl_xor_sync(0xffffffffu, n_full, o);
        if (TAGS) {
            n_flag += __shfl_xor_sync(0xffffffffu, n_flag, o);
            donor_kills += __shfl_xor_sync(0xffffffffu, donor_kills, o);
            tag_kills += __shfl_xor_sync(0xffffffffu, tag_kills, o);
        }
    }
    if (lane == 0) {
        if (n_first) {
            atomicAdd(&ctr->s1_first, (unsigned long long)n_first);
            atomicAdd(&ctr->pairs_eval, (unsigned long long)n_first);   // every pair of the task's columns is looked at
        }
        if (n_full) atomicAdd(&ctr->s1_full, (unsigned long long)n_full);
        if (TAGS && n_flag) atomicAdd(&ctr->n_s1_flag, (unsigned long long)n_flag);
        if (TAGS && donor_kills) atomicAdd(&ctr->n_s1_donor, (unsigned long long)donor_kills);
        if (TAGS && tag_kills) atomicAdd(&ctr->n_s1_kill, (unsigned long long)tag_kills);
    }
}

}  // namespace ejb
