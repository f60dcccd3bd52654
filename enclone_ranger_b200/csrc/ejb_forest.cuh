// ejb_forest.cuh — lexicographic spanning forest (Boruvka), finalize kernels, pipe-peak microbenchmarks.
#pragma once
#include "ejb_p1.cuh"
#include "ejb_stage1.cuh"
#include "ejb_types.cuh"

namespace ejb {

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
constexpr int BOR_MAX_ITERS = 60;

// grid-wide barrier of a cooperatively launched kernel (all CTAs co-resident): monotone arrival counter, zeroed by the host
// before the launch; `gen` counts the barriers this CTA has passed
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int& gen) {
    __syncthreads();
    if (threadIdx.x == 0) {
        gen++;
        __threadfence();
        atomicAdd(bar, 1u);
        const unsigned int target = gen * gridDim.x;
        while (*reinterpret_cast<volatile unsigned int*>(bar) < target) {
        }
        __threadfence();
    }
    __syncthreads();
}

struct ForestArgs {
    const unsigned long long* pass_w;   // pass edges of this round, weight (k1 << 28) | k2
    const PassTuple* pass_t;
    const unsigned long long* n_pass_p;
    unsigned long long cap;
    int32_t* parent;
    unsigned long long* best;           // per row (only roots are used)
    int2* roots;                        // per pass edge: roots of its endpoints at the start of the iteration
    unsigned long long* forest;
    PassTuple* forest_t;
    unsigned long long forest_cap;
    Counters* ctr;
    unsigned int* bar;
    // class labels refreshed at the end of the round
    const int32_t* rowq;
    int64_t n_pos;
    int32_t* labq;
    const SubBucket* subs;
    int32_t* blk_lab;
    const int32_t* blk_sub;
    int n_blocks;
};

// One scheduler round of the forest: Boruvka iterations to convergence on this round's pass edges (every component's lightest
// outgoing edge is in the minimum spanning forest; weights are distinct), then the class labels of every packed row and the
// per-128-row-block uniform labels.  ONE cooperative launch per round: the iterations are separated by grid barriers instead of
// host read-backs.  Nothing is committed when the round overflowed its work buffers.
__global__ void __launch_bounds__(256) k_forest_round(ForestArgs A) {
    unsigned int gen = 0;
    const unsigned long long tid0 = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long nthr = (unsigned long long)gridDim.x * blockDim.x;
    if (A.ctr->overflow) return;   // uniform: written by earlier kernels only
    const unsigned long long n = min(*A.n_pass_p, A.cap);
    if (n == 0) return;            // no new edge: labels are unchanged
    // reset the per-root minima this round can touch
    for (unsigned long long i = tid0; i < n; i += nthr) {
        const unsigned long long e = A.pass_w[i];
        A.best[uf_find_ro(A.parent, (int)(e >> 28))] = ~0ull;
        A.best[uf_find_ro(A.parent, (int)(e & 0xfffffffull))] = ~0ull;
    }
    grid_barrier(A.bar, gen);
    int iter = 0;
    for (;; iter++) {
        // newer iterations carry smaller tags, so their minima override the stale ones without a reset
        const unsigned long long tag = (unsigned long long)(63 - iter) << BOR_TAG_SHIFT;
        unsigned long long alive = 0;
        for (unsigned long long i0 = tid0 - (threadIdx.x & 31); i0 < n; i0 += nthr) {   // warp-uniform trip count
            const unsigned long long i = i0 + (threadIdx.x & 31);
            int ra = -1, rb = -1;
            unsigned long long e = 0;
            if (i < n) {
                e = A.pass_w[i];
                const int a = (int)(e >> 28), b = (int)(e & 0xfffffffull);
                ra = uf_find_ro(A.parent, a);
                rb = uf_find_ro(A.parent, b);
                // pointer jumping: no hook runs concurrently with this phase, a root is a valid parent
                if (A.parent[a] != ra) A.parent[a] = ra;
                if (A.parent[b] != rb) A.parent[b] = rb;
                A.roots[i] = make_int2(ra, rb);
            }
            const bool live = i < n && ra != rb;
            const unsigned int lm = __ballot_sync(0xffffffffu, live);
            if (live) {
                // Consecutive pass edges mostly share a root (one clone): the lanes that update the same root reduce their
                // minimum first, so that a root costs one atomic per warp instead of up to 32 serialized ones.
                const unsigned long long v = tag | e;
                const int lane = threadIdx.x & 31;
#pragma unroll
                for (int side = 0; side < 2; side++) {
                    const int root = side ? rb : ra;
                    const unsigned int g = __match_any_sync(lm, root);
                    const unsigned int hi = __reduce_min_sync(g, (unsigned int)(v >> 32));
                    const unsigned int lo = __reduce_min_sync(g, (unsigned int)(v >> 32) == hi ? (unsigned int)v : 0xffffffffu);
                    if (lane == __ffs(g) - 1) atomicMin(&A.best[root], ((unsigned long long)hi << 32) | lo);
                }
                alive++;
            }
        }
        for (int o = 16; o > 0; o >>= 1) alive += __shfl_xor_sync(0xffffffffu, alive, o);
        if ((threadIdx.x & 31) == 0 && alive) atomicAdd(&A.ctr->alive_it[iter], alive);
        grid_barrier(A.bar, gen);
        if (*reinterpret_cast<volatile unsigned long long*>(&A.ctr->alive_it[iter]) == 0 || iter >= BOR_MAX_ITERS) break;   // uniform
        // hook: an edge that is the minimum of one of its components joins the forest
        for (unsigned long long i = tid0; i < n; i += nthr) {
            const int2 r = A.roots[i];
            if (r.x == r.y) continue;
            const unsigned long long e = A.pass_w[i], te = tag | e;
            if (A.best[r.x] == te || A.best[r.y] == te) {
                const unsigned long long o = atomicAdd(&A.ctr->n_forest, 1ull);
                if (o < A.forest_cap) {
                    A.forest[o] = e;
                    A.forest_t[o] = A.pass_t[i];
                } else {
                    A.ctr->overflow = 2ull;
                }
                uf_union(A.parent, (int)(e >> 28), (int)(e & 0xfffffffull));
            }
        }
        grid_barrier(A.bar, gen);
        // re-point every old root straight at its new root
        for (unsigned long long i = tid0; i < n; i += nthr) {
            const int2 r = A.roots[i];
            if (r.x == r.y) continue;
            const int nx = uf_find_ro(A.parent, r.x), ny = uf_find_ro(A.parent, r.y);
            if (A.parent[r.x] != nx) A.parent[r.x] = nx;
            if (A.parent[r.y] != ny) A.parent[r.y] = ny;
        }
        grid_barrier(A.bar, gen);
    }
    if (tid0 == 0) atomicAdd(&A.ctr->bor_iters, (unsigned long long)iter);
    // class label of every packed row
    for (unsigned long long q = tid0; q < (unsigned long long)A.n_pos; q += nthr) {
        const int k = A.rowq[q];
        if (k < 0) continue;  // padding position
        const int r = uf_find_ro(A.parent, k);
        if (A.parent[k] != r) A.parent[k] = r;
        A.labq[q] = r;
    }
    grid_barrier(A.bar, gen);
    // per 128-row block: the common label or -1 (one warp per block)
    const int lane = threadIdx.x & 31;
    for (unsigned long long b = tid0 >> 5; b < (unsigned long long)A.n_blocks; b += nthr >> 5) {
        const SubBucket sb = A.subs[A.blk_sub[b]];
        const int rb = (int)((long long)b - sb.blk0);
        const int r0 = rb * TILE, r1 = min(sb.n, r0 + TILE);
        const int first = A.labq[sb.q0 + r0];
        bool same = true;
        for (int r = r0 + lane; r < r1; r += 32)
            if (A.labq[sb.q0 + r] != first) same = false;
        same = __all_sync(0xffffffffu, same);
        if (lane == 0) A.blk_lab[b] = same ? first : -1;
    }
}

// Unit truncation after stage 1 (was a host step between two stream synchronisations): a unit whose counted candidates exceed
// its allowance keeps whole 16-row BANDS (the rows one stage-1 warp covers) up to and including the one that crosses it; stage 2
// ignores candidates of the rows cut away (qcut) and the scheduler resumes from the cut.  A jump in candidate density (a clone
// that is not connected yet) switches to the strict allowance: the fewer rows of such a clone are scored before the class
// labels are refreshed, the fewer redundant pass edges (every one of its rows joins every other).  One warp per unit; the
// bands of a unit are consecutive and in row order, 8 per task (row block).
struct UnitDesc {
    int32_t sub, t0, nt, r1;   // sub-bucket, first task, number of tasks, planned end row
    double allow, strict, rho_ref;
};
struct UnitOut {
    int32_t r1, cut;           // committed end row; 1 when the unit was truncated
    double cand;               // candidates of the committed rows
    double rho_last;           // candidates per row of the last committed band that had rows
    unsigned int pass, pad;    // pass edges of the unit's sub-bucket in this round (filled by k_round_report)
};
__global__ void k_truncate(const UnitDesc* __restrict__ units, int n_units, const RowTask* __restrict__ tasks, const unsigned long long* __restrict__ unit_cand,
                           const SubBucket* __restrict__ subs, int32_t* __restrict__ qcut, unsigned int* __restrict__ sub_pass, UnitOut* __restrict__ out,
                           Counters* __restrict__ ctr) {
    const int lane = threadIdx.x & 31;
    const int ui = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (ui >= n_units) return;
    const UnitDesc u = units[ui];
    const int n_bands = u.nt * S1_WARPS;
    double carry_cand = 0.0, carry_aj = 0.0, cand_out = 0.0, rho_last = -1.0;
    bool jumped = false;
    int r1 = u.r1, cut = 0;
    constexpr int PF = 4;   // chunks of 32 bands whose loads are in flight together (a 33k-row unit has 2100 bands: the walk is
                            // bound by the latency of these dependent loads, 92 us before the batching)
    for (int base0 = 0; base0 < n_bands && !cut; base0 += 32 * PF) {
      double cnt_pf[PF];
      int tr0_pf[PF], tr1_pf[PF];
#pragma unroll
      for (int j = 0; j < PF; j++) {
        const int bi = base0 + 32 * j + lane;
        cnt_pf[j] = 0.0;
        tr0_pf[j] = tr1_pf[j] = 0;   // rows of this band that belong to the unit (empty when the band lies outside [row_lo, row_hi))
        if (bi < n_bands) {
            const RowTask rt = tasks[u.t0 + bi / S1_WARPS];
            const int b = bi % S1_WARPS;
            tr0_pf[j] = rt.rb * TILE + max(rt.row_lo, 16 * b);
            tr1_pf[j] = rt.rb * TILE + min(rt.row_hi, 16 * b + 16);
            if (tr1_pf[j] > tr0_pf[j]) cnt_pf[j] = (double)unit_cand[(size_t)rt.tid * S1_WARPS + b];
        }
      }
#pragma unroll
      for (int j = 0; j < PF; j++) {
        if (cut || base0 + 32 * j >= n_bands) break;
        const double cnt = cnt_pf[j];
        const int tr0 = tr0_pf[j], tr1 = tr1_pf[j];
        const bool valid = tr1 > tr0;
        const bool dj = valid && (u.rho_ref < 0.0 || cnt / (double)(tr1 - tr0) > 8.0 * u.rho_ref + 16.0);
        const unsigned int jm = __ballot_sync(0xffffffffu, dj);
        const int J = jumped ? 0 : (jm ? __ffs(jm) - 1 : 32);   // first lane of this chunk that counts towards the strict allowance
        double incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        const double excl_at_J = (J < 32) ? __shfl_sync(0xffffffffu, incl - cnt, J & 31) : 0.0;
        const double aj = carry_aj + (lane >= J ? incl - excl_at_J : 0.0);
        const double cum = carry_cand + incl;
        const bool cross = valid && (cum > u.allow || (lane >= J && aj > u.strict)) && tr1 < u.r1;
        const unsigned int cm = __ballot_sync(0xffffffffu, cross);
        const unsigned int vm = __ballot_sync(0xffffffffu, valid);
        if (cm) {
            const int L = __ffs(cm) - 1;
            r1 = __shfl_sync(0xffffffffu, tr1, L);
            cand_out = __shfl_sync(0xffffffffu, cum, L);
            rho_last = __shfl_sync(0xffffffffu, cnt / (double)max(1, tr1 - tr0), L);
            cut = 1;
        } else {
            carry_cand = __shfl_sync(0xffffffffu, cum, 31);
            carry_aj = __shfl_sync(0xffffffffu, aj, 31);
            if (J < 32) jumped = true;
            cand_out = carry_cand;
            if (vm) {
                const int L = 31 - __clz(vm);   // last band of this chunk with rows
                rho_last = __shfl_sync(0xffffffffu, cnt / (double)max(1, tr1 - tr0), L);
            }
        }
      }
    }
    if (lane == 0) {
        qcut[u.sub] = cut ? (int32_t)(subs[u.sub].q0 + r1) : 0x7fffffff;
        sub_pass[u.sub] = 0u;
        out[ui].r1 = r1;
        out[ui].cut = cut;
        out[ui].cand = cand_out;
        out[ui].rho_last = rho_last;
        out[ui].pass = 0u;
        if (cut) ctr->any_cut = 1ull;
    }
}
__global__ void k_round_report(const UnitDesc* __restrict__ units, int n_units, const unsigned int* __restrict__ sub_pass, UnitOut* __restrict__ out) {
    const int ui = blockIdx.x * blockDim.x + threadIdx.x;
    if (ui < n_units) out[ui].pass = sub_pass[units[ui].sub];
}

// first launch of a run: every row its own class
__global__ void k_init_labels(int32_t* __restrict__ parent, int64_t n_rows, const int32_t* __restrict__ rowq, int64_t n_pos, int32_t* __restrict__ labq,
                              int32_t* __restrict__ blk_lab, const int32_t* __restrict__ blk_sub, const SubBucket* __restrict__ subs, int n_blocks) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_rows) parent[i] = (int32_t)i;
    if (i < n_pos) labq[i] = rowq[i];                      // padding positions carry -1 already
    if (i < n_blocks) {
        const SubBucket sb = subs[blk_sub[i]];
        const int rb = (int)(i - sb.blk0);
        blk_lab[i] = (min(sb.n, rb * TILE + TILE) - rb * TILE == 1) ? rowq[sb.q0 + rb * TILE] : -1;   // distinct rows: never uniform unless alone
    }
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
// Finalize (every count stays on the device: no host read-back until the results are downloaded)
// ------------------------------------------------------------------------------------------------
// Merge of the gathered partial forests, first pass: an edge of a sub-bucket that ONE rank joined as a whole is final — it goes
// straight to the merged forest and its endpoints are united; the edges of split sub-buckets (overlapping partial forests) are
// compacted into the pass-edge list for the Boruvka pass that follows.
__global__ void k_merge_split(const unsigned long long* __restrict__ w, const PassTuple* __restrict__ t, unsigned long long n, int32_t* __restrict__ parent,
                              unsigned long long* __restrict__ forest, PassTuple* __restrict__ forest_t, unsigned long long* __restrict__ pass_w,
                              PassTuple* __restrict__ pass_t, Counters* ctr) {
    const int lane = threadIdx.x & 31;
    for (unsigned long long i0 = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x - lane; i0 < n; i0 += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long i = i0 + lane;
        bool whole = false, split = false;
        unsigned long long e = 0;
        PassTuple pt;
        if (i < n) {
            e = w[i];
            pt = t[i];
            split = (pt.nrefs_err & PT_SPLIT) != 0;
            whole = !split;
        }
        const unsigned int mw = __ballot_sync(0xffffffffu, whole), ms = __ballot_sync(0xffffffffu, split);
        unsigned long long bw = 0, bs = 0;
        if (lane == 0) {
            if (mw) bw = atomicAdd(&ctr->n_forest, (unsigned long long)__popc(mw));
            if (ms) bs = atomicAdd(&ctr->n_pass, (unsigned long long)__popc(ms));
        }
        bw = __shfl_sync(0xffffffffu, bw, 0);
        bs = __shfl_sync(0xffffffffu, bs, 0);
        const unsigned int lt = (1u << lane) - 1u;
        if (whole) {
            const unsigned long long o = bw + __popc(mw & lt);
            forest[o] = e;
            forest_t[o] = pt;
            uf_union(parent, (int)(e >> 28), (int)(e & 0xfffffffull));
        } else if (split) {
            const unsigned long long o = bs + __popc(ms & lt);
            pass_w[o] = e;
            pass_t[o] = pt;
        }
    }
}
__global__ void k_init_parent(int32_t* __restrict__ parent, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) parent[i] = (int32_t)i;
}
// forest degree of every row (two-cell filter)
__global__ void k_forest_deg(const unsigned long long* __restrict__ forest, const unsigned long long* __restrict__ n_p, int32_t* __restrict__ deg) {
    const unsigned long long n = *n_p;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long e = forest[i];
        atomicAdd(&deg[(int)(e >> 28)], 1);
        atomicAdd(&deg[(int)(e & 0xfffffffull)], 1);
    }
}
// join.rs:120-178: an orbit of exactly two rows holding two cells in total loses its edge when
// cd > min(shares) / 2 (unless EASY).  Kept edges are compacted into sort keys (weight) + source index; parent[] is un-merged
// for the deleted ones.
__global__ void k_two_cell(DevIn in, DevParams P, const unsigned long long* __restrict__ forest, const unsigned long long* __restrict__ n_p,
                           const int32_t* __restrict__ deg, const PassTuple* __restrict__ tuples, int32_t* __restrict__ parent,
                           unsigned long long* __restrict__ keys, uint32_t* __restrict__ idx, Counters* ctr, const uint8_t* __restrict__ roles) {
    const unsigned long long n = *n_p;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long e = forest[i];
        const int a = (int)(e >> 28), b = (int)(e & 0xfffffffull);
        bool kp = true;
        // rows of a sub-bucket that was split over several contexts: this forest is partial, the caller filters after merging
        const bool partial = roles && (roles[a] & ROLE_SPLIT);
        if (!partial && deg[a] == 1 && deg[b] == 1) {
            const int ea = in.row_exact[a], eb = in.row_exact[b];
            const long long ncells = (in.ex_cell_off[ea + 1] - in.ex_cell_off[ea]) + (in.ex_cell_off[eb + 1] - in.ex_cell_off[eb]);
            if (ncells == 2) {
                const PassTuple t = tuples[i];
                const int nrefs = (int)(t.nrefs_err & 0xff);
                const int cd = (int)(t.cd12 & 0xffff) + (int)(t.cd12 >> 16);
                const int min_sh = (nrefs == 2) ? min(t.sh0, t.sh1) : t.sh0;
                if (cd > min_sh / 2 && !P.easy) {
                    kp = false;
                    parent[a] = a;
                    parent[b] = b;
                    atomicAdd(&ctr->n_two_cell, 1ull);
                }
            }
        }
        if (kp) {
            const unsigned long long o = atomicAdd(&ctr->n_final, 1ull);
            keys[o] = e;
            idx[o] = (uint32_t)i;
        }
    }
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
// rows per exact subclonotype, then the (exact << 32 | row) list of the exacts that own two or more rows: all the host needs
// for the clonotype_id merge of finish_join (join2.rs:69-84) — about 2 % of the rows instead of a pass over all of them
__global__ void k_exact_count(DevIn in, int32_t* __restrict__ count) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < in.n_rows) atomicAdd(&count[in.row_exact[k]], 1);
}
__global__ void k_multi_rows(DevIn in, const int32_t* __restrict__ count, unsigned long long* __restrict__ out, Counters* ctr) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool multi = k < in.n_rows && count[in.row_exact[k]] >= 2;
    const unsigned int m = __ballot_sync(0xffffffffu, multi);
    if (!m) return;
    const int lane = threadIdx.x & 31;
    unsigned long long b = 0;
    if (lane == __ffs(m) - 1) b = atomicAdd(&ctr->n_multi, (unsigned long long)__popc(m));
    b = __shfl_sync(0xffffffffu, b, __ffs(m) - 1);
    if (multi) out[b + __popc(m & ((1u << lane) - 1u))] = ((unsigned long long)(uint32_t)in.row_exact[k] << 32) | (unsigned long long)(uint32_t)k;
}
__global__ void k_labels(const int32_t* __restrict__ parent, int64_t n, int32_t* __restrict__ labels) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) labels[k] = uf_find_ro(parent, (int)k);
}
__global__ void k_fill_i32(int32_t* p, int64_t n, int32_t v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
struct FinalOut {
    int32_t *k1, *k2, *cd, *nrefs, *shares, *indeps, *k, *d, *n;
    double *p1, *mult, *score;
    uint8_t* err;
};
// edges in raw_joins order + the PotentialJoin members rebuilt from the kept PassTuple (join_one.rs:444-447, 499-544: same
// formulas and tables as stage 2b, which accepted the edge with exactly these values; p1 is the value it used)
__global__ void k_final_gather(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ idx, const unsigned long long* __restrict__ n_p,
                               const PassTuple* __restrict__ tuples, DevIn in, const double* __restrict__ powtab,
                               const int32_t* __restrict__ pow_off, FinalOut o) {
    const unsigned long long n = *n_p;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long e = keys[i];
        const PassTuple t = tuples[idx[i]];
        const int k1 = (int)(e >> 28), k2 = (int)(e & 0xfffffffull);
        // the edge's sub-bucket facts straight from its first row (rows of other shards are not packed on this context)
        const int n3 = 3 * (in.row_len[2 * (int64_t)k1] + in.row_len[2 * (int64_t)k1 + 1]);               // join_one.rs:499
        const int ch = in.row_cdr3_len[2 * (int64_t)k1], cl = in.row_cdr3_len[2 * (int64_t)k1 + 1];
        const int cd1 = (int)(t.cd12 & 0xffff), cd2 = (int)(t.cd12 >> 16);
        const int nrefs = (int)(t.nrefs_err & 0xff);
        const int min_sh = (nrefs == 2) ? min(t.sh0, t.sh1) : t.sh0;
        const int min_in = (nrefs == 2) ? min(t.in0, t.in1) : t.in0;
        const int kk = min_in + 2 * min_sh;
        const double p1 = t.p1;
        const double mult = __dmul_rn(powtab[pow_off[ch] + cd1], powtab[pow_off[cl] + cd2]);
        o.k1[i] = k1;
        o.k2[i] = k2;
        o.cd[i] = cd1 + cd2;
        o.nrefs[i] = nrefs;
        o.shares[2 * i] = t.sh0;
        o.shares[2 * i + 1] = t.sh1;
        o.indeps[2 * i] = t.in0;
        o.indeps[2 * i + 1] = t.in1;
        o.k[i] = kk;
        o.d[i] = min_sh;
        o.n[i] = n3;
        o.p1[i] = p1;
        o.mult[i] = mult;
        o.score[i] = __dmul_rn(p1, mult);
        o.err[i] = (uint8_t)(t.nrefs_err >> 8);
    }
}

// ------------------------------------------------------------------------------------------------
// Integer-pipe peak microbenchmarks (roofline denominators, SURVEY §8 d)
// ------------------------------------------------------------------------------------------------
// All operands are per-lane values (anything block-uniform would be moved to the uniform datapath and
// inflate the figure); 8 independent chains cover the pipe latency.
__global__ void k_peak_popc(uint32_t* out, int iters) {
    const uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    uint32_t a0 = t, a1 = t ^ 0x9e3779b9u, a2 = t + 0x7f4a7c15u, a3 = t * 3u, a4 = t ^ 0x55555555u, a5 = t + 0x33333333u, a6 = ~t, a7 = t * 5u;
    uint32_t s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0, s5 = 0, s6 = 0, s7 = 0;
#pragma unroll 4
    for (int i = 0; i < iters; i++) {
        s0 += __popc(a0); s1 += __popc(a1); s2 += __popc(a2); s3 += __popc(a3);
        s4 += __popc(a4); s5 += __popc(a5); s6 += __popc(a6); s7 += __popc(a7);
        a0 += s4; a1 += s5; a2 += s6; a3 += s7; a4 += s0; a5 += s1; a6 += s2; a7 += s3;   // data-dependent: nothing can be hoisted
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s0 + s1 + s2 + s3 + s4 + s5 + s6 + s7;
}
__global__ void k_peak_lop3(uint32_t* out, int iters) {
    const uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    uint32_t a0 = t, a1 = t ^ 0x9e3779b9u, a2 = t + 0x7f4a7c15u, a3 = t * 3u, a4 = t ^ 0x55555555u, a5 = t + 0x33333333u, a6 = ~t, a7 = t * 5u;
    const uint32_t m0 = t * 7u, m1 = t * 11u;
#pragma unroll 4
    for (int i = 0; i < iters; i++) {
        // 8 three-input logic ops, one LOP3 each (two interleaved groups of 4 independent chains)
        a0 = (a0 ^ m0) | a4; a1 = (a1 & m0) ^ a5; a2 = (a2 | m0) ^ a6; a3 = (a3 ^ m1) & a7;
        a4 = (a4 ^ m1) | a0; a5 = (a5 & m1) ^ a1; a6 = (a6 | m1) ^ a2; a7 = (a7 ^ m0) & a3;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
}

}  // namespace ejb
