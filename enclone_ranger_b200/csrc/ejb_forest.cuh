// ejb_forest.cuh — lexicographic spanning forest (Boruvka), finalize kernels, pipe-peak microbenchmarks.
#pragma once
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
                           const EdgeTuple* __restrict__ tuples, int32_t* __restrict__ parent, uint8_t* __restrict__ keep, Counters* ctr,
                           const uint8_t* __restrict__ roles) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long e = forest[i];
    const int a = (int)(e >> 28), b = (int)(e & 0xfffffffull);
    int kp = 1;
    // rows of a sub-bucket that was split over several contexts: this forest is partial, the caller filters after merging
    const bool partial = roles && (roles[a] & ROLE_SPLIT);
    if (!partial && deg[a] == 1 && deg[b] == 1) {
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
