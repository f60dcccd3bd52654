// ejb_p1.cuh — the per-join p1 memo, its term groups and the double-double evaluation kernels (join_one.rs:22-43).
#pragma once
#include "ejb_types.cuh"

namespace ejb {

// ------------------------------------------------------------------------------------------------
// p1 cache: open-addressing hash (n, k, d) -> f64, filled by k_p1_sum.
// ------------------------------------------------------------------------------------------------
struct P1Cache {
    unsigned long long* keys;  // ~0 = empty
    double* vals;
    uint32_t* newslots;        // slots inserted since the last evaluation (enqueue_p1)
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
// Probe sequences are bounded by the table size: a full table raises Counters::p1_full instead of spinning (the host keeps the
// load factor below 1/4 by growing the table between rounds, so this never triggers in practice).
__device__ __forceinline__ void p1_request(const P1Cache& c, int n, int k, int d, Counters* ctr) {
    const unsigned long long key = p1_key(n, k, d);
    uint32_t h = p1_hash(key) & c.cap_mask;
    for (uint32_t probes = 0; probes <= c.cap_mask; probes++) {
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
    ctr->p1_full = 1ull;
}
__device__ __forceinline__ double p1_lookup(const P1Cache& c, int n, int k, int d) {
    const unsigned long long key = p1_key(n, k, d);
    uint32_t h = p1_hash(key) & c.cap_mask;
    for (uint32_t probes = 0; probes <= c.cap_mask; probes++) {
        const unsigned long long cur = c.keys[h];
        if (cur == key) return c.vals[h];
        if (cur == ~0ull) break;
        h = (h + 1) & c.cap_mask;
    }
    return __longlong_as_double(0x7ff8000000000000ll);  // NaN: must not happen (every survivor requested its key in stage 2a)
}
// grow: re-insert every (key, value) of the old table into a larger, empty one
__global__ void k_p1_rehash(const unsigned long long* __restrict__ old_keys, const double* __restrict__ old_vals, uint32_t old_cap, P1Cache dst) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= old_cap) return;
    const unsigned long long key = old_keys[i];
    if (key == ~0ull) return;
    uint32_t h = p1_hash(key) & dst.cap_mask;
    for (;;) {   // the new table is at most 1/4 full
        if (atomicCAS(&dst.keys[h], ~0ull, key) == ~0ull) {
            dst.vals[h] = old_vals[i];
            return;
        }
        h = (h + 1) & dst.cap_mask;
    }
}

// Quotient table: the factors (n-v+1)/(u-v+1) of join_one.rs:37 are double-double quotients of two small integers a / b with
// a - b = n - u constant along a term.  QT[delta][b] = dd(delta + b) / dd(b), evaluated by the same dd_div, is a constant
// table like the Stirling ratios (built once in ejb_create, 134 MB): a term then costs one 16-byte load + one dd product per
// factor instead of a three-quotient division (110 fp64 instructions).  A term walks b = u..1 on one row: contiguous loads.
constexpr int QT_DELTA = 8192;   // rows: n - u
constexpr int QT_B = 1024;       // columns: b = u - v + 1
__global__ void k_quot_table(double2* __restrict__ qt) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)QT_DELTA * QT_B) return;
    const int delta = (int)(i / QT_B), b = (int)(i % QT_B);
    dd_t q = dd_make(0.0);
    if (b >= 1) q = dd_div(dd_make((double)(delta + b)), dd_make((double)b));
    qt[i] = make_double2(q.hi, q.lo);
}
__device__ __forceinline__ dd_t p1_quot(const double2* __restrict__ qt, int n, int u, int v) {
    const int delta = n - u, b = u - v + 1;
    if (qt != nullptr && delta >= 0 && delta < QT_DELTA && b < QT_B) {
        const double2 q = __ldg(qt + (size_t)delta * QT_B + b);
        return dd_make(q.x, q.y);
    }
    return dd_div(dd_make((double)(n - v + 1)), dd_make((double)b));
}

// p(n, k, d) = 1 - sum_{u = k-d+1 .. k} T(n, k, u),  T(n, k, u) = sr[k][u] * (u/n)^k * prod_{v=1..u} (n-v+1)/(u-v+1)
// (p_at_most_m_distinct_in_sample_of_x_from_n_double, join_one.rs:22-43, x = k, m = k - d).  One term, with the reference's
// operation order: start from the Stirling ratio, k products by u/n, then the u quotient factors in increasing v.
__device__ __forceinline__ dd_t p1_term(const dd_t* __restrict__ sr, const double2* __restrict__ qt, int n, int k, int u) {
    dd_t z = sr[(size_t)k * (k + 1) / 2 + u];
    const dd_t f = dd_div(dd_make((double)u), dd_make((double)n));
    for (int i = 0; i < k; i++) z = dd_mul(z, f);
    // the quotients do not depend on z: the next four are loaded while the chain consumes the current four
    constexpr int QB = 4;
    dd_t q[QB], qn[QB];
    int v = 1;
    if (u >= QB) {
#pragma unroll
        for (int i = 0; i < QB; i++) q[i] = p1_quot(qt, n, u, v + i);
        for (; v + 2 * QB - 1 <= u; v += QB) {
#pragma unroll
            for (int i = 0; i < QB; i++) qn[i] = p1_quot(qt, n, u, v + QB + i);
#pragma unroll
            for (int i = 0; i < QB; i++) z = dd_mul(z, q[i]);
#pragma unroll
            for (int i = 0; i < QB; i++) q[i] = qn[i];
        }
#pragma unroll
        for (int i = 0; i < QB; i++) z = dd_mul(z, q[i]);
        v += QB;
    }
    for (; v <= u; v++) z = dd_mul(z, p1_quot(qt, n, u, v));
    return z;
}
// One warp, one key: lanes evaluate the terms, the subtraction runs in increasing u so that the double-double rounding
// sequence is the reference's (explicit triples: k_p1_batch, ejb_join_one_batch).
__device__ __forceinline__ double p1_eval_warp(const dd_t* __restrict__ sr, const double2* __restrict__ qt, int k, int d, int n, int lane) {
    const int x = k, m = k - d;
    dd_t p = dd_make(1.0);
    for (int u0 = m + 1; u0 <= x; u0 += 32) {
        const int u = u0 + lane;
        dd_t z = dd_make(0.0);
        if (u <= x) z = p1_term(sr, qt, n, k, u);
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

// ------------------------------------------------------------------------------------------------
// Term groups.  T(n, k, u) does not depend on d: every key (n, k, d) of one (n, k) GROUP sums a suffix u = k-d+1 .. k of the
// same term sequence, and a repertoire asks for many d per group (configs[1]: 2.7 requested terms per distinct term).  A group
// owns k consecutive entries of the term pool, entry j = the term u = k - j; `have` of them are evaluated.  Per round:
//   k_p1_groups  thread per new key: find / create the group, want = max(want, d), list the group once (stamp)
//   k_p1_expand  warp per listed group: list the missing terms [have, want) for evaluation, have = want
//   k_p1_terms   THREAD per listed term (neighbouring threads: neighbouring u of one group, equal chain lengths): p1_term
//   k_p1_sum     thread per new key: p = 1 - terms in increasing u (= decreasing j), clamp, value into the key cache
// Nothing here can fail: a key whose group could not be created, or whose terms could not be listed (a structure was full; the
// host grows it after the round), is evaluated by the direct path of k_p1_sum.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t P1_NOPOOL = 0xffffffffu;
struct P1Groups {
    unsigned long long* gkeys;   // (n << 12) | k, ~0 = empty
    uint32_t* goff;              // first pool entry of the group, P1_NOPOOL when the pool was full
    uint32_t* ghave;             // evaluated terms (j < have)
    uint32_t* gwant;             // largest d requested
    uint32_t* gstamp;            // round in which the group was last listed
    uint32_t* touched;           // groups listed this round
    uint2* desc;                 // (group slot, j) of the terms to evaluate this round
    double2* pool;
    uint32_t gmask, round;
    unsigned long long pool_cap, desc_cap, touched_cap;
};
__device__ __forceinline__ unsigned long long p1_gkey(int n, int k) { return ((unsigned long long)(uint32_t)n << 12) | (unsigned long long)(uint32_t)k; }
constexpr uint32_t P1_GPROBES = 128;   // bounded: a crowded table sends the key to the direct path
__device__ __forceinline__ uint32_t p1_group_find(const P1Groups& G, unsigned long long gkey) {
    uint32_t h = p1_hash(gkey) & G.gmask;
    for (uint32_t probes = 0; probes < P1_GPROBES; probes++) {
        const unsigned long long cur = G.gkeys[h];
        if (cur == gkey) return h;
        if (cur == ~0ull) break;
        h = (h + 1) & G.gmask;
    }
    return P1_NOPOOL;
}
__global__ void k_p1_groups(P1Cache c, P1Groups G, const unsigned long long* __restrict__ n_new, Counters* ctr) {
    const unsigned long long n = *n_new;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long key = c.keys[c.newslots[i]];
        const int d = (int)(key & 0xfff), k = (int)((key >> 12) & 0xfff);
        if (d == 0) continue;   // no term: p = 1
        const unsigned long long gkey = key >> 12;   // (n << 12) | k
        uint32_t h = p1_hash(gkey) & G.gmask, slot = P1_NOPOOL;
        for (uint32_t probes = 0; probes < P1_GPROBES; probes++) {
            unsigned long long cur = G.gkeys[h];
            if (cur == ~0ull) {
                cur = atomicCAS(&G.gkeys[h], ~0ull, gkey);
                if (cur == ~0ull) {   // created: k pool entries (ghave / gwant / gstamp are zero since the last reset)
                    const unsigned long long off = atomicAdd(&ctr->p1_pool_used, (unsigned long long)k);
                    G.goff[h] = (off + (unsigned long long)k <= G.pool_cap) ? (uint32_t)off : P1_NOPOOL;
                    atomicAdd(&ctr->p1_groups, 1ull);
                    slot = h;
                    break;
                }
            }
            if (cur == gkey) { slot = h; break; }
            h = (h + 1) & G.gmask;
        }
        if (slot == P1_NOPOOL) continue;
        atomicMax(&G.gwant[slot], (uint32_t)d);
        if (atomicExch(&G.gstamp[slot], G.round) != G.round) {
            const unsigned long long t = atomicAdd(&ctr->p1_touched, 1ull);
            if (t < G.touched_cap) G.touched[t] = slot;
        }
    }
}
__global__ void k_p1_expand(P1Groups G, Counters* ctr) {
    const int lane = threadIdx.x & 31;
    const unsigned long long nt = min(ctr->p1_touched, G.touched_cap);
    const unsigned long long nwarps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    for (unsigned long long w = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < nt; w += nwarps) {
        const uint32_t h = G.touched[w];
        const uint32_t have = G.ghave[h], want = G.gwant[h];
        if (G.goff[h] == P1_NOPOOL || want <= have) continue;
        const uint32_t cnt = want - have;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(&ctr->p1_desc, (unsigned long long)cnt);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base + cnt > G.desc_cap) {   // not listed: have stays, the keys take the direct path
            for (uint32_t i = lane; i < cnt && base + i < G.desc_cap; i += 32) G.desc[base + i] = make_uint2(P1_NOPOOL, 0u);
            continue;
        }
        for (uint32_t i = lane; i < cnt; i += 32) G.desc[base + i] = make_uint2(h, have + i);
        __syncwarp();
        if (lane == 0) G.ghave[h] = want;
    }
}
__global__ void __launch_bounds__(256, 3) k_p1_terms(P1Groups G, const dd_t* __restrict__ sr, const double2* __restrict__ qt, Counters* ctr) {
    const unsigned long long n = min(ctr->p1_desc, G.desc_cap);
    if (blockIdx.x == 0 && threadIdx.x == 0 && n) atomicAdd(&ctr->p1_terms, n);
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (unsigned long long)gridDim.x * blockDim.x) {
        const uint2 ds = G.desc[t];
        if (ds.x == P1_NOPOOL) continue;
        const unsigned long long gkey = G.gkeys[ds.x];
        const int k = (int)(gkey & 0xfff), nn = (int)(gkey >> 12);
        const dd_t z = p1_term(sr, qt, nn, k, k - (int)ds.y);
        G.pool[(size_t)G.goff[ds.x] + ds.y] = make_double2(z.hi, z.lo);
    }
}
__global__ void k_p1_sum(P1Cache c, P1Groups G, const dd_t* __restrict__ sr, const double2* __restrict__ qt, const unsigned long long* __restrict__ n_new,
                         Counters* ctr) {
    const unsigned long long n = *n_new;
    if (blockIdx.x == 0 && threadIdx.x == 0 && n) atomicAdd(&ctr->p1_total, n);
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t slot = c.newslots[i];
        const unsigned long long key = c.keys[slot];
        const int d = (int)(key & 0xfff), k = (int)((key >> 12) & 0xfff), nn = (int)(key >> 24);
        dd_t p = dd_make(1.0);
        if (d > 0) {
            const uint32_t h = p1_group_find(G, key >> 12);
            if (h != P1_NOPOOL && G.goff[h] != P1_NOPOOL && (uint32_t)d <= G.ghave[h]) {
                const double2* t = G.pool + G.goff[h];
                for (int j = d - 1; j >= 0; j--) {   // u = k - j increasing
                    const double2 z = t[j];
                    p = dd_sub(p, dd_make(z.x, z.y));
                }
            } else {
                atomicAdd(&ctr->p1_direct, 1ull);
                for (int u = k - d + 1; u <= k; u++) p = dd_sub(p, p1_term(sr, qt, nn, k, u));
            }
        }
        if (dd_lt(p, dd_make(0.0))) p = dd_make(0.0);
        c.vals[slot] = p.hi;
    }
}
// grow: re-insert every group into a larger, empty table (pool offsets stay)
__global__ void k_p1_groups_rehash(P1Groups src, uint32_t old_cap, P1Groups dst) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= old_cap) return;
    const unsigned long long gkey = src.gkeys[i];
    if (gkey == ~0ull) return;
    uint32_t h = p1_hash(gkey) & dst.gmask;
    for (;;) {   // the new table is at most 1/4 full
        if (atomicCAS(&dst.gkeys[h], ~0ull, gkey) == ~0ull) {
            dst.goff[h] = src.goff[i];
            dst.ghave[h] = src.ghave[i];
            dst.gwant[h] = src.gwant[i];
            dst.gstamp[h] = src.gstamp[i];
            return;
        }
        h = (h + 1) & dst.gmask;
    }
}

// stirling2_ratio_table_double (enclone_stuff/src/start.rs:50-99) on the device.  Triangular layout, row n at n(n+1)/2.
// The recurrence over n is sequential, the 3000 columns of a row are independent: ONE CTA walks n = 1..n_max with a barrier per
// row; thread t owns the columns k = t, t + 1024, t + 2048 and carries their z_k = ((k-1)/k)^(n-1) in registers, updated by
// the same multiplication sequence as the reference (powi at n = k + 1, one multiplication by (k-1)/k per later row), so that
// every entry has the bits of the host loop.  s[n][n] = n!/n^n is the product of j/n in increasing j (start.rs:71-80).
constexpr int SR_KPT = 3;
__global__ void __launch_bounds__(1024) k_sr_table(dd_t* s, int n_max) {
    const int t = threadIdx.x;
    dd_t k1k[SR_KPT], njn[SR_KPT], z[SR_KPT];
#pragma unroll
    for (int i = 0; i < SR_KPT; i++) {
        const int k = t + 1024 * i;
        k1k[i] = njn[i] = z[i] = dd_make(0.0);
        if (k >= 1 && k <= n_max) {
            k1k[i] = dd_div(dd_make((double)(k - 1)), dd_make((double)k));
            dd_t p = dd_make(1.0);
            for (int j = 1; j <= k; j++) p = dd_mul(p, dd_div(dd_make((double)j), dd_make((double)k)));
            njn[i] = p;
        }
    }
    if (t == 0) s[0] = dd_make(1.0);
    __syncthreads();
    for (int n = 1; n <= n_max; n++) {
        dd_t* row = s + (size_t)n * (n + 1) / 2;
        const dd_t* prev = s + (size_t)(n - 1) * n / 2;
#pragma unroll
        for (int i = 0; i < SR_KPT; i++) {
            const int k = t + 1024 * i;
            if (k == 0) {
                row[0] = dd_make(0.0);
            } else if (k < n) {
                z[i] = (n == k + 1) ? dd_powi(k1k[i], k) : dd_mul(z[i], k1k[i]);
                row[k] = dd_add(prev[k], dd_mul(prev[k - 1], z[i]));
            } else if (k == n) {
                row[k] = njn[i];
            }
        }
        __syncthreads();
    }
}

// p1 for explicit triples (parity tests): one warp per triple, the same p1_term as the join
__global__ void k_p1_batch(const dd_t* __restrict__ sr, const double2* __restrict__ qt, const int32_t* kk, const int32_t* ddv, const int32_t* nn, int64_t count, double* out) {
    const int lane = threadIdx.x & 31;
    const int64_t wi = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (wi >= count) return;
    const double v = p1_eval_warp(sr, qt, kk[wi], ddv[wi], nn[wi], lane);
    if (lane == 0) out[wi] = v;
}


}  // namespace ejb
