// ejb_p1.cuh — p1 cache and the double-double evaluation kernel (join_one.rs:22-43).
#pragma once
#include "ejb_types.cuh"

namespace ejb {

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
__global__ void k_p1(P1Cache c, const dd_t* __restrict__ sr, const unsigned long long* n_new, Counters* ctr) {
    const int lane = threadIdx.x & 31;
    const unsigned long long warp0 = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned long long nwarps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    const unsigned long long n = *n_new;
    if (blockIdx.x == 0 && threadIdx.x == 0 && n) atomicAdd(&ctr->p1_total, n);
    for (unsigned long long wi = warp0; wi < n; wi += nwarps) {
        const uint32_t slot = c.newslots[wi];
        const unsigned long long key = c.keys[slot];
        const int d = (int)(key & 0xfff), k = (int)((key >> 12) & 0xfff), nn = (int)(key >> 24);
        const double v = p1_eval_warp(sr, k, d, nn, lane);
        if (lane == 0) c.vals[slot] = v;
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

// p1 for explicit triples (parity tests): one warp per triple, same code path as k_p1
__global__ void k_p1_batch(const dd_t* __restrict__ sr, const int32_t* kk, const int32_t* ddv, const int32_t* nn, int64_t count, double* out) {
    const int lane = threadIdx.x & 31;
    const int64_t wi = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (wi >= count) return;
    const double v = p1_eval_warp(sr, kk[wi], ddv[wi], nn[wi], lane);
    if (lane == 0) out[wi] = v;
}


}  // namespace ejb
