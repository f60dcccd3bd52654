// equiv.hpp — host-side mirror of the reference's union-find OUTPUT TYPE (product code).
//
// join_exacts returns an `EquivRel` (equiv/src/lib.rs:34-138) whose internal state — class ids
// y[], circular orbit lists x[], sizes z[] — is observable downstream (filter_umi.rs:45-51,
// disintegrate.rs:106-121).  It depends on the ORDER of the join() calls, so the library replays
// the device-computed edges sequentially, exactly like finish_join (join2.rs:45-54, 69-84), and
// hands x/y/z back for EquivRel::from_raw (equiv/src/lib.rs:57).
#pragma once
#include <algorithm>
#include <cstdint>
#include <thread>
#include <utility>
#include <vector>

namespace ejbhost {

struct EquivRel {
    // the state lives in caller-provided arrays (the library's pinned result arena): no allocation, no copy-out
    int32_t *x, *y, *z;
    std::vector<int32_t> own;   // backing store when constructed without external arrays
    EquivRel(int32_t n, int32_t* xs, int32_t* ys, int32_t* zs) : x(xs), y(ys), z(zs) { init(n); }
    explicit EquivRel(int32_t n) : own(3 * (size_t)n) {
        x = own.data();
        y = x + n;
        z = y + n;
        init(n);
    }
    EquivRel(int32_t* xs, int32_t* ys, int32_t* zs) : x(xs), y(ys), z(zs) {}   // arrays initialised by the caller (init_range)
    void init(int32_t n) { init_range(0, n); }
    void init_range(int32_t lo, int32_t hi) {
        for (int32_t i = lo; i < hi; i++) {
            x[i] = y[i] = i;
            z[i] = 1;
        }
    }
    int32_t class_id(int32_t a) const { return y[a]; }
    int32_t orbit_size(int32_t a) const { return z[y[a]]; }
    // relabels the smaller orbit; on a tie b's orbit takes a's class id (equiv/src/lib.rs:64-93)
    void join(int32_t a, int32_t b) {
        if (y[a] == y[b]) return;
        if (orbit_size(a) < orbit_size(b)) std::swap(a, b);
        const int32_t total = orbit_size(a) + orbit_size(b);
        std::swap(x[a], x[b]);
        const int32_t cls = y[a];
        for (int32_t n = x[a]; y[n] != cls; n = x[n]) y[n] = cls;
        z[y[b]] = total;
    }
};

// join2.rs:69-84: sort (clonotype_id, class id) and join consecutive entries of one clonotype_id;
// note that the joined values are CLASS IDS used as element ids (join2.rs:81).
// Only clonotype_ids owning two or more rows can cause a join, so the sort is restricted to those
// rows (same join sequence as sorting everything: ids ascending, class ids ascending inside an id).
inline void finish_join_merge(EquivRel& eq, const int32_t* row_exact, int64_t n_rows, int64_t n_exacts = -1) {
    if (n_exacts < 0) {
        int32_t max_id = -1;
        for (int64_t i = 0; i < n_rows; i++) max_id = std::max(max_id, row_exact[i]);
        n_exacts = (int64_t)max_id + 1;
    }
    std::vector<uint8_t> cnt((size_t)n_exacts + 1, 0);   // saturating at 2: all that matters
    for (int64_t i = 0; i < n_rows; i++) {
        uint8_t& c = cnt[(size_t)row_exact[i]];
        if (c < 2) c++;
    }
    std::vector<std::pair<int32_t, int32_t>> ox;
    for (int64_t i = 0; i < n_rows; i++)
        if (cnt[(size_t)row_exact[i]] >= 2) ox.emplace_back(row_exact[i], eq.class_id((int32_t)i));
    std::sort(ox.begin(), ox.end());
    size_t i = 0;
    while (i < ox.size()) {
        size_t j = i + 1;
        while (j < ox.size() && ox[j].first == ox[i].first) j++;
        for (size_t k = i; k + 1 < j; k++) eq.join(ox[k].second, ox[k + 1].second);
        i = j;
    }
}

// The same merge from the device-built list of the rows whose clonotype_id owns two or more rows, as (id << 32 | row) sorted
// ascending: identical join sequence (ids ascending, class ids ascending inside an id), O(list) instead of O(rows).
inline void finish_join_merge_list(EquivRel& eq, const uint64_t* multi, int64_t n_multi) {
    // the class ids are those BEFORE the first of these joins (join2.rs:70-74 collects them all, then joins)
    std::vector<int32_t> cls((size_t)n_multi);
    for (int64_t k = 0; k < n_multi; k++) cls[(size_t)k] = eq.class_id((int32_t)(multi[k] & 0xffffffffu));
    int64_t i = 0;
    while (i < n_multi) {
        int64_t j = i + 1;
        while (j < n_multi && (multi[j] >> 32) == (multi[i] >> 32)) j++;
        std::sort(cls.begin() + i, cls.begin() + j);
        for (int64_t k = i; k + 1 < j; k++) eq.join(cls[(size_t)k], cls[(size_t)k + 1]);
        i = j;
    }
}

// Parallel replay: edges never cross lens buckets and the rows of a bucket are contiguous, so edge ranges cut at bucket
// boundaries touch disjoint parts of x / y / z and can be replayed by different threads; the state is the sequential one.
inline void replay_edges(EquivRel& eq, const int32_t* k1, const int32_t* k2, int64_t n_edges);
// Replay of an ordered edge list (join2.rs:45-54).  The walk is bound by the latency of the scattered x / y / z reads, so the
// entries of the edges a few steps ahead are prefetched.
inline void replay_edges(EquivRel& eq, const int32_t* k1, const int32_t* k2, int64_t n_edges) {
    constexpr int64_t AHEAD = 12;
    for (int64_t i = 0; i < n_edges; i++) {
        if (i + AHEAD < n_edges) {
            const int32_t a = k1[i + AHEAD], b = k2[i + AHEAD];
            __builtin_prefetch(&eq.y[a]);
            __builtin_prefetch(&eq.y[b]);
            __builtin_prefetch(&eq.x[a], 1);
            __builtin_prefetch(&eq.x[b], 1);
        }
        eq.join(k1[i], k2[i]);
    }
}

inline void replay_edges_parallel(EquivRel& eq, int64_t n_rows, const int32_t* k1, const int32_t* k2, int64_t n_edges, const int32_t* bucket_start,
                                  int64_t n_buckets, int threads) {
    if (threads <= 1 || n_buckets < 2 || n_rows < 400000) {
        eq.init((int32_t)n_rows);
        replay_edges(eq, k1, k2, n_edges);
        return;
    }
    // thread t owns the rows [cut[t], cut[t+1]) (cut at bucket starts, balanced by edge count) and the edges whose k1 lies there
    std::vector<int64_t> row_cut((size_t)threads + 1, 0), edge_cut((size_t)threads + 1, 0);
    row_cut[(size_t)threads] = n_rows;
    edge_cut[(size_t)threads] = n_edges;
    for (int t = 1; t < threads; t++) {
        int64_t r;
        if (n_edges > 0) {
            const int64_t e = n_edges * t / threads;
            const int32_t row = k1[e];
            const int32_t* b = std::upper_bound(bucket_start, bucket_start + n_buckets, row);   // first bucket starting after `row`
            r = (b == bucket_start) ? 0 : *(b - 1);
        } else {
            const int32_t* b = std::upper_bound(bucket_start, bucket_start + n_buckets, (int32_t)(n_rows * t / threads));
            r = (b == bucket_start) ? 0 : *(b - 1);
        }
        r = std::max<int64_t>(r, row_cut[(size_t)t - 1]);
        row_cut[(size_t)t] = r;
        edge_cut[(size_t)t] = std::lower_bound(k1, k1 + n_edges, (int32_t)r) - k1;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++)
        th.emplace_back([&, t] {
            eq.init_range((int32_t)row_cut[(size_t)t], (int32_t)row_cut[(size_t)t + 1]);
            replay_edges(eq, k1 + edge_cut[(size_t)t], k2 + edge_cut[(size_t)t], edge_cut[(size_t)t + 1] - edge_cut[(size_t)t]);
        });
    for (auto& x : th) x.join();
}

}  // namespace ejbhost
