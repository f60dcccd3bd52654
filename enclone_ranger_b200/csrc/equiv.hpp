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
    void init(int32_t n) {
        for (int32_t i = 0; i < n; i++) {
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

}  // namespace ejbhost
