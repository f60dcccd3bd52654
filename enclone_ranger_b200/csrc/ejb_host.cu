// ejb_host.cu — host orchestration and C ABI of the B200 clonotype join.
//
// Replaces, behind include/enclone_join_b200.h, the reference path
//   join_exacts (enclone/src/join.rs:31-590) -> join_core (join_core.rs:11-51) ->
//   join_one (enclone_core/src/join_one.rs:66-887) -> finish_join (join2.rs:36-87).
// There is no CPU fallback: every entry point fails with EJB_ERR_NO_DEVICE / EJB_ERR_CUDA when the
// device path cannot run.  See DESIGN.md for the round scheduler and the forest argument.
#include <cub/cub.cuh>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../../include/enclone_join_b200.h"
#include "ejb_forest.cuh"
#include "ejb_pack.cuh"
#include "ejb_stage1.cuh"
#include "ejb_stage2.cuh"
#include "equiv.hpp"

using namespace ejb;

namespace {

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), cap(o.cap) { o.p = nullptr; o.cap = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) {
            release();
            p = o.p;
            cap = o.cap;
            o.p = nullptr;
            o.cap = 0;
        }
        return *this;
    }
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap && p) return cudaSuccess;
        release();
        size_t want = std::max<size_t>(bytes, 256);
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want; else p = nullptr;
        return e;
    }
    template <class T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

struct PinBuf {   // page-locked host scratch, grow-only
    void* p = nullptr;
    size_t cap = 0;
    ~PinBuf() { if (p) cudaFreeHost(p); }
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap && p) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        const size_t want = std::max<size_t>(bytes + bytes / 2, 4096);
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want; else p = nullptr;
        return e;
    }
    template <class T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

// std::vector storage in page-locked memory: the sub-bucket table is built in place and uploaded by a truly asynchronous copy
template <class T>
struct PinAlloc {
    using value_type = T;
    PinAlloc() = default;
    template <class U> PinAlloc(const PinAlloc<U>&) {}
    T* allocate(size_t n) {
        void* p = nullptr;
        if (cudaMallocHost(&p, n * sizeof(T)) != cudaSuccess) throw std::bad_alloc();
        return static_cast<T*>(p);
    }
    void deallocate(T* p, size_t) { cudaFreeHost(p); }
    template <class U> bool operator==(const PinAlloc<U>&) const { return true; }
    template <class U> bool operator!=(const PinAlloc<U>&) const { return false; }
};
using SubVec = std::vector<SubBucket, PinAlloc<SubBucket>>;

struct Unit {  // rows [r0, r1) of one sub-bucket, processed against all columns to their right
    int32_t sub, r0, r1;
    double allow;      // candidates after which the unit is truncated (at a row-range boundary) once stage 1 has counted them
    double strict;     // ... allowance that applies from the first row range whose density jumps above 8x rho_ref
    double rho_ref;    // candidates per row of the sub-bucket's last committed row range (< 0: unknown)
    double cand, pass; // measured: candidates of the committed rows, pass edges
    double rho_last;   // measured: candidates per row of the last committed 16-row band (< 0: none)
};

// Developer / test knobs: environment read once in ejb_create, overridable with ejb_set_option.
struct Options {
    long long first_unit = 2, growth = 6;    // unit schedule (round 2, with per-band cuts: configs[4] 12 -> 84 ms, 6 -> 72 ms; configs[1] indifferent)
    int s1_mode = S1_RUNI | S1_TAGS;
    int no_tags = 0;
    int trace = 0;
    int thin = 16;             // stage 1: units of at most this many rows go to k_stage1_thin (0: always k_stage1; measured: 2- and 12-row
                               // units 3-4x faster than the tile kernel, 72-row units 25 % slower)
    int quot_table = 1;        // p1: quotients from the constant table (0: three-quotient divisions, A/B and parity knob)
    int sr_host = 0;                         // build the Stirling table on the host (A/B against the device kernel)
    long long strict_div = 2048;             // strict allowance of a unit = cand_cap / strict_div candidates
    int early_return = 0;                    // stage-1 CTAs return when the previous block of their unit is certain to be cut (measured: no gain)
};

}  // namespace

struct ejb_ctx {
    int device = 0;
    int sm_count = 148;
    int coop_grid = 148;        // co-resident CTAs of k_forest_round
    cudaStream_t stream = nullptr;
    // Large host->device copies are spread over a few copy streams: on this platform one cudaMemcpyAsync from pinned memory
    // reaches ~40 GB/s at 16 MB and ~55 GB/s at 512 MB, i.e. every copy pays ~0.1 ms of setup that concurrent copies overlap.
    static constexpr int N_UP = 4;
    cudaStream_t up_stream[N_UP] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t up_event[N_UP] = {nullptr, nullptr, nullptr, nullptr};
    int up_next = 0;
    std::string err;
    Options opt;
    double ms_create = 0.0;
    int shard_rank = 0, shard_world = 1;
    unsigned long long cand_cap = 1ull << 27;
    bool cap_explicit = false;
    unsigned long long pairs_eval_seen = 0;
    double cand_rate = 0.0;   // largest candidates/pairs ratio seen in a completed round of this run
    unsigned long long pair_budget = 1ull << 32;
    uint32_t p1_cap = 1u << 22;
    bool have_input = false, have_run = false;

    // persistent
    // p1 term groups (ejb_p1.cuh): group hash, term pool, per-round lists
    DevBuf d_gkeys, d_goff, d_ghave, d_gwant, d_gstamp, d_gtouched, d_gdesc, d_gpool;
    uint32_t g_cap = 1u << 20;                    // group slots
    unsigned long long pool_cap = 1ull << 25;     // term pool entries (16 B)
    unsigned long long desc_cap = 1ull << 24;     // terms listed per round
    uint32_t p1_round = 0;
    unsigned long long p1_groups = 0, p1_pool_used = 0;   // host tallies (from the round reports)
    DevBuf d_slowrows;   // rows the 8-lane bucketer left to the generic kernel
    DevBuf d_qt;   // quotient table (ejb_p1.cuh), nullptr-safe: option "quot_table" 0 evaluates the divisions
    const double2* qt() const { return opt.quot_table ? reinterpret_cast<const double2*>(d_qt.p) : nullptr; }
    DevBuf d_sr, d_p1_keys, d_p1_vals, d_p1_new, d_err, d_bar;
    unsigned long long p1_count = 0;   // keys in the p1 cache (host tally of the per-round insert counts)
    char* h_res = nullptr;      // pinned result arena (grow-only); ejb_result arrays point into it
    size_t h_res_cap = 0;

    // uploaded input
    ejb_params params;
    DevIn din;
    std::vector<DevBuf> in_bufs;
    size_t in_used = 0;
    int64_t h2d_bytes = 0;
    double ms_h2d = 0;
    int64_t n_canon_segs = 0;   // distinct seg contents (canonical ids are 0 .. n_canon_segs-1)
    const char* arena_base = nullptr;   // ejb_set_input_arena: one host allocation holding (most of) the input arrays
    size_t arena_bytes = 0;
    size_t arena_slice = 0;     // bytes per slice of the last sliced upload (0: whole arena uploaded)
    int arena_parts = 1;
    DevBuf d_arena;
    DevBuf d_roles, d_subroles; // ejb_set_row_roles (per canonical row), per-sub-bucket summary
    std::vector<uint8_t> roles; // host copy; in force for every run until cleared
    bool have_roles = false;    // roles are in force for the current run

    // derived (per run)
    DevBuf d_head, d_scan, d_keys, d_keys2, d_vals, d_vals2, d_cubtmp, d_subhead, d_subincl, d_subinfo;
    DevBuf d_subs, d_cdr3w, d_rowstore, d_m_row, d_m_sub, d_aux, d_packoff, d_packsub, d_bstart;
    DevBuf d_donor, d_bckeys, d_bckeys2, d_powtab, d_powoff, d_labq, d_blklab, d_blksub, d_parent, d_best;
    DevBuf d_qcut, d_subpass, d_segpl, d_seglen, d_memo, d_tagq, d_deadq, d_donq, d_flagq, d_subdom;
    // per-round report: [Counters | UnitOut x units | candidate count x tasks], read back with ONE copy per round
    DevBuf d_round;
    PinBuf h_round, h_prep, h_stage, h_tab;
    size_t round_units_cap = 0, round_tasks_cap = 0, cur_round_tasks = 0;
    double cur_strict = 0.0;    // strict allowance of the units of the round being queued
    size_t cur_cta_off = 0;     // offset of the current launch in the CTA -> task map
    bool tags_on = false;      // reference-set tags in use (every canonical seg id < 65535)
    struct TaskMeta { int32_t unit, r0, r1; };
    std::vector<TaskMeta> task_meta;
    std::vector<uint8_t> h_taskskip;              // ... tasks of the last round that returned early (not counted)
    std::vector<unsigned long long> h_unitcand;   // candidates of each task (row range) of the last round, valid even when it overflowed
    DevBuf d_mergew, d_merget;   // gathered partial forests (multi-GPU merge)
    DevBuf d_ctatask;            // CTA -> task map of the stage-1 launches of a round
    DevBuf d_tasks, d_units, d_cand, d_slow, d_slowslot, d_surv, d_pass, d_passt, d_roots, d_forest, d_forestt;
    DevBuf d_deg, d_first, d_labels, d_excount, d_multi, d_multi2, d_fkeys, d_fkeys2, d_fidx, d_fidx2;
    DevBuf d_o_k1, d_o_k2, d_o_cd, d_o_nrefs, d_o_sh, d_o_in, d_o_k, d_o_d, d_o_n, d_o_p1, d_o_mult, d_o_score, d_o_err;
    // pair-batch mode (ejb_join_one_batch)
    DevBuf d_bpairs, d_bslots, d_btuples, d_bk1, d_bk2, d_qofrow;
    SubVec subs;
    std::map<int, std::vector<double>> pow_rows;   // mult_pow ^ (cdr3_normal_len * cd / n) per CDR3 length n, kept across joins
    double pow_base = -1.0, pow_norm = -1.0;
    int64_t n_active = 0, n_pos = 0, n_blocks = 0, n_final = 0, n_buckets = 0, n_multi = 0, n_part = 0;
    bool have_part = false;     // a sharded run left its partial forest on the device (ejb_part_forest)
    unsigned long long forest_cap = 0;
    bool prepared = false;      // the packed row store of the uploaded input is resident (ejb_run got at least that far)
    ejb_stats stats;

    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> s2a_events, p1_events;
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    cudaEvent_t ev() {
        if (ev_used == ev_pool.size()) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            ev_pool.push_back(e);
        }
        return ev_pool[ev_used++];
    }
    Counters* d_ctr() const { return d_round.as<Counters>(); }

    // multi-GPU (ejb_create with n_devices > 1): this context drives device_ids[0]; one peer context per further device
    std::vector<ejb_ctx*> peers;
    struct MultiState;
    MultiState* multi = nullptr;
};

namespace {

int fail(ejb_ctx* c, int code, const char* fmt, ...) {
    char buf[768];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    return code;
}

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t _e = (call);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return fail(c, _e == cudaErrorMemoryAllocation ? EJB_ERR_OOM : EJB_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, \
                        cudaGetErrorString(_e));                                                         \
    } while (0)

inline unsigned int nblk(int64_t n, int t) { return (unsigned int)std::max<int64_t>(1, (n + t - 1) / t); }

// every kernel of this library is launched through here: the launch count of a join is what was launched, not a tally kept by hand
template <class... KA, class... A>
inline void launch(ejb_ctx* c, void (*k)(KA...), unsigned int grid, unsigned int block, A&&... a) {
    k<<<grid, block, 0, c->stream>>>(std::forward<A>(a)...);
    c->stats.kernel_launches++;
}
int sync_stream(ejb_ctx* c) {
    CK(cudaStreamSynchronize(c->stream));
    c->stats.host_syncs++;
    return EJB_OK;
}

// start.rs:50-99 with the product's dd arithmetic on the HOST (developer A/B knob "sr_host"; the default is k_sr_table)
void build_sr_table_host(std::vector<dd_t>& s, int n_max) {
    s.assign((size_t)(n_max + 1) * (n_max + 2) / 2, dd_make(0.0));
    auto at = [&](int n, int k) -> dd_t& { return s[(size_t)n * (n + 1) / 2 + k]; };
    at(0, 0) = dd_make(1.0);
    std::vector<dd_t> z, n2n1(n_max + 1, dd_make(0.0)), k1k(n_max, dd_make(0.0)), njn(n_max + 1, dd_make(0.0));
    for (int n = 2; n <= n_max; n++) n2n1[n] = dd_div(dd_make((double)(n - 2)), dd_make((double)(n - 1)));
    for (int k = 1; k < n_max; k++) k1k[k] = dd_div(dd_make((double)(k - 1)), dd_make((double)k));
    for (int n = 1; n <= n_max; n++) {
        dd_t p = dd_make(1.0);
        for (int j = 1; j <= n; j++) p = dd_mul(p, dd_div(dd_make((double)j), dd_make((double)n)));
        njn[n] = p;
    }
    for (int n = 1; n <= n_max; n++) {
        at(n, 0) = dd_make(0.0);
        for (int k = 1; k + 1 < n; k++) z[k - 1] = dd_mul(z[k - 1], k1k[k]);
        if (n >= 2) z.push_back(dd_powi(n2n1[n], n - 1));
        for (int k = 1; k < n; k++) at(n, k) = dd_add(at(n - 1, k), dd_mul(at(n - 1, k - 1), z[k - 1]));
        at(n, n) = njn[n];
    }
}

template <class T>
int upload_array(ejb_ctx* c, const T* src, int64_t n, const T** dst) {
    // inside the caller's arena: already on its way with the single arena copy
    if (n > 0 && src && c->arena_base && c->d_arena.p) {
        const char* p = reinterpret_cast<const char*>(src);
        if (p >= c->arena_base && p + (size_t)n * sizeof(T) <= c->arena_base + c->arena_bytes) {
            *dst = reinterpret_cast<const T*>(c->d_arena.as<char>() + (p - c->arena_base));
            return EJB_OK;
        }
    }
    if (c->in_used == c->in_bufs.size()) c->in_bufs.emplace_back();
    DevBuf& b = c->in_bufs[c->in_used++];
    const size_t bytes = (size_t)std::max<int64_t>(n, 1) * sizeof(T);
    CK(b.ensure(bytes + 16));
    if (n > 0) {
        if (!src) return fail(c, EJB_ERR_INVALID, "null input array with %lld elements", (long long)n);
        cudaStream_t st = c->stream;
        if ((size_t)n * sizeof(T) >= (1u << 20)) st = c->up_stream[c->up_next++ % ejb_ctx::N_UP];
        CK(cudaMemcpyAsync(b.p, src, (size_t)n * sizeof(T), cudaMemcpyHostToDevice, st));
        c->h2d_bytes += (int64_t)n * sizeof(T);
    }
    *dst = b.as<const T>();
    return EJB_OK;
}

int check_params(ejb_ctx* c, const ejb_params* p) {
    if (p->unsupported_modes != 0)
        return fail(c, EJB_ERR_UNSUPPORTED,
                    "join mode mask 0x%x requested; only the enclone_ranger default join is implemented "
                    "(bcjoin/basic/basic_h/basicx/join_full_diff/old_mult/force/super_comp_filt are rejected)",
                    p->unsupported_modes);
    if (p->ref_v_trim < 0 || p->ref_j_trim < 0 || p->max_degradation < 0) return fail(c, EJB_ERR_INVALID, "negative heuristic");
    return EJB_OK;
}

DevParams make_dev_params(const ejb_params& p) {
    DevParams P;
    P.is_bcr = p.is_bcr;
    P.mix_donors = p.mix_donors;
    P.easy = p.easy;
    P.old_light = p.old_light;
    P.max_score = p.max_score;
    P.cdr3_mult = p.cdr3_mult;
    P.cdr3_ident_thr = 1.0 - p.join_cdr3_ident / 100.0;  // join_one.rs:231, evaluated once in f64
    P.fwr1_cdr12_delta = p.fwr1_cdr12_delta;
    P.max_cdr3_diffs = p.max_cdr3_diffs;
    P.auto_share = p.auto_share;
    P.comp_filt = p.comp_filt;
    P.comp_filt_bound = p.comp_filt_bound;
    P.max_diffs = p.max_diffs;
    P.max_degradation = p.max_degradation;
    P.ref_v_trim = p.ref_v_trim;
    P.ref_j_trim = p.ref_j_trim;
    return P;
}

#define CTR_FIELD(f) (reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(c->d_round.p) + offsetof(Counters, f)))

template <int W1, int MODE>
void launch_stage1_wm(ejb_ctx* c, int64_t n_ctas, const RowTask* tasks, int n_tasks, unsigned long long* unit_cand, int strip) {
    launch(c, k_stage1<W1, MODE>, (unsigned int)n_ctas, S1_THREADS, c->d_subs.as<SubBucket>(), tasks, n_tasks, c->d_cdr3w.as<uint32_t>(), c->d_labq.as<int32_t>(),
           c->d_blklab.as<int32_t>(), c->d_cand.as<uint2>(), c->cand_cap, c->d_ctr(), unit_cand, c->d_tagq.as<unsigned long long>(),
           c->d_deadq.as<unsigned long long>(), c->d_donq.as<unsigned long long>(), c->d_flagq.as<uint32_t>(), strip,
           reinterpret_cast<uint8_t*>(unit_cand + c->cur_round_tasks * S1_WARPS),   // the per-task skip flags follow the band counters in the round report
           c->opt.early_return ? c->cur_strict : 0.0, c->d_ctatask.as<const int32_t>() + c->cur_cta_off);
}
template <int W1>
void launch_stage1_w(ejb_ctx* c, int64_t n_ctas, const RowTask* tasks, int n_tasks, unsigned long long* unit_cand, int strip) {
    switch (c->opt.s1_mode & (c->tags_on ? 3 : 1)) {
        case 0: launch_stage1_wm<W1, 0>(c, n_ctas, tasks, n_tasks, unit_cand, strip); break;
        case 1: launch_stage1_wm<W1, S1_RUNI>(c, n_ctas, tasks, n_tasks, unit_cand, strip); break;
        case 2: launch_stage1_wm<W1, S1_TAGS>(c, n_ctas, tasks, n_tasks, unit_cand, strip); break;
        default: launch_stage1_wm<W1, S1_RUNI | S1_TAGS>(c, n_ctas, tasks, n_tasks, unit_cand, strip); break;
    }
}

template <int W1>
void launch_stage1_thin(ejb_ctx* c, int64_t n_ctas, const RowTask* tasks, unsigned long long* unit_cand, int strip) {
    const bool tags = c->tags_on && (c->opt.s1_mode & S1_TAGS);
    if (tags)
        launch(c, k_stage1_thin<W1, true>, (unsigned int)n_ctas, S1_THREADS, c->d_subs.as<SubBucket>(), tasks, c->d_cdr3w.as<uint32_t>(), c->d_labq.as<int32_t>(),
               c->d_cand.as<uint2>(), c->cand_cap, c->d_ctr(), unit_cand, c->d_tagq.as<unsigned long long>(), c->d_deadq.as<unsigned long long>(),
               c->d_donq.as<unsigned long long>(), c->d_flagq.as<uint32_t>(), strip, c->d_ctatask.as<const int32_t>() + c->cur_cta_off);
    else
        launch(c, k_stage1_thin<W1, false>, (unsigned int)n_ctas, S1_THREADS, c->d_subs.as<SubBucket>(), tasks, c->d_cdr3w.as<uint32_t>(), c->d_labq.as<int32_t>(),
               c->d_cand.as<uint2>(), c->cand_cap, c->d_ctr(), unit_cand, (const unsigned long long*)nullptr, (const unsigned long long*)nullptr,
               (const unsigned long long*)nullptr, (const uint32_t*)nullptr, strip, c->d_ctatask.as<const int32_t>() + c->cur_cta_off);
}

// CTA -> task map of one stage-1 launch: thread per (task, CTA of the task); replaces a binary search over the task list in the
// prologue of every CTA (eleven dependent L2 loads) by one load
__global__ void k_expand_tasks(const RowTask* __restrict__ tasks, int n_tasks, const SubBucket* __restrict__ subs, int strip, int32_t* __restrict__ cta_task) {
    const int t = blockIdx.x;
    if (t >= n_tasks) return;
    const RowTask rt = tasks[t];
    const int nb = (subs[rt.sub].n + TILE - 1) / TILE;
    const int n = (nb - rt.rb + strip - 1) / strip;
    for (int j = threadIdx.x; j < n; j += blockDim.x) cta_task[rt.cta0 + j] = t;
}

// stage-1 for sub-buckets whose CDR3s exceed the fast path: every not-yet-connected pair is a candidate
__global__ void k_stage1_nofilter(const SubBucket* __restrict__ subs, const RowTask* __restrict__ tasks, int n_tasks,
                                  const int32_t* __restrict__ labq, uint2* __restrict__ cand, unsigned long long cand_cap, Counters* ctr,
                                  unsigned long long* __restrict__ unit_cand, int strip) {
    const int64_t cta = blockIdx.x;
    int lo = 0, hi = n_tasks - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (tasks[mid].cta0 <= cta) lo = mid; else hi = mid - 1;
    }
    const RowTask t = tasks[lo];
    const SubBucket sb = subs[t.sub];
    const int nb = (sb.n + TILE - 1) / TILE;
    const int cb_first = t.rb + (int)(cta - t.cta0) * strip;
    unsigned long long ev = 0;
    for (int cb = cb_first; cb < nb && cb < cb_first + strip; cb++) {
        for (int idx = threadIdx.x; idx < TILE * TILE; idx += blockDim.x) {
            const int rl = idx >> 7;
            const int r = t.rb * TILE + rl, cc = cb * TILE + (idx & 127);
            if (rl < t.row_lo || rl >= t.row_hi || r >= sb.n || cc >= sb.n || r >= cc) continue;
            ev++;
            const uint32_t q1 = (uint32_t)(sb.q0 + r), q2 = (uint32_t)(sb.q0 + cc);
            if (labq[q1] == labq[q2]) continue;
            const unsigned long long o = atomicAdd(&ctr->n_cand, 1ull);
            atomicAdd(&unit_cand[(size_t)t.tid * S1_WARPS + (rl >> 4)], 1ull);   // per 16-row band, like k_stage1
            if (o < cand_cap) cand[o] = make_uint2(q1, q2); else ctr->overflow = 1ull;
        }
    }
    if (ev) atomicAdd(&ctr->pairs_eval, ev);
}

// pair-batch mode (ejb_join_one_batch): rows -> packed positions; a pair whose rows cannot pass join_one's first tests (a row
// without exactly two chains, join_one.rs:83-96; different lens or CDR3 lengths, :101-109) is answered `false` right here
__global__ void k_q_of_row(const int32_t* __restrict__ rowq, int64_t n_pos, int32_t* __restrict__ q_of_row) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q < n_pos && rowq[q] >= 0) q_of_row[rowq[q]] = (int32_t)q;
}
__global__ void k_batch_pairs(const int32_t* __restrict__ k1, const int32_t* __restrict__ k2, int64_t n, int64_t n_rows, const int32_t* __restrict__ q_of_row,
                              const int32_t* __restrict__ subq, uint2* __restrict__ pairs, uint32_t* __restrict__ slots, unsigned long long* __restrict__ n_out,
                              unsigned int* __restrict__ bad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int a = k1[i], b = k2[i];
    if (a < 0 || a >= n_rows || b < 0 || b >= n_rows) {
        atomicOr(bad, 1u);
        return;
    }
    const int qa = q_of_row[a], qb = q_of_row[b];
    if (qa < 0 || qb < 0 || subq[qa] != subq[qb]) return;   // join_one returns false before looking at any sequence
    const unsigned long long o = atomicAdd(n_out, 1ull);
    pairs[o] = make_uint2((uint32_t)qa, (uint32_t)qb);
    slots[o] = (uint32_t)i;
}

struct RoundTimes {
    cudaEvent_t e0, e1, e2, e3;
};

S2Ctx make_s2ctx(ejb_ctx* c, int tuple_mode) {
    S2Ctx X;
    X.in = c->din;
    X.P = make_dev_params(c->params);
    X.subs = c->d_subs.as<SubBucket>();
    X.aux = c->d_aux.as<int32_t>();
    X.rowq = c->d_m_row.as<int32_t>();
    X.subq = c->d_m_sub.as<int32_t>();
    X.qcut = c->d_qcut.as<int32_t>();
    X.sub_pass = c->d_subpass.as<unsigned int>();
    X.memo = c->d_memo.as<int32_t>();
    X.tagq = c->tags_on ? c->d_tagq.as<unsigned long long>() : nullptr;
    X.deadq = c->tags_on ? c->d_deadq.as<unsigned long long>() : nullptr;
    X.cdr3w = c->d_cdr3w.as<uint32_t>();
    X.rowstore = c->d_rowstore.as<uint4>();
    X.segpl = c->d_segpl.as<uint32_t>();
    X.seglen = c->d_seglen.as<int32_t>();
    X.bc_sorted = c->d_bckeys2.as<uint64_t>();
    X.powtab = c->d_powtab.as<double>();
    X.pow_off = c->d_powoff.as<int32_t>();
    X.p1.keys = c->d_p1_keys.as<unsigned long long>();
    X.p1.vals = c->d_p1_vals.as<double>();
    X.p1.newslots = c->d_p1_new.as<uint32_t>();
    X.p1.cap_mask = c->p1_cap - 1;
    X.ctr = c->d_ctr();
    X.surv = c->d_surv.as<Survivor>();
    X.surv_cap = c->cand_cap;
    X.slow = c->d_slow.as<uint2>();
    X.slow_slot = c->d_slowslot.as<uint32_t>();
    X.slow_cap = c->cand_cap;
    X.tuple_mode = tuple_mode;
    X.trace = c->opt.trace;
    return X;
}

P1Groups make_p1_groups(ejb_ctx* c) {
    P1Groups G;
    G.gkeys = c->d_gkeys.as<unsigned long long>();
    G.goff = c->d_goff.as<uint32_t>();
    G.ghave = c->d_ghave.as<uint32_t>();
    G.gwant = c->d_gwant.as<uint32_t>();
    G.gstamp = c->d_gstamp.as<uint32_t>();
    G.touched = c->d_gtouched.as<uint32_t>();
    G.desc = c->d_gdesc.as<uint2>();
    G.pool = c->d_gpool.as<double2>();
    G.gmask = c->g_cap - 1;
    G.round = c->p1_round;
    G.pool_cap = c->pool_cap;
    G.desc_cap = c->desc_cap;
    G.touched_cap = c->g_cap;
    return G;
}
// p1 of the keys requested since the last call (Counters::n_newkeys of them in pc.newslots): groups -> missing terms -> sums.
// The per-round counters p1_touched / p1_desc were zeroed with the round block.
int enqueue_p1(ejb_ctx* c, const P1Cache& pc) {
    c->p1_round++;
    const P1Groups G = make_p1_groups(c);
    const int grid = c->sm_count * 8;
    const unsigned long long* n_new = (const unsigned long long*)CTR_FIELD(n_newkeys);
    launch(c, k_p1_groups, grid, 256, pc, G, n_new, c->d_ctr());
    launch(c, k_p1_expand, grid, 256, G, c->d_ctr());
    launch(c, k_p1_terms, grid, 256, G, c->d_sr.as<const dd_t>(), c->qt(), c->d_ctr());
    launch(c, k_p1_sum, grid, 256, pc, G, c->d_sr.as<const dd_t>(), c->qt(), n_new, c->d_ctr());
    CK(cudaGetLastError());
    return EJB_OK;
}
int p1_groups_alloc(ejb_ctx* c) {
    const size_t g = c->g_cap;
    CK(c->d_gkeys.ensure(g * 8)); CK(c->d_goff.ensure(g * 4)); CK(c->d_ghave.ensure(g * 4)); CK(c->d_gwant.ensure(g * 4));
    CK(c->d_gstamp.ensure(g * 4)); CK(c->d_gtouched.ensure(g * 4));
    CK(c->d_gdesc.ensure((size_t)c->desc_cap * sizeof(uint2)));
    CK(c->d_gpool.ensure((size_t)c->pool_cap * sizeof(double2)));
    return EJB_OK;
}
int p1_groups_clear(ejb_ctx* c) {
    const size_t g = c->g_cap;
    CK(cudaMemsetAsync(c->d_gkeys.p, 0xff, g * 8, c->stream));
    CK(cudaMemsetAsync(c->d_ghave.p, 0, g * 4, c->stream));
    CK(cudaMemsetAsync(c->d_gwant.p, 0, g * 4, c->stream));
    CK(cudaMemsetAsync(c->d_gstamp.p, 0, g * 4, c->stream));
    c->p1_round = 0;
    c->p1_groups = c->p1_pool_used = 0;
    return EJB_OK;
}
// between rounds (rare): keep the group table below 1/4 load, the pool and the per-round term list below 1/2
int p1_groups_grow(ejb_ctx* c, unsigned long long desc_seen) {
    cudaStream_t st = c->stream;
    if (c->p1_groups > c->g_cap / 4 && c->g_cap < (1u << 29)) {
        const P1Groups src = make_p1_groups(c);
        const uint32_t old_cap = c->g_cap, new_cap = c->g_cap << 2;
        DevBuf k, o, h, w, s2, t;
        CK(k.ensure((size_t)new_cap * 8)); CK(o.ensure((size_t)new_cap * 4)); CK(h.ensure((size_t)new_cap * 4)); CK(w.ensure((size_t)new_cap * 4));
        CK(s2.ensure((size_t)new_cap * 4)); CK(t.ensure((size_t)new_cap * 4));
        CK(cudaMemsetAsync(k.p, 0xff, (size_t)new_cap * 8, st));
        P1Groups dst = src;
        dst.gkeys = k.as<unsigned long long>(); dst.goff = o.as<uint32_t>(); dst.ghave = h.as<uint32_t>(); dst.gwant = w.as<uint32_t>();
        dst.gstamp = s2.as<uint32_t>(); dst.touched = t.as<uint32_t>();
        dst.gmask = new_cap - 1;
        launch(c, k_p1_groups_rehash, nblk(old_cap, 256), 256, src, old_cap, dst);
        CK(cudaStreamSynchronize(st));
        std::swap(c->d_gkeys, k); std::swap(c->d_goff, o); std::swap(c->d_ghave, h); std::swap(c->d_gwant, w); std::swap(c->d_gstamp, s2);
        std::swap(c->d_gtouched, t);
        c->g_cap = new_cap;
    }
    if (c->p1_pool_used > c->pool_cap / 2 && c->pool_cap < (1ull << 32)) {
        unsigned long long nc = c->pool_cap;
        while (c->p1_pool_used > nc / 2 && nc < (1ull << 32)) nc <<= 1;
        nc = std::min(nc, (1ull << 32) - 1);   // goff is 32 bits
        DevBuf np;
        CK(np.ensure((size_t)nc * sizeof(double2)));
        CK(cudaMemcpyAsync(np.p, c->d_gpool.p, (size_t)std::min(c->p1_pool_used, c->pool_cap) * sizeof(double2), cudaMemcpyDeviceToDevice, st));
        CK(cudaStreamSynchronize(st));
        std::swap(c->d_gpool, np);
        c->pool_cap = nc;
    }
    if (desc_seen > c->desc_cap / 2 && c->desc_cap < (1ull << 30)) {
        while (desc_seen > c->desc_cap / 2 && c->desc_cap < (1ull << 30)) c->desc_cap <<= 1;
        DevBuf nd;
        CK(nd.ensure((size_t)c->desc_cap * sizeof(uint2)));
        CK(cudaStreamSynchronize(st));
        std::swap(c->d_gdesc, nd);
    }
    return EJB_OK;
}

// stage 2 on a candidate list (count on the device): 2a -> slow -> p1 -> 2b.  The per-round counters it uses were zeroed with
// the round (filter mode) or by the caller (pair-batch mode).
int run_stage2(ejb_ctx* c, const uint2* cand, const uint32_t* cand_slot, const unsigned long long* n_cand_dev, unsigned long long list_cap, int tuple_mode,
               EdgeTuple* tuples_out) {
    S2Ctx X = make_s2ctx(c, tuple_mode);
    const int grid = c->sm_count * 8;
    cudaEvent_t ea = c->ev(), eb = c->ev();
    CK(cudaEventRecord(ea, c->stream));
    launch(c, k_stage2a, grid, 256, X, cand, cand_slot, n_cand_dev, list_cap);
    CK(cudaEventRecord(eb, c->stream));
    c->s2a_events.push_back({ea, eb});
    c->stats.stage2a_launches++;
    launch(c, k_stage2_slow, grid, 256, X, (const unsigned long long*)CTR_FIELD(n_slow));
    cudaEvent_t pa = c->ev(), pb = c->ev();
    CK(cudaEventRecord(pa, c->stream));
    int rc = enqueue_p1(c, X.p1);
    if (rc) return rc;
    CK(cudaEventRecord(pb, c->stream));
    c->p1_events.push_back({pa, pb});
    launch(c, k_stage2b, grid, 256, X, (const unsigned long long*)CTR_FIELD(n_surv), c->d_pass.as<unsigned long long>(), c->d_passt.as<PassTuple>(), c->cand_cap,
           tuples_out);
    CK(cudaGetLastError());
    return EJB_OK;
}

int p1_cache_clear(ejb_ctx* c) {
    CK(cudaMemsetAsync(c->d_p1_keys.p, 0xff, (size_t)c->p1_cap * 8, c->stream));
    c->p1_count = 0;
    return p1_groups_clear(c);
}
// keep the load factor of the p1 hash below 1/4: grow by 4x and re-insert (between rounds; rare)
int p1_cache_grow(ejb_ctx* c) {
    if (c->p1_cap >= (1u << 30)) return fail(c, EJB_ERR_OOM, "p1 cache cannot grow beyond 2^30 slots (%llu keys)", c->p1_count);
    const uint32_t new_cap = c->p1_cap << 2;
    DevBuf nk, nv, nn;
    CK(nk.ensure((size_t)new_cap * 8));
    CK(nv.ensure((size_t)new_cap * 8));
    CK(nn.ensure((size_t)new_cap * 4));
    CK(cudaMemsetAsync(nk.p, 0xff, (size_t)new_cap * 8, c->stream));
    P1Cache dst;
    dst.keys = nk.as<unsigned long long>();
    dst.vals = nv.as<double>();
    dst.newslots = nn.as<uint32_t>();
    dst.cap_mask = new_cap - 1;
    launch(c, k_p1_rehash, nblk(c->p1_cap, 256), 256, c->d_p1_keys.as<const unsigned long long>(), c->d_p1_vals.as<const double>(), c->p1_cap, dst);
    CK(cudaStreamSynchronize(c->stream));
    std::swap(c->d_p1_keys, nk);
    std::swap(c->d_p1_vals, nv);
    std::swap(c->d_p1_new, nn);
    c->p1_cap = new_cap;
    return EJB_OK;
}

void destroy_ctx(ejb_ctx* c);

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
namespace {

void destroy_ctx(ejb_ctx* c) {
    if (!c) return;
    for (ejb_ctx* p : c->peers) destroy_ctx(p);
    c->peers.clear();
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (auto e : c->ev_pool) cudaEventDestroy(e);
    if (c->h_res) cudaFreeHost(c->h_res);
    c->in_bufs.clear();
    for (int i = 0; i < ejb_ctx::N_UP; i++) {
        if (c->up_event[i]) cudaEventDestroy(c->up_event[i]);
        if (c->up_stream[i]) cudaStreamDestroy(c->up_stream[i]);
    }
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;   // DevBuf / PinBuf members release their memory (the device is current)
}

void read_env_options(Options& o) {
    if (const char* e = getenv("EJB_FIRST_UNIT")) o.first_unit = std::max(1ll, atoll(e));
    if (const char* e = getenv("EJB_GROWTH")) o.growth = std::max(1ll, atoll(e));
    if (const char* e = getenv("EJB_S1_MODE")) o.s1_mode = atoi(e) & 3;
    if (const char* e = getenv("EJB_STRICT_DIV")) o.strict_div = std::max(1ll, atoll(e));
    if (getenv("EJB_NO_TAGS")) o.no_tags = 1;
    if (getenv("EJB_TRACE")) o.trace = 1;
    if (getenv("EJB_SR_HOST")) o.sr_host = 1;
}

// one context bound to one device; every failure path runs the same teardown as ejb_destroy
int create_one(ejb_ctx** out, int dev) {
    ejb_ctx* c = new ejb_ctx();
    c->device = dev;
    auto bail = [&](int code) {
        destroy_ctx(c);
        return code;
    };
    if (cudaSetDevice(dev) != cudaSuccess) {
        delete c;
        return EJB_ERR_NO_DEVICE;
    }
    const double t0 = now_ms();
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return bail(EJB_ERR_NO_DEVICE);
    c->sm_count = prop.multiProcessorCount;
    read_env_options(c->opt);
    if (const char* e = getenv("EJB_CAND_CAP")) {
        c->cand_cap = std::max<unsigned long long>(1024, strtoull(e, nullptr, 10));
        c->cap_explicit = true;
    }
    if (const char* e = getenv("EJB_PAIR_BUDGET")) c->pair_budget = std::max<unsigned long long>(1ull << 16, strtoull(e, nullptr, 10));
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(EJB_ERR_CUDA);
    for (int i = 0; i < ejb_ctx::N_UP; i++)
        if (cudaStreamCreateWithFlags(&c->up_stream[i], cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&c->up_event[i], cudaEventDisableTiming) != cudaSuccess)
            return bail(EJB_ERR_CUDA);
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_forest_round, 256, 0) != cudaSuccess || per_sm < 1) return bail(EJB_ERR_CUDA);
    c->coop_grid = c->sm_count * std::min(per_sm, 4);
    int coop = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    if (!coop) return bail(EJB_ERR_NO_DEVICE);
    // Stirling ratio table (start.rs:50-99): built by one CTA on the device, 72 MB, once per context
    const size_t sr_entries = (size_t)(SR_NMAX + 1) * (SR_NMAX + 2) / 2;
    if (c->d_sr.ensure(sr_entries * sizeof(dd_t)) != cudaSuccess) return bail(EJB_ERR_OOM);
    if (c->opt.sr_host) {
        std::vector<dd_t> t;
        build_sr_table_host(t, SR_NMAX);
        if (cudaMemcpy(c->d_sr.p, t.data(), sr_entries * sizeof(dd_t), cudaMemcpyHostToDevice) != cudaSuccess) return bail(EJB_ERR_CUDA);
    } else {
        static_assert(SR_NMAX < 1024 * SR_KPT, "k_sr_table: one thread owns SR_KPT columns");
        k_sr_table<<<1, 1024, 0, c->stream>>>(c->d_sr.as<dd_t>(), SR_NMAX);
    }
    if (c->d_qt.ensure((size_t)QT_DELTA * QT_B * sizeof(double2)) != cudaSuccess) return bail(EJB_ERR_OOM);
    k_quot_table<<<(unsigned)(((size_t)QT_DELTA * QT_B + 255) / 256), 256, 0, c->stream>>>(reinterpret_cast<double2*>(c->d_qt.p));
    if (c->d_p1_keys.ensure((size_t)c->p1_cap * 8) != cudaSuccess || c->d_p1_vals.ensure((size_t)c->p1_cap * 8) != cudaSuccess ||
        c->d_p1_new.ensure((size_t)c->p1_cap * 4) != cudaSuccess || c->d_round.ensure(sizeof(Counters) + 4096) != cudaSuccess ||
        c->d_err.ensure(256) != cudaSuccess || c->d_bar.ensure(256) != cudaSuccess)
        return bail(EJB_ERR_OOM);
    if (p1_groups_alloc(c) != EJB_OK || p1_cache_clear(c) != EJB_OK) return bail(EJB_ERR_OOM);
    cudaMemsetAsync(c->d_round.p, 0, sizeof(Counters), c->stream);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) return bail(EJB_ERR_CUDA);
    memset(&c->stats, 0, sizeof(c->stats));
    c->ms_create = now_ms() - t0;
    *out = c;
    return EJB_OK;
}

}  // namespace

extern "C" {

int ejb_abi_version(void) { return EJB_ABI_VERSION; }
void ejb_abi_sizes(int64_t out[4]) {
    out[0] = sizeof(ejb_params);
    out[1] = sizeof(ejb_input);
    out[2] = sizeof(ejb_stats);
    out[3] = sizeof(ejb_result);
}

void ejb_params_default(ejb_params* p, int is_bcr) {
    memset(p, 0, sizeof(*p));
    p->is_bcr = is_bcr ? 1 : 0;
    p->max_score = 100000.0;
    p->cdr3_mult = 5.0;
    p->mult_pow = 80.0;
    p->join_cdr3_ident = 85.0;
    p->fwr1_cdr12_delta = 20.0;
    p->max_cdr3_diffs = 1000;
    p->cdr3_normal_len = 42;
    p->auto_share = 15;
    p->comp_filt = 8;
    p->comp_filt_bound = 80;
    p->max_diffs = 1000000;
    p->max_degradation = 2;
    p->ref_v_trim = 15;
    p->ref_j_trim = 15;
}

const char* ejb_last_error(const ejb_ctx* c) { return c ? c->err.c_str() : "null context"; }
double ejb_create_ms(const ejb_ctx* c) { return c ? c->ms_create : 0.0; }

int ejb_create(ejb_ctx** out, const int* device_ids, int n_devices) {
    if (!out) return EJB_ERR_INVALID;
    *out = nullptr;
    if (n_devices < 1 || n_devices > 64) return EJB_ERR_INVALID;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return EJB_ERR_NO_DEVICE;
    std::vector<int> devs((size_t)n_devices);
    for (int i = 0; i < n_devices; i++) {
        devs[i] = device_ids ? device_ids[i] : i;
        if (devs[i] < 0 || devs[i] >= count) return EJB_ERR_NO_DEVICE;
    }
    ejb_ctx* c = nullptr;
    int rc = create_one(&c, devs[0]);
    if (rc) return rc;
    // one peer context per further device; created concurrently (each builds its Stirling table on its own GPU)
    if (n_devices > 1) {
        std::vector<ejb_ctx*> peers((size_t)n_devices - 1, nullptr);
        std::vector<int> rcs((size_t)n_devices - 1, 0);
        std::vector<std::thread> th;
        for (int i = 1; i < n_devices; i++) th.emplace_back([&, i] { rcs[i - 1] = create_one(&peers[i - 1], devs[i]); });
        for (auto& t : th) t.join();
        for (int i = 1; i < n_devices; i++) {
            if (peers[i - 1]) c->peers.push_back(peers[i - 1]);
            if (rcs[i - 1] && !rc) rc = rcs[i - 1];
        }
        if (rc) {
            destroy_ctx(c);
            return rc;
        }
        // direct GPU-to-GPU copies (NVLink) between every pair of devices of the context; "already enabled" is fine
        for (int i = 0; i < n_devices; i++)
            for (int j = 0; j < n_devices; j++) {
                if (devs[i] == devs[j]) continue;
                int can = 0;
                if (cudaDeviceCanAccessPeer(&can, devs[i], devs[j]) == cudaSuccess && can) {
                    cudaSetDevice(devs[i]);
                    cudaDeviceEnablePeerAccess(devs[j], 0);
                    cudaGetLastError();
                }
            }
        cudaSetDevice(devs[0]);
    }
    *out = c;
    return EJB_OK;
}

void ejb_destroy(ejb_ctx* c) { destroy_ctx(c); }

int ejb_set_option(ejb_ctx* c, const char* name, long long value) {
    if (!c || !name) return EJB_ERR_INVALID;
    const std::string n(name);
    auto apply = [&](ejb_ctx* x) {
        if (n == "first_unit") x->opt.first_unit = std::max(1ll, value);
        else if (n == "growth") x->opt.growth = std::max(1ll, value);
        else if (n == "strict_div") x->opt.strict_div = std::max(1ll, value);
        else if (n == "early_return") x->opt.early_return = value != 0;
        else if (n == "s1_mode") x->opt.s1_mode = (int)value & 3;
        else if (n == "no_tags") x->opt.no_tags = value != 0;
        else if (n == "trace") x->opt.trace = value != 0;
        else if (n == "quot_table") x->opt.quot_table = value != 0;
        else if (n == "thin") x->opt.thin = (int)std::max(0ll, std::min(128ll, value));
        else return false;
        return true;
    };
    if (!apply(c)) return fail(c, EJB_ERR_INVALID, "unknown option '%s'", name);
    for (ejb_ctx* p : c->peers) apply(p);
    return EJB_OK;
}

// Multi-GPU: this context only joins the sub-buckets it owns (LPT by pair count over `world`
// owners, identical on every rank); the caller gathers the per-rank edge lists.
int ejb_set_shard(ejb_ctx* c, int rank, int world) {
    if (!c || world < 1 || rank < 0 || rank >= world) return EJB_ERR_INVALID;
    c->shard_rank = rank;
    c->shard_world = world;
    return EJB_OK;
}

int ejb_set_input_arena(ejb_ctx* c, const void* base, uint64_t bytes) {
    if (!c) return EJB_ERR_INVALID;
    c->arena_base = nullptr;
    c->arena_bytes = 0;
    if (!base || bytes == 0) return EJB_OK;
    if ((reinterpret_cast<uintptr_t>(base) & 255u) != 0) return fail(c, EJB_ERR_INVALID, "input arena must be 256-byte aligned");
    c->arena_base = reinterpret_cast<const char*>(base);
    c->arena_bytes = (size_t)bytes;
    return EJB_OK;
}

int ejb_set_capacity(ejb_ctx* c, unsigned long long cand_cap, unsigned long long pair_budget) {
    if (!c) return EJB_ERR_INVALID;
    auto apply = [&](ejb_ctx* x) {
        if (cand_cap) {
            x->cand_cap = std::max<unsigned long long>(1024, cand_cap);
            x->cap_explicit = true;
        }
        if (pair_budget) x->pair_budget = std::max<unsigned long long>(1ull << 14, pair_budget);
    };
    apply(c);
    for (ejb_ctx* p : c->peers) apply(p);
    return EJB_OK;
}

}  // extern "C"
namespace {
// part / nparts: copy only slice `part` of the declared input arena (the other slices arrive GPU-to-GPU, ejb_upload_slice)
int upload_impl(ejb_ctx* c, const ejb_params* p, const ejb_input* in, int part, int nparts) {
    if (!c || !p || !in) return EJB_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    c->have_input = c->have_run = c->prepared = false;
    int rc = check_params(c, p);
    if (rc) return rc;
    if (in->n_rows < 0 || in->n_rows >= (1ll << 28)) return fail(c, EJB_ERR_INVALID, "n_rows out of range (need 0 <= n < 2^28)");
    if (in->n_exacts < 0 || in->n_chains < 0 || in->n_cells < 0 || in->n_segs < 0 || in->n_refs < 0) return fail(c, EJB_ERR_INVALID, "negative count");
    if (in->n_rows > 0 && (in->n_exacts == 0 || in->n_segs == 0)) return fail(c, EJB_ERR_INVALID, "rows without exacts/segs");
    c->params = *p;
    const double t0 = now_ms();
    if (c->in_bufs.capacity() < 64) c->in_bufs.reserve(64);   // buffers persist across uploads (no cudaFree/cudaMalloc per join)
    c->in_used = 0;
    c->h2d_bytes = 0;
    DevIn& d = c->din;
    memset(&d, 0, sizeof(d));
    d.n_rows = in->n_rows; d.n_exacts = in->n_exacts; d.n_chains = in->n_chains; d.n_cells = in->n_cells;
    d.n_segs = in->n_segs; d.n_refs = in->n_refs;
    d.n_tig_bytes = in->n_tig_bytes; d.n_cdr3_bytes = in->n_cdr3_bytes; d.n_amino_bytes = in->n_amino_bytes;
    d.n_tig2_words = in->row_tig2_off ? in->n_tig2_words : 0;
    d.n_cdr32_words = in->row_cdr32_off ? in->n_cdr32_words : 0;
    const int64_t N = in->n_rows;
    // host-side checks come before anything is dereferenced or queued
    if (in->n_segs > 0 && (!in->seg_off || !in->seg_bytes)) return fail(c, EJB_ERR_INVALID, "seg_off / seg_bytes is null");
    if (in->n_refs > 0 && !in->ref_off) return fail(c, EJB_ERR_INVALID, "ref_off is null");
    d.n_seg_bytes = in->n_segs > 0 ? in->seg_off[in->n_segs] : 0;
    d.n_ref_bytes = in->n_refs > 0 ? in->ref_off[in->n_refs] : 0;
    if (d.n_seg_bytes < 0 || d.n_ref_bytes < 0) return fail(c, EJB_ERR_INVALID, "negative seg/ref byte count");
    c->arena_slice = 0;
    c->arena_parts = 1;
    if (nparts > 1 && !c->arena_base) return fail(c, EJB_ERR_INVALID, "a sliced upload needs the input arena (ejb_set_input_arena)");
    if (c->arena_base) {   // one DMA for everything the caller keeps in its arena
        if (nparts > 1) {
            // equal 256-byte aligned slices, so that the caller's all-gather works on one contiguous device buffer
            const size_t slice = ((c->arena_bytes + (size_t)nparts - 1) / (size_t)nparts + 255) & ~(size_t)255;
            CK(c->d_arena.ensure(slice * (size_t)nparts + 256));
            c->arena_slice = slice;
            c->arena_parts = nparts;
            const size_t lo = std::min(slice * (size_t)part, c->arena_bytes), hi = std::min(lo + slice, c->arena_bytes);
            if (hi > lo) {
                CK(cudaMemcpyAsync(c->d_arena.as<char>() + lo, c->arena_base + lo, hi - lo, cudaMemcpyHostToDevice, c->up_stream[c->up_next++ % ejb_ctx::N_UP]));
                c->h2d_bytes += (int64_t)(hi - lo);
            }
        } else {
            CK(c->d_arena.ensure(c->arena_bytes + 256));
            CK(cudaMemcpyAsync(c->d_arena.p, c->arena_base, c->arena_bytes, cudaMemcpyHostToDevice, c->up_stream[c->up_next++ % ejb_ctx::N_UP]));
            c->h2d_bytes += (int64_t)c->arena_bytes;
        }
    }
    // content-canonical seg ids: DnaString equality is by content (join_one.rs:331)
    std::vector<int32_t> canon((size_t)std::max<int64_t>(in->n_segs, 1), 0);
    std::vector<uint8_t> seg_canon_bytes((size_t)std::max<int64_t>(d.n_seg_bytes, 1), 'A');   // DnaString content: upper-case ACGT, anything else -> A
    {
        std::unordered_map<std::string, int> ids;
        for (int64_t s = 0; s < in->n_segs; s++) {
            const int64_t a = in->seg_off[s], b = in->seg_off[s + 1];
            if (a < 0 || b < a || b > d.n_seg_bytes) return fail(c, EJB_ERR_INVALID, "seg_off not monotone");
            std::string key((const char*)in->seg_bytes + a, (size_t)(b - a));
            for (auto& ch : key) {  // 2-bit content: non-ACGT -> A, case-insensitive
                switch (ch) {
                    case 'C': case 'c': ch = 'C'; break;
                    case 'G': case 'g': ch = 'G'; break;
                    case 'T': case 't': ch = 'T'; break;
                    default: ch = 'A';
                }
            }
            memcpy(seg_canon_bytes.data() + a, key.data(), key.size());
            auto it = ids.find(key);
            if (it == ids.end()) it = ids.emplace(std::move(key), (int)ids.size()).first;
            canon[s] = it->second;
        }
        c->n_canon_segs = (int64_t)ids.size();
    }
#define UP(field, n) if ((rc = upload_array(c, in->field, (n), &d.field)) != EJB_OK) return rc
    UP(row_nchains, N); UP(row_len, 2 * N); UP(row_has_del, 2 * N); UP(row_vs, 2 * N); UP(row_js, 2 * N);
    UP(row_cdr3_len, 2 * N); UP(row_exact, N); UP(row_exact_col, 2 * N);
    UP(ex_chain_off, in->n_exacts + 1); UP(ex_cell_off, in->n_exacts + 1);
    UP(ch_left, in->n_chains); UP(ch_c_ref, in->n_chains); UP(ch_v_ref, in->n_chains); UP(ch_u_ref, in->n_chains);
    UP(ch_hcomp, in->n_chains); UP(ch_amino_off, in->n_chains); UP(ch_amino_len, in->n_chains);
    UP(amino_bytes, in->n_amino_bytes); UP(cell_donor, in->n_cells); UP(cell_barcode, in->n_cells);
    UP(seg_off, in->n_segs + 1); UP(ref_off, in->n_refs + 1); UP(ref_bytes, d.n_ref_bytes);
    UP(ref_name_id, in->n_refs);
#undef UP
    // ASCII contigs / CDR3s may be absent altogether when every row supplies the 2-bit packed form
    if (in->n_tig_bytes > 0 || !in->row_tig2_off) {
        if ((rc = upload_array(c, in->row_tig_off, 2 * N, &d.row_tig_off)) != EJB_OK) return rc;
        if ((rc = upload_array(c, in->tig_bytes, in->n_tig_bytes, &d.tig_bytes)) != EJB_OK) return rc;
    } else {
        if ((rc = upload_array(c, (const uint8_t*)nullptr, 0, &d.tig_bytes)) != EJB_OK) return rc;
        d.row_tig_off = nullptr;
    }
    if (in->n_cdr3_bytes > 0 || !in->row_cdr32_off) {
        if ((rc = upload_array(c, in->row_cdr3_off, 2 * N, &d.row_cdr3_off)) != EJB_OK) return rc;
        if ((rc = upload_array(c, in->cdr3_bytes, in->n_cdr3_bytes, &d.cdr3_bytes)) != EJB_OK) return rc;
    } else {
        if ((rc = upload_array(c, (const uint8_t*)nullptr, 0, &d.cdr3_bytes)) != EJB_OK) return rc;
        d.row_cdr3_off = nullptr;
    }
    if ((rc = upload_array(c, in->ch_fr1_start, in->n_chains, &d.ch_fr1)) != EJB_OK) return rc;
    if ((rc = upload_array(c, in->ch_cdr1_start, in->n_chains, &d.ch_cdr1)) != EJB_OK) return rc;
    if ((rc = upload_array(c, in->ch_fr2_start, in->n_chains, &d.ch_fr2)) != EJB_OK) return rc;
    if ((rc = upload_array(c, in->ch_cdr2_start, in->n_chains, &d.ch_cdr2)) != EJB_OK) return rc;
    if ((rc = upload_array(c, in->ch_fr3_start, in->n_chains, &d.ch_fr3)) != EJB_OK) return rc;
    if (in->row_tig2_off) {
        if ((rc = upload_array(c, in->row_tig2_off, 2 * N, &d.row_tig2_off)) != EJB_OK) return rc;
        if ((rc = upload_array(c, in->tig2_words, in->n_tig2_words, &d.tig2_words)) != EJB_OK) return rc;
    }
    if (in->row_cdr32_off) {
        if ((rc = upload_array(c, in->row_cdr32_off, 2 * N, &d.row_cdr32_off)) != EJB_OK) return rc;
        if ((rc = upload_array(c, in->cdr32_words, in->n_cdr32_words, &d.cdr32_words)) != EJB_OK) return rc;
    }
    if ((rc = upload_array(c, canon.data(), in->n_segs, &d.seg_canon)) != EJB_OK) return rc;
    if ((rc = upload_array(c, (const uint8_t*)seg_canon_bytes.data(), d.n_seg_bytes, &d.seg_bytes)) != EJB_OK) return rc;
    for (int i = 0; i < ejb_ctx::N_UP; i++) {   // the join's stream continues after every copy stream
        CK(cudaEventRecord(c->up_event[i], c->up_stream[i]));
        CK(cudaStreamWaitEvent(c->stream, c->up_event[i], 0));
    }
    CK(cudaStreamSynchronize(c->stream));   // canon / seg_canon_bytes are host temporaries
    c->ms_h2d = now_ms() - t0;
    c->have_input = true;
    return EJB_OK;
}
}  // namespace

extern "C" {
int ejb_upload(ejb_ctx* c, const ejb_params* p, const ejb_input* in) { return upload_impl(c, p, in, 0, 1); }

// Sharded upload for contexts that all need the WHOLE input (ejb_set_shard): context `part` of `nparts` copies only its slice of
// the declared input arena over PCIe; the caller then completes every context's device arena GPU-to-GPU (an in-place all-gather
// over NVLink on the buffer ejb_device_arena returns: slice i lives at offset i * slice_bytes) before ejb_run.  The host -> device
// traffic of N GPUs is the input once instead of N times.
int ejb_upload_slice(ejb_ctx* c, const ejb_params* p, const ejb_input* in, int part, int nparts) {
    if (nparts < 1 || part < 0 || part >= nparts) return EJB_ERR_INVALID;
    return upload_impl(c, p, in, part, nparts);
}
int ejb_device_arena(ejb_ctx* c, void** d_base, uint64_t* total_bytes, uint64_t* slice_bytes) {
    if (!c) return EJB_ERR_INVALID;
    if (!c->have_input || !c->d_arena.p) return fail(c, EJB_ERR_INVALID, "ejb_device_arena without an uploaded input arena");
    if (d_base) *d_base = c->d_arena.p;
    if (slice_bytes) *slice_bytes = (uint64_t)c->arena_slice;
    if (total_bytes) *total_bytes = c->arena_slice ? (uint64_t)c->arena_slice * (uint64_t)c->arena_parts : (uint64_t)c->arena_bytes;
    return EJB_OK;
}
}  // extern "C"

// ================================================================================================
// ejb_run
// ================================================================================================
namespace {

int cub_sort_pairs64(ejb_ctx* c, const uint64_t* kin, uint64_t* kout, const uint32_t* vin, uint32_t* vout, int64_t n, int end_bit = 64) {
    size_t tmp = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp, kin, kout, vin, vout, (int)n, 0, end_bit, c->stream));
    CK(c->d_cubtmp.ensure(tmp));
    CK(cub::DeviceRadixSort::SortPairs(c->d_cubtmp.p, tmp, kin, kout, vin, vout, (int)n, 0, end_bit, c->stream));
    return EJB_OK;
}
inline int bits_for(int64_t n) {   // bits needed for values 0 .. n
    int b = 1;
    while (b < 63 && (1ll << b) <= n) b++;
    return b;
}
int cub_sort_keys64(ejb_ctx* c, const uint64_t* kin, uint64_t* kout, int64_t n) {
    size_t tmp = 0;
    CK(cub::DeviceRadixSort::SortKeys(nullptr, tmp, kin, kout, (int)n, 0, 64, c->stream));
    CK(c->d_cubtmp.ensure(tmp));
    CK(cub::DeviceRadixSort::SortKeys(c->d_cubtmp.p, tmp, kin, kout, (int)n, 0, 64, c->stream));
    return EJB_OK;
}
int cub_incl_sum(ejb_ctx* c, const int32_t* in, int32_t* out, int64_t n) {
    size_t tmp = 0;
    CK(cub::DeviceScan::InclusiveSum(nullptr, tmp, in, out, (int)n, c->stream));
    CK(c->d_cubtmp.ensure(tmp));
    CK(cub::DeviceScan::InclusiveSum(c->d_cubtmp.p, tmp, in, out, (int)n, c->stream));
    return EJB_OK;
}

// One scheduler round, queued as ONE uninterrupted piece of stream work: stage 1 over `units`, the truncation of the units from
// the counted candidates (k_truncate), stage 2, the forest update (k_forest_round: Boruvka to convergence on the device), and a
// single read-back of the round's report.  Returns 1 when a work buffer overflowed (nothing committed; the caller splits the
// round and retries).
int run_round(ejb_ctx* c, std::vector<Unit>& units, std::vector<RoundTimes>& times) {
    const double th0 = now_ms();
    const unsigned long long eval_before = c->pairs_eval_seen;
    const SubVec& subs = c->subs;
    // ---- tasks grouped by W1; task ids number them in launch order ----
    // groups 0 .. MAXW1: k_stage1 per CDR3 width (0: no filter); groups MAXW1+1+W1: the THIN units of that width (at most
    // opt.thin rows: the first rounds of every sub-bucket), scored by k_stage1_thin
    constexpr int NG = 2 * (MAXW1 + 1);
    auto group_of = [&](const Unit& u, const SubBucket& sb) {
        return (sb.W1 >= 1 && u.r1 - u.r0 <= std::min(c->opt.thin, S1_THIN_ROWS)) ? MAXW1 + 1 + sb.W1 : sb.W1;
    };
    std::vector<RowTask> tasks[NG];
    std::vector<ejb_ctx::TaskMeta> metas[NG];
    int64_t nctas[NG] = {0};
    // column tiles per CTA, per launch: short strips keep the early rounds (a few rows of every sub-bucket) wide enough to fill
    // the GPU, long strips amortise the per-CTA prologue (task lookup, row fragment) once a launch has tens of thousands of CTAs
    int strip_of[NG];
    {
        int64_t ctas4[NG] = {0};
        for (const Unit& u : units) {
            const SubBucket& sb = subs[u.sub];
            const int nb = (sb.n + TILE - 1) / TILE;
            for (int rb = u.r0 / TILE; rb * TILE < u.r1; rb++) ctas4[group_of(u, sb)] += (nb - rb + 3) / 4;
        }
        for (int w = 0; w <= MAXW1; w++) strip_of[w] = ctas4[w] > 65536 ? S1_STRIP : (ctas4[w] > 16384 ? 8 : 4);
        for (int w = MAXW1 + 1; w < NG; w++) strip_of[w] = 8;   // thin: a CTA walks 1024 columns, four per thread
    }
    std::vector<UnitDesc> udesc(units.size());
    std::vector<int> ugroup(units.size());
    for (size_t ui = 0; ui < units.size(); ui++) {
        const Unit& u = units[ui];
        const SubBucket& sb = subs[u.sub];
        const int nb = (sb.n + TILE - 1) / TILE;
        const int grp = group_of(u, sb);
        const int strip = strip_of[grp];
        const size_t first = tasks[grp].size();
        for (int rb = u.r0 / TILE; rb * TILE < u.r1; rb++) {
            RowTask t;
            t.tid = 0;
            t.allow = (int32_t)std::min(u.allow, 2.0e9);
            t.rho_ref = (float)u.rho_ref;
            t.sub = u.sub;
            t.rb = rb;
            t.row_lo = std::max(u.r0 - rb * TILE, 0);
            t.row_hi = std::min(u.r1 - rb * TILE, TILE);
            t.cta0 = (int32_t)nctas[grp];
            tasks[grp].push_back(t);
            metas[grp].push_back({(int32_t)ui, rb * TILE + t.row_lo, rb * TILE + t.row_hi});
            nctas[grp] += (nb - rb + strip - 1) / strip;
        }
        UnitDesc& d = udesc[ui];
        d.sub = u.sub;
        d.t0 = (int32_t)first;
        d.nt = (int32_t)(tasks[grp].size() - first);
        d.r1 = u.r1;
        d.allow = u.allow;
        d.strict = u.strict;
        d.rho_ref = u.rho_ref;
        ugroup[ui] = grp;
    }
    c->task_meta.clear();
    std::vector<RowTask> all_tasks;
    size_t goff[NG + 1] = {0};
    for (int w = 0; w < NG; w++) {
        goff[w] = all_tasks.size();
        for (size_t i = 0; i < tasks[w].size(); i++) {
            tasks[w][i].tid = (int32_t)c->task_meta.size();
            c->task_meta.push_back(metas[w][i]);
            all_tasks.push_back(tasks[w][i]);
        }
    }
    for (size_t ui = 0; ui < units.size(); ui++) udesc[ui].t0 += (int32_t)goff[ugroup[ui]];
    const size_t total_tasks = all_tasks.size(), n_units = units.size();
    if (n_units > c->round_units_cap || total_tasks > c->round_tasks_cap) return fail(c, EJB_ERR_CUDA, "round report buffer too small (%zu units, %zu tasks)", n_units, total_tasks);
    // report layout of this round: [Counters | UnitOut x n_units | candidate count x total_tasks]
    char* const rbase = c->d_round.as<char>();
    UnitOut* const d_uout = reinterpret_cast<UnitOut*>(rbase + sizeof(Counters));
    unsigned long long* const d_ucand = reinterpret_cast<unsigned long long*>(rbase + sizeof(Counters) + n_units * sizeof(UnitOut));
    const size_t n_bands = total_tasks * S1_WARPS;   // candidate counters: one per 16-row band of every task
    const size_t report_bytes = sizeof(Counters) + n_units * sizeof(UnitOut) + n_bands * 8;
    const size_t zero_bytes = report_bytes + total_tasks;   // + one skip flag per task
    c->cur_round_tasks = total_tasks;
    c->cur_strict = units.empty() ? 0.0 : units[0].strict;   // the same for every unit of a run (target_pass)
    for (int w = 0; w < NG; w++)
        if (nctas[w] >= (1ll << 31)) return fail(c, EJB_ERR_INVALID, "stage-1 launch too large (%lld CTAs)", (long long)nctas[w]);
    cudaStream_t st = c->stream;
    CK(c->d_tasks.ensure(total_tasks * sizeof(RowTask) + 64));
    CK(c->d_units.ensure(n_units * sizeof(UnitDesc) + 64));
    const double th1 = now_ms();
    CK(cudaMemsetAsync(rbase, 0, offsetof(Counters, round_end), st));
    CK(cudaMemsetAsync(rbase + sizeof(Counters), 0, zero_bytes - sizeof(Counters), st));
    // tasks and unit descriptors go through page-locked staging: the copies are queued, not waited for
    CK(c->h_stage.ensure(total_tasks * sizeof(RowTask) + n_units * sizeof(UnitDesc) + 64));
    memcpy(c->h_stage.p, all_tasks.data(), total_tasks * sizeof(RowTask));
    memcpy(c->h_stage.as<char>() + total_tasks * sizeof(RowTask), udesc.data(), n_units * sizeof(UnitDesc));
#ifdef EJB_NO_STAGE
    CK(cudaMemcpyAsync(c->d_tasks.p, all_tasks.data(), total_tasks * sizeof(RowTask), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(c->d_units.p, udesc.data(), n_units * sizeof(UnitDesc), cudaMemcpyHostToDevice, st));
#else
    CK(cudaMemcpyAsync(c->d_tasks.p, c->h_stage.p, total_tasks * sizeof(RowTask), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(c->d_units.p, c->h_stage.as<char>() + total_tasks * sizeof(RowTask), n_units * sizeof(UnitDesc), cudaMemcpyHostToDevice, st));
#endif
    RoundTimes rt;
    rt.e0 = c->ev(); rt.e1 = c->ev(); rt.e2 = c->ev(); rt.e3 = c->ev();
    CK(cudaEventRecord(rt.e0, st));
    if (c->tags_on && (c->opt.s1_mode & S1_TAGS))
        launch(c, k_row_flags, nblk(c->n_pos, 256), 256, c->d_m_sub.as<const int32_t>(), c->d_m_row.as<const int32_t>(), c->n_pos, c->d_subdom.as<const SubDom>(),
               c->d_tagq.as<const unsigned long long>(), c->d_deadq.as<const unsigned long long>(), c->d_donq.as<const unsigned long long>(), c->d_flagq.as<uint32_t>());
    {
        size_t total_ctas = 0;
        for (int w = 1; w < NG; w++) total_ctas += (size_t)nctas[w];
        CK(c->d_ctatask.ensure(total_ctas * 4 + 64));
    }
    size_t cta_off = 0;
    for (int w = 0; w < NG; w++) {
        if (tasks[w].empty()) continue;
        const RowTask* dt = c->d_tasks.as<RowTask>() + goff[w];
        const int nt = (int)tasks[w].size();
        if (w > 0) {
            launch(c, k_expand_tasks, (unsigned int)nt, 64, dt, nt, c->d_subs.as<const SubBucket>(), strip_of[w], c->d_ctatask.as<int32_t>() + cta_off);
            c->cur_cta_off = cta_off;
            cta_off += (size_t)nctas[w];
        }
        switch (w) {
            case 0:
                launch(c, k_stage1_nofilter, (unsigned int)nctas[0], 256, c->d_subs.as<const SubBucket>(), dt, nt, c->d_labq.as<const int32_t>(), c->d_cand.as<uint2>(),
                       c->cand_cap, c->d_ctr(), d_ucand, strip_of[0]);
                break;
            case 1: launch_stage1_w<1>(c, nctas[1], dt, nt, d_ucand, strip_of[1]); break;
            case 2: launch_stage1_w<2>(c, nctas[2], dt, nt, d_ucand, strip_of[2]); break;
            case 3: launch_stage1_w<3>(c, nctas[3], dt, nt, d_ucand, strip_of[3]); break;
            case 4: launch_stage1_w<4>(c, nctas[4], dt, nt, d_ucand, strip_of[4]); break;
            case 5: launch_stage1_w<5>(c, nctas[5], dt, nt, d_ucand, strip_of[5]); break;
            case 6: launch_stage1_w<6>(c, nctas[6], dt, nt, d_ucand, strip_of[6]); break;
            case MAXW1 + 2: launch_stage1_thin<1>(c, nctas[w], dt, d_ucand, strip_of[w]); break;
            case MAXW1 + 3: launch_stage1_thin<2>(c, nctas[w], dt, d_ucand, strip_of[w]); break;
            case MAXW1 + 4: launch_stage1_thin<3>(c, nctas[w], dt, d_ucand, strip_of[w]); break;
            case MAXW1 + 5: launch_stage1_thin<4>(c, nctas[w], dt, d_ucand, strip_of[w]); break;
            case MAXW1 + 6: launch_stage1_thin<5>(c, nctas[w], dt, d_ucand, strip_of[w]); break;
            case MAXW1 + 7: launch_stage1_thin<6>(c, nctas[w], dt, d_ucand, strip_of[w]); break;
        }
        c->stats.stage1_launches++;
    }
    // truncate units whose counted candidates exceed their allowance (on the device: no read-back between stage 1 and stage 2)
    launch(c, k_truncate, nblk((int64_t)n_units * 32, 128), 128, c->d_units.as<const UnitDesc>(), (int)n_units, c->d_tasks.as<const RowTask>(),
           (const unsigned long long*)d_ucand, c->d_subs.as<const SubBucket>(), c->d_qcut.as<int32_t>(), c->d_subpass.as<unsigned int>(), d_uout, c->d_ctr());
    CK(cudaGetLastError());
    CK(cudaEventRecord(rt.e1, st));
    int rc = run_stage2(c, c->d_cand.as<uint2>(), nullptr, CTR_FIELD(n_cand), c->cand_cap, 0, nullptr);
    if (rc) return rc;
    CK(cudaEventRecord(rt.e2, st));
    {
        ForestArgs A;
        A.pass_w = c->d_pass.as<unsigned long long>();
        A.pass_t = c->d_passt.as<PassTuple>();
        A.n_pass_p = CTR_FIELD(n_pass);
        A.cap = c->cand_cap;
        A.parent = c->d_parent.as<int32_t>();
        A.best = c->d_best.as<unsigned long long>();
        A.roots = c->d_roots.as<int2>();
        A.forest = c->d_forest.as<unsigned long long>();
        A.forest_t = c->d_forestt.as<PassTuple>();
        A.forest_cap = c->forest_cap;
        A.ctr = c->d_ctr();
        A.bar = c->d_bar.as<unsigned int>();
        A.rowq = c->d_m_row.as<int32_t>();
        A.n_pos = c->n_pos;
        A.labq = c->d_labq.as<int32_t>();
        A.subs = c->d_subs.as<SubBucket>();
        A.blk_lab = c->d_blklab.as<int32_t>();
        A.blk_sub = c->d_blksub.as<int32_t>();
        A.n_blocks = (int)c->n_blocks;
        CK(cudaMemsetAsync(c->d_bar.p, 0, 4, st));
        void* args[] = {&A};
        CK(cudaLaunchCooperativeKernel((const void*)k_forest_round, dim3((unsigned int)c->coop_grid), dim3(256), args, 0, st));
        c->stats.kernel_launches++;
    }
    launch(c, k_round_report, nblk((int64_t)n_units, 128), 128, c->d_units.as<const UnitDesc>(), (int)n_units, c->d_subpass.as<const unsigned int>(), d_uout);
    CK(cudaEventRecord(rt.e3, st));
    CK(c->h_round.ensure(zero_bytes));
    CK(cudaMemcpyAsync(c->h_round.p, rbase, zero_bytes, cudaMemcpyDeviceToHost, st));
    const double th2 = now_ms();
    rc = sync_stream(c);
    if (rc) return rc;
    if (c->opt.trace)
        fprintf(stderr, "[ejb trace] round %lld host: plan %.3f ms, enqueue %.3f ms, wait %.3f ms (%zu units, %zu tasks)\n", (long long)c->stats.rounds, th1 - th0, th2 - th1,
                now_ms() - th2, n_units, total_tasks);
    c->stats.d2h_bytes += (int64_t)report_bytes;
    const Counters& h = *c->h_round.as<Counters>();
    const UnitOut* uo = reinterpret_cast<const UnitOut*>(c->h_round.as<char>() + sizeof(Counters));
    const unsigned long long* uc = reinterpret_cast<const unsigned long long*>(c->h_round.as<char>() + sizeof(Counters) + n_units * sizeof(UnitOut));
    c->h_unitcand.assign(total_tasks, 0ull);
    for (size_t t = 0; t < total_tasks; t++)
        for (int b = 0; b < S1_WARPS; b++) c->h_unitcand[t] += uc[t * S1_WARPS + b];
    {   // tasks that returned early behind a dense block were not counted: their numbers say nothing about those rows
        const uint8_t* skipped = reinterpret_cast<const uint8_t*>(uc + n_bands);
        c->h_taskskip.assign(skipped, skipped + total_tasks);
    }
    c->p1_count += h.n_newkeys;
    c->stats.p1_unique += (int64_t)h.n_newkeys;
    c->p1_groups = h.p1_groups;
    c->p1_pool_used = h.p1_pool_used;
    if (h.p1_full) return fail(c, EJB_ERR_OOM, "p1 cache full (%llu keys in %u slots)", c->p1_count, c->p1_cap);
    if ((rc = p1_groups_grow(c, h.p1_desc)) != EJB_OK) return rc;
    if (h.overflow == 2ull) return fail(c, EJB_ERR_CUDA, "forest buffer overflow");
    if (h.overflow || h.n_cand > c->cand_cap || h.n_surv > c->cand_cap || h.n_pass > c->cand_cap || h.n_slow > c->cand_cap) {
        if (h.pairs_eval > eval_before) c->cand_rate = std::max(c->cand_rate, (double)h.n_cand / (double)(h.pairs_eval - eval_before));
        c->pairs_eval_seen = h.pairs_eval;
        c->stats.overflow_retries++;
        if (c->p1_count > c->p1_cap / 4 && (rc = p1_cache_grow(c)) != EJB_OK) return rc;
        return 1;
    }
    if (h.n_errflag_range) return fail(c, EJB_ERR_RANGE, "k = indeps + 2*shares exceeds the Stirling table (k > 3000); the reference panics here (join_one.rs:30)");
    if (h.n_errflag_panic) return fail(c, EJB_ERR_INVALID, "input on which the reference panics (J range longer than a contig, seg shorter than ref trim, or inconsistent feature starts)");
    if (h.alive_it[BOR_MAX_ITERS]) return fail(c, EJB_ERR_CUDA, "forest did not converge");
    if (h.pairs_eval > eval_before) c->cand_rate = (double)h.n_cand / (double)(h.pairs_eval - eval_before);   // most recent round
    c->pairs_eval_seen = h.pairs_eval;
    c->stats.stage1_candidates += (int64_t)h.n_cand;
    c->stats.stage2_slow += (int64_t)h.n_slow;
    c->stats.stage2_survivors += (int64_t)h.n_surv_total;
    c->stats.pass_edges += (int64_t)h.n_pass;
    for (size_t ui = 0; ui < n_units; ui++) {
        units[ui].r1 = uo[ui].r1;
        units[ui].cand = uo[ui].cand;
        units[ui].pass = (double)uo[ui].pass;
        units[ui].rho_last = uo[ui].rho_last;
    }
    times.push_back(rt);
    if (c->opt.trace) {   // developer aid: what the round was made of
        size_t top = 0;
        for (size_t i = 1; i < units.size(); i++)
            if (units[i].cand > units[top].cand) top = i;
        fprintf(stderr, "[ejb trace] round %lld: %zu units, cand %llu surv %llu pass %llu; top unit: sub %d (n %d) rows [%d,%d) cand %.0f pass %.0f\n",
                (long long)c->stats.rounds, units.size(), h.n_cand, h.n_surv_total, h.n_pass, units.empty() ? -1 : units[top].sub,
                units.empty() ? 0 : subs[units[top].sub].n, units.empty() ? 0 : units[top].r0, units.empty() ? 0 : units[top].r1,
                units.empty() ? 0.0 : units[top].cand, units.empty() ? 0.0 : units[top].pass);
    }
    c->stats.rounds++;
    if (c->p1_count > c->p1_cap / 4 && (rc = p1_cache_grow(c)) != EJB_OK) return rc;
    return EJB_OK;
}

// Multi-GPU ownership.  Sub-buckets are independent units (join.rs:70-88), so ranks shard them with no collective on the data
// path.  Longest-processing-time assignment on a cost model (a row acting as first row costs the rows to its right + 64); a
// sub-bucket whose cost exceeds `split_above` x a rank's share is cut into row ranges of about `piece` x the share (each part =
// the pairs whose FIRST row lies in the range; it returns the lexicographic spanning forest of ITS pass edges, and the merge
// keeps the minimum spanning forest of the union of the partial forests: MSF(G) = MSF(MSF(G1) u MSF(G2) u ...)).  Parts of one
// sub-bucket go to different ranks.  Deterministic: every rank computes the same assignment from the same table.
void shard_assign(SubBucket* subs, int n_subs, int rank, int world, double split_above = 0.25, double piece = 0.125) {
    struct Item { double cost; int sub, lo, hi, nparts; };
    auto cum = [](double n, double r) { return r * (n - 1.0) - r * (r - 1.0) / 2.0 + 64.0 * r; };   // cost of first rows [0, r)
    double total = 0.0;
    for (int s = 0; s < n_subs; s++)
        if (subs[s].n >= 2) total += cum(subs[s].n, subs[s].n);
    const double share = total / std::max(world, 1);
    // the items that matter for the balance are sorted (LPT); the many small ones (each below 1/256 of a share) follow in table
    // order, each to the least loaded rank: the same bound on the imbalance without sorting ten thousand items
    std::vector<Item> big, small;
    small.reserve((size_t)n_subs);
    for (int s = 0; s < n_subs; s++) {
        const SubBucket& sb = subs[s];
        if (sb.n < 2) continue;
        const double c = cum(sb.n, sb.n);
        if (sb.n >= 1024 && c > split_above * share) {
            const int parts = (int)std::min<double>(world, std::ceil(c / (piece * share)));
            std::vector<int> cuts;
            cuts.push_back(0);
            for (int i = 1; i < parts; i++) {   // smallest r with cum(r) >= c i / parts
                int lo = 0, hi = sb.n;
                const double want = c * i / parts;
                while (lo < hi) {
                    const int mid = (lo + hi) / 2;
                    if (cum(sb.n, mid) >= want) hi = mid; else lo = mid + 1;
                }
                if (lo > cuts.back() && lo < sb.n) cuts.push_back(lo);
            }
            cuts.push_back(sb.n);
            for (size_t i = 0; i + 1 < cuts.size(); i++)
                big.push_back({cum(sb.n, cuts[i + 1]) - cum(sb.n, cuts[i]), s, cuts[i], cuts[i + 1], (int)cuts.size() - 1});
        } else if (c * 256.0 >= share) {
            big.push_back({c, s, 0, sb.n, 1});
        } else {
            small.push_back({c, s, 0, sb.n, 1});
        }
    }
    std::stable_sort(big.begin(), big.end(), [](const Item& a, const Item& b) { return a.cost > b.cost; });
    std::vector<double> load((size_t)world, 0.0);
    std::vector<uint64_t> holder((size_t)n_subs, 0);   // ranks already holding a part of the sub-bucket (world <= 64)
    for (int s = 0; s < n_subs; s++) {
        subs[s].owner = -1;
        subs[s].r_beg = subs[s].n_act = 0;
    }
    auto place = [&](const Item& it) {
        int best = -1;
        for (int r = 0; r < world; r++) {
            if (holder[(size_t)it.sub] >> r & 1) continue;
            if (best < 0 || load[(size_t)r] < load[(size_t)best]) best = r;
        }
        if (best < 0) best = 0;   // cannot happen: parts <= world
        holder[(size_t)it.sub] |= 1ull << best;
        load[(size_t)best] += it.cost;
        if (best == rank) {
            subs[it.sub].r_beg = it.lo;
            subs[it.sub].n_act = it.hi;
            subs[it.sub].owner = it.nparts == 1 ? rank : -1;
        }
    };
    for (const Item& it : big) place(it);
    // small items (never split): the least loaded rank keeps receiving them until it is no longer the least loaded
    int best = 0;
    double next_load = 0.0;
    auto pick = [&]() {
        best = 0;
        for (int r = 1; r < world; r++)
            if (load[(size_t)r] < load[(size_t)best]) best = r;
        next_load = 1e300;
        for (int r = 0; r < world; r++)
            if (r != best) next_load = std::min(next_load, load[(size_t)r]);
    };
    pick();
    for (const Item& it : small) {
        if (load[(size_t)best] > next_load) pick();
        load[(size_t)best] += it.cost;
        if (best == rank) {
            subs[it.sub].r_beg = it.lo;
            subs[it.sub].n_act = it.hi;
            subs[it.sub].owner = rank;
        }
    }
}

// ---- prepare: bucket scan, sub-bucket table, bucketer.  ONE read-back (the run header + the sub-bucket table). ----
int prepare(ejb_ctx* c, std::vector<std::pair<const char*, cudaEvent_t>>& tr) {
    const DevIn& din = c->din;
    const int64_t N = din.n_rows;
    cudaStream_t st = c->stream;
    auto mark = [&](const char* name) {
        if (!c->opt.trace) return;
        cudaEvent_t e = c->ev();
        cudaEventRecord(e, st);
        tr.emplace_back(name, e);
    };
    mark("start");
    RunHdr* const hdr = c->d_err.as<RunHdr>();
    CK(cudaMemsetAsync(c->d_err.p, 0, 256, st));
    // Every join evaluates the p1 values it needs (what a caller that joins once per process sees, and what a benchmark must
    // time): the memo of p1(n, k, d) and its term groups only live for one join.
    {
        int rc = p1_cache_clear(c);
        if (rc) return rc;
    }
    c->have_roles = false;
    if (!c->roles.empty()) {
        if ((int64_t)c->roles.size() != N) return fail(c, EJB_ERR_INVALID, "row roles were set for %zu rows, the input has %lld", c->roles.size(), (long long)N);
        CK(c->d_roles.ensure((size_t)N));
        CK(cudaMemcpyAsync(c->d_roles.p, c->roles.data(), (size_t)N, cudaMemcpyHostToDevice, st));
        c->have_roles = true;
    }
    c->subs.clear();
    c->n_active = c->n_pos = c->n_blocks = c->n_final = 0;
    if (N == 0) return EJB_OK;
    constexpr int64_t SPEC = 16384;   // sub-bucket records fetched together with the header (more need a second copy)
    // ---- bucket scan + sub-bucket sort (join.rs:68-88) ----
    CK(c->d_head.ensure((size_t)N * 4));
    CK(c->d_scan.ensure((size_t)N * 4));
    CK(c->d_keys.ensure((size_t)N * 8));
    CK(c->d_keys2.ensure((size_t)N * 8));
    CK(c->d_vals.ensure((size_t)N * 4));
    CK(c->d_vals2.ensure((size_t)N * 4));
    launch(c, k_heads, nblk(N, 256), 256, din, c->d_head.as<int32_t>(), &hdr->err);
    const int64_t tmax = std::max({din.n_exacts, din.n_chains, din.n_cells, din.n_segs, din.n_refs});
    launch(c, k_validate_tables, nblk(tmax, 256), 256, din, &hdr->err);
    int rc = cub_incl_sum(c, c->d_head.as<int32_t>(), c->d_scan.as<int32_t>(), N);
    if (rc) return rc;
    launch(c, k_keys, nblk(N, 256), 256, din, c->d_scan.as<const int32_t>(), c->d_keys.as<uint64_t>(), c->d_vals.as<uint32_t>(), hdr);
    // keys: (bucket id << 32) | (c_h << 16) | c_l, bucket id < N; rows that never join carry ~0 (one extra bit keeps them last)
    rc = cub_sort_pairs64(c, c->d_keys.as<uint64_t>(), c->d_keys2.as<uint64_t>(), c->d_vals.as<uint32_t>(), c->d_vals2.as<uint32_t>(), N, std::min(64, 33 + bits_for(N)));
    if (rc) return rc;
    mark("scan+sort");
    const uint64_t* keys_sorted = c->d_keys2.as<uint64_t>();
    const uint32_t* row_of = c->d_vals2.as<uint32_t>();
    // work that only needs the raw input: per exact donor masks + sorted barcode lists, per segment reference planes
    CK(c->d_donor.ensure((size_t)std::max<int64_t>(din.n_exacts, 1) * 8));
    CK(c->d_bckeys.ensure((size_t)std::max<int64_t>(din.n_cells, 1) * 8));
    CK(c->d_bckeys2.ensure((size_t)std::max<int64_t>(din.n_cells, 1) * 8));
    launch(c, k_exact_prep, nblk(din.n_exacts, 128), 128, din, c->d_donor.as<unsigned long long>(), c->d_bckeys.as<uint64_t>(), c->d_bckeys2.as<uint64_t>(), hdr);
    CK(c->d_segpl.ensure((size_t)std::max<int64_t>(din.n_segs, 1) * 2 * SEGPL_WORDS * 4));
    CK(c->d_seglen.ensure((size_t)std::max<int64_t>(din.n_segs, 1) * 4));
    launch(c, k_seg_planes, nblk(din.n_segs * 32, 256), 256, din, c->d_segpl.as<uint32_t>(), c->d_seglen.as<int32_t>());
    // sub-bucket heads, ids and records; the number of active rows stays on the device (kernels run over all N rows)
    CK(c->d_subhead.ensure((size_t)N * 4));
    CK(c->d_subincl.ensure((size_t)N * 4));
    CK(c->d_subinfo.ensure((size_t)N * sizeof(SubInfo)));
    launch(c, k_subheads, nblk(N, 256), 256, keys_sorted, (const RunHdr*)hdr, N, c->d_subhead.as<int32_t>());
    rc = cub_incl_sum(c, c->d_subhead.as<int32_t>(), c->d_subincl.as<int32_t>(), N);
    if (rc) return rc;
    CK(cudaMemsetAsync(c->d_subinfo.p, 0, (size_t)N * sizeof(SubInfo), st));
    launch(c, k_subinfo, nblk(N, 256), 256, din, keys_sorted, row_of, c->d_subhead.as<const int32_t>(), c->d_subincl.as<const int32_t>(), hdr, c->d_subinfo.as<SubInfo>());
    if (c->have_roles) {
        CK(c->d_subroles.ensure((size_t)N * sizeof(SubRole)));
        CK(cudaMemsetAsync(c->d_subroles.p, 0, (size_t)N * sizeof(SubRole), st));
        launch(c, k_subroles, nblk(N, 256), 256, c->d_roles.as<const uint8_t>(), row_of, c->d_subhead.as<const int32_t>(), c->d_subincl.as<const int32_t>(),
               (const RunHdr*)hdr, c->d_subroles.as<SubRole>(), &hdr->err);
    }
    const int64_t spec = std::min<int64_t>(N, SPEC);
    const size_t off_si = 256, off_sr = off_si + (size_t)spec * sizeof(SubInfo);
    CK(c->h_prep.ensure(off_sr + (size_t)spec * sizeof(SubRole)));
    CK(cudaMemcpyAsync(c->h_prep.p, c->d_err.p, sizeof(RunHdr), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(c->h_prep.as<char>() + off_si, c->d_subinfo.p, (size_t)spec * sizeof(SubInfo), cudaMemcpyDeviceToHost, st));
    if (c->have_roles) CK(cudaMemcpyAsync(c->h_prep.as<char>() + off_sr, c->d_subroles.p, (size_t)spec * sizeof(SubRole), cudaMemcpyDeviceToHost, st));
    mark("subinfo");
    rc = sync_stream(c);
    if (rc) return rc;
    const double th0 = now_ms();
    const RunHdr hh = *c->h_prep.as<RunHdr>();
    if (hh.err & 16384u) return fail(c, EJB_ERR_INVALID, "row roles: column-only rows must form a suffix of their sub-bucket");
    if (hh.err) return fail(c, EJB_ERR_INVALID, "input validation failed (mask 0x%x: 1 nchains, 2 exact id, 4 tig offsets, 8 cdr3 offsets, 16 seg ids, 32 exact_cols, 64/128 exact tables, 256 ref ids, 512 amino, 1024 donor id >= 64, 2048 seg_off, 4096 ref_off, 8192 packed contig)", hh.err);
    c->stats.n_buckets = (int64_t)hh.n_buckets;
    c->n_active = (int64_t)hh.n_active;
    if (hh.big_exacts && din.n_cells > 0) {   // an exact with many cells: one global sort of all barcode lists instead of the per-thread sorts
        rc = cub_sort_keys64(c, c->d_bckeys.as<uint64_t>(), c->d_bckeys2.as<uint64_t>(), din.n_cells);
        if (rc) return rc;
    }
    const int64_t NA = c->n_active;
    const int n_subs = (int)hh.n_subs;
    c->n_buckets = (int64_t)hh.n_buckets;
    CK(c->d_bstart.ensure((size_t)std::max<int64_t>(c->n_buckets, 1) * 4));
    launch(c, k_bucket_starts, nblk(N, 256), 256, c->d_head.as<const int32_t>(), c->d_scan.as<const int32_t>(), N, c->d_bstart.as<int32_t>());
    if (NA == 0) return EJB_OK;
    std::vector<SubInfo> si((size_t)n_subs);
    std::vector<SubRole> sroles;
    memcpy(si.data(), c->h_prep.as<char>() + off_si, (size_t)std::min<int64_t>(n_subs, spec) * sizeof(SubInfo));
    if (c->have_roles) {
        sroles.resize((size_t)n_subs);
        memcpy(sroles.data(), c->h_prep.as<char>() + off_sr, (size_t)std::min<int64_t>(n_subs, spec) * sizeof(SubRole));
    }
    if (n_subs > spec) {   // rare: more sub-buckets than the speculative read covered
        CK(cudaMemcpyAsync(si.data() + spec, c->d_subinfo.as<SubInfo>() + spec, (size_t)(n_subs - spec) * sizeof(SubInfo), cudaMemcpyDeviceToHost, st));
        if (c->have_roles) CK(cudaMemcpyAsync(sroles.data() + spec, c->d_subroles.as<SubRole>() + spec, (size_t)(n_subs - spec) * sizeof(SubRole), cudaMemcpyDeviceToHost, st));
        rc = sync_stream(c);
        if (rc) return rc;
    }
    c->stats.d2h_bytes += (int64_t)n_subs * sizeof(SubInfo);

    // ---- sub-bucket table: sizes and thresholds first, then ownership (multi-GPU), then offsets of what this context packs ----
    const DevParams P = make_dev_params(c->params);
    SubVec& subs = c->subs;
    subs.resize(n_subs);
    int max_cdr3 = 0;
    for (int s = 0; s < n_subs; s++) {
        SubBucket& sb = subs[s];
        sb.j0 = si[s].j0;
        sb.n = (int32_t)(((s + 1 < n_subs) ? si[s + 1].j0 : NA) - si[s].j0);
        sb.npad = (sb.n + 3) & ~3;
        sb.ch = (int32_t)((si[s].key >> 16) & 0xffff);
        sb.cl = (int32_t)(si[s].key & 0xffff);
        sb.Lh = si[s].Lh;
        sb.Ll = si[s].Ll;
        const int T = sb.ch + sb.cl;
        const int w1 = (T + 31) / 32;
        sb.W1 = (w1 >= 1 && w1 <= MAXW1) ? w1 : 0;
        // uint4 per plane: 4 up to 512 nt (compile-time stride on the common path), 8 up to 1024 nt, else byte path only
        sb.W4h = (sb.Lh >= 1 && sb.Lh <= 512) ? 4 : (sb.Lh <= 128 * MAXW4 && sb.Lh >= 1) ? MAXW4 : 0;
        sb.W4l = (sb.Ll >= 1 && sb.Ll <= 512) ? 4 : (sb.Ll <= 128 * MAXW4 && sb.Ll >= 1) ? MAXW4 : 0;
        sb.gap = si[s].gap & 3;
        sb.rec_u4 = rec_planes(0, sb.gap & 1) * sb.W4h + rec_planes(1, sb.gap & 2) * sb.W4l;
        sb.rot = (sb.W1 >= 2 && sb.ch > 32) ? (sb.ch - 32) / 2 : 0;   // word 0 = the middle of the heavy junction
        sb.pad2 = 0;
        sb.r_beg = 0;
        sb.n_act = c->have_roles ? sroles[(size_t)s].n_act : sb.n;
        sb.raw = c->have_roles ? sroles[(size_t)s].raw : 0;
        sb.n3 = 3 * (sb.Lh + sb.Ll);
        sb.owner = 0;
        // stage-1 threshold: a superset of join_one.rs:231 and :286-290 (stage 2 re-tests exactly)
        int maxcd = T;
        if (P.is_bcr) {
            // the largest cd with !(cd / T > thr): the predicate is monotone in cd, so start just above thr * T
            maxcd = 0;
            int cd_hi = T;
            if (P.cdr3_ident_thr >= 0.0 && P.cdr3_ident_thr < 1.0) cd_hi = std::min(T, (int)(P.cdr3_ident_thr * (double)T) + 2);
            for (int cd = cd_hi; cd >= 0; cd--)
                if (!((double)cd / (double)T > P.cdr3_ident_thr)) { maxcd = cd; break; }
            if (T == 0) maxcd = 0;
            if (P.max_cdr3_diffs < 1000) maxcd = (int)std::min<long long>(maxcd, P.max_cdr3_diffs);
        } else {
            maxcd = 0;  // TCR: cd > 0 rejects
        }
        sb.maxcd = maxcd;
        max_cdr3 = std::max({max_cdr3, sb.ch, sb.cl});
    }
    const double th1 = now_ms();
    // ---- multi-GPU ownership (identical on every rank: same table, same deterministic assignment) ----
    if (c->shard_world > 1) shard_assign(subs.data(), n_subs, c->shard_rank, c->shard_world);
    const double th2 = now_ms();
    int64_t q = 0, s1 = 0, rs = 0, blk = 0;
    int64_t bucket_rows = 0;
    uint64_t cur_bucket = ~0ull;
    std::vector<int64_t> pack_off;
    std::vector<int32_t> pack_sub;
    pack_off.push_back(0);
    for (int s = 0; s < n_subs; s++) {
        SubBucket& sb = subs[s];
        const bool packed = sb.n_act > sb.r_beg || (c->shard_world == 1);   // single context: everything is packed (labels cover all rows)
        sb.q0 = q;
        sb.s1_off = s1;
        sb.rs_off = rs;
        sb.blk0 = (int32_t)blk;
        if (packed) {
            q += sb.npad;
            s1 += (int64_t)2 * sb.W1 * sb.npad;
            rs += (int64_t)sb.n * sb.rec_u4;
            blk += (sb.n + TILE - 1) / TILE;
            pack_sub.push_back(s);
            pack_off.push_back(pack_off.back() + sb.n);
        } else {
            sb.r_beg = sb.n_act = 0;
        }
        // pairs whose first row is one of the rows [r_beg, n_act) (all n(n-1)/2 when the whole sub-bucket is joined here)
        {
            const int64_t a0 = sb.r_beg, a1 = sb.n_act, m = a1 - a0;
            if (m > 0) c->stats.pairs_scored += m * (sb.n - 1) - (a0 + a1 - 1) * m / 2;
        }
        const uint64_t b = si[s].key >> 32;
        if (b != cur_bucket) {
            c->stats.pairs_lens += bucket_rows * (bucket_rows - 1) / 2;
            bucket_rows = 0;
            cur_bucket = b;
        }
        bucket_rows += sb.n;
    }
    c->stats.pairs_lens += bucket_rows * (bucket_rows - 1) / 2;
    c->stats.n_subbuckets = n_subs;
    c->n_pos = q;
    c->n_blocks = blk;
    if (q >= (1ll << 31)) return fail(c, EJB_ERR_INVALID, "too many packed rows");
    if (rs >= (1ll << 32) || s1 >= (1ll << 32)) return fail(c, EJB_ERR_INVALID, "packed row store exceeds 32-bit indexing (64 GB)");

    const double th3 = now_ms();
    // ---- pow table: mult_pow ^ (cdr3_normal_len * cd / n), join_one.rs:531-539 ----
    std::vector<int32_t> pow_off((size_t)max_cdr3 + 2, 0);
    std::vector<char> len_used((size_t)max_cdr3 + 1, 0);
    for (auto& sb : subs) len_used[sb.ch] = len_used[sb.cl] = 1;
    std::vector<double> powtab;
    // rows of the table are kept across joins (std::pow is ~25 ns a call and a repertoire uses ~100 CDR3 lengths x their length)
    if (c->pow_base != c->params.mult_pow || c->pow_norm != (double)c->params.cdr3_normal_len) {
        c->pow_rows.clear();
        c->pow_base = c->params.mult_pow;
        c->pow_norm = (double)c->params.cdr3_normal_len;
    }
    for (int n = 0; n <= max_cdr3; n++) {
        pow_off[n] = (int32_t)powtab.size();
        if (!len_used[n]) continue;
        std::vector<double>& row = c->pow_rows[n];
        if (row.empty())
            for (int cd = 0; cd <= n; cd++) row.push_back(std::pow(c->params.mult_pow, (double)c->params.cdr3_normal_len * (double)cd / (double)n));
        powtab.insert(powtab.end(), row.begin(), row.end());
    }
    if (powtab.empty()) powtab.push_back(1.0);
    std::vector<int32_t> blk_sub((size_t)blk);
    for (int s : pack_sub)
        for (int b = 0; b < (subs[s].n + TILE - 1) / TILE; b++) blk_sub[subs[s].blk0 + b] = s;
    // small tables go through one pinned staging buffer (a copy from pageable memory would synchronise the stream first)
    const size_t o_pow = 0, o_poff = o_pow + ((powtab.size() * 8 + 15) & ~(size_t)15), o_blk = o_poff + ((pow_off.size() * 4 + 15) & ~(size_t)15),
                 o_poffs = o_blk + ((blk_sub.size() * 4 + 15) & ~(size_t)15), o_psub = o_poffs + ((pack_off.size() * 8 + 15) & ~(size_t)15),
                 o_end = o_psub + pack_sub.size() * 4 + 16;
    CK(c->h_tab.ensure(o_end));
    char* const ht = c->h_tab.as<char>();
    memcpy(ht + o_pow, powtab.data(), powtab.size() * 8);
    memcpy(ht + o_poff, pow_off.data(), pow_off.size() * 4);
    memcpy(ht + o_blk, blk_sub.data(), blk_sub.size() * 4);
    memcpy(ht + o_poffs, pack_off.data(), pack_off.size() * 8);
    memcpy(ht + o_psub, pack_sub.data(), pack_sub.size() * 4);
    CK(c->d_powtab.ensure(powtab.size() * 8));
    CK(c->d_powoff.ensure(pow_off.size() * 4));
    CK(cudaMemcpyAsync(c->d_powtab.p, ht + o_pow, powtab.size() * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(c->d_powoff.p, ht + o_poff, pow_off.size() * 4, cudaMemcpyHostToDevice, st));

    const double th4 = now_ms();
    // ---- device arrays ----
    CK(c->d_qcut.ensure((size_t)n_subs * 4 + 4));
    CK(c->d_subpass.ensure((size_t)n_subs * 4 + 4));
    CK(cudaMemsetAsync(c->d_qcut.p, 0x7f, (size_t)n_subs * 4, st));   // 0x7f7f7f7f: no cut
    CK(cudaMemsetAsync(c->d_subpass.p, 0, (size_t)n_subs * 4, st));
    CK(c->d_subs.ensure((size_t)n_subs * sizeof(SubBucket)));
    CK(cudaMemcpyAsync(c->d_subs.p, subs.data(), (size_t)n_subs * sizeof(SubBucket), cudaMemcpyHostToDevice, st));
    CK(c->d_blksub.ensure((size_t)std::max<int64_t>(blk, 1) * 4));
    CK(cudaMemcpyAsync(c->d_blksub.p, ht + o_blk, (size_t)blk * 4, cudaMemcpyHostToDevice, st));
    CK(c->d_cdr3w.ensure((size_t)(s1 + 4) * 4));
    CK(c->d_rowstore.ensure((size_t)(rs + 1) * 16));
    const size_t np = (size_t)c->n_pos + 4;
    CK(c->d_m_row.ensure(np * 4)); CK(c->d_m_sub.ensure(np * 4)); CK(c->d_aux.ensure(np * AUX_INTS * 4));
    CK(c->d_labq.ensure(np * 4)); CK(c->d_blklab.ensure((size_t)std::max<int64_t>(blk, 1) * 4));
    CK(cudaMemsetAsync(c->d_m_row.p, 0xff, np * 4, st));   // -1 = padding position
    CK(cudaMemsetAsync(c->d_labq.p, 0xff, np * 4, st));
    CK(cudaMemsetAsync(c->d_cdr3w.p, 0, (size_t)(s1 + 4) * 4, st));
    CK(c->d_memo.ensure(np * 32));
    CK(cudaMemsetAsync(c->d_memo.p, 0, np * 32, st));
    c->tags_on = c->n_canon_segs < 65535 && !c->opt.no_tags && (c->opt.s1_mode & S1_TAGS);
    if (c->tags_on) {
        CK(c->d_tagq.ensure(np * 8)); CK(c->d_deadq.ensure(np * 8)); CK(c->d_donq.ensure(np * 8));
        CK(cudaMemsetAsync(c->d_donq.p, 0, np * 8, st));       // padding positions
        CK(c->d_flagq.ensure(np * 4));
        CK(c->d_subdom.ensure((size_t)n_subs * sizeof(SubDom)));
        CK(cudaMemsetAsync(c->d_tagq.p, 0xff, np * 8, st));    // padding positions: never part of a valid pair
        CK(cudaMemsetAsync(c->d_deadq.p, 0xff, np * 8, st));   // DEAD_NONE
    }
    mark("tables+memsets+bcsort");
    const int n_pack = (int)pack_sub.size();
    CK(c->d_packoff.ensure(pack_off.size() * 8));
    CK(c->d_packsub.ensure((size_t)std::max(n_pack, 1) * 4));
    CK(cudaMemcpyAsync(c->d_packoff.p, ht + o_poffs, pack_off.size() * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(c->d_packsub.p, ht + o_psub, (size_t)n_pack * 4, cudaMemcpyHostToDevice, st));
    if (n_pack > 0) {
        // fast path first (8 lanes per row); the rows it cannot take are listed and packed by the generic warp-per-row kernel
        CK(c->d_slowrows.ensure((size_t)pack_off.back() * 4 + 4));
        unsigned long long* const n_slow_rows = reinterpret_cast<unsigned long long*>(c->d_err.as<char>() + 64);   // zeroed with the run header
        unsigned long long* const tq = c->tags_on ? c->d_tagq.as<unsigned long long>() : (unsigned long long*)nullptr;
        unsigned long long* const dq = c->tags_on ? c->d_donq.as<unsigned long long>() : (unsigned long long*)nullptr;
        launch(c, k_pack_rows8, nblk(((pack_off.back() + 3) / 4) * 32, 256), 256, din, P, c->d_subs.as<const SubBucket>(), c->d_packoff.as<const int64_t>(),
               c->d_packsub.as<const int32_t>(), n_pack, row_of, c->d_donor.as<const unsigned long long>(), c->d_segpl.as<const uint32_t>(),
               c->d_seglen.as<const int32_t>(), c->d_cdr3w.as<uint32_t>(), c->d_rowstore.as<uint4>(), c->d_aux.as<int32_t>(), c->d_m_row.as<int32_t>(),
               c->d_m_sub.as<int32_t>(), tq, dq, c->d_slowrows.as<uint32_t>(), n_slow_rows);
        launch(c, k_pack_rows, c->sm_count * 8, 256, din, P, c->d_subs.as<const SubBucket>(), c->d_packoff.as<const int64_t>(),
               c->d_packsub.as<const int32_t>(), n_pack, row_of, c->d_donor.as<const unsigned long long>(), c->d_segpl.as<const uint32_t>(),
               c->d_seglen.as<const int32_t>(), c->d_cdr3w.as<uint32_t>(), c->d_rowstore.as<uint4>(), c->d_aux.as<int32_t>(), c->d_m_row.as<int32_t>(),
               c->d_m_sub.as<int32_t>(), tq, dq, c->d_slowrows.as<const uint32_t>(), (const unsigned long long*)n_slow_rows);
    }
    if (c->tags_on)
        launch(c, k_sub_dominant, nblk(n_subs, 128), 128, c->d_subs.as<const SubBucket>(), n_subs, c->d_tagq.as<const unsigned long long>(),
               c->d_donq.as<const unsigned long long>(), c->d_subdom.as<SubDom>());
    mark("pack_rows");
    if (c->opt.trace)
        fprintf(stderr, "[ejb trace] prepare host: table %.3f ms, shard plan %.3f ms, offsets %.3f ms, pow %.3f ms, uploads + enqueue %.3f ms\n", th1 - th0, th2 - th1,
                th3 - th2, th4 - th3, now_ms() - th4);
    CK(cudaGetLastError());
    return EJB_OK;
}

}  // namespace

namespace {
// ---- finalize: two-cell filter, ordering, PotentialJoin members, clonotype_id unions, labels.  Every count stays on the
//      device; the only read-back is the final counter block (queued by the caller). ----
int finalize(ejb_ctx* c) {
    const DevIn& din = c->din;
    const int64_t N = din.n_rows;
    cudaStream_t st = c->stream;
    int rc = EJB_OK;
    CK(c->d_labels.ensure((size_t)std::max<int64_t>(N, 1) * 4));
    const int fgrid = c->sm_count * 8;
    if (N > 0) {
        const size_t nf = (size_t)c->forest_cap;
        CK(c->d_deg.ensure((size_t)N * 4));
        CK(c->d_fkeys.ensure(nf * 8)); CK(c->d_fkeys2.ensure(nf * 8));
        CK(c->d_fidx.ensure(nf * 4)); CK(c->d_fidx2.ensure(nf * 4));
        CK(cudaMemsetAsync(c->d_deg.p, 0, (size_t)N * 4, st));
        CK(cudaMemsetAsync(c->d_fkeys.p, 0xff, nf * 8, st));   // unused tail sorts to the end
        launch(c, k_forest_deg, fgrid, 256, c->d_forest.as<const unsigned long long>(), (const unsigned long long*)CTR_FIELD(n_forest), c->d_deg.as<int32_t>());
        launch(c, k_two_cell, fgrid, 256, din, make_dev_params(c->params), c->d_forest.as<const unsigned long long>(), (const unsigned long long*)CTR_FIELD(n_forest),
               c->d_deg.as<const int32_t>(), c->d_forestt.as<const PassTuple>(), c->d_parent.as<int32_t>(), c->d_fkeys.as<unsigned long long>(),
               c->d_fidx.as<uint32_t>(), c->d_ctr(), c->have_roles ? c->d_roles.as<const uint8_t>() : (const uint8_t*)nullptr);
        // keys: (k1 << 28) | k2 with k1 < N; the unused tail is ~0 (one extra bit keeps it last)
        rc = cub_sort_pairs64(c, c->d_fkeys.as<uint64_t>(), c->d_fkeys2.as<uint64_t>(), c->d_fidx.as<uint32_t>(), c->d_fidx2.as<uint32_t>(), (int64_t)nf, std::min(64, 29 + bits_for(N)));
        if (rc) return rc;
        CK(c->d_o_k1.ensure(nf * 4)); CK(c->d_o_k2.ensure(nf * 4)); CK(c->d_o_cd.ensure(nf * 4)); CK(c->d_o_nrefs.ensure(nf * 4));
        CK(c->d_o_sh.ensure(nf * 8)); CK(c->d_o_in.ensure(nf * 8)); CK(c->d_o_k.ensure(nf * 4)); CK(c->d_o_d.ensure(nf * 4));
        CK(c->d_o_n.ensure(nf * 4)); CK(c->d_o_p1.ensure(nf * 8)); CK(c->d_o_mult.ensure(nf * 8)); CK(c->d_o_score.ensure(nf * 8));
        CK(c->d_o_err.ensure(nf));
        FinalOut fo;
        fo.k1 = c->d_o_k1.as<int32_t>(); fo.k2 = c->d_o_k2.as<int32_t>(); fo.cd = c->d_o_cd.as<int32_t>(); fo.nrefs = c->d_o_nrefs.as<int32_t>();
        fo.shares = c->d_o_sh.as<int32_t>(); fo.indeps = c->d_o_in.as<int32_t>(); fo.k = c->d_o_k.as<int32_t>(); fo.d = c->d_o_d.as<int32_t>();
        fo.n = c->d_o_n.as<int32_t>(); fo.p1 = c->d_o_p1.as<double>(); fo.mult = c->d_o_mult.as<double>(); fo.score = c->d_o_score.as<double>();
        fo.err = c->d_o_err.as<uint8_t>();
        launch(c, k_final_gather, fgrid, 256, c->d_fkeys2.as<const unsigned long long>(), c->d_fidx2.as<const uint32_t>(), (const unsigned long long*)CTR_FIELD(n_final),
               c->d_forestt.as<const PassTuple>(), din, c->d_powtab.as<const double>(), c->d_powoff.as<const int32_t>(), fo);
    }
    if (N > 0) {
        // join2.rs:69-84 + canonical labels
        CK(c->d_first.ensure((size_t)std::max<int64_t>(din.n_exacts, 1) * 4));
        launch(c, k_fill_i32, nblk(din.n_exacts, 256), 256, c->d_first.as<int32_t>(), din.n_exacts, 0x7fffffff);
        launch(c, k_exact_first, nblk(N, 256), 256, din, c->d_first.as<int32_t>());
        launch(c, k_exact_union, nblk(N, 256), 256, din, c->d_first.as<const int32_t>(), c->d_parent.as<int32_t>());
        launch(c, k_labels, nblk(N, 256), 256, c->d_parent.as<const int32_t>(), N, c->d_labels.as<int32_t>());
        // what the host needs for the clonotype_id merge of its EquivRel replay: the rows of exacts that own several rows
        CK(c->d_excount.ensure((size_t)std::max<int64_t>(din.n_exacts, 1) * 4));
        CK(c->d_multi.ensure((size_t)N * 8));
        CK(c->d_multi2.ensure((size_t)N * 8));
        CK(cudaMemsetAsync(c->d_excount.p, 0, (size_t)std::max<int64_t>(din.n_exacts, 1) * 4, st));
        launch(c, k_exact_count, nblk(N, 256), 256, din, c->d_excount.as<int32_t>());
        launch(c, k_multi_rows, nblk(N, 256), 256, din, c->d_excount.as<const int32_t>(), c->d_multi.as<unsigned long long>(), c->d_ctr());
    }
    return EJB_OK;
}

}  // namespace

extern "C" int ejb_run(ejb_ctx* c) {
    if (!c) return EJB_ERR_INVALID;
    if (!c->have_input) return fail(c, EJB_ERR_INVALID, "ejb_run without ejb_upload");
    CK(cudaSetDevice(c->device));
    c->have_run = c->prepared = c->have_part = false;
    const double t_wall0 = now_ms();
    memset(&c->stats, 0, sizeof(c->stats));
    c->ev_used = 0;
    c->s2a_events.clear();
    c->p1_events.clear();
    const DevIn& din = c->din;
    const int64_t N = din.n_rows;
    cudaStream_t st = c->stream;
    cudaEvent_t ev_start = c->ev(), ev_packed = c->ev(), ev_rounds = c->ev(), ev_end = c->ev();
    std::vector<std::pair<const char*, cudaEvent_t>> tr;   // opt.trace: finer CUDA events inside the prepare phase (developer aid)
    CK(cudaEventRecord(ev_start, st));
    std::vector<RoundTimes> times;
    int rc = prepare(c, tr);
    if (rc) return rc;
    const int n_subs = (int)c->subs.size();

    // ---- round report buffer (its head is the Counters block), union-find state, work buffers ----
    c->round_units_cap = (size_t)n_subs + 8;
    c->round_tasks_cap = (size_t)c->n_blocks + (size_t)n_subs + 8;
    CK(c->d_round.ensure(sizeof(Counters) + c->round_units_cap * sizeof(UnitOut) + c->round_tasks_cap * (8 * S1_WARPS + 1) + 64));
    CK(cudaMemsetAsync(c->d_round.p, 0, sizeof(Counters), st));
    CK(c->d_parent.ensure((size_t)std::max<int64_t>(N, 1) * 4));
    CK(c->d_best.ensure((size_t)std::max<int64_t>(N, 1) * 8));
    if (N > 0) {
        const int64_t m = std::max({N, c->n_pos, c->n_blocks});
        launch(c, k_init_labels, nblk(m, 256), 256, c->d_parent.as<int32_t>(), N, c->d_m_row.as<const int32_t>(), c->n_pos, c->d_labq.as<int32_t>(),
               c->d_blklab.as<int32_t>(), c->d_blksub.as<const int32_t>(), c->d_subs.as<const SubBucket>(), (int)c->n_blocks);
    }
    {
        unsigned long long need = 0;
        for (const SubBucket& sb : c->subs)
            if (sb.n_act > sb.r_beg) need = std::max<unsigned long long>(need, (unsigned long long)TILE * (unsigned long long)sb.n);
        if (need > c->cand_cap && !c->cap_explicit) c->cand_cap = need;   // worst case of ONE row block: all of its pairs are candidates
    }
    const unsigned long long cap = c->cand_cap;
    c->cand_rate = 0.0;
    c->pairs_eval_seen = 0;
    c->forest_cap = (unsigned long long)std::max<int64_t>(N, 1);
    CK(c->d_cand.ensure(cap * 8)); CK(c->d_slow.ensure(cap * 8)); CK(c->d_slowslot.ensure(cap * 4));
    CK(c->d_surv.ensure(cap * sizeof(Survivor))); CK(c->d_pass.ensure(cap * 8)); CK(c->d_passt.ensure(cap * sizeof(PassTuple)));
    CK(c->d_roots.ensure(cap * 8));
    CK(c->d_forest.ensure(c->forest_cap * 8));
    CK(c->d_forestt.ensure(c->forest_cap * sizeof(PassTuple)));
    CK(cudaEventRecord(ev_packed, st));
    c->prepared = true;

    // ---- round scheduler ----
    {
        const SubVec& subs = c->subs;
        // Units: rows [r0, r1) of a sub-bucket against every column to their right, in row order.  A unit may only be
        // split by ROWS (splitting by columns and contracting after each part is not exact: a lighter edge from a later
        // column part could have to evict a heavier edge already in the forest).
        // Sizing: optimistic growth, bounded by the candidate density (per row) of the last row range processed.  When
        // a round does overflow the candidate buffer (nothing committed), the per-row-range
        // candidate counts of that attempt are exact upper bounds for the retry, which is planned from them.
        struct Known { int r0, r1; double cand; };
        std::vector<int> next_row(subs.size(), 0), live;
        for (size_t s = 0; s < subs.size(); s++) next_row[s] = subs[s].r_beg;
        std::vector<double> rho(subs.size(), -1.0);          // candidates per row of the last committed row range
        std::vector<double> prate(subs.size(), 1.0);         // pass edges per candidate of the last committed unit
        std::vector<std::vector<Known>> known(subs.size());  // counted, not yet committed row ranges (ascending)
        for (int s = 0; s < (int)subs.size(); s++)
            if (subs[s].n >= 2 && subs[s].n_act > subs[s].r_beg && subs[s].r_beg < subs[s].n - 1) live.push_back(s);
        const double cap_round = 0.5 * (double)c->cand_cap;                      // expected candidates allowed per round
        const double target_pass = std::max(4096.0, (double)c->cand_cap / (double)c->opt.strict_div);   // pass edges one unit should not exceed by much
        int stalls = 0;
        const long long first_unit = c->opt.first_unit, growth = c->opt.growth;
        while (!live.empty()) {
            std::vector<Unit> units;
            unsigned long long pairs = 0;
            double expect = 0.0;
            const unsigned long long budget = std::max<unsigned long long>(c->pair_budget, 1ull << 14);
            std::vector<int> still;
            for (int s : live) {
                const SubBucket& sb = subs[s];
                auto unit_pairs = [&](long long r0, long long r1) {
                    const long long rows = r1 - r0;
                    return (unsigned long long)(rows * (sb.n - r1) + rows * (rows - 1) / 2);
                };
                const int r0 = next_row[s];
                int r1;
                double ex = 0.0;
                // where the pairs found mostly JOIN, a unit is truncated early (the rows behind it get connected and are then
                // skipped by the class test); where they mostly fail, every pair has to be scored once anyway
                const double allow = target_pass / std::max(prate[s], 0.02);
                std::vector<Known>& kn = known[s];
                while (!kn.empty() && kn.front().r1 <= r0) kn.erase(kn.begin());
                if (!kn.empty() && kn.front().r0 <= r0) {
                    // counted rows (after an overflow): take as many consecutive ranges as fit, at least a piece of the first
                    const double room = std::min(std::max(cap_round - expect, 0.0), allow);
                    r1 = r0;
                    size_t i = 0;
                    while (i < kn.size() && kn[i].r0 <= r1) {
                        const double per_row = kn[i].cand / (double)std::max(1, kn[i].r1 - kn[i].r0);
                        const double c_i = per_row * (double)(kn[i].r1 - r1);
                        if (ex + c_i <= room) {
                            ex += c_i;
                            r1 = kn[i].r1;
                            i++;
                        } else {
                            const int rows = (int)((room - ex) / std::max(per_row, 1e-9));
                            if (rows > 0) { r1 += rows; ex += per_row * rows; }
                            break;
                        }
                    }
                    if (r1 == r0) {
                        if (!units.empty()) { still.push_back(s); continue; }   // no room left in this round
                        r1 = r0 + 1;                                             // a single row must always be tried
                        ex = kn.front().cand / (double)std::max(1, kn.front().r1 - kn.front().r0);
                    }
                } else {
                    const long long done = r0 - sb.r_beg;   // rows this context has processed as first rows
                    const long long want = (done == 0) ? first_unit : std::max<long long>(first_unit, growth * done);   // optimistic growth: rows done so far x growth
                    r1 = (int)std::min<long long>((long long)r0 + want, sb.n_act);
                    if (sb.n_act - r1 < (r1 - r0) / 4) r1 = sb.n_act;   // no separate round for a short tail
                    ex = (rho[s] > 0.0 ? rho[s] : std::min<double>(64.0, (double)sb.n)) * (double)(r1 - r0);
                    // capacity: never plan more expected candidates than a round may hold
                    if (ex > 0.5 * cap_round) {
                        r1 = r0 + std::max(1, (int)((double)(r1 - r0) * 0.5 * cap_round / ex));
                        ex = 0.5 * cap_round;
                    }
                }
                r1 = std::min(r1, sb.n_act);
                while (r1 - r0 > 1 && unit_pairs(r0, r1) > budget / 2) r1 = r0 + (r1 - r0) / 2;
                const unsigned long long up = unit_pairs(r0, r1);
                if (!units.empty() && (pairs + up > budget || expect + ex > cap_round)) {
                    still.push_back(s);
                    continue;
                }
                Unit u;
                u.sub = s;
                u.r0 = r0;
                u.r1 = r1;
                u.allow = allow;
                u.strict = target_pass;
                u.rho_ref = rho[s];
                u.cand = u.pass = 0.0;
                u.rho_last = -1.0;
                units.push_back(u);
                pairs += up;
                expect += ex;
            }
            const std::vector<Unit> planned = units;
            rc = run_round(c, units, times);
            if (rc != 0 && rc != 1) return rc;
            if (rc == 1) {
                // nothing was committed: the per-range counts (valid upper bounds) drive the retry
                if (units.size() == 1 && units[0].r1 - units[0].r0 <= 1)
                    return fail(c, EJB_ERR_OOM, "work buffers (%llu entries) too small for one row of a %d-row sub-bucket; raise EJB_CAND_CAP",
                                c->cand_cap, subs[units[0].sub].n);
                if (++stalls > 256) return fail(c, EJB_ERR_OOM, "scheduler cannot fit a round into the work buffers; raise EJB_CAND_CAP");
                for (const Unit& u : planned) {
                    still.push_back(u.sub);
                    known[u.sub].clear();
                }
                for (size_t t = 0; t < c->task_meta.size(); t++) {
                    const auto& m = c->task_meta[t];
                    if (c->h_taskskip[t]) continue;
                    known[planned[m.unit].sub].push_back({m.r0, m.r1, 1.25 * (double)c->h_unitcand[t] + 1.0});
                }
                for (const Unit& u : planned) std::sort(known[u.sub].begin(), known[u.sub].end(), [](const Known& a, const Known& b) { return a.r0 < b.r0; });
            } else {
                stalls = 0;
                for (size_t ui = 0; ui < units.size(); ui++) {
                    const Unit& u = units[ui];       // u.r1 may have been truncated by the device (k_truncate)
                    next_row[u.sub] = u.r1;
                    known[u.sub].clear();            // stale after a commit
                    if (u.cand > 0.0) prate[u.sub] = std::min(1.0, u.pass / u.cand);
                    if (u.r1 < std::min(subs[u.sub].n_act, subs[u.sub].n - 1)) still.push_back(u.sub);   // the last row has no pair to its right
                }
                // candidate density of the last committed 16-row band predicts the rows that follow (capacity planning)
                // (not below the unit's own average: one sparse band must not make the next ordinary band look like a jump)
                for (const Unit& u : units)
                    if (u.rho_last >= 0.0) rho[u.sub] = std::max({1e-3, u.rho_last, u.cand / (double)std::max(1, u.r1 - u.r0)});
            }
            std::sort(still.begin(), still.end());
            still.erase(std::unique(still.begin(), still.end()), still.end());
            live.swap(still);
        }
    }
    CK(cudaEventRecord(ev_rounds, st));

    // a shard returns its partial forest (ejb_part_forest); the merging context finalizes the union (ejb_merge_parts)
    if (c->shard_world == 1 && (rc = finalize(c)) != EJB_OK) return rc;
    CK(cudaEventRecord(ev_end, st));
    CK(c->h_round.ensure(sizeof(Counters)));
    CK(cudaMemcpyAsync(c->h_round.p, c->d_round.p, sizeof(Counters), cudaMemcpyDeviceToHost, st));
    rc = sync_stream(c);
    if (rc) return rc;
    CK(cudaGetLastError());
    const Counters& h = *c->h_round.as<Counters>();
    c->n_final = (int64_t)h.n_final;
    c->n_multi = (int64_t)h.n_multi;
    c->n_part = (int64_t)h.n_forest;
    c->stats.forest_edges = (int64_t)h.n_forest;
    c->stats.pairs_evaluated = (int64_t)h.pairs_eval;
    c->stats.two_cell_deleted = (int64_t)h.n_two_cell;
    c->stats.stage1_ref_kills = (int64_t)(h.n_s1_kill - h.n_s1_donor);
    c->stats.stage1_donor_kills = (int64_t)h.n_s1_donor;
    c->stats.stage1_flag_skips = (int64_t)h.n_s1_flag;
    c->stats.stage2_two_ref = (int64_t)h.n_two_ref;
    c->stats.stage2_memo_dead = (int64_t)h.n_memo_dead;
    c->stats.stage2_degr_dead = (int64_t)h.n_degr_dead;
    c->stats.pairs_distance_computed = (int64_t)h.s1_first;
    c->stats.pairs_distance_full = (int64_t)h.s1_full;
    {   // executed popc of stage 1: one per first-word evaluation + the remaining words where they were evaluated (per width)
        double rest = 0.0, pairs_w = 0.0;
        for (const SubBucket& sb : c->subs) {
            const double p = (double)sb.n * (sb.n - 1) / 2;
            rest += p * cdr3_rest_popc(sb.W1);
            pairs_w += p;
        }
        const double rest_avg = pairs_w > 0 ? rest / pairs_w : 0.0;
        c->stats.popc_executed_stage1 = (int64_t)((double)h.s1_first + rest_avg * (double)h.s1_full);
    }
    c->stats.boruvka_iterations = (int64_t)h.bor_iters;
    c->stats.p1_cache_slots = (int64_t)c->p1_cap;
    c->stats.n_devices = 1;
    c->stats.joins = c->n_final;
    float ms = 0;
    cudaEventElapsedTime(&ms, ev_start, ev_packed); c->stats.ms_pack = ms;
    cudaEventElapsedTime(&ms, ev_rounds, ev_end); c->stats.ms_finalize = ms;
    cudaEventElapsedTime(&ms, ev_start, ev_end); c->stats.ms_device_total = ms;
    for (auto& pe : c->s2a_events) {
        cudaEventElapsedTime(&ms, pe.first, pe.second); c->stats.ms_stage2a += ms;
    }
    for (auto& t : times) {
        cudaEventElapsedTime(&ms, t.e0, t.e1); c->stats.ms_stage1 += ms;
        cudaEventElapsedTime(&ms, t.e1, t.e2); c->stats.ms_stage2 += ms;
        cudaEventElapsedTime(&ms, t.e2, t.e3); c->stats.ms_forest += ms;
    }
    if (c->opt.trace) {
        fprintf(stderr, "[ejb trace] prepare:");
        for (size_t i = 1; i < tr.size(); i++) {
            cudaEventElapsedTime(&ms, tr[i - 1].second, tr[i].second);
            fprintf(stderr, " %s %.3f ms;", tr[i].first, ms);
        }
        fprintf(stderr, " total pack %.3f ms\n", c->stats.ms_pack);
        fprintf(stderr, "[ejb trace] integer-test rejections: ident %llu max_diffs %llu max_cdr3 %llu donors %llu cd>=mult*indeps %llu cref %llu\n", h.n_rej[0],
                h.n_rej[1], h.n_rej[2], h.n_rej[3], h.n_rej[4], h.n_rej[5]);
        for (size_t i = 0; i < times.size(); i++) {
            float a, b, d;
            cudaEventElapsedTime(&a, times[i].e0, times[i].e1);
            cudaEventElapsedTime(&b, times[i].e1, times[i].e2);
            cudaEventElapsedTime(&d, times[i].e2, times[i].e3);
            fprintf(stderr, "[ejb trace] round %zu: stage1 %.3f stage2 %.3f forest %.3f ms\n", i, a, b, d);
        }
        fprintf(stderr, "[ejb trace] finalize %.3f ms, %lld syncs, %lld boruvka iterations\n", c->stats.ms_finalize, (long long)c->stats.host_syncs,
                (long long)c->stats.boruvka_iterations);
        float p1ms = 0;
        for (auto& pe : c->p1_events) {
            float x = 0;
            cudaEventElapsedTime(&x, pe.first, pe.second);
            p1ms += x;
        }
        fprintf(stderr, "[ejb trace] p1: %.3f ms; %llu keys in %llu groups, %llu terms evaluated (%llu pool entries), %llu keys by the direct path\n", p1ms,
                h.p1_total, h.p1_groups, h.p1_terms, h.p1_pool_used, h.p1_direct);
    }
    c->stats.ms_h2d = c->ms_h2d;
    c->stats.h2d_bytes = c->h2d_bytes;
    c->stats.ms_total = now_ms() - t_wall0;
    c->have_run = c->shard_world == 1;
    c->have_part = c->shard_world > 1;
    return EJB_OK;
}

namespace {
int multi_join(ejb_ctx* c, const ejb_params* p, const ejb_input* in, ejb_result* out);
}

extern "C" {

int ejb_download(ejb_ctx* c, ejb_result* out) {
    if (!c || !out) return EJB_ERR_INVALID;
    if (!c->have_run) return fail(c, EJB_ERR_INVALID, "ejb_download without ejb_run");
    CK(cudaSetDevice(c->device));
    memset(out, 0, sizeof(*out));
    const double t0 = now_ms();
    const int64_t E = c->n_final, N = c->din.n_rows;
    cudaStream_t st = c->stream;
    // ---- carve the pinned arena: 13 edge arrays + labels + x/y/z + row_exact scratch ----
    const size_t e1 = (size_t)std::max<int64_t>(E, 1), n1 = (size_t)std::max<int64_t>(N, 1);
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t need = 7 * al(e1 * 4) + 2 * al(e1 * 8) + 3 * al(e1 * 8) + al(e1) + 5 * al(n1 * 4) + al(n1 * 8) + 4096;
    if (need > c->h_res_cap) {
        if (c->h_res) cudaFreeHost(c->h_res);
        c->h_res = nullptr;
        c->h_res_cap = 0;
        const size_t want = need + need / 4;
        CK(cudaMallocHost(&c->h_res, want));
        c->h_res_cap = want;
    }
    char* p = c->h_res;
    auto take = [&](size_t bytes) { char* r = p; p += al(bytes); return r; };
    int32_t* k1 = (int32_t*)take(e1 * 4); int32_t* k2 = (int32_t*)take(e1 * 4); int32_t* cd = (int32_t*)take(e1 * 4);
    int32_t* nrefs = (int32_t*)take(e1 * 4); int32_t* kk = (int32_t*)take(e1 * 4); int32_t* dd_ = (int32_t*)take(e1 * 4); int32_t* nn = (int32_t*)take(e1 * 4);
    int32_t* sh = (int32_t*)take(e1 * 8); int32_t* in = (int32_t*)take(e1 * 8);
    double* p1 = (double*)take(e1 * 8); double* mult = (double*)take(e1 * 8); double* score = (double*)take(e1 * 8);
    uint8_t* err = (uint8_t*)take(e1);
    int32_t* labels = (int32_t*)take(n1 * 4); int32_t* ex = (int32_t*)take(n1 * 4); int32_t* ey = (int32_t*)take(n1 * 4); int32_t* ez = (int32_t*)take(n1 * 4);
    // scratch for the replay: bucket starts + the (exact << 32 | row) list of multi-row exacts (together never more than n1 * 4 + n1 * 8 bytes)
    int32_t* bstart = (int32_t*)take(n1 * 4);
    uint64_t* multi = (uint64_t*)take(n1 * 8);
    auto get = [&](void* dst, const void* src, size_t bytes) -> cudaError_t {
        if (bytes == 0) return cudaSuccess;
        c->stats.d2h_bytes += (int64_t)bytes;
        return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st);
    };
    // what the EquivRel replay needs comes first; the PotentialJoin numerics follow while the host replays
    if (E > 0) {
        CK(get(k1, c->d_o_k1.p, (size_t)E * 4));
        CK(get(k2, c->d_o_k2.p, (size_t)E * 4));
    }
    const int64_t NB = c->n_buckets, NM = c->n_multi;
    if (NB > 0) CK(get(bstart, c->d_bstart.p, (size_t)NB * 4));
    if (NM > 0) {
        int rc = cub_sort_keys64(c, c->d_multi.as<uint64_t>(), c->d_multi2.as<uint64_t>(), NM);
        if (rc) return rc;
        CK(get(multi, c->d_multi2.p, (size_t)NM * 8));
    }
    cudaEvent_t ev_edges = c->ev();
    CK(cudaEventRecord(ev_edges, st));
    if (E > 0) {
        const size_t e = (size_t)E;
        CK(get(cd, c->d_o_cd.p, e * 4)); CK(get(nrefs, c->d_o_nrefs.p, e * 4));
        CK(get(sh, c->d_o_sh.p, e * 8)); CK(get(in, c->d_o_in.p, e * 8)); CK(get(kk, c->d_o_k.p, e * 4)); CK(get(dd_, c->d_o_d.p, e * 4));
        CK(get(nn, c->d_o_n.p, e * 4)); CK(get(p1, c->d_o_p1.p, e * 8)); CK(get(mult, c->d_o_mult.p, e * 8)); CK(get(score, c->d_o_score.p, e * 8));
        CK(get(err, c->d_o_err.p, e));
    }
    if (N > 0) CK(get(labels, c->d_labels.p, (size_t)N * 4));
    CK(cudaEventSynchronize(ev_edges));
    // replay into the reference's EquivRel (join2.rs:45-54, 69-84) so that x/y/z match from_raw; the state is built in place
    // in the result arena
    const double t1 = now_ms();
    if (N > 0) {
        ejbhost::EquivRel eq(ex, ey, ez);
        const int threads = (int)std::min<unsigned>(16u, std::max(1u, std::thread::hardware_concurrency()));
        ejbhost::replay_edges_parallel(eq, N, k1, k2, E, bstart, NB, threads);
        ejbhost::finish_join_merge_list(eq, multi, NM);
    }
    c->stats.ms_replay = now_ms() - t1;
    CK(cudaStreamSynchronize(st));
    c->stats.ms_d2h = now_ms() - t0 - c->stats.ms_replay;
    int64_t errors = 0;
    for (int64_t i = 0; i < E; i++) errors += err[i];
    c->stats.errors = errors;
    out->n_edges = E;
    out->edge_k1 = k1; out->edge_k2 = k2; out->edge_cd = cd; out->edge_nrefs = nrefs;
    out->edge_shares = sh; out->edge_indeps = in; out->edge_k = kk; out->edge_d = dd_; out->edge_n = nn;
    out->edge_p1 = p1; out->edge_mult = mult; out->edge_score = score; out->edge_err = err;
    out->n_rows = N;
    out->labels = labels;
    out->eq_x = ex; out->eq_y = ey; out->eq_z = ez;
    out->stats = c->stats;
    out->opaque = nullptr;
    return EJB_OK;
}

void ejb_result_free(ejb_ctx* c, ejb_result* out) {
    (void)c;   // the arrays live in the context's pinned arena, reused by the next download
    if (out) memset(out, 0, sizeof(*out));
}

int ejb_join(ejb_ctx* c, const ejb_params* p, const ejb_input* in, ejb_result* out) {
    if (!c) return EJB_ERR_INVALID;
    if (!c->peers.empty()) return multi_join(c, p, in, out);
    const double t0 = now_ms();
    int rc = ejb_upload(c, p, in);
    if (rc) return rc;
    rc = ejb_run(c);
    if (rc) return rc;
    rc = ejb_download(c, out);
    if (rc) return rc;
    out->stats.ms_total = now_ms() - t0;
    c->stats.ms_total = out->stats.ms_total;
    return EJB_OK;
}

int ejb_device_edges(ejb_ctx* c, int64_t* n_edges, const int32_t** d_k1, const int32_t** d_k2, const int32_t** d_labels) {
    if (!c) return EJB_ERR_INVALID;
    if (!c->have_run) return fail(c, EJB_ERR_INVALID, "ejb_device_edges without ejb_run");
    if (n_edges) *n_edges = c->n_final;
    if (d_k1) *d_k1 = c->n_final > 0 ? c->d_o_k1.as<const int32_t>() : nullptr;
    if (d_k2) *d_k2 = c->n_final > 0 ? c->d_o_k2.as<const int32_t>() : nullptr;
    if (d_labels) *d_labels = c->d_labels.as<const int32_t>();
    return EJB_OK;
}

int ejb_device_tuples(ejb_ctx* c, const int32_t** d_cd, const int32_t** d_nrefs, const int32_t** d_shares) {
    if (!c) return EJB_ERR_INVALID;
    if (!c->have_run) return fail(c, EJB_ERR_INVALID, "ejb_device_tuples without ejb_run");
    const bool any = c->n_final > 0;
    if (d_cd) *d_cd = any ? c->d_o_cd.as<const int32_t>() : nullptr;
    if (d_nrefs) *d_nrefs = any ? c->d_o_nrefs.as<const int32_t>() : nullptr;
    if (d_shares) *d_shares = any ? c->d_o_sh.as<const int32_t>() : nullptr;
    return EJB_OK;
}

int ejb_set_row_roles(ejb_ctx* c, const uint8_t* roles, int64_t n_rows) {
    if (!c || n_rows < 0) return EJB_ERR_INVALID;
    c->roles.clear();
    if (!roles || n_rows == 0) return EJB_OK;
    for (int64_t i = 0; i < n_rows; i++)
        if (roles[i] > 3) return fail(c, EJB_ERR_INVALID, "row role %lld out of range", (long long)i);
    c->roles.assign(roles, roles + n_rows);
    return EJB_OK;
}

// Kruskal over an edge list sorted by (k1, k2): keep[i] = 1 iff edge i joins two components of the edges before it.
// Merges the partial forests of a sub-bucket that was split by row ranges: MSF(G) = MSF(MSF(G1) u MSF(G2) u ...).
int ejb_forest_merge(int32_t n_rows, const int32_t* k1, const int32_t* k2, int64_t n_edges, uint8_t* keep) {
    if (n_rows < 0 || n_edges < 0 || (n_edges > 0 && (!k1 || !k2 || !keep))) return EJB_ERR_INVALID;
    std::vector<int32_t> parent((size_t)n_rows);
    for (int32_t i = 0; i < n_rows; i++) parent[(size_t)i] = i;
    auto find = [&](int32_t x) {
        while (parent[(size_t)x] != x) {
            parent[(size_t)x] = parent[(size_t)parent[(size_t)x]];
            x = parent[(size_t)x];
        }
        return x;
    };
    for (int64_t i = 0; i < n_edges; i++) {
        if (k1[i] < 0 || k1[i] >= n_rows || k2[i] < 0 || k2[i] >= n_rows) return EJB_ERR_INVALID;
        if (i > 0 && (k1[i] < k1[i - 1] || (k1[i] == k1[i - 1] && k2[i] < k2[i - 1]))) return EJB_ERR_INVALID;   // not sorted
        const int32_t a = find(k1[i]), b = find(k2[i]);
        keep[i] = a != b;
        if (a != b) parent[(size_t)std::max(a, b)] = std::min(a, b);
    }
    return EJB_OK;
}

int ejb_last_stats(const ejb_ctx* c, ejb_stats* out) {
    if (!c || !out) return EJB_ERR_INVALID;
    *out = c->stats;
    return EJB_OK;
}

int ejb_p1_batch(ejb_ctx* c, const int32_t* k, const int32_t* d, const int32_t* n, int64_t count, double* p1_out) {
    if (!c || count < 0) return EJB_ERR_INVALID;
    if (count == 0) return EJB_OK;
    CK(cudaSetDevice(c->device));
    for (int64_t i = 0; i < count; i++)
        if (k[i] < 0 || k[i] > SR_NMAX || d[i] < 0 || d[i] > k[i] || n[i] < 1) return fail(c, EJB_ERR_RANGE, "p1 triple %lld out of range", (long long)i);
    DevBuf dk, dd_, dn, dout;
    CK(dk.ensure(count * 4)); CK(dd_.ensure(count * 4)); CK(dn.ensure(count * 4)); CK(dout.ensure(count * 8));
    CK(cudaMemcpyAsync(dk.p, k, count * 4, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(dd_.p, d, count * 4, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(dn.p, n, count * 4, cudaMemcpyHostToDevice, c->stream));
    k_p1_batch<<<nblk(count * 32, 128), 128, 0, c->stream>>>(c->d_sr.as<dd_t>(), c->qt(), dk.as<int32_t>(), dd_.as<int32_t>(), dn.as<int32_t>(), count, dout.as<double>());
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(p1_out, dout.p, count * 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return EJB_OK;
}

// entries [first, first + count) of the device-built Stirling ratio table in its triangular layout (row n at n(n+1)/2), as
// (hi, lo) pairs — for the bit-level comparison with the oracle's host build (tests)
int ejb_sr_table_read(ejb_ctx* c, int64_t first, int64_t count, double* hi_lo_out) {
    if (!c || first < 0 || count < 0 || !hi_lo_out) return EJB_ERR_INVALID;
    const int64_t total = (int64_t)(SR_NMAX + 1) * (SR_NMAX + 2) / 2;
    if (first + count > total) return EJB_ERR_RANGE;
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpy(hi_lo_out, c->d_sr.as<dd_t>() + first, (size_t)count * sizeof(dd_t), cudaMemcpyDeviceToHost));
    return EJB_OK;
}

int ejb_measure_peaks(ejb_ctx* c, double* popc_per_s, double* lop3_per_s, double* sm_clock_mhz) {
    if (!c) return EJB_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    const int blocks = c->sm_count * 8, threads = 256, iters = 1 << 14;
    DevBuf out;
    CK(out.ensure((size_t)blocks * threads * 4));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    double best_popc = 0, best_lop = 0;
    for (int rep = 0; rep < 4; rep++) {
        CK(cudaEventRecord(a, c->stream));
        k_peak_popc<<<blocks, threads, 0, c->stream>>>(out.as<uint32_t>(), iters);
        CK(cudaEventRecord(b, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        if (rep > 0) best_popc = std::max(best_popc, (double)blocks * threads * iters * 8.0 / (ms * 1e-3));
        CK(cudaEventRecord(a, c->stream));
        k_peak_lop3<<<blocks, threads, 0, c->stream>>>(out.as<uint32_t>(), iters);
        CK(cudaEventRecord(b, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        cudaEventElapsedTime(&ms, a, b);
        if (rep > 0) best_lop = std::max(best_lop, (double)blocks * threads * iters * 8.0 / (ms * 1e-3));
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    if (popc_per_s) *popc_per_s = best_popc;
    if (lop3_per_s) *lop3_per_s = best_lop;
    if (sm_clock_mhz) {
        int khz = 0;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, c->device);
        *sm_clock_mhz = khz / 1000.0;
    }
    return EJB_OK;
}

// host-side EquivRel replay for callers that merge per-rank edge lists (multi-GPU)
int ejb_equiv_replay(int32_t n_rows, const int32_t* k1, const int32_t* k2, int64_t n_edges, const int32_t* row_exact, int32_t* x, int32_t* y, int32_t* z,
                     int32_t* labels) {
    if (n_rows < 0 || n_edges < 0) return EJB_ERR_INVALID;
    for (int64_t i = 0; i < n_edges; i++)
        if (k1[i] < 0 || k1[i] >= n_rows || k2[i] < 0 || k2[i] >= n_rows) return EJB_ERR_INVALID;
    ejbhost::EquivRel eq(n_rows);
    ejbhost::replay_edges(eq, k1, k2, n_edges);
    if (row_exact) ejbhost::finish_join_merge(eq, row_exact, n_rows);
    if (x) memcpy(x, eq.x, (size_t)n_rows * 4);
    if (y) memcpy(y, eq.y, (size_t)n_rows * 4);
    if (z) memcpy(z, eq.z, (size_t)n_rows * 4);
    if (labels) {
        std::vector<int32_t> mn((size_t)std::max(n_rows, 1), 0x7fffffff);
        for (int32_t i = 0; i < n_rows; i++) mn[eq.y[i]] = std::min(mn[eq.y[i]], i);
        for (int32_t i = 0; i < n_rows; i++) labels[i] = mn[eq.y[i]];
    }
    return EJB_OK;
}

}  // extern "C"

// ================================================================================================
// f2: join_one for arbitrary pairs (define_mat's second use of the predicate, enclone_process/src/define_mat.rs:172,247)
// ================================================================================================
extern "C" int ejb_join_one_batch(ejb_ctx* c, const int32_t* k1, const int32_t* k2, int64_t n_pairs, const ejb_pair_out* out) {
    if (!c || n_pairs < 0 || !out || (n_pairs > 0 && (!k1 || !k2))) return EJB_ERR_INVALID;
    if (!c->prepared || c->shard_world != 1) return fail(c, EJB_ERR_INVALID, "ejb_join_one_batch needs the input of a completed single-context ejb_run / ejb_join");
    if (n_pairs == 0) return EJB_OK;
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const int64_t N = c->din.n_rows;
    CK(c->d_qofrow.ensure((size_t)std::max<int64_t>(N, 1) * 4));
    CK(cudaMemsetAsync(c->d_qofrow.p, 0xff, (size_t)std::max<int64_t>(N, 1) * 4, st));
    if (c->n_pos > 0) launch(c, k_q_of_row, nblk(c->n_pos, 256), 256, c->d_m_row.as<const int32_t>(), c->n_pos, c->d_qofrow.as<int32_t>());
    const int64_t chunk = (int64_t)std::min<unsigned long long>(c->cand_cap, 1ull << 24);
    CK(c->d_bk1.ensure((size_t)chunk * 4)); CK(c->d_bk2.ensure((size_t)chunk * 4));
    CK(c->d_bpairs.ensure((size_t)chunk * 8)); CK(c->d_bslots.ensure((size_t)chunk * 4));
    CK(c->d_btuples.ensure((size_t)chunk * sizeof(EdgeTuple)));
    std::vector<EdgeTuple> host((size_t)std::min<int64_t>(chunk, n_pairs));
    unsigned int* const d_bad = c->d_err.as<unsigned int>() + 40;   // scratch word of the header block
    for (int64_t o = 0; o < n_pairs; o += chunk) {
        const int64_t m = std::min<int64_t>(chunk, n_pairs - o);
        char* const rbase = c->d_round.as<char>();
        CK(cudaMemsetAsync(rbase, 0, offsetof(Counters, round_end), st));   // n_cand counts the pairs that reach stage 2
        CK(cudaMemsetAsync(d_bad, 0, 4, st));
        CK(cudaMemsetAsync(c->d_btuples.p, 0, (size_t)m * sizeof(EdgeTuple), st));   // pass = 0 unless stage 2 says otherwise
        CK(cudaMemcpyAsync(c->d_bk1.p, k1 + o, (size_t)m * 4, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(c->d_bk2.p, k2 + o, (size_t)m * 4, cudaMemcpyHostToDevice, st));
        launch(c, k_batch_pairs, nblk(m, 256), 256, c->d_bk1.as<const int32_t>(), c->d_bk2.as<const int32_t>(), m, N, c->d_qofrow.as<const int32_t>(),
               c->d_m_sub.as<const int32_t>(), c->d_bpairs.as<uint2>(), c->d_bslots.as<uint32_t>(), CTR_FIELD(n_cand), d_bad);
        int rc = run_stage2(c, c->d_bpairs.as<uint2>(), c->d_bslots.as<uint32_t>(), CTR_FIELD(n_cand), (unsigned long long)chunk, 1, c->d_btuples.as<EdgeTuple>());
        if (rc) return rc;
        CK(cudaMemcpyAsync(host.data(), c->d_btuples.p, (size_t)m * sizeof(EdgeTuple), cudaMemcpyDeviceToHost, st));
        CK(c->h_round.ensure(sizeof(Counters) + 64));
        CK(cudaMemcpyAsync(c->h_round.p, c->d_round.p, sizeof(Counters), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(c->h_round.as<char>() + sizeof(Counters), d_bad, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        const Counters& h = *c->h_round.as<Counters>();
        c->p1_count += h.n_newkeys;
        if (*reinterpret_cast<const unsigned int*>(c->h_round.as<char>() + sizeof(Counters))) return fail(c, EJB_ERR_INVALID, "pair batch: row index out of range");
        if (h.overflow || h.p1_full) return fail(c, EJB_ERR_OOM, "pair batch overflowed the work buffers");
        if (h.n_errflag_panic) return fail(c, EJB_ERR_INVALID, "pair batch: input on which the reference panics");
c->p1_groups = h.p1_groups;
        c->p1_pool_used = h.p1_pool_used;
        if ((rc = p1_groups_grow(c, h.p1_desc)) != EJB_OK) return rc;
                if (c->p1_count > c->p1_cap / 4 && (rc = p1_cache_grow(c)) != EJB_OK) return rc;
        for (int64_t i = 0; i < m; i++) {
            const EdgeTuple& t = host[(size_t)i];
            const int64_t g = o + i;
            if (out->pass) out->pass[g] = (uint8_t)(t.pass != 0);
            if (out->cd) out->cd[g] = t.cd;
            if (out->nrefs) out->nrefs[g] = t.nrefs;
            if (out->shares) { out->shares[2 * g] = t.sh0; out->shares[2 * g + 1] = t.sh1; }
            if (out->indeps) { out->indeps[2 * g] = t.in0; out->indeps[2 * g + 1] = t.in1; }
            if (out->k) out->k[g] = t.k;
            if (out->d) out->d[g] = t.d;
            if (out->n) out->n[g] = t.n;
            if (out->p1) out->p1[g] = t.p1;
            if (out->mult) out->mult[g] = t.mult;
            if (out->score) out->score[g] = t.score;
            if (out->err) out->err[g] = (uint8_t)t.err;
        }
    }
    return EJB_OK;
}

// ================================================================================================
// Multi-GPU: partial forests of sharded runs, merged on one context
// ================================================================================================
extern "C" {

// The shard assignment as a pure host function (tests; no device needed): sub_rows[s] = rows of sub-bucket s.  r_beg[s] / n_act[s]
// receive the first-row range [r_beg, n_act) of sub-bucket s that `rank` of `world` joins (empty: another rank owns it).
int ejb_shard_plan(const int32_t* sub_rows, int32_t n_subs, int world, int rank, int32_t* r_beg, int32_t* n_act) {
    if (!sub_rows || n_subs < 0 || world < 1 || world > 64 || rank < 0 || rank >= world || !r_beg || !n_act) return EJB_ERR_INVALID;
    std::vector<SubBucket> subs((size_t)n_subs);
    for (int s = 0; s < n_subs; s++) {
        memset(&subs[(size_t)s], 0, sizeof(SubBucket));
        subs[(size_t)s].n = sub_rows[s];
    }
    shard_assign(subs.data(), n_subs, rank, world);
    for (int s = 0; s < n_subs; s++) {
        r_beg[s] = subs[(size_t)s].r_beg;
        n_act[s] = subs[(size_t)s].n_act;
    }
    return EJB_OK;
}

// Partial forest of the last sharded ejb_run (ejb_set_shard with world > 1): device pointers, valid until the next run.
int ejb_part_forest(ejb_ctx* c, int64_t* n_edges, const unsigned long long** d_edges, const void** d_tuples) {
    if (!c) return EJB_ERR_INVALID;
    if (!c->have_part) return fail(c, EJB_ERR_INVALID, "ejb_part_forest without a sharded ejb_run");
    if (n_edges) *n_edges = c->n_part;
    if (d_edges) *d_edges = c->d_forest.as<const unsigned long long>();
    if (d_tuples) *d_tuples = c->d_forestt.p;
    return EJB_OK;
}

// Device buffers (on the merging context) that receive the concatenated partial forests of ALL parts, `total` entries:
// 8-byte edge weights and 32-byte tuples.  The caller fills them (NCCL gather, peer copies), then calls ejb_merge_parts.
int ejb_merge_buffers(ejb_ctx* c, int64_t total, unsigned long long** d_edges, void** d_tuples) {
    if (!c || total < 0) return EJB_ERR_INVALID;
    if (!c->have_part) return fail(c, EJB_ERR_INVALID, "ejb_merge_buffers without a sharded ejb_run on the merging context");
    CK(cudaSetDevice(c->device));
    const size_t n = (size_t)std::max<int64_t>(total, 1);
    CK(c->d_mergew.ensure(n * 8));
    CK(c->d_merget.ensure(n * sizeof(PassTuple)));
    if (d_edges) *d_edges = c->d_mergew.as<unsigned long long>();
    if (d_tuples) *d_tuples = c->d_merget.p;
    return EJB_OK;
}

// Merge: the gathered edges are the union of the parts' spanning forests.  Sub-buckets that were not split contribute disjoint
// forests (every edge stays); the parts of a split sub-bucket overlap, and the lexicographic spanning forest of the whole
// sub-bucket is the minimum spanning forest of the union of the partial forests.  The edges carry a flag (PT_SPLIT): those of
// whole sub-buckets are united directly (k_merge_split), one Boruvka pass runs over the edges of split sub-buckets only; then the
// ordinary finalize (two-cell filter on the merged forest, ordering, PotentialJoin members, clonotype_id unions, labels).
// ejb_download returns the result of the WHOLE join.
int ejb_merge_parts(ejb_ctx* c, int64_t total) {
    if (!c || total < 0) return EJB_ERR_INVALID;
    if (!c->have_part) return fail(c, EJB_ERR_INVALID, "ejb_merge_parts without a sharded ejb_run on the merging context");
    CK(cudaSetDevice(c->device));
    const double t0 = now_ms();
    const int64_t N = c->din.n_rows;
    cudaStream_t st = c->stream;
    if ((unsigned long long)total > c->forest_cap * 64ull) return fail(c, EJB_ERR_INVALID, "more gathered edges than the parts of %lld rows can hold", (long long)N);
    cudaEvent_t e0 = c->ev(), e1 = c->ev();
    CK(cudaEventRecord(e0, st));
    // fresh round counters, forest and final counts; the union-find starts from singletons
    char* const rbase = c->d_round.as<char>();
    CK(cudaMemsetAsync(rbase, 0, offsetof(Counters, round_end), st));
    CK(cudaMemsetAsync(CTR_FIELD(n_forest), 0, 8, st));
    CK(cudaMemsetAsync(CTR_FIELD(n_final), 0, 8, st));
    CK(cudaMemsetAsync(CTR_FIELD(n_two_cell), 0, 8, st));
    CK(cudaMemsetAsync(CTR_FIELD(n_multi), 0, 8, st));
    if (N > 0) launch(c, k_init_parent, nblk(N, 256), 256, c->d_parent.as<int32_t>(), N);
    CK(c->d_roots.ensure((size_t)std::max<int64_t>(total, 1) * 8));
    CK(c->d_pass.ensure((size_t)std::max<int64_t>(total, 1) * 8));
    CK(c->d_passt.ensure((size_t)std::max<int64_t>(total, 1) * sizeof(PassTuple)));
    if (total > 0) {
        // edges of whole sub-buckets: final, united directly; edges of split sub-buckets: compacted for the Boruvka pass
        launch(c, k_merge_split, c->sm_count * 8, 256, c->d_mergew.as<const unsigned long long>(), c->d_merget.as<const PassTuple>(), (unsigned long long)total,
               c->d_parent.as<int32_t>(), c->d_forest.as<unsigned long long>(), c->d_forestt.as<PassTuple>(), c->d_pass.as<unsigned long long>(),
               c->d_passt.as<PassTuple>(), c->d_ctr());
        ForestArgs A;
        A.pass_w = c->d_pass.as<unsigned long long>();
        A.pass_t = c->d_passt.as<PassTuple>();
        A.n_pass_p = CTR_FIELD(n_pass);
        A.cap = (unsigned long long)total;
        A.parent = c->d_parent.as<int32_t>();
        A.best = c->d_best.as<unsigned long long>();
        A.roots = c->d_roots.as<int2>();
        A.forest = c->d_forest.as<unsigned long long>();
        A.forest_t = c->d_forestt.as<PassTuple>();
        A.forest_cap = c->forest_cap;
        A.ctr = c->d_ctr();
        A.bar = c->d_bar.as<unsigned int>();
        A.rowq = c->d_m_row.as<int32_t>();
        A.n_pos = c->n_pos;
        A.labq = c->d_labq.as<int32_t>();
        A.subs = c->d_subs.as<SubBucket>();
        A.blk_lab = c->d_blklab.as<int32_t>();
        A.blk_sub = c->d_blksub.as<int32_t>();
        A.n_blocks = (int)c->n_blocks;
        CK(cudaMemsetAsync(c->d_bar.p, 0, 4, st));
        void* args[] = {&A};
        CK(cudaLaunchCooperativeKernel((const void*)k_forest_round, dim3((unsigned int)c->coop_grid), dim3(256), args, 0, st));
        c->stats.kernel_launches++;
        // p1 of every merged edge travels in its PassTuple: nothing is re-evaluated here
    }
    int rc = finalize(c);
    if (rc) return rc;
    CK(cudaEventRecord(e1, st));
    CK(c->h_round.ensure(sizeof(Counters)));
    CK(cudaMemcpyAsync(c->h_round.p, c->d_round.p, sizeof(Counters), cudaMemcpyDeviceToHost, st));
    rc = sync_stream(c);
    if (rc) return rc;
    CK(cudaGetLastError());
    const Counters& h = *c->h_round.as<Counters>();
    if (h.p1_full) return fail(c, EJB_ERR_OOM, "p1 cache full during the merge");
    if (h.overflow) return fail(c, EJB_ERR_CUDA, "forest buffer overflow during the merge");
    if (h.alive_it[BOR_MAX_ITERS]) return fail(c, EJB_ERR_CUDA, "merge forest did not converge");
    c->p1_count += h.n_newkeys;
    c->n_final = (int64_t)h.n_final;
    c->n_multi = (int64_t)h.n_multi;
    c->stats.forest_edges = (int64_t)h.n_forest;
    c->stats.two_cell_deleted = (int64_t)h.n_two_cell;
    c->stats.joins = c->n_final;
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    c->stats.ms_merge = ms;
    c->stats.ms_device_total += ms;
    (void)t0;
    c->have_run = true;
    return EJB_OK;
}

}  // extern "C"

namespace {
// ejb_create with n_devices > 1: one worker thread per GPU uploads the input and joins its shard; the partial forests are
// copied GPU-to-GPU into the first context, merged and finalized there; the caller sees ONE result (ejb_join is unchanged).
int multi_join(ejb_ctx* c, const ejb_params* p, const ejb_input* in, ejb_result* out) {
    const double t0 = now_ms();
    std::vector<ejb_ctx*> ctxs;
    ctxs.push_back(c);
    for (ejb_ctx* x : c->peers) ctxs.push_back(x);
    const int n = (int)ctxs.size();
    std::vector<int> rcs((size_t)n, 0);
    std::vector<std::thread> th;
    const bool sliced = c->arena_base != nullptr;   // with a declared arena the input crosses PCIe once: 1/n per GPU, the rest over NVLink
    for (int i = 0; i < n; i++)
        th.emplace_back([&, i] {
            ejb_ctx* x = ctxs[i];
            x->shard_rank = i;
            x->shard_world = n;
            x->arena_base = c->arena_base;
            x->arena_bytes = c->arena_bytes;
            rcs[i] = sliced ? ejb_upload_slice(x, p, in, i, n) : ejb_upload(x, p, in);
        });
    for (auto& t : th) t.join();
    th.clear();
    for (int i = 0; i < n; i++)
        if (rcs[i]) {
            if (i > 0) c->err = ctxs[i]->err;
            return rcs[i];
        }
    if (sliced) {   // all-gather of the arena slices by peer copies: context j pulls slice i from context i
        const size_t slice = c->arena_slice;
        for (int j = 0; j < n; j++) {
            CK(cudaSetDevice(ctxs[j]->device));
            for (int i = 0; i < n; i++) {
                if (i == j) continue;
                const size_t lo = std::min(slice * (size_t)i, c->arena_bytes), hi = std::min(lo + slice, c->arena_bytes);
                if (hi > lo) CK(cudaMemcpyPeerAsync(ctxs[j]->d_arena.as<char>() + lo, ctxs[j]->device, ctxs[i]->d_arena.as<char>() + lo, ctxs[i]->device, hi - lo, ctxs[j]->stream));
            }
        }
        CK(cudaSetDevice(c->device));
    }
    for (int i = 0; i < n; i++)
        th.emplace_back([&, i] { rcs[i] = ejb_run(ctxs[i]); });
    for (auto& t : th) t.join();
    for (int i = 0; i < n; i++)
        if (rcs[i]) {
            if (i > 0) c->err = ctxs[i]->err;
            return rcs[i];
        }
    const double t1 = now_ms();
    int64_t total = 0;
    for (ejb_ctx* x : ctxs) total += x->n_part;
    unsigned long long* dw = nullptr;
    void* dt = nullptr;
    int rc = ejb_merge_buffers(c, total, &dw, &dt);
    if (rc) return rc;
    CK(cudaSetDevice(c->device));
    int64_t off = 0;
    for (ejb_ctx* x : ctxs) {
        if (x->n_part > 0) {
            CK(cudaMemcpyPeerAsync(dw + off, c->device, x->d_forest.p, x->device, (size_t)x->n_part * 8, c->stream));
            CK(cudaMemcpyPeerAsync(reinterpret_cast<char*>(dt) + (size_t)off * sizeof(PassTuple), c->device, x->d_forestt.p, x->device,
                                   (size_t)x->n_part * sizeof(PassTuple), c->stream));
        }
        off += x->n_part;
    }
    rc = ejb_merge_parts(c, total);
    if (rc) return rc;
    // whole-join statistics: the work counters are sums over the parts, the times those of the slowest part
    ejb_stats agg = c->stats;
    for (int i = 1; i < n; i++) {
        const ejb_stats& s = ctxs[i]->stats;
        agg.pairs_scored += s.pairs_scored; agg.pairs_evaluated += s.pairs_evaluated; agg.stage1_candidates += s.stage1_candidates;
        agg.stage2_slow += s.stage2_slow; agg.stage2_survivors += s.stage2_survivors; agg.pass_edges += s.pass_edges;
        agg.p1_unique += s.p1_unique; agg.kernel_launches += s.kernel_launches; agg.h2d_bytes += s.h2d_bytes;
        agg.stage1_ref_kills += s.stage1_ref_kills; agg.stage1_flag_skips += s.stage1_flag_skips; agg.stage1_donor_kills += s.stage1_donor_kills; agg.stage2_two_ref += s.stage2_two_ref;
        agg.pairs_distance_computed += s.pairs_distance_computed; agg.pairs_distance_full += s.pairs_distance_full;
        agg.popc_executed_stage1 += s.popc_executed_stage1; agg.overflow_retries += s.overflow_retries;
        agg.rounds = std::max(agg.rounds, s.rounds); agg.host_syncs = std::max(agg.host_syncs, s.host_syncs);
        agg.ms_device_total = std::max(agg.ms_device_total, s.ms_device_total + c->stats.ms_merge);
        agg.ms_pack = std::max(agg.ms_pack, s.ms_pack); agg.ms_stage1 = std::max(agg.ms_stage1, s.ms_stage1);
        agg.ms_stage2 = std::max(agg.ms_stage2, s.ms_stage2); agg.ms_forest = std::max(agg.ms_forest, s.ms_forest);
        agg.ms_h2d = std::max(agg.ms_h2d, s.ms_h2d);
    }
    agg.n_devices = n;
    c->stats = agg;
    c->stats.ms_partition = 0.0;
    rc = ejb_download(c, out);
    if (rc) return rc;
    out->stats.ms_total = now_ms() - t0;
    out->stats.ms_merge = now_ms() - t1 - out->stats.ms_d2h - out->stats.ms_replay;
    c->stats = out->stats;
    return EJB_OK;
}
}  // namespace
