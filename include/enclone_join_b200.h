/*
 * enclone_join_b200.h — C ABI of the B200-native clonotype join.
 *
 * This is the drop-in boundary for ONE hot path of 10XGenomics/enclone_ranger:
 *
 *     enclone::join::join_exacts(to_bc, refdata, ctl, exact_clonotypes, info,
 *                                join_info, raw_joins, sr, dref) -> EquivRel
 *                                            (reference: enclone/src/join.rs:31-41)
 *
 * i.e. join_exacts -> join_core (enclone/src/join_core.rs:11-51) -> join_one
 * (enclone_core/src/join_one.rs:66-887) -> finish_join (enclone/src/join2.rs:36-87)
 * -> EquivRel (equiv/src/lib.rs:34-138).
 *
 * The reference has no FFI of its own (it is 100 % safe Rust).  A thin Rust `-sys`
 * crate (rust/enclone_join_b200/) flattens the borrowed Rust
 * inputs into the plain arrays below, calls ejb_join(), appends the returned edges
 * to `raw_joins` and rebuilds the `EquivRel` with `EquivRel::from_raw(x, y, z)`
 * (equiv/src/lib.rs:57).
 *
 * Conventions: every function returns 0 (EJB_OK) or a negative ejb_status; nothing
 * unwinds across the ABI; ejb_last_error() gives a human-readable message; one
 * ejb_ctx is used from one thread at a time; all input pointers are caller-owned
 * HOST memory (pinned or pageable), read-only for the duration of the call; result
 * arrays live in a pinned arena owned by the context and stay valid until ejb_result_free(),
 * the next ejb_join / ejb_download on the same context, or ejb_destroy().  Absent Option<usize> values are
 * encoded as EJB_NONE (-1).
 */
#ifndef ENCLONE_JOIN_B200_H
#define ENCLONE_JOIN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EJB_ABI_VERSION 4
#define EJB_NONE (-1)

typedef enum ejb_status {
    EJB_OK = 0,
    EJB_ERR_INVALID = -1,     /* malformed input (bad offsets, |V| < ref_v_trim, ...)        */
    EJB_ERR_UNSUPPORTED = -2, /* a join mode outside the ranger default path was requested   */
    EJB_ERR_CUDA = -3,        /* CUDA runtime / launch failure                               */
    EJB_ERR_OOM = -4,         /* host or device allocation failed                            */
    EJB_ERR_RANGE = -5,       /* k = indeps + 2*shares exceeds the Stirling table (k > 3000);
                                 the reference panics on sr[x][u] (join_one.rs:30)           */
    EJB_ERR_NO_DEVICE = -6    /* no CUDA device / extension cannot run (no CPU fallback)     */
} ejb_status;

/* Bits of ejb_params.unsupported_modes: reference modes that are NOT reachable from
 * enclone_ranger (main_enclone.rs:35-53 admits none of them) and are rejected.        */
#define EJB_MODE_BCJOIN 1u         /* ctl.join_alg_opt.bcjoin          join.rs:45-64      */
#define EJB_MODE_BASIC 2u          /* join_alg_opt.basic.is_some()     join_one.rs:113    */
#define EJB_MODE_BASIC_H 4u        /* join_alg_opt.basic_h.is_some()   join_one.rs:113    */
#define EJB_MODE_BASICX 8u         /* join_alg_opt.basicx              join_one.rs:155    */
#define EJB_MODE_JOIN_FULL_DIFF 16u /* join_alg_opt.join_full_diff     join_one.rs:186    */
#define EJB_MODE_OLD_MULT 32u      /* join_alg_opt.old_mult            join_one.rs:510    */
#define EJB_MODE_FORCE 64u         /* ctl.force                        join_core.rs:32    */
#define EJB_MODE_SUPER_COMP_FILT 128u /* join_alg_opt.super_comp_filt > 0 join_one.rs:568 */

/* Flat copy of the knobs join_one/join_exacts read from EncloneControl
 * (enclone_core/src/defs.rs:344-370 JoinAlgOpt, :48-54 ClonotypeHeuristics;
 *  defaults: enclone_args/src/proc_args.rs:36-41,152-163).                            */
typedef struct ejb_params {
    int32_t is_bcr;          /* ctl.gen_opt.is_bcr()  (defs.rs:243-253; !tcr)             */
    int32_t mix_donors;      /* ctl.clono_filt_opt_def.donor (MIX_DONORS)                  */
    int32_t easy;            /* join_alg_opt.easy        (join.rs:171)                     */
    int32_t old_light;       /* join_alg_opt.old_light   (join_one.rs:481)                 */
    uint32_t unsupported_modes; /* EJB_MODE_* bitmask; must be 0                           */
    int32_t reserved0;
    double max_score;        /* 100000.0 */
    double cdr3_mult;        /* 5.0      */
    double mult_pow;         /* 80.0     */
    double join_cdr3_ident;  /* 85.0     */
    double fwr1_cdr12_delta; /* 20.0     */
    int64_t max_cdr3_diffs;  /* 1000     */
    int64_t cdr3_normal_len; /* 42       */
    int64_t auto_share;      /* 15       */
    int64_t comp_filt;       /* 8        */
    int64_t comp_filt_bound; /* 80       */
    int64_t max_diffs;       /* 1000000  (heur.max_diffs)        */
    int64_t max_degradation; /* 2        (heur.max_degradation)  */
    int64_t ref_v_trim;      /* 15       */
    int64_t ref_j_trim;      /* 15       */
} ejb_params;

/* Fill `p` with the defaults enclone_ranger runs with; is_bcr: 1 = BCR (incl. TCRGD), 0 = TCR. */
void ejb_params_default(ejb_params* p, int is_bcr);

/*
 * Flattened join input.  Index spaces:
 *   row   k  in [0, n_rows)    one per info[k]               (CloneInfo, defs.rs:736-758)
 *   exact e  in [0, n_exacts)  exact_clonotypes[e]           (ExactClonotype, defs.rs:698-728)
 *   chain c  in [0, n_chains)  exact_clonotypes[e].share[j]  (TigData1, defs.rs:639-686),
 *                              c = ex_chain_off[e] + j
 *   cell  x  in [0, n_cells)   exact_clonotypes[e].clones[j][0] (TigData0, defs.rs:605-626),
 *                              x = ex_cell_off[e] + j
 *   seg   s  in [0, n_segs)    distinct *contents* of info[k].vs[m] / info[k].js[m]
 *                              (need not be de-duplicated: equality is by content, join_one.rs:331)
 *   ref   r  in [0, n_refs)    refdata.refs[r] / refdata.name[r]  (vdj_ann/src/refx.rs:24-38)
 * Per-row-per-chain arrays are [2*k + m], m = 0 (left/heavy) or 1 (right/light); entries
 * for m >= row_nchains[k] are ignored.
 */
typedef struct ejb_input {
    /* ---- rows: info[] in the reference's order (it defines k1 < k2, buckets, edge order) ---- */
    int64_t n_rows;
    const int32_t* row_nchains;   /* [n_rows]   info[k].lens.len()                             */
    const int32_t* row_len;       /* [2*n_rows] info[k].lens[m] == info[k].tigs[m].len()       */
    const int64_t* row_tig_off;   /* [2*n_rows] offset of info[k].tigs[m] in tig_bytes         */
    const uint8_t* row_has_del;   /* [2*n_rows] info[k].has_del[m]                             */
    const int32_t* row_vs;        /* [2*n_rows] seg index of info[k].vs[m]                     */
    const int32_t* row_js;        /* [2*n_rows] seg index of info[k].js[m]                     */
    const int64_t* row_cdr3_off;  /* [2*n_rows] offset of info[k].cdr3s[m] in cdr3_bytes       */
    const int32_t* row_cdr3_len;  /* [2*n_rows] info[k].cdr3s[m].len()                         */
    const int32_t* row_exact;     /* [n_rows]   info[k].clonotype_id (== clonotype_index)      */
    const int32_t* row_exact_col; /* [2*n_rows] info[k].exact_cols[m]                          */
    const uint8_t* tig_bytes;     /* ASCII A,C,G,T and '-' (info.rs:78-110), n_tig_bytes long   */
    int64_t n_tig_bytes;
    const uint8_t* cdr3_bytes;    /* ASCII, n_cdr3_bytes long                                   */
    int64_t n_cdr3_bytes;

    /* ---- exact subclonotypes ---- */
    int64_t n_exacts;
    const int64_t* ex_chain_off;  /* [n_exacts+1]  share.len() = ex_chain_off[e+1]-ex_chain_off[e] */
    const int64_t* ex_cell_off;   /* [n_exacts+1]  ncells()    = ex_cell_off[e+1]-ex_cell_off[e]   */

    /* ---- chains: exact_clonotypes[e].share[j] ---- */
    int64_t n_chains;
    const uint8_t* ch_left;       /* share[j].left                                              */
    const int32_t* ch_c_ref;      /* c_ref_id or EJB_NONE                                        */
    const int32_t* ch_v_ref;      /* v_ref_id (ref index)                                        */
    const int32_t* ch_u_ref;      /* u_ref_id or EJB_NONE                                        */
    const int32_t* ch_fr1_start;  /* fr1_start                                                   */
    const int32_t* ch_cdr1_start; /* cdr1_start or EJB_NONE                                      */
    const int32_t* ch_fr2_start;  /* fr2_start or EJB_NONE                                       */
    const int32_t* ch_cdr2_start; /* cdr2_start or EJB_NONE                                      */
    const int32_t* ch_fr3_start;  /* fr3_start or EJB_NONE                                       */
    const int32_t* ch_hcomp;      /* jun.hcomp (heavy_complexity, start.rs:348-358)              */
    const int64_t* ch_amino_off;  /* offset of seq_del_amino in amino_bytes, or EJB_NONE when it
                                     is byte-identical to the row's tig (the usual case: it only
                                     differs for a V deletion not starting at a codon boundary,
                                     info.rs:84-95)                                              */
    const int32_t* ch_amino_len;  /* seq_del_amino.len() (ignored when ch_amino_off == EJB_NONE)  */
    const uint8_t* amino_bytes;
    int64_t n_amino_bytes;

    /* ---- cells: exact_clonotypes[e].clones[j][0] ---- */
    int64_t n_cells;
    const int32_t* cell_donor;    /* donor_index or EJB_NONE                                     */
    const uint32_t* cell_barcode; /* interned barcode STRING id: equal ids <=> equal strings
                                     (the reference compares strings across datasets,
                                      join_one.rs:451-470)                                       */

    /* ---- segs: the V / J reference segments the rows are compared against ---- */
    int64_t n_segs;
    const int64_t* seg_off;       /* [n_segs+1] into seg_bytes                                    */
    const uint8_t* seg_bytes;     /* ASCII ACGT (DnaString::to_ascii_vec)                         */

    /* ---- refdata: only used by the V-gene-name test, join_one.rs:722-758 ---- */
    int64_t n_refs;
    const int64_t* ref_off;       /* [n_refs+1] into ref_bytes                                    */
    const uint8_t* ref_bytes;     /* ASCII ACGT                                                   */
    const int32_t* ref_name_id;   /* [n_refs] interned id of refdata.name[r] with any "*..." suffix
                                     removed (TextUtils::before("*"), join_one.rs:727-732)        */

    /* ---- optional: 2-bit packed contigs (the role of info[k].tigsp, defs.rs:744) ----
     * A row-chain whose contig is pure upper-case ACGT and has_del == 0 may be supplied packed instead of
     * as ASCII: row_tig2_off[2*k+m] >= 0 is the index of its first word in tig2_words (16 bases per
     * uint32_t, base i in bits [2i%32, 2i%32+2) of word i/16, A,C,G,T = 0,1,2,3); row_tig_off is then
     * ignored for that row-chain.  row_tig2_off == NULL or an entry == EJB_NONE means ASCII.
     * A quarter of the bytes to copy to the device.                                              */
    const int64_t* row_tig2_off;  /* [2*n_rows] or NULL                                           */
    const uint32_t* tig2_words;
    int64_t n_tig2_words;

    /* ---- optional: 2-bit packed CDR3s (same packing as tig2_words; info[k].cdr3s[m] is a substring of the contig,
     * so a caller that holds tigsp can cut it out without touching ASCII) ----
     * row_cdr32_off[2*k+m] >= 0 is the index of the first word of that CDR3 in cdr32_words; row_cdr3_off / cdr3_bytes are
     * then ignored for it (cdr3_bytes may be NULL when every CDR3 is packed).  Pure upper-case ACGT only.          */
    const int64_t* row_cdr32_off; /* [2*n_rows] or NULL                                           */
    const uint32_t* cdr32_words;
    int64_t n_cdr32_words;
} ejb_input;

/* Counters and timings of one join (all optional diagnostics). */
typedef struct ejb_stats {
    int64_t n_buckets;          /* runs of equal info[].lens              (join.rs:68-88)         */
    int64_t n_subbuckets;       /* 2-chain buckets split by CDR3 lengths                          */
    int64_t pairs_lens;         /* sum over 2-chain buckets of n(n-1)/2                           */
    int64_t pairs_scored;       /* sum over sub-buckets of n(n-1)/2  (the BASELINE metric's unit) */
    int64_t pairs_evaluated;    /* pairs of the tiles stage 1 staged (pairs_scored minus whole 128 x 128 tiles skipped
                                   because both row blocks already carry ONE class label); NOT all of them had their
                                   distance computed: see pairs_distance_computed                 */
    int64_t stage1_candidates;  /* pairs passing the CDR3 pre-filter and not already connected    */
    int64_t stage2_slow;        /* candidates scored by the byte-exact slow path                  */
    int64_t stage2_survivors;   /* candidates passing the integer tests (reach the score)         */
    int64_t pass_edges;         /* candidates for which join_one would return true                */
    int64_t forest_edges;       /* edges of the lexicographic spanning forest (= pot, join.rs:105)*/
    int64_t two_cell_deleted;   /* forest edges removed by the two-cell filter (join.rs:120-178)  */
    int64_t joins;              /* edges returned (= sum of r.joins, join.rs:210)                 */
    int64_t errors;             /* returned edges with err == true  (join.rs:211-213)             */
    int64_t p1_unique;          /* distinct (k, d, n) triples whose p1 was evaluated              */
    int64_t rounds;             /* scheduler rounds                                               */
    int64_t kernel_launches;    /* CUDA kernels launched by this join                             */
    int64_t h2d_bytes, d2h_bytes;
    double ms_h2d, ms_pack, ms_stage1, ms_stage2, ms_forest, ms_finalize, ms_d2h, ms_replay;
    double ms_stage2a;          /* k_stage2a alone (inside ms_stage2), CUDA events on the launching stream */
    int64_t stage1_launches, stage2a_launches;
    int64_t overflow_retries;   /* rounds that overflowed the work buffers and were split + re-run   */
    double ms_device_total;     /* pack .. finalize, CUDA events                                  */
    double ms_total;            /* wall clock of the whole call                                   */
    /* two-reference candidates (rows whose V/J reference sets differ, join_one.rs:331-341) and how they ended */
    int64_t stage1_ref_kills;   /* CDR3-close pairs dropped in stage 1: one row is already known to fail the
                                   degradation test (join_one.rs:434-440) against the other's reference set */
    int64_t stage2_two_ref;     /* stage-2a candidates with two reference sets                    */
    int64_t stage2_memo_dead;   /* ... rejected on memoised totals                                */
    int64_t stage2_degr_dead;   /* ... rejected by the degradation test on computed totals        */
    int64_t stage1_donor_kills; /* CDR3-close pairs dropped in stage 1: different donor sets (join_one.rs:301-323) */
    /* executed work of stage 1 (what the popc pipe really ran, lane granular) */
    int64_t pairs_distance_computed; /* lane-pairs whose FIRST 32 CDR3 positions went through XOR + popc (16 x 16 groups that
                                        carry one class label are skipped before that)                          */
    int64_t pairs_distance_full;     /* ... of which the remaining CDR3 words were evaluated too (some lane of the warp was
                                        still within max distance after the first word)                         */
    int64_t popc_executed_stage1;    /* 32-bit popc lane-operations stage 1 executed                              */
    int64_t stage1_flag_skips;  /* pairs stage 1 skipped on the per-row flag nibbles alone (row known dead against the sub-bucket's
                                   dominant reference set x row carrying it; foreign x dominant donor set), BEFORE any distance:
                                   join_one rejects them whatever their CDR3s (join_one.rs:301-323, 434-440)     */
    int64_t host_syncs;         /* stream synchronisations of this ejb_run                                      */
    int64_t boruvka_iterations; /* forest iterations over all rounds (device-side loop, no read-back)           */
    int64_t p1_cache_slots;     /* current size of the device p1 hash                                           */
    int64_t n_devices;          /* GPUs that took part in this join                                             */
    double ms_partition;        /* multi-GPU: host partitioning of the buckets                                  */
    double ms_merge;            /* multi-GPU: gather + merge of the per-device forests                          */
} ejb_stats;

/*
 * Join result.  Edges are `raw_joins` (join.rs:584-588): bucket order, then lexicographic
 * (k1, k2) inside a bucket.  The parallel arrays are the numeric members of the
 * PotentialJoin (defs.rs:870-891) each edge was accepted with.
 */
typedef struct ejb_result {
    int64_t n_edges;
    const int32_t* edge_k1;     /* [n_edges] */
    const int32_t* edge_k2;     /* [n_edges] */
    const int32_t* edge_cd;     /* pot.cd                                                         */
    const int32_t* edge_nrefs;  /* pot.nrefs (1 or 2)                                             */
    const int32_t* edge_shares; /* [2*n_edges] pot.shares[u]  (entry u=1 is 0 when nrefs == 1)    */
    const int32_t* edge_indeps; /* [2*n_edges] pot.indeps[u]                                      */
    const int32_t* edge_k;      /* pot.k = min_indeps + 2*min_shares                              */
    const int32_t* edge_d;      /* pot.d = min_shares                                             */
    const int32_t* edge_n;      /* pot.n = 3 * (len_heavy + len_light)                            */
    const double* edge_p1;      /* pot.p1                                                         */
    const double* edge_mult;    /* pot.mult                                                       */
    const double* edge_score;   /* pot.score = p1 * mult                                          */
    const uint8_t* edge_err;    /* pot.err                                                        */

    int64_t n_rows;
    const int32_t* labels;      /* [n_rows] canonical clonotype id = smallest row index of the
                                   orbit in the EquivRel finish_join returns (device union-find)  */
    /* The EquivRel finish_join returns (join2.rs:36-87), as raw x/y/z (equiv/src/lib.rs:34-38),
       obtained by replaying the edges in order; suitable for EquivRel::from_raw.                 */
    const int32_t* eq_x;
    const int32_t* eq_y;
    const int32_t* eq_z;

    ejb_stats stats;
    void* opaque;               /* library bookkeeping */
} ejb_result;

typedef struct ejb_ctx ejb_ctx;

/* Create a context bound to the given CUDA devices (n_devices >= 1; device_ids == NULL means
 * devices 0..n_devices-1).  Builds the double-double Stirling ratio table (start.rs:50-99)
 * on the device, once per device.  Fails with EJB_ERR_NO_DEVICE when CUDA is unavailable.
 * With n_devices > 1 ejb_join shards the lens buckets over the devices (one worker thread and stream per GPU, no
 * collective on the data path), gathers the per-device forests and merges them; the caller sees one result.      */
int ejb_create(ejb_ctx** out, const int* device_ids, int n_devices);
/* Developer / test knobs (none is needed in production).  Names: "first_unit", "growth" (unit schedule), "s1_mode"
 * (bit 0 uniform-label vote, bit 1 tags), "no_tags", "trace", "thin" (largest unit, in rows, scored by the thread-per-column kernel; 0: always the tile kernel),
 * "quot_table" (0: p1 with three-quotient divisions instead of the constant quotient table).  The EJB_* environment variables of the same meaning are
 * read ONCE, in ejb_create; this call overrides them for the following joins.  Unknown name: EJB_ERR_INVALID.     */
int ejb_set_option(ejb_ctx* ctx, const char* name, long long value);
/* Milliseconds ejb_create spent building + uploading the Stirling ratio table (device build: start.rs:50-99).   */
double ejb_create_ms(const ejb_ctx* ctx);
void ejb_destroy(ejb_ctx* ctx);
const char* ejb_last_error(const ejb_ctx* ctx);

/* The drop-in for join_exacts: host buffers in, edges + EquivRel out.                        */
int ejb_join(ejb_ctx* ctx, const ejb_params* p, const ejb_input* in, ejb_result* out);
void ejb_result_free(ejb_ctx* ctx, ejb_result* out);

/* The same join in three steps, so that the device-resident part can be timed on its own:
 * ejb_upload copies + validates the input into HBM; ejb_run executes bucketer, pair kernels,
 * forest and union-find entirely on the device(s); ejb_download brings edges/labels back and
 * replays the EquivRel.  ejb_join == upload; run; download.                                  */
int ejb_upload(ejb_ctx* ctx, const ejb_params* p, const ejb_input* in);
int ejb_run(ejb_ctx* ctx);
int ejb_download(ejb_ctx* ctx, ejb_result* out);

/* Statistics of the most recent ejb_run on this context.                                     */
int ejb_last_stats(const ejb_ctx* ctx, ejb_stats* out);

/* p1 = p_at_most_m_distinct_in_sample_of_x_from_n_double(k-d, k, n, sr) (join_one.rs:22-43),
 * evaluated by the library's device kernel for `count` triples (for parity tests).           */
int ejb_p1_batch(ejb_ctx* ctx, const int32_t* k, const int32_t* d, const int32_t* n,
                 int64_t count, double* p1_out);

/* Entries [first, first + count) of the device-built Stirling ratio table (start.rs:50-99) in its triangular layout (row n
 * at n(n+1)/2, n <= 3000), as (hi, lo) pairs of the double-double values — for bit-level parity tests of the device build. */
int ejb_sr_table_read(ejb_ctx* ctx, int64_t first, int64_t count, double* hi_lo_out);

/* Measured integer-pipe peaks of device 0 (dependency-free __popc / LOP3 microbenchmarks),
 * in 32-bit lane-operations per second; used as the roofline denominator (SURVEY §8 d).     */
int ejb_measure_peaks(ejb_ctx* ctx, double* popc_per_s, double* lop3_per_s, double* sm_clock_mhz);

/* Multi-GPU: buckets are independent units (join.rs:70-88), so ranks shard them with no collective
 * on the data path.  After ejb_set_shard(ctx, rank, world) the context joins only the sub-buckets it
 * owns (longest-processing-time assignment by pair count, identical on every rank) and
 * ejb_download returns only their edges (labels/eq_* then describe the local edges only); the
 * caller concatenates the per-rank edge lists, sorts them by (k1, k2) and calls ejb_equiv_replay. */
int ejb_set_shard(ejb_ctx* ctx, int rank, int world);

/* join_one for ARBITRARY pairs of rows of the input of the last completed ejb_join / ejb_run on this context: the second use the
 * reference makes of the pair predicate, define_mat's extra raw joins (enclone_process/src/define_mat.rs:172,247).  Same
 * kernels as the join (k_stage2a / k_stage2_slow / k_p1_* / k_stage2b), without the same-class skip: every pair is decided on its
 * own, in the ORDER given (join_one(k1, k2) reads a few facts from its first argument only).  pass[i] = what join_one returns
 * (the caller filters on equal lens first, as define_mat.rs does; pairs with different lens or a row without exactly two
 * chains are answered 0).  The other arrays receive the PotentialJoin members join_one would have pushed (meaningful where
 * pass[i] = 1; any of them may be NULL).  Caller-allocated: [n_pairs], shares / indeps [2 * n_pairs].                    */
typedef struct ejb_pair_out {
    uint8_t* pass;
    int32_t *cd, *nrefs, *shares, *indeps, *k, *d, *n;
    double *p1, *mult, *score;
    uint8_t* err;
} ejb_pair_out;
int ejb_join_one_batch(ejb_ctx* ctx, const int32_t* k1, const int32_t* k2, int64_t n_pairs, const ejb_pair_out* out);

/* Sharded join across several contexts (one per GPU; one process per GPU under torchrun, or the worker threads of a context
 * created with n_devices > 1).  EVERY context uploads the whole input and, after ejb_set_shard(rank, world), ejb_run joins only
 * its share: whole sub-buckets by longest-processing-time assignment on a pair-count cost model, the heaviest sub-buckets cut
 * into row ranges (identical, deterministic assignment on every rank; no collective on the data path).  A sharded ejb_run ends
 * before the finalize step and leaves its PARTIAL FOREST on the device:
 *   ejb_part_forest    -> device pointers to the part's edges (8-byte weights (k1 << 28) | k2, global row ids) and 32-byte tuples
 *   ejb_merge_buffers  -> on the merging context (any rank that also ran its part): device buffers for the concatenation of all
 *                         parts; the caller fills them GPU-to-GPU (NCCL gather, peer copies)
 *   ejb_merge_parts    -> forest of the partial forests (Kruskal of a union of graphs == Kruskal of the union of their forests),
 *                         two-cell filter on the merged forest, ordering, PotentialJoin members, labels
 *   ejb_download       -> edges + EquivRel of the WHOLE join.                                                               */
/* Sharded upload: every context of a sharded join needs the whole input on its GPU, but it need not cross PCIe N times.  With an
 * input arena declared (ejb_set_input_arena), context `part` of `nparts` copies only ITS slice of the arena host -> device; the
 * caller completes the device arenas GPU-to-GPU — an in-place all-gather (NCCL / peer copies over NVLink) on the buffer
 * ejb_device_arena returns, slice i at offset i * slice_bytes, total_bytes = nparts * slice_bytes — and then calls ejb_run.
 * ejb_join on a multi-device context does exactly this by itself.                                                          */
int ejb_upload_slice(ejb_ctx* ctx, const ejb_params* p, const ejb_input* in, int part, int nparts);
int ejb_device_arena(ejb_ctx* ctx, void** d_base, uint64_t* total_bytes, uint64_t* slice_bytes);

/* The assignment itself, as a pure host function (no device; tests and planning tools): sub_rows[s] = rows of sub-bucket s in the
 * library's order; r_beg[s] / n_act[s] receive the first-row range [r_beg, n_act) that `rank` of `world` joins (empty when
 * another rank owns the sub-bucket).                                                                                       */
int ejb_shard_plan(const int32_t* sub_rows, int32_t n_subs, int world, int rank, int32_t* r_beg, int32_t* n_act);
int ejb_part_forest(ejb_ctx* ctx, int64_t* n_edges, const unsigned long long** d_edges, const void** d_tuples);
int ejb_merge_buffers(ejb_ctx* ctx, int64_t total_edges, unsigned long long** d_edges, void** d_tuples);
int ejb_merge_parts(ejb_ctx* ctx, int64_t total_edges);

/* Input arena.  Every host->device copy pays a fixed setup cost (about 0.1 ms on PCIe Gen5 here), so ~36 separate input
 * arrays upload at half the bandwidth of one large copy.  A caller that flattens its records into ONE page-locked
 * allocation (the Rust shim does: it owns every buffer it passes) declares it here; ejb_upload / ejb_join then move the
 * whole arena with a single DMA and resolve every ejb_input array that lies inside it by offset.  Arrays outside the
 * arena are copied one by one as before.  base must be 256-byte aligned and readable for `bytes` bytes; the setting stays
 * in force until cleared with base == NULL.  Keep each array aligned to its element size inside the arena.            */
int ejb_set_input_arena(ejb_ctx* ctx, const void* base, uint64_t bytes);

/* Work-buffer sizing: cand_cap = capacity (entries) of the candidate / survivor / pass-edge
 * buffers; pair_budget = upper bound on the pairs one scheduler round evaluates.  0 keeps the
 * current value.  Defaults: 2^26 and 2^32 (env EJB_CAND_CAP / EJB_PAIR_BUDGET).                */
int ejb_set_capacity(ejb_ctx* ctx, unsigned long long cand_cap, unsigned long long pair_budget);

/* Host-side replay of an ordered edge list into the reference's EquivRel (join2.rs:45-54) followed,
 * when row_exact != NULL, by the clonotype_id merge (join2.rs:69-84).  Any of x/y/z/labels may be
 * NULL.  Pure host code (the EquivRel state is order-dependent and therefore sequential).       */
int ejb_equiv_replay(int32_t n_rows, const int32_t* k1, const int32_t* k2, int64_t n_edges,
                     const int32_t* row_exact, int32_t* x, int32_t* y, int32_t* z, int32_t* labels);

/* Device-resident result of the last ejb_run (valid until the next ejb_run on this context): final edge
 * list in raw_joins order and the canonical labels, as DEVICE pointers — lets a multi-GPU caller gather the
 * per-rank edges GPU-to-GPU (NCCL) without a host round trip.                                       */
int ejb_device_edges(ejb_ctx* ctx, int64_t* n_edges, const int32_t** d_k1, const int32_t** d_k2, const int32_t** d_labels);

/* PotentialJoin numerics of the same edges (device pointers, same lifetime): cd, nrefs, shares[2*i + u].      */
int ejb_device_tuples(ejb_ctx* ctx, const int32_t** d_cd, const int32_t** d_nrefs, const int32_t** d_shares);

/* Multi-GPU, one LARGE sub-bucket (equal lens and CDR3 lengths) split over several contexts by row ranges.
 * The pairs (k1 < k2) of a sub-bucket are partitioned by their FIRST row: a context that is given the rows
 * [lo, n) of the sub-bucket with roles[] = 0 for [lo, hi) and EJB_ROLE_COLUMN_ONLY for [hi, n) evaluates exactly
 * the pairs whose first row lies in [lo, hi) and returns the lexicographic spanning forest of THOSE pass edges
 * (join_core.rs:23-50 restricted to them).  Rows of a split sub-bucket also carry EJB_ROLE_SPLIT: their edges are
 * returned without the two-cell filter (join.rs:120-178), which only makes sense on the merged forest.  The
 * caller concatenates the parts' edges, sorts them by (k1, k2), keeps what ejb_forest_merge keeps (the minimum
 * spanning forest of a union of graphs is the minimum spanning forest of the union of their forests), applies
 * the two-cell filter and replays the EquivRel (enclone_ranger_b200/multi.py does all of this).
 * roles: one byte per row of the input the NEXT ejb_run / ejb_join calls work on; they stay in force on this
 * context until cleared with roles == NULL (a run on an input with another row count fails).                 */
#define EJB_ROLE_COLUMN_ONLY 1
#define EJB_ROLE_SPLIT 2
int ejb_set_row_roles(ejb_ctx* ctx, const uint8_t* roles, int64_t n_rows);
int ejb_forest_merge(int32_t n_rows, const int32_t* k1, const int32_t* k2, int64_t n_edges, uint8_t* keep);

int ejb_abi_version(void);
/* sizeof(ejb_params), sizeof(ejb_input), sizeof(ejb_stats), sizeof(ejb_result) as the library was
 * compiled — lets a binding verify its struct mirrors.                                          */
void ejb_abi_sizes(int64_t out[4]);

#ifdef __cplusplus
}
#endif
#endif /* ENCLONE_JOIN_B200_H */
