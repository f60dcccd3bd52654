//! Drop-in replacement for `enclone::join::join_exacts` (enclone/src/join.rs:31-41) backed by the
//! B200 CUDA library through `include/enclone_join_b200.h`.
//!
//! Same signature, same out-params (`raw_joins` appended in the reference's order; `join_info` stays
//! empty, as under the default `quiet = true`), same return type: the `EquivRel` is rebuilt with
//! `EquivRel::from_raw(x, y, z)` (equiv/src/lib.rs:57) from the vectors the library replayed.
//!
//! Written against the reference's types; NOT compiled in the development image (no Rust toolchain).  The struct mirrors are
//! checked at run time against the sizes the library reports (ejb_abi_sizes) and the ABI version (4).

use debruijn::Mer;
use enclone_core::defs::{CloneInfo, EncloneControl, ExactClonotype};
use enclone_core::enclone_structs::JoinInfo;
use enclone_proto::types::DonorReferenceItem;
use equiv::EquivRel;
use qd::Double;
use std::collections::HashMap;
use std::os::raw::{c_char, c_int, c_void};
use vdj_ann::refx::RefData;

pub mod sys {
    use super::*;

    pub const EJB_NONE: i32 = -1;

    #[repr(C)]
    #[derive(Default, Clone, Copy)]
    pub struct EjbParams {
        pub is_bcr: i32, pub mix_donors: i32, pub easy: i32, pub old_light: i32,
        pub unsupported_modes: u32, pub reserved0: i32,
        pub max_score: f64, pub cdr3_mult: f64, pub mult_pow: f64, pub join_cdr3_ident: f64, pub fwr1_cdr12_delta: f64,
        pub max_cdr3_diffs: i64, pub cdr3_normal_len: i64, pub auto_share: i64, pub comp_filt: i64, pub comp_filt_bound: i64,
        pub max_diffs: i64, pub max_degradation: i64, pub ref_v_trim: i64, pub ref_j_trim: i64,
    }

    #[repr(C)]
    pub struct EjbInput {
        pub n_rows: i64,
        pub row_nchains: *const i32, pub row_len: *const i32, pub row_tig_off: *const i64, pub row_has_del: *const u8,
        pub row_vs: *const i32, pub row_js: *const i32, pub row_cdr3_off: *const i64, pub row_cdr3_len: *const i32,
        pub row_exact: *const i32, pub row_exact_col: *const i32,
        pub tig_bytes: *const u8, pub n_tig_bytes: i64, pub cdr3_bytes: *const u8, pub n_cdr3_bytes: i64,
        pub n_exacts: i64, pub ex_chain_off: *const i64, pub ex_cell_off: *const i64,
        pub n_chains: i64, pub ch_left: *const u8, pub ch_c_ref: *const i32, pub ch_v_ref: *const i32, pub ch_u_ref: *const i32,
        pub ch_fr1_start: *const i32, pub ch_cdr1_start: *const i32, pub ch_fr2_start: *const i32, pub ch_cdr2_start: *const i32,
        pub ch_fr3_start: *const i32, pub ch_hcomp: *const i32, pub ch_amino_off: *const i64, pub ch_amino_len: *const i32,
        pub amino_bytes: *const u8, pub n_amino_bytes: i64,
        pub n_cells: i64, pub cell_donor: *const i32, pub cell_barcode: *const u32,
        pub n_segs: i64, pub seg_off: *const i64, pub seg_bytes: *const u8,
        pub n_refs: i64, pub ref_off: *const i64, pub ref_bytes: *const u8, pub ref_name_id: *const i32,
        pub row_tig2_off: *const i64, pub tig2_words: *const u32, pub n_tig2_words: i64,
        pub row_cdr32_off: *const i64, pub cdr32_words: *const u32, pub n_cdr32_words: i64,
    }

    #[repr(C)]
    pub struct EjbStats {
        pub counters: [i64; 18],      // n_buckets .. d2h_bytes, in header order
        pub ms: [f64; 8],             // ms_h2d, ms_pack, ms_stage1, ms_stage2, ms_forest, ms_finalize, ms_d2h, ms_replay
        pub ms_stage2a: f64,
        pub launches: [i64; 3],       // stage1_launches, stage2a_launches, overflow_retries
        pub ms_device_total: f64,
        pub ms_total: f64,
        pub two_ref: [i64; 5],        // stage1_ref_kills, stage2_two_ref, stage2_memo_dead, stage2_degr_dead, stage1_donor_kills
        pub executed: [i64; 8],       // pairs_distance_computed, pairs_distance_full, popc_executed_stage1, stage1_flag_skips, host_syncs,
                                      // boruvka_iterations, p1_cache_slots, n_devices
        pub ms_partition: f64,
        pub ms_merge: f64,
    }

    /// Caller-allocated outputs of ejb_join_one_batch (any pointer may be null).
    #[repr(C)]
    pub struct EjbPairOut {
        pub pass: *mut u8,
        pub cd: *mut i32, pub nrefs: *mut i32, pub shares: *mut i32, pub indeps: *mut i32, pub k: *mut i32, pub d: *mut i32, pub n: *mut i32,
        pub p1: *mut f64, pub mult: *mut f64, pub score: *mut f64,
        pub err: *mut u8,
    }

    #[repr(C)]
    pub struct EjbResult {
        pub n_edges: i64,
        pub edge_k1: *const i32, pub edge_k2: *const i32, pub edge_cd: *const i32, pub edge_nrefs: *const i32,
        pub edge_shares: *const i32, pub edge_indeps: *const i32, pub edge_k: *const i32, pub edge_d: *const i32, pub edge_n: *const i32,
        pub edge_p1: *const f64, pub edge_mult: *const f64, pub edge_score: *const f64, pub edge_err: *const u8,
        pub n_rows: i64, pub labels: *const i32, pub eq_x: *const i32, pub eq_y: *const i32, pub eq_z: *const i32,
        pub stats: EjbStats, pub opaque: *mut c_void,
    }

    #[repr(C)]
    pub struct EjbCtx { _private: [u8; 0] }

    extern "C" {
        pub fn ejb_create(out: *mut *mut EjbCtx, device_ids: *const c_int, n_devices: c_int) -> c_int;
        pub fn ejb_destroy(ctx: *mut EjbCtx);
        pub fn ejb_last_error(ctx: *const EjbCtx) -> *const c_char;
        pub fn ejb_join(ctx: *mut EjbCtx, p: *const EjbParams, input: *const EjbInput, out: *mut EjbResult) -> c_int;
        pub fn ejb_result_free(ctx: *mut EjbCtx, out: *mut EjbResult);
        pub fn ejb_abi_sizes(out: *mut i64);
        /// One page-locked allocation holding the flattened input arrays: uploaded with a single DMA.
        pub fn ejb_set_input_arena(ctx: *mut EjbCtx, base: *const c_void, bytes: u64) -> c_int;
        /// Multi-GPU only (one process per GPU): row ranges of a sub-bucket split over several contexts, and the merge of
        /// their partial forests on rank 0.
        pub fn ejb_set_row_roles(ctx: *mut EjbCtx, roles: *const u8, n_rows: i64) -> c_int;
        pub fn ejb_forest_merge(n_rows: i32, k1: *const i32, k2: *const i32, n_edges: i64, keep: *mut u8) -> c_int;
        /// join_one for arbitrary pairs of the rows of the last ejb_join on this context (define_mat.rs:172,247).
        pub fn ejb_join_one_batch(ctx: *mut EjbCtx, k1: *const i32, k2: *const i32, n_pairs: i64, out: *const EjbPairOut) -> c_int;
        pub fn ejb_abi_version() -> c_int;
    }
}

/// ONE context per process (ejb_create builds the Stirling table on the device and allocates the work buffers: not per call).
/// `ENCLONE_JOIN_GPUS=n` makes it span the first n devices; ejb_join then shards the buckets over them by itself.
struct Ctx(*mut sys::EjbCtx);
unsafe impl Send for Ctx {}
unsafe impl Sync for Ctx {}
static CTX: std::sync::OnceLock<std::sync::Mutex<Ctx>> = std::sync::OnceLock::new();
fn context() -> &'static std::sync::Mutex<Ctx> {
    CTX.get_or_init(|| unsafe {
        assert!(sys::ejb_abi_version() == 4, "enclone_join_b200: library / binding ABI mismatch");
        let mut sizes = [0i64; 4];
        sys::ejb_abi_sizes(sizes.as_mut_ptr());
        assert!(
            sizes == [std::mem::size_of::<sys::EjbParams>() as i64, std::mem::size_of::<sys::EjbInput>() as i64,
                      std::mem::size_of::<sys::EjbStats>() as i64, std::mem::size_of::<sys::EjbResult>() as i64],
            "enclone_join_b200: struct mirrors out of date"
        );
        let n: c_int = std::env::var("ENCLONE_JOIN_GPUS").ok().and_then(|v| v.parse().ok()).unwrap_or(1);
        let mut ctx: *mut sys::EjbCtx = std::ptr::null_mut();
        let rc = sys::ejb_create(&mut ctx, std::ptr::null(), n);
        assert!(rc == 0, "enclone_join_b200: ejb_create failed ({rc}); a CUDA device is required (no CPU fallback)");
        std::sync::Mutex::new(Ctx(ctx))
    })
}

/// The second use of the pair predicate (enclone_process/src/define_mat.rs:172,247): `join_one(k1, k2, ..)` for many pairs of
/// `info` rows at once, on the rows of the last `join_exacts` call.  Returns what join_one returns for each pair.
pub fn join_one_batch(pairs: &[(usize, usize)]) -> Vec<bool> {
    let (k1, k2): (Vec<i32>, Vec<i32>) = pairs.iter().map(|(a, b)| (*a as i32, *b as i32)).unzip();
    let mut pass = vec![0u8; pairs.len()];
    let out = sys::EjbPairOut {
        pass: pass.as_mut_ptr(), cd: std::ptr::null_mut(), nrefs: std::ptr::null_mut(), shares: std::ptr::null_mut(), indeps: std::ptr::null_mut(),
        k: std::ptr::null_mut(), d: std::ptr::null_mut(), n: std::ptr::null_mut(), p1: std::ptr::null_mut(), mult: std::ptr::null_mut(),
        score: std::ptr::null_mut(), err: std::ptr::null_mut(),
    };
    let g = context().lock().unwrap();
    let rc = unsafe { sys::ejb_join_one_batch(g.0, k1.as_ptr(), k2.as_ptr(), pairs.len() as i64, &out) };
    assert!(rc == 0, "enclone_join_b200: ejb_join_one_batch failed ({rc})");
    pass.into_iter().map(|b| b != 0).collect()
}

/// EJB_MODE_* bits for the join modes the CUDA path rejects (they are unreachable from enclone_ranger).
fn unsupported_modes(ctl: &EncloneControl) -> u32 {
    let j = &ctl.join_alg_opt;
    (j.bcjoin as u32) | (j.basic.is_some() as u32) << 1 | (j.basic_h.is_some() as u32) << 2 | (j.basicx as u32) << 3
        | (j.join_full_diff as u32) << 4 | (j.old_mult as u32) << 5 | (ctl.force as u32) << 6
        | ((j.super_comp_filt > 0) as u32) << 7
}

/// Same signature as enclone::join::join_exacts.  `sr` and `dref` are accepted for signature
/// compatibility: the library owns its own Stirling table, `dref` only feeds the disabled
/// SUPER_COMP_FILT branch.
#[allow(clippy::too_many_arguments)]
pub fn join_exacts(
    _to_bc: &HashMap<(usize, usize), Vec<String>>,
    refdata: &RefData,
    ctl: &EncloneControl,
    exact_clonotypes: &[ExactClonotype],
    info: &[CloneInfo],
    _join_info: &mut Vec<JoinInfo>,
    raw_joins: &mut Vec<(i32, i32)>,
    _sr: &[Vec<Double>],
    _dref: &[DonorReferenceItem],
) -> EquivRel {
    use sys::*;
    // ---- knobs ----
    let j = &ctl.join_alg_opt;
    let p = EjbParams {
        is_bcr: ctl.gen_opt.is_bcr() as i32, mix_donors: ctl.clono_filt_opt_def.donor as i32, easy: j.easy as i32,
        old_light: j.old_light as i32, unsupported_modes: unsupported_modes(ctl), reserved0: 0,
        max_score: j.max_score, cdr3_mult: j.cdr3_mult, mult_pow: j.mult_pow, join_cdr3_ident: j.join_cdr3_ident,
        fwr1_cdr12_delta: j.fwr1_cdr12_delta, max_cdr3_diffs: j.max_cdr3_diffs.min(i64::MAX as usize) as i64,
        cdr3_normal_len: j.cdr3_normal_len as i64, auto_share: j.auto_share as i64, comp_filt: j.comp_filt.min(i64::MAX as usize) as i64,
        comp_filt_bound: j.comp_filt_bound as i64, max_diffs: ctl.heur.max_diffs as i64, max_degradation: ctl.heur.max_degradation as i64,
        ref_v_trim: ctl.heur.ref_v_trim as i64, ref_j_trim: ctl.heur.ref_j_trim as i64,
    };

    // ---- flatten info[] ----
    let n = info.len();
    let (mut row_nchains, mut row_exact) = (Vec::with_capacity(n), Vec::with_capacity(n));
    let mut row_len = vec![0i32; 2 * n];
    let mut row_tig_off = vec![0i64; 2 * n];
    let mut row_has_del = vec![0u8; 2 * n];
    let (mut row_vs, mut row_js) = (vec![0i32; 2 * n], vec![0i32; 2 * n]);
    let (mut row_cdr3_off, mut row_cdr3_len) = (vec![0i64; 2 * n], vec![0i32; 2 * n]);
    let mut row_exact_col = vec![0i32; 2 * n];
    let (mut tig_bytes, mut cdr3_bytes) = (Vec::<u8>::new(), Vec::<u8>::new());
    let mut row_tig2_off = vec![-1i64; 2 * n];
    let mut tig2_words = Vec::<u32>::new();   // 2-bit packed contigs (from tigsp) for rows without indel marks
    let mut row_cdr32_off = vec![-1i64; 2 * n];
    let mut cdr32_words = Vec::<u32>::new();  // 2-bit packed CDR3s (same packing); a CDR3 with another byte stays ASCII
    let mut segs: HashMap<Vec<u8>, i32> = HashMap::new();   // DnaString content -> seg id (join_one.rs:331 compares content)
    let (mut seg_off, mut seg_bytes) = (vec![0i64], Vec::<u8>::new());
    let mut seg_id = |s: &debruijn::dna_string::DnaString| -> i32 {
        let a = s.to_ascii_vec();
        let next = segs.len() as i32;
        *segs.entry(a.clone()).or_insert_with(|| {
            seg_bytes.extend_from_slice(&a);
            seg_off.push(seg_bytes.len() as i64);
            next
        })
    };
    for (k, x) in info.iter().enumerate() {
        // the reference indexes exact_clonotypes with clonotype_index for donors / ncells / shares (join_one.rs:303-304,
        // join.rs:146) and with clonotype_id for to_bc / finish_join; build_info sets both to the same value (info.rs:297-298)
        assert!(x.clonotype_id == x.clonotype_index, "enclone_join_b200: info[{k}].clonotype_id != clonotype_index");
        row_nchains.push(x.lens.len() as i32);
        row_exact.push(x.clonotype_id as i32);
        for m in 0..x.lens.len().min(2) {
            let i = 2 * k + m;
            row_len[i] = x.lens[m] as i32;
            if !x.has_del[m] && x.tigs[m].iter().all(|b| matches!(b, b'A' | b'C' | b'G' | b'T')) {
                // tigsp[m] holds the same bases 2-bit packed (info.rs:64); re-pack 16 per u32, base i at bits 2*(i%16)
                row_tig2_off[i] = tig2_words.len() as i64;
                row_tig_off[i] = -1;
                let sp = &x.tigsp[m];
                for w in 0..(sp.len() + 15) / 16 {
                    let mut v = 0u32;
                    for j in 0..16.min(sp.len() - 16 * w) {
                        v |= (sp.get(16 * w + j) as u32) << (2 * j);
                    }
                    tig2_words.push(v);
                }
            } else {
                row_tig_off[i] = tig_bytes.len() as i64;
                tig_bytes.extend_from_slice(&x.tigs[m]);
            }
            row_has_del[i] = x.has_del[m] as u8;
            row_vs[i] = seg_id(&x.vs[m]);
            row_js[i] = seg_id(&x.js[m]);
            row_cdr3_len[i] = x.cdr3s[m].len() as i32;
            let c3 = x.cdr3s[m].as_bytes();
            if c3.iter().all(|b| matches!(b, b'A' | b'C' | b'G' | b'T')) {
                row_cdr32_off[i] = cdr32_words.len() as i64;
                row_cdr3_off[i] = -1;
                for w in c3.chunks(16) {
                    let mut v = 0u32;
                    for (j, b) in w.iter().enumerate() {
                        let code = match b { b'C' => 1u32, b'G' => 2, b'T' => 3, _ => 0 };
                        v |= code << (2 * j);
                    }
                    cdr32_words.push(v);
                }
            } else {
                row_cdr3_off[i] = cdr3_bytes.len() as i64;
                cdr3_bytes.extend_from_slice(c3);
            }
            row_exact_col[i] = x.exact_cols[m] as i32;
        }
    }

    // ---- flatten exact_clonotypes[] (share = chains, clones = cells); barcodes interned as strings (join_one.rs:451-470) ----
    let opt = |v: Option<usize>| v.map_or(EJB_NONE, |x| x as i32);
    let (mut ex_chain_off, mut ex_cell_off) = (vec![0i64], vec![0i64]);
    let (mut ch_left, mut ch_c_ref, mut ch_v_ref, mut ch_u_ref) = (vec![], vec![], vec![], vec![]);
    let (mut fr1, mut cdr1, mut fr2, mut cdr2, mut fr3, mut hcomp) = (vec![], vec![], vec![], vec![], vec![], vec![]);
    let (mut amino_off, mut amino_len, mut amino_bytes) = (Vec::<i64>::new(), Vec::<i32>::new(), Vec::<u8>::new());
    let (mut cell_donor, mut cell_barcode) = (Vec::<i32>::new(), Vec::<u32>::new());
    let mut barcodes: HashMap<&str, u32> = HashMap::new();
    for ex in exact_clonotypes {
        for s in &ex.share {
            ch_left.push(s.left as u8);
            ch_c_ref.push(opt(s.c_ref_id));
            ch_v_ref.push(s.v_ref_id as i32);
            ch_u_ref.push(opt(s.u_ref_id));
            fr1.push(s.fr1_start as i32);
            cdr1.push(opt(s.cdr1_start));
            fr2.push(opt(s.fr2_start));
            cdr2.push(opt(s.cdr2_start));
            fr3.push(opt(s.fr3_start));
            hcomp.push(s.jun.hcomp as i32);
            if s.seq_del_amino == s.seq_del {
                amino_off.push(EJB_NONE as i64);   // byte-identical to the row's contig (info.rs:276-277)
                amino_len.push(0);
            } else {
                amino_off.push(amino_bytes.len() as i64);
                amino_len.push(s.seq_del_amino.len() as i32);
                amino_bytes.extend_from_slice(&s.seq_del_amino);
            }
        }
        ex_chain_off.push(ch_left.len() as i64);
        for c in &ex.clones {
            cell_donor.push(opt(c[0].donor_index));
            let next = barcodes.len() as u32;
            cell_barcode.push(*barcodes.entry(c[0].barcode.as_str()).or_insert(next));
        }
        ex_cell_off.push(cell_donor.len() as i64);
    }

    // ---- refdata: sequences + names with any "*..." suffix removed (join_one.rs:727-732) ----
    let (mut ref_off, mut ref_bytes, mut ref_name_id) = (vec![0i64], Vec::<u8>::new(), Vec::<i32>::new());
    let mut names: HashMap<String, i32> = HashMap::new();
    for (r, name) in refdata.refs.iter().zip(refdata.name.iter()) {
        ref_bytes.extend_from_slice(&r.to_ascii_vec());
        ref_off.push(ref_bytes.len() as i64);
        let stripped = name.split('*').next().unwrap().to_string();
        let next = names.len() as i32;
        ref_name_id.push(*names.entry(stripped).or_insert(next));
    }

    let input = EjbInput {
        n_rows: n as i64,
        row_nchains: row_nchains.as_ptr(), row_len: row_len.as_ptr(), row_tig_off: row_tig_off.as_ptr(), row_has_del: row_has_del.as_ptr(),
        row_vs: row_vs.as_ptr(), row_js: row_js.as_ptr(), row_cdr3_off: row_cdr3_off.as_ptr(), row_cdr3_len: row_cdr3_len.as_ptr(),
        row_exact: row_exact.as_ptr(), row_exact_col: row_exact_col.as_ptr(),
        tig_bytes: tig_bytes.as_ptr(), n_tig_bytes: tig_bytes.len() as i64, cdr3_bytes: cdr3_bytes.as_ptr(), n_cdr3_bytes: cdr3_bytes.len() as i64,
        n_exacts: exact_clonotypes.len() as i64, ex_chain_off: ex_chain_off.as_ptr(), ex_cell_off: ex_cell_off.as_ptr(),
        n_chains: ch_left.len() as i64, ch_left: ch_left.as_ptr(), ch_c_ref: ch_c_ref.as_ptr(), ch_v_ref: ch_v_ref.as_ptr(), ch_u_ref: ch_u_ref.as_ptr(),
        ch_fr1_start: fr1.as_ptr(), ch_cdr1_start: cdr1.as_ptr(), ch_fr2_start: fr2.as_ptr(), ch_cdr2_start: cdr2.as_ptr(), ch_fr3_start: fr3.as_ptr(),
        ch_hcomp: hcomp.as_ptr(), ch_amino_off: amino_off.as_ptr(), ch_amino_len: amino_len.as_ptr(),
        amino_bytes: amino_bytes.as_ptr(), n_amino_bytes: amino_bytes.len() as i64,
        n_cells: cell_donor.len() as i64, cell_donor: cell_donor.as_ptr(), cell_barcode: cell_barcode.as_ptr(),
        n_segs: (seg_off.len() - 1) as i64, seg_off: seg_off.as_ptr(), seg_bytes: seg_bytes.as_ptr(),
        n_refs: refdata.refs.len() as i64, ref_off: ref_off.as_ptr(), ref_bytes: ref_bytes.as_ptr(), ref_name_id: ref_name_id.as_ptr(),
        row_tig2_off: row_tig2_off.as_ptr(), tig2_words: tig2_words.as_ptr(), n_tig2_words: tig2_words.len() as i64,
        row_cdr32_off: row_cdr32_off.as_ptr(), cdr32_words: cdr32_words.as_ptr(), n_cdr32_words: cdr32_words.len() as i64,
    };

    // ---- call the library; errors become panics, like the reference's own invariant failures ----
    let g = context().lock().unwrap();
    unsafe {
        let ctx = g.0;
        let mut res: EjbResult = std::mem::zeroed();
        let rc = ejb_join(ctx, &p, &input, &mut res);
        if rc != 0 {
            let msg = std::ffi::CStr::from_ptr(ejb_last_error(ctx)).to_string_lossy().into_owned();
            panic!("enclone_join_b200: join failed ({rc}): {msg}");
        }
        let e = res.n_edges as usize;
        let (k1, k2) = (std::slice::from_raw_parts(res.edge_k1, e), std::slice::from_raw_parts(res.edge_k2, e));
        raw_joins.extend(k1.iter().zip(k2.iter()).map(|(a, b)| (*a, *b)));   // join.rs:584-588
        let eq = EquivRel::from_raw(
            std::slice::from_raw_parts(res.eq_x, n).to_vec(),
            std::slice::from_raw_parts(res.eq_y, n).to_vec(),
            std::slice::from_raw_parts(res.eq_z, n).to_vec(),
        );
        ejb_result_free(ctx, &mut res);   // the context (and the packed rows join_one_batch works on) stays alive for the process
        eq
    }
}
