"""ctypes mirror of include/enclone_join_b200.h (+ synth.h).

Field order and types must match the C headers exactly; tests/test_abi.py checks sizes against
the values the shared library reports.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

EJB_OK = 0
EJB_ERR_INVALID = -1
EJB_ERR_UNSUPPORTED = -2
EJB_ERR_CUDA = -3
EJB_ERR_OOM = -4
EJB_ERR_RANGE = -5
EJB_ERR_NO_DEVICE = -6
EJB_NONE = -1

EJB_MODE_BCJOIN = 1
EJB_MODE_BASIC = 2
EJB_MODE_BASIC_H = 4
EJB_MODE_BASICX = 8
EJB_MODE_JOIN_FULL_DIFF = 16
EJB_MODE_OLD_MULT = 32
EJB_MODE_FORCE = 64
EJB_MODE_SUPER_COMP_FILT = 128

_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_f64p = C.POINTER(C.c_double)


class EjbParams(C.Structure):
    _fields_ = [
        ("is_bcr", C.c_int32),
        ("mix_donors", C.c_int32),
        ("easy", C.c_int32),
        ("old_light", C.c_int32),
        ("unsupported_modes", C.c_uint32),
        ("reserved0", C.c_int32),
        ("max_score", C.c_double),
        ("cdr3_mult", C.c_double),
        ("mult_pow", C.c_double),
        ("join_cdr3_ident", C.c_double),
        ("fwr1_cdr12_delta", C.c_double),
        ("max_cdr3_diffs", C.c_int64),
        ("cdr3_normal_len", C.c_int64),
        ("auto_share", C.c_int64),
        ("comp_filt", C.c_int64),
        ("comp_filt_bound", C.c_int64),
        ("max_diffs", C.c_int64),
        ("max_degradation", C.c_int64),
        ("ref_v_trim", C.c_int64),
        ("ref_j_trim", C.c_int64),
    ]


class EjbInput(C.Structure):
    _fields_ = [
        ("n_rows", C.c_int64),
        ("row_nchains", _i32p),
        ("row_len", _i32p),
        ("row_tig_off", _i64p),
        ("row_has_del", _u8p),
        ("row_vs", _i32p),
        ("row_js", _i32p),
        ("row_cdr3_off", _i64p),
        ("row_cdr3_len", _i32p),
        ("row_exact", _i32p),
        ("row_exact_col", _i32p),
        ("tig_bytes", _u8p),
        ("n_tig_bytes", C.c_int64),
        ("cdr3_bytes", _u8p),
        ("n_cdr3_bytes", C.c_int64),
        ("n_exacts", C.c_int64),
        ("ex_chain_off", _i64p),
        ("ex_cell_off", _i64p),
        ("n_chains", C.c_int64),
        ("ch_left", _u8p),
        ("ch_c_ref", _i32p),
        ("ch_v_ref", _i32p),
        ("ch_u_ref", _i32p),
        ("ch_fr1_start", _i32p),
        ("ch_cdr1_start", _i32p),
        ("ch_fr2_start", _i32p),
        ("ch_cdr2_start", _i32p),
        ("ch_fr3_start", _i32p),
        ("ch_hcomp", _i32p),
        ("ch_amino_off", _i64p),
        ("ch_amino_len", _i32p),
        ("amino_bytes", _u8p),
        ("n_amino_bytes", C.c_int64),
        ("n_cells", C.c_int64),
        ("cell_donor", _i32p),
        ("cell_barcode", _u32p),
        ("n_segs", C.c_int64),
        ("seg_off", _i64p),
        ("seg_bytes", _u8p),
        ("n_refs", C.c_int64),
        ("ref_off", _i64p),
        ("ref_bytes", _u8p),
        ("ref_name_id", _i32p),
        ("row_tig2_off", _i64p),
        ("tig2_words", _u32p),
        ("n_tig2_words", C.c_int64),
        ("row_cdr32_off", _i64p),
        ("cdr32_words", _u32p),
        ("n_cdr32_words", C.c_int64),
    ]


class EjbStats(C.Structure):
    _fields_ = [
        ("n_buckets", C.c_int64),
        ("n_subbuckets", C.c_int64),
        ("pairs_lens", C.c_int64),
        ("pairs_scored", C.c_int64),
        ("pairs_evaluated", C.c_int64),
        ("stage1_candidates", C.c_int64),
        ("stage2_slow", C.c_int64),
        ("stage2_survivors", C.c_int64),
        ("pass_edges", C.c_int64),
        ("forest_edges", C.c_int64),
        ("two_cell_deleted", C.c_int64),
        ("joins", C.c_int64),
        ("errors", C.c_int64),
        ("p1_unique", C.c_int64),
        ("rounds", C.c_int64),
        ("kernel_launches", C.c_int64),
        ("h2d_bytes", C.c_int64),
        ("d2h_bytes", C.c_int64),
        ("ms_h2d", C.c_double),
        ("ms_pack", C.c_double),
        ("ms_stage1", C.c_double),
        ("ms_stage2", C.c_double),
        ("ms_forest", C.c_double),
        ("ms_finalize", C.c_double),
        ("ms_d2h", C.c_double),
        ("ms_replay", C.c_double),
        ("ms_stage2a", C.c_double),
        ("stage1_launches", C.c_int64),
        ("stage2a_launches", C.c_int64),
        ("overflow_retries", C.c_int64),
        ("ms_device_total", C.c_double),
        ("ms_total", C.c_double),
        ("stage1_ref_kills", C.c_int64),
        ("stage2_two_ref", C.c_int64),
        ("stage2_memo_dead", C.c_int64),
        ("stage2_degr_dead", C.c_int64),
        ("stage1_donor_kills", C.c_int64),
        ("pairs_distance_computed", C.c_int64),
        ("pairs_distance_full", C.c_int64),
        ("popc_executed_stage1", C.c_int64),
        ("stage1_flag_skips", C.c_int64),
        ("host_syncs", C.c_int64),
        ("boruvka_iterations", C.c_int64),
        ("p1_cache_slots", C.c_int64),
        ("n_devices", C.c_int64),
        ("ms_partition", C.c_double),
        ("ms_merge", C.c_double),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class EjbResult(C.Structure):
    _fields_ = [
        ("n_edges", C.c_int64),
        ("edge_k1", _i32p),
        ("edge_k2", _i32p),
        ("edge_cd", _i32p),
        ("edge_nrefs", _i32p),
        ("edge_shares", _i32p),
        ("edge_indeps", _i32p),
        ("edge_k", _i32p),
        ("edge_d", _i32p),
        ("edge_n", _i32p),
        ("edge_p1", _f64p),
        ("edge_mult", _f64p),
        ("edge_score", _f64p),
        ("edge_err", _u8p),
        ("n_rows", C.c_int64),
        ("labels", _i32p),
        ("eq_x", _i32p),
        ("eq_y", _i32p),
        ("eq_z", _i32p),
        ("stats", EjbStats),
        ("opaque", C.c_void_p),
    ]


class EjbPairOut(C.Structure):
    _fields_ = [
        ("passed", _u8p),
        ("cd", _i32p), ("nrefs", _i32p), ("shares", _i32p), ("indeps", _i32p), ("k", _i32p), ("d", _i32p), ("n", _i32p),
        ("p1", _f64p), ("mult", _f64p), ("score", _f64p),
        ("err", _u8p),
    ]


class SynthConfig(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64),
        ("n_cells", C.c_int64),
        ("max_clone", C.c_int64),
        ("zipf_s", C.c_double),
        ("shm_min", C.c_double),
        ("shm_max", C.c_double),
        ("frac_has_del", C.c_double),
        ("frac_three_chain", C.c_double),
        ("frac_onesie", C.c_double),
        ("v_skew", C.c_double),
        ("cells_per_exact_p", C.c_double),
        ("frac_alt_allele", C.c_double),
        ("frac_contam", C.c_double),
        ("is_bcr", C.c_int32),
        ("n_datasets", C.c_int32),
        ("n_donors", C.c_int32),
        ("threads", C.c_int32),
    ]


class SynthSummary(C.Structure):
    _fields_ = [
        ("n_cells", C.c_int64),
        ("n_clones", C.c_int64),
        ("n_exacts", C.c_int64),
        ("n_rows", C.c_int64),
        ("max_clone_cells", C.c_int64),
        ("max_clone_exacts", C.c_int64),
        ("tig_bytes", C.c_int64),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


LIB_JOIN = os.path.join(HERE, "libenclone_join_b200.so")
LIB_SYNTH = os.path.join(HERE, "libejb_synth.so")

_cache = {}


def load_synth():
    if "synth" not in _cache:
        if not os.path.exists(LIB_SYNTH):
            raise RuntimeError(f"{LIB_SYNTH} is missing — run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(LIB_SYNTH)
        L.ejb_synth_preset.argtypes = [C.c_int, C.POINTER(SynthConfig)]
        L.ejb_synth_preset.restype = None
        L.ejb_synth_generate.argtypes = [C.POINTER(SynthConfig), C.POINTER(C.c_void_p)]
        L.ejb_synth_generate.restype = C.c_int
        L.ejb_synth_input.argtypes = [C.c_void_p]
        L.ejb_synth_input.restype = C.POINTER(EjbInput)
        L.ejb_synth_row_clone.argtypes = [C.c_void_p]
        L.ejb_synth_row_clone.restype = _i32p
        L.ejb_synth_get_summary.argtypes = [C.c_void_p, C.POINTER(SynthSummary)]
        L.ejb_synth_get_summary.restype = None
        L.ejb_synth_free.argtypes = [C.c_void_p]
        L.ejb_synth_free.restype = None
        _cache["synth"] = L
    return _cache["synth"]


def load_join():
    """Load the CUDA join library.  There is no CPU fallback: a missing library is an error."""
    if "join" not in _cache:
        if not os.path.exists(LIB_JOIN):
            raise RuntimeError(
                f"{LIB_JOIN} is missing — the CUDA extension must be built "
                "(`python -c 'import __graft_entry__ as g; g.build()'`); there is no CPU fallback")
        L = C.CDLL(os.environ.get("EJB_LIB", LIB_JOIN))   # EJB_LIB: developer override to A/B a differently built library
        L.ejb_abi_version.restype = C.c_int
        L.ejb_abi_sizes.argtypes = [_i64p]
        L.ejb_abi_sizes.restype = None
        L.ejb_params_default.argtypes = [C.POINTER(EjbParams), C.c_int]
        L.ejb_params_default.restype = None
        L.ejb_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_int]
        L.ejb_create.restype = C.c_int
        L.ejb_destroy.argtypes = [C.c_void_p]
        L.ejb_destroy.restype = None
        L.ejb_last_error.argtypes = [C.c_void_p]
        L.ejb_last_error.restype = C.c_char_p
        L.ejb_join.argtypes = [C.c_void_p, C.POINTER(EjbParams), C.POINTER(EjbInput), C.POINTER(EjbResult)]
        L.ejb_join.restype = C.c_int
        L.ejb_result_free.argtypes = [C.c_void_p, C.POINTER(EjbResult)]
        L.ejb_result_free.restype = None
        L.ejb_upload.argtypes = [C.c_void_p, C.POINTER(EjbParams), C.POINTER(EjbInput)]
        L.ejb_upload.restype = C.c_int
        L.ejb_run.argtypes = [C.c_void_p]
        L.ejb_run.restype = C.c_int
        L.ejb_download.argtypes = [C.c_void_p, C.POINTER(EjbResult)]
        L.ejb_download.restype = C.c_int
        L.ejb_last_stats.argtypes = [C.c_void_p, C.POINTER(EjbStats)]
        L.ejb_last_stats.restype = C.c_int
        L.ejb_p1_batch.argtypes = [C.c_void_p, _i32p, _i32p, _i32p, C.c_int64, _f64p]
        L.ejb_p1_batch.restype = C.c_int
        L.ejb_measure_peaks.argtypes = [C.c_void_p, _f64p, _f64p, _f64p]
        L.ejb_measure_peaks.restype = C.c_int
        L.ejb_set_input_arena.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        L.ejb_set_input_arena.restype = C.c_int
        L.ejb_set_row_roles.argtypes = [C.c_void_p, _u8p, C.c_int64]
        L.ejb_set_row_roles.restype = C.c_int
        L.ejb_forest_merge.argtypes = [C.c_int32, _i32p, _i32p, C.c_int64, _u8p]
        L.ejb_forest_merge.restype = C.c_int
        L.ejb_device_tuples.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        L.ejb_device_tuples.restype = C.c_int
        L.ejb_join_one_batch.argtypes = [C.c_void_p, _i32p, _i32p, C.c_int64, C.POINTER(EjbPairOut)]
        L.ejb_join_one_batch.restype = C.c_int
        L.ejb_upload_slice.argtypes = [C.c_void_p, C.POINTER(EjbParams), C.POINTER(EjbInput), C.c_int, C.c_int]
        L.ejb_upload_slice.restype = C.c_int
        L.ejb_device_arena.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.ejb_device_arena.restype = C.c_int
        L.ejb_set_shard.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ejb_set_shard.restype = C.c_int
        L.ejb_part_forest.argtypes = [C.c_void_p, _i64p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        L.ejb_part_forest.restype = C.c_int
        L.ejb_merge_buffers.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        L.ejb_merge_buffers.restype = C.c_int
        L.ejb_merge_parts.argtypes = [C.c_void_p, C.c_int64]
        L.ejb_merge_parts.restype = C.c_int
        L.ejb_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_longlong]
        L.ejb_set_option.restype = C.c_int
        L.ejb_create_ms.argtypes = [C.c_void_p]
        L.ejb_create_ms.restype = C.c_double
        L.ejb_sr_table_read.argtypes = [C.c_void_p, C.c_int64, C.c_int64, _f64p]
        L.ejb_sr_table_read.restype = C.c_int
        _cache["join"] = L
    return _cache["join"]
