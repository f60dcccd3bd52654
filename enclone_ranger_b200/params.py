"""Join knobs as enclone_ranger runs them (flat POD mirror of the fields join_one/join_exacts read
from EncloneControl; defaults: enclone_args/src/proc_args.rs:36-41,152-163)."""
from ._abi import EjbParams


def default_params(is_bcr=True, **overrides):
    p = EjbParams()
    p.is_bcr = 1 if is_bcr else 0
    p.mix_donors = 0
    p.easy = 0
    p.old_light = 0
    p.unsupported_modes = 0
    p.reserved0 = 0
    p.max_score = 100000.0
    p.cdr3_mult = 5.0
    p.mult_pow = 80.0
    p.join_cdr3_ident = 85.0
    p.fwr1_cdr12_delta = 20.0
    p.max_cdr3_diffs = 1000
    p.cdr3_normal_len = 42
    p.auto_share = 15
    p.comp_filt = 8
    p.comp_filt_bound = 80
    p.max_diffs = 1000000
    p.max_degradation = 2
    p.ref_v_trim = 15
    p.ref_j_trim = 15
    for k, v in overrides.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p
