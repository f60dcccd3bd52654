/* synth.h — C interface of the synthetic repertoire generator (bench / test infrastructure).
 * Produces the flattened join input of include/enclone_join_b200.h for the BASELINE.json
 * configs (SURVEY.md §8 d).  See synth.cpp. */
#ifndef EJB_SYNTH_H
#define EJB_SYNTH_H
#include <stdint.h>
#include "../../include/enclone_join_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ejb_synth_config {
    uint64_t seed;
    int64_t n_cells;
    int64_t max_clone;        /* Zipf truncation (cells)                              */
    double zipf_s;            /* clone-size law exponent                              */
    double shm_min, shm_max;  /* per-clone nt divergence from germline, uniform range */
    double frac_has_del;      /* exacts with a V indel (has_del)                      */
    double frac_three_chain;  /* clones whose root exact carries a second light chain */
    double frac_onesie;       /* exacts with a single chain                           */
    double v_skew;            /* 0 = uniform V usage, 1 = top-5 V genes ~ 50 %        */
    double cells_per_exact_p; /* cells per exact ~ Geometric(p)                       */
    double frac_alt_allele;   /* clones whose heavy vs is a donor allele              */
    double frac_contam;       /* cells assigned to a random donor                     */
    int32_t is_bcr;
    int32_t n_datasets, n_donors;
    int32_t threads;          /* 0 = all host cores                                   */
} ejb_synth_config;

typedef struct ejb_synth_summary {
    int64_t n_cells, n_clones, n_exacts, n_rows, max_clone_cells, max_clone_exacts, tig_bytes;
} ejb_synth_summary;

typedef struct ejb_synth_data ejb_synth_data;

/* config index = position in BASELINE.json "configs" (0..4) */
void ejb_synth_preset(int config, ejb_synth_config* c);
int ejb_synth_generate(const ejb_synth_config* cfg, ejb_synth_data** out);
const ejb_input* ejb_synth_input(const ejb_synth_data* d);
const int32_t* ejb_synth_row_clone(const ejb_synth_data* d); /* ground-truth clone of each row */
void ejb_synth_get_summary(const ejb_synth_data* d, ejb_synth_summary* s);
void ejb_synth_free(ejb_synth_data* d);

#ifdef __cplusplus
}
#endif
#endif
