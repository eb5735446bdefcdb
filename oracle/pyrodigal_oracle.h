/* pyrodigal_oracle.h -- TEST INFRASTRUCTURE ONLY (not shipped, never on the product path).
 *
 * Plain-C CPU restatement of the Pyrodigal / Prodigal gene-finding hot path:
 *   encode -> add_nodes -> sort -> score_nodes -> record_overlapping_starts ->
 *   connection-scoring DP (with the skip filter) -> traceback -> eliminate_bad_genes ->
 *   gene extraction -> tweak_final_starts -> meta-mode model loop.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this.  Parity is PINNED: tests/test_oracle_vs_reference.py checks every function
 * below against the unmodified reference built in oracle/_ref (when present) and against
 * the golden fixtures in tests/golden/ (generated from the reference by
 * tests/golden/make_golden.py).
 *
 * Citations are relative to /root/reference (pyrodigal v3.7.1):
 *   lib.pyx = src/pyrodigal/lib.pyx, _connection.h = src/pyrodigal/_connection.h,
 *   _sequence.h = src/pyrodigal/_sequence.h, node.c / dprog.c = vendor/Prodigal/...
 */
#ifndef PYRODIGAL_ORACLE_H
#define PYRODIGAL_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_ATG 0
#define ORC_GTG 1
#define ORC_TTG 2
#define ORC_STOP 3

/* Binary-compatible with `struct _training` (vendor/Prodigal/training.h:29-51, 558392 B) so
 * that bytes(memoryview(TrainingInfo)) blobs can be passed straight through. */
typedef struct {
    double gc;
    int32_t trans_table;
    int32_t _pad0;
    double st_wt;
    double bias[3];
    double type_wt[3];
    int32_t uses_sd;
    int32_t _pad1;
    double rbs_wt[28];
    double ups_comp[32][4];
    double mot_wt[4][4][4096];
    double no_mot;
    double gene_dc[4096];
} orc_training;

/* Our own node record (NOT the reference's 128-byte layout; every field the reference's
 * `struct _node` (src/Prodigal/node.h:48-76) carries is present). */
typedef struct {
    int32_t ndx, stop_val;
    int32_t strand;   /* +1 / -1 */
    int32_t type;     /* ORC_ATG.. ORC_STOP */
    int32_t edge, elim, gc_bias;
    int32_t star_ptr[3];
    int32_t traceb, tracef, ov_mark;
    int32_t rbs[2];
    int32_t mot_ndx, mot_len, mot_spacer, mot_spacendx;
    double mot_score;
    double gc_score[3];
    double cscore, uscore, tscore, rscore, sscore, score;
    float gc_cont;
    int32_t _pad;
} orc_node;

typedef struct {
    int32_t begin, end, start_ndx, stop_ndx;
} orc_gene;

typedef struct {
    int32_t closed, min_gene, min_edge_gene, max_overlap;
    int32_t n_masks;
    const int32_t *masks; /* n_masks x [begin,end) */
} orc_opts;

/* lib.pyx:664-697 -- returns unknown count, writes gc_count */
int orc_encode(const uint8_t *ascii, int n, uint8_t *digits, int *gc_count);
/* lib.pyx:699-713 -- returns number of masks written (pairs) */
int orc_find_masks(const uint8_t *digits, int slen, int mask_size, int32_t *out, int cap);
/* lib.pyx:1905-2117 -- returns number of nodes, -1 if cap too small */
int orc_extract(const uint8_t *digits, int slen, int tt, const orc_opts *o, orc_node *out, int cap);
/* lib.pyx:2489 / node.c:1578 */
void orc_sort(orc_node *nodes, int nn);
/* node.c:176-197 */
void orc_reset_scores(orc_node *nodes, int nn);
/* lib.pyx:2331-2487 (includes _calc_orf_gc, _raw_coding_score, _rbs_score / upstream motif) */
void orc_score(const uint8_t *digits, int slen, orc_node *nodes, int nn, const orc_training *t, int closed,
               int is_meta);
/* individual pieces, exposed for unit parity tests */
void orc_calc_orf_gc(const uint8_t *digits, int slen, orc_node *nodes, int nn);
void orc_raw_coding_score(const uint8_t *digits, int slen, orc_node *nodes, int nn, const orc_training *t);
void orc_rbs_score(const uint8_t *digits, int slen, orc_node *nodes, int nn, const orc_training *t);
int orc_shine_dalgarno_exact(const uint8_t *digits, int slen, int pos, int start, const double *rbs_wt, int strand);
int orc_shine_dalgarno_mm(const uint8_t *digits, int slen, int pos, int start, const double *rbs_wt, int strand);
/* lib.pyx:2279-2329 */
void orc_record_overlapping_starts(orc_node *nodes, int nn, const orc_training *t, int flag, int max_overlap);
/* lib.pyx:1205-1237 + _connection.h:94-408 + impl/generic.h:13-49 (filter folded in).
 * `pairs` (may be NULL) receives sum_i (i - min_i). */
void orc_score_connections(orc_node *nodes, int nn, const orc_training *t, int final, int64_t *pairs);
/* skip predicate of impl/generic.h:29-36 for one (j,i) pair; 1 = skipped */
int orc_skippable(const orc_node *nodes, int j, int i);
/* lib.pyx:1297-1311 -- returns ipath or -1 */
int orc_dynamic_programming(orc_node *nodes, int nn, const orc_training *t, int final);
/* dprog.c:306-335 */
void orc_eliminate_bad_genes(orc_node *nodes, int ipath, const orc_training *t);
/* lib.pyx:3231-3270 -- returns gene count */
int orc_genes_extract(const orc_node *nodes, int ipath, orc_gene *genes, int cap);
/* lib.pyx:3272-3401 */
void orc_tweak_final_starts(orc_gene *genes, int ng, const orc_node *nodes, int nn, const orc_training *t,
                            int max_overlap);

/* lib.pyx:5281-5315.  nodes/genes are caller-provided buffers. returns ng (or -1: cap) */
int orc_find_genes_single(const uint8_t *digits, int slen, const orc_training *t, const orc_opts *o,
                          orc_node *nodes, int node_cap, int *nn_out, orc_gene *genes, int gene_cap,
                          int *ipath_out);
/* lib.pyx:5317-5396.  bins = n_bins contiguous orc_training. `gc` = Sequence.gc.
 * returns ng; *winner = bin index or -1.  nodes hold the final re-scored winner nodes. */
int orc_find_genes_meta(const uint8_t *digits, int slen, double gc, const orc_training *bins, int n_bins,
                        const orc_opts *o, orc_node *nodes, int node_cap, int *nn_out, orc_gene *genes,
                        int gene_cap, int *winner, int64_t *pairs);

/* ---- training (GeneFinder.train, lib.pyx:5236-5279) ---- */
/* lib.pyx:724-768: gp[i] = GC-richest frame (0..2) of the codon triplet holding i, -1 in the tail */
void orc_gc_frame_plot(const uint8_t *digits, int slen, int8_t *gp);
/* vendor/Prodigal/node.c:263-317: fills gc_bias / gc_score of every start and t->bias */
void orc_record_gc_bias(const int8_t *gp, orc_node *nodes, int nn, orc_training *t);
/* lib.pyx:4284-4358 */
void orc_calc_dicodon_gene(const uint8_t *digits, int slen, const orc_node *nodes, int ipath, orc_training *t);
/* lib.pyx:4391-4599 */
void orc_train_starts_sd(const uint8_t *digits, int slen, orc_node *nodes, int nn, orc_training *t);
/* vendor/Prodigal/node.c:686-693 */
void orc_determine_sd_usage(orc_training *t);
/* lib.pyx:4601-4826 */
void orc_train_starts_nonsd(const uint8_t *digits, int slen, orc_node *nodes, int nn, orc_training *t);
/* lib.pyx:5236-5279 on a zero-initialised TrainingInfo(gc, tt, st_wt).  returns nn or -1 (cap) */
int orc_train(const uint8_t *digits, int slen, double gc, int tt, double st_wt, int force_nonsd, const orc_opts *o,
              orc_node *nodes, int node_cap, orc_training *t);

int orc_sizeof_node(void);
int orc_sizeof_training(void);

#ifdef __cplusplus
}
#endif
#endif
