/* pyrodigal_b200.h -- C ABI of the B200-native gene-finding hot path.
 *
 * Drop-in boundary for ONE path of Pyrodigal (reference = /root/reference, v3.7.1):
 *   Sequence._build -> Nodes._extract -> Nodes._sort -> Nodes._score ->
 *   Nodes._record_overlapping_starts -> ConnectionScorer._dynamic_programming ->
 *   eliminate_bad_genes -> Genes._extract -> Genes._tweak_final_starts
 * i.e. everything GeneFinder.find_genes() runs inside its `with nogil:` block
 * (src/pyrodigal/lib.pyx:5400-5469, _find_genes_meta 5317-5396, _find_genes_single 5281-5315).
 *
 * Conventions: every function returns 0 on success or a negative PGPU_E* code; the message is
 * available from pgpu_last_error().  Inputs are borrowed for the duration of the call, outputs
 * are written into caller-owned buffers; pgpu_result objects are owned by the library until
 * pgpu_result_free().  Plain pointers and sizes only -- no torch / CUDA types in signatures.
 * All compute runs on the GPU: there is NO CPU fallback; without a CUDA device pgpu_create fails.
 * Threading: a context owns one stream; calls on the SAME context must not overlap (the Python mirror
 * holds a lock per context), different contexts may be used from different threads concurrently.
 * (The reference's find_genes is re-entrant because every call allocates its own scorer / nodes / genes,
 * lib.pyx:5424-5426; here the per-call state lives in the context's stream-ordered workspace.)
 */
#ifndef PYRODIGAL_B200_H
#define PYRODIGAL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PGPU_OK 0
#define PGPU_ENODEV (-1)   /* no usable CUDA device            (reference: n/a)                  */
#define PGPU_ENOMEM (-2)   /* allocation failed                (reference: MemoryError)          */
#define PGPU_EINVAL (-3)   /* bad argument                     (reference: ValueError)           */
#define PGPU_ESTATE (-4)   /* e.g. no models loaded            (reference: RuntimeError)         */
#define PGPU_ECUDA (-5)    /* CUDA runtime error               (reference: n/a)                  */

#define PGPU_TRAINING_SIZE 558392 /* sizeof(struct _training), vendor/Prodigal/training.h:29-51 */

typedef struct pgpu_ctx pgpu_ctx;
typedef struct pgpu_result pgpu_result;

/* Mirrors the GeneFinder constructor keywords (lib.pyx:5102-5115). */
typedef struct {
    int32_t meta;          /* 1 = try every loaded model inside the GC window (lib.pyx:5335-5342) */
    int32_t single_model;  /* meta == 0: index of the loaded model to use                         */
    int32_t closed;        /* lib.pyx:5108 */
    int32_t mask;          /* lib.pyx:5109 : mask runs of >= min_mask unknown bases               */
    int32_t min_mask;      /* default 50 */
    int32_t min_gene;      /* default 90 */
    int32_t min_edge_gene; /* default 60 */
    int32_t max_overlap;   /* default 60 */
    int32_t want_nodes;    /* 1 = keep the final node arrays for pgpu_result_nodes()              */
    int32_t reserved[7];
} pgpu_opts;

/* `_gene` of lib.pyx:2604-2608: 1-based inclusive coordinates + node indices. */
typedef struct {
    int32_t begin, end, start_ndx, stop_ndx;
} pgpu_gene;

/* One node of the final (winning model) node array: the fields of `struct _node`
 * (src/Prodigal/node.h:48-76) that Node / Gene accessors read (lib.pyx:1440-1552, 2644-3047). */
typedef struct {
    int32_t ndx, stop_val;
    int32_t traceb, tracef;
    int32_t star_ptr[3];
    int8_t strand;    /* +1 / -1 */
    uint8_t type;     /* 0 ATG, 1 GTG, 2 TTG, 3 STOP */
    uint8_t edge, elim;
    int8_t ov_mark;
    uint8_t rbs[2];
    uint8_t mot_len;
    uint16_t mot_ndx;
    uint8_t mot_spacer, mot_spacendx;
    float gc_cont;
    double mot_score;
    double cscore, uscore, tscore, rscore, sscore, score;
} pgpu_node;

/* Per-contig summary. */
typedef struct {
    int32_t n_genes;
    int32_t n_nodes;   /* nodes of the winning model's translation table (0 if no winner)        */
    int32_t winner;    /* index of the winning model, -1 if none (lib.pyx:5396)                   */
    int32_t ipath;     /* last node of the best path, -1 if none                                  */
    int32_t unknown;   /* Sequence.unknown                                                        */
    int32_t gc_count;  /* number of G/C letters; Sequence.gc = gc_count / slen                    */
    double score;      /* DP score of the winning path (nodes[ipath].score at selection time)    */
} pgpu_contig_summary;

/* Work/timing counters of one call (device times from CUDA events on the library's stream). */
typedef struct {
    int64_t n_contigs, total_bp, total_nodes, total_chain_nodes, n_chains, total_genes;
    int64_t pairs;           /* sum over chains of sum_i (i - min_i): SURVEY.md 8(d) node pairs   */
    int64_t dp_steps;        /* sum over chains of nodes (one DP step = one node of one model)    */
    int64_t h2d_bytes, d2h_bytes;
    int64_t kernel_launches;
    double ms_total_device;  /* first kernel start -> last kernel end                             */
    double ms_encode, ms_extract, ms_score, ms_overlap, ms_dp, ms_trace, ms_final;
    double ms_h2d, ms_d2h;
    double reserved[4];
} pgpu_stats;

/* -------------------------------------------------------------------------------------------- */
/* context                                                                                      */
/* -------------------------------------------------------------------------------------------- */

/* One context per device (owns a stream, a workspace arena and the uploaded models).
 * Replaces: per-call ConnectionScorer/Nodes/Genes allocation, lib.pyx:5424-5426. */
int pgpu_create(int device, pgpu_ctx **out);
void pgpu_destroy(pgpu_ctx *ctx);
/* Message of the last failing call on this context ("" if none). ctx may be NULL (creation). */
const char *pgpu_last_error(const pgpu_ctx *ctx);

/* Upload `n` training structs in the reference's raw layout (bytes(memoryview(TrainingInfo)),
 * lib.pyx:4047-4063).  Replaces MetagenomicBins / TrainingInfo objects held by GeneFinder
 * (lib.pyx:5187-5196).  Call again to replace the model set. */
int pgpu_set_models(pgpu_ctx *ctx, const void *blobs, int n, size_t stride);
int pgpu_num_models(const pgpu_ctx *ctx);

/* CUDA-event stopwatch on the library's own stream (the stream every kernel of this context is launched
 * on): start records an event, stop records a second one, waits for it and returns the elapsed device
 * time.  Used by bench.py so that the timed region is measured on the launching stream. */
int pgpu_timer_start(pgpu_ctx *ctx);
int pgpu_timer_stop(pgpu_ctx *ctx, double *ms);

/* Page-locked host memory for input buffers (cudaHostAlloc): a sequence buffer that lives in it is copied to the device
 * by asynchronous DMA at link speed, and -- with two lanes -- the copy of one half of a batch overlaps the kernels of the
 * other; a buffer in ordinary (pageable) memory is staged by the driver at roughly a third of that.  Used by the FASTA
 * reader of the Python mirror (pyrodigal_b200/fasta.py; the reference's reader is tests/fasta.py:61-86 / cli.py:283-284).
 * Returns NULL when the allocation fails (the caller falls back to ordinary memory). */
void *pgpu_host_alloc(size_t bytes);
void pgpu_host_free(void *p);

/* Upper bound for the workspace the library may allocate on the device (bytes; 0 = default). */
int pgpu_set_workspace_limit(pgpu_ctx *ctx, size_t bytes);

/* -------------------------------------------------------------------------------------------- */
/* the hot path                                                                                 */
/* -------------------------------------------------------------------------------------------- */

/* find_genes over a batch of contigs: `seq` holds the concatenated ASCII nucleotides, contig k
 * is seq[offsets[k] .. offsets[k+1]).  Replaces GeneFinder.find_genes (lib.pyx:5400-5469) mapped
 * over contigs (cli.py:286-300).  Host buffers; H2D / D2H copies happen inside the call. */
int pgpu_find_genes_batch(pgpu_ctx *ctx, const uint8_t *seq, const int64_t *offsets, int n_contigs,
                          const pgpu_opts *opts, pgpu_result **out);

/* Two-step form for device-resident inputs: upload once, run many times (bench "value" leg). */
typedef struct pgpu_batch pgpu_batch;
int pgpu_batch_upload(pgpu_ctx *ctx, const uint8_t *seq, const int64_t *offsets, int n_contigs,
                      pgpu_batch **out);
/* The same for input that ALREADY sits in device memory of ctx's device (e.g. a sequence shard received from
 * another GPU over NCCL, pyrodigal_b200/distributed.py): `d_seq` is a device pointer, borrowed -- the caller keeps
 * it alive until pgpu_batch_free and frees it itself.  `offsets` is a host array as above. */
int pgpu_batch_wrap_device(pgpu_ctx *ctx, const uint8_t *d_seq, const int64_t *offsets, int n_contigs,
                           pgpu_batch **out);
int pgpu_batch_run(pgpu_ctx *ctx, pgpu_batch *batch, const pgpu_opts *opts, pgpu_result **out);
void pgpu_batch_free(pgpu_batch *batch);

/* result accessors */
int pgpu_result_num_contigs(const pgpu_result *res);
int pgpu_result_summaries(const pgpu_result *res, pgpu_contig_summary *dst /* [n_contigs] */);
/* genes of one contig -> dst[n_genes] */
int pgpu_result_genes(const pgpu_result *res, int contig, pgpu_gene *dst);
/* all genes of all contigs, contig-major -> dst[sum n_genes] */
int pgpu_result_all_genes(const pgpu_result *res, pgpu_gene *dst);
/* start/stop node records of every gene, contig-major: dst[2 * sum n_genes] (start, stop) */
int pgpu_result_gene_nodes(const pgpu_result *res, pgpu_node *dst);
/* zero-copy access: the result is stored as one segment per sub-batch (normally one), each holding the
 * genes of a contiguous contig range and their (start, stop) node records in page-locked host memory owned
 * by the result.  pgpu_result_segment returns the number of genes of segment k (or <0) and the pointers. */
int pgpu_result_num_segments(const pgpu_result *res);
long long pgpu_result_segment(const pgpu_result *res, int k, long long *first_gene, const pgpu_gene **genes,
                              const pgpu_node **gene_nodes);
/* final node array of one contig (requires opts.want_nodes) -> dst[n_nodes] */
int pgpu_result_nodes(const pgpu_result *res, int contig, pgpu_node *dst);
/* the same nodes in the reference's own `struct _node` layout (src/Prodigal/node.h:41-76, 128 bytes as packed by
 * Pyrodigal), so that the Cython side fills `Nodes.nodes` with one memcpy -> dst[n_nodes * PGPU_NODE_STRUCT_SIZE] */
#define PGPU_NODE_STRUCT_SIZE 128
int pgpu_result_nodes_struct(const pgpu_result *res, int contig, void *dst);
int pgpu_result_stats(const pgpu_result *res, pgpu_stats *dst);
void pgpu_result_free(pgpu_result *res);

/* -------------------------------------------------------------------------------------------- */
/* training (the caller side of the hot path in single mode)                */
/* -------------------------------------------------------------------------------------------- */

/* Keyword arguments of GeneFinder.train (lib.pyx:5471-5478). */
typedef struct {
    int32_t translation_table; /* default 11; must be one of lib.pyx:172                        */
    int32_t force_nonsd;       /* 1 = skip the Shine-Dalgarno usage heuristic (node.c:686-693)  */
    double start_weight;       /* default 4.35                                                  */
    int32_t reserved[4];
} pgpu_train_opts;

/* GeneFinder.train (lib.pyx:5471-5575, _train 5236-5279): extract the nodes of `seq` (several training
 * sequences are joined by the caller with TTAATTAATTAA linkers, lib.pyx:5534-5541), GC frame bias, training DP,
 * dicodon statistics, SD / non-SD start training.  Writes the resulting `struct _training`
 * (PGPU_TRAINING_SIZE bytes, the reference's raw layout) to out_training.  opts supplies closed / mask /
 * min_mask / min_gene / min_edge_gene / max_overlap; opts->meta must be 0 (RuntimeError in the reference).
 * Sequences shorter than 20000 bases are rejected with PGPU_EINVAL (ValueError, lib.pyx:5547-5550).
 * Does not touch the loaded model set.  stats may be NULL. */
int pgpu_train(pgpu_ctx *ctx, const uint8_t *seq, int64_t slen, const pgpu_opts *opts, const pgpu_train_opts *topts,
               void *out_training, pgpu_stats *stats);

/* -------------------------------------------------------------------------------------------- */
/* operator-level twins (what the reference's own backend tests drive)                          */
/* -------------------------------------------------------------------------------------------- */

/* Nodes.extract + Nodes.sort (lib.pyx:2518-2560): returns the number of nodes (>= 0) or an error;
 * arrays may be NULL to only count.  `masks` = n_masks x [begin,end) or NULL. */
int pgpu_extract_nodes(pgpu_ctx *ctx, const uint8_t *seq, int slen, int translation_table,
                       const pgpu_opts *opts, int cap, int32_t *ndx, int32_t *stop_val,
                       int8_t *strand, uint8_t *type, uint8_t *edge);

/* Sequence.max_gc_frame_plot (lib.pyx:724-768, 1001-1026): for every base the codon position (0..2) with the highest
 * GC content in the 120-base window around its triplet, -1 for the one or two bases after the last full triplet.
 * out[slen].  (The reference ignores its window_size argument and always uses 120 bases; so does this.) */
int pgpu_max_gc_frame_plot(pgpu_ctx *ctx, const uint8_t *seq, int slen, int8_t *out);

/* Sequence.shine_dalgarno (lib.pyx:1028-1072): bin of the best Shine-Dalgarno motif in the window that starts at
 * strand coordinate `pos` upstream of the start codon at `start`, for the rbs weights of loaded model `model`;
 * exact != 0: AGGAGG sub-motifs without mismatch, else with exactly one.  PGPU_EINVAL (ValueError) on a bad strand or
 * negative coordinates. */
int pgpu_shine_dalgarno(pgpu_ctx *ctx, const uint8_t *seq, int slen, int pos, int start, int model, int strand,
                        int exact, int32_t *out);

/* Nodes.reset_scores + Nodes.score (lib.pyx:2563-2595) for one loaded model, followed by
 * _record_overlapping_starts(flag=1).  `first_pass`: 1 = nodes freshly extracted.
 * dst[cap] receives the scored node array. returns number of nodes. */
int pgpu_score_nodes(pgpu_ctx *ctx, const uint8_t *seq, int slen, int model, const pgpu_opts *opts,
                     int is_meta, int first_pass, int cap, pgpu_node *dst);

/* ConnectionScorer.index + score_connections (lib.pyx:1315-1357; kernel = _connection.h:94-408 +
 * impl/generic.h:29-36): caller supplies the node arrays (SoA), gets score / traceb / ov_mark.
 * gc_score = n x 3 doubles (only read when final == 0), star_ptr = n x 3. */
int pgpu_score_connections(pgpu_ctx *ctx, int n, const int32_t *ndx, const int32_t *stop_val,
                           const int8_t *strand, const uint8_t *type, const double *cscore,
                           const double *sscore, const double *rscore, const double *uscore,
                           const double *gc_score, const int32_t *star_ptr, int model, int final,
                           double *out_score, int32_t *out_traceb, int8_t *out_ov_mark,
                           int64_t *out_pairs, double *out_ms);

/* The skip filter as an operator (ConnectionScorer.compute_skippable, lib.pyx:1321-1334;
 * plug-in ABI skippable_t, lib.pxd:120): skip[j] for j in [min, i). */
int pgpu_compute_skippable(pgpu_ctx *ctx, int n, const int8_t *strand, const uint8_t *type,
                           const int32_t *ndx, int min, int i, uint8_t *skip);
/* The same filter with exactly the signature of the plug-in ABI (`skippable_t`, lib.pxd:120: node strands (1 / 255),
 * types, frames = ndx % 3, window start, target, output), so that it can be stored in
 * BaseConnectionScorer.skippable (lib.pyx:1149-1162) next to the SIMD back-ends.  It has no context argument: it
 * runs on a process-wide context (device $PGPU_DEVICE, default 0).  On failure skip[min .. i) is cleared, which is
 * always a valid answer. */
void pgpu_skippable(const uint8_t *strands, const uint8_t *types, const uint8_t *frames, const int min, const int i,
                    uint8_t *skip);

#ifdef __cplusplus
}
#endif
#endif
