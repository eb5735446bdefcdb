// kernels.cuh -- launch wrappers shared between the translation units of libpyrodigal_b200.so
#pragma once
#include <functional>
#include "common.cuh"
#include "train_device.cuh"

namespace pgpu {

// seq_kernels.cu
void launch_encode(const DevBatch &B, const int2 *tiles, int n_tiles, cudaStream_t st);
void launch_dicodon_index(const DevBatch &B, const int2 *tiles, int n_tiles, cudaStream_t st);
void launch_find_masks(const DevBatch &B, const int2 *tiles, int n_tiles, int min_mask, int4 *out, int cap,
                       int *count, cudaStream_t st);
void launch_extract_mark(const DevBatch &B, int n_ext, int total_chunks, RunOpts o, cudaStream_t st);
void launch_codon_bits(const DevBatch &B, int n_ext, int total_chunks, cudaStream_t st);
void launch_extract_bits(const DevBatch &B, int n_ext, int total_chunks, RunOpts o, bool fill, cudaStream_t st);
void launch_extract_fill(const DevBatch &B, int n_ext, int total_chunks, RunOpts o, cudaStream_t st);
int scan_num_blocks(int64_t nwords);
void launch_gc_scan(const DevBatch &B, int64_t nwords, int *block_sums, int *total_out, cudaStream_t st);
void launch_word_scan(const DevBatch &B, int64_t nwords, int *block_sums, int *total_out, cudaStream_t st);

// score_kernels.cu
void launch_block_owner_off(const int64_t *off, int n, int shift, int32_t *tab, cudaStream_t st);
void launch_block_owner_chains(const ChainInfo *chains, int n, int32_t *tab, cudaStream_t st);
void launch_block_owner_exts(const ExtractInfo *exts, int n, int32_t *tab, cudaStream_t st);
void launch_node_prep(const DevBatch &B, int n_ext, int total_nodes, int seq_parts, cudaStream_t st);
void launch_score_chains(const DevBatch &B, const DevModel *models, int n_chains, int64_t total_chain_nodes,
                         RunOpts o, void *mot_out, int n_ext, int total_nodes, cudaStream_t st);
void launch_coding(const DevBatch &B, const DevModel *models, int n_chains, int64_t total, int n_ext, int total_nodes,
                   cudaStream_t st);
void launch_start_score(const DevBatch &B, const DevModel *models, int n_chains, int64_t total, RunOpts o, void *mot_out,
                        cudaStream_t st);
void launch_start_score_lean(const DevBatch &B, const DevModel *models, int n_chains, int64_t total, RunOpts o, cudaStream_t st);
void launch_overlap(const DevBatch &B, const DevModel *models, int n_chains, int64_t total_chain_nodes, int64_t total_il,
                    int n_ext, RunOpts o, int flag, cudaStream_t st);
void launch_opv(const DevBatch &B, const DevModel *models, int n_chains, int64_t total, cudaStream_t st);
void launch_dp_index(const DevBatch &B, int n_ext, int total_nodes, cudaStream_t st);
void launch_pairs(const DevBatch &B, int n_ext, int total_nodes, unsigned long long *ext_pairs, cudaStream_t st);

// dp_kernels.cu
void launch_dp_ml(const DevBatch &B, const DevModel *models, const int4 *groups, int n_groups, int n_chains, int minb,
                  cudaStream_t st);
void launch_dp_compare(const double *sa, const double *sb, const int32_t *ta, const int32_t *tb, const int8_t *oa,
                       const int8_t *ob, int64_t n, unsigned long long *bad, cudaStream_t st);
void launch_dp(const DevBatch &B, const DevModel *models, const int32_t *order, int n_chains, int final, int algo,
               cudaStream_t st);
void launch_trace(const DevBatch &B, const DevModel *models, int n_contigs, const int32_t *contig_chain_begin,
                  int32_t *tracef, uint8_t *elim, pgpu_gene *genes, pgpu_gene *genes_raw, const int64_t *gene_off,
                  int64_t total_gene_slots, pgpu_contig_summary *summary, int32_t *winner_chain, int meta, int max_overlap,
                  const DevBatch *W, const std::function<void()> &score_winners, const int64_t *slot_off, cudaStream_t st);
void launch_pack_nodes(const DevBatch &B, int n_chains, int64_t total, const void *mot, const int32_t *tracef,
                       const uint8_t *elim, int dp_state, pgpu_node *out, cudaStream_t st);
void launch_shine_dalgarno(const DevBatch &B, const DevModel *models, int model, int pos, int start, int strand, int exact,
                           int32_t *out, cudaStream_t st);
void launch_score_genes(const DevBatch &B, const DevModel *models, int n_contigs, const void *summary, const void *genes,
                        const int64_t *gene_off, int64_t gene_cap, int2 *list, int *count, RunOpts o, void *mot_out,
                        cudaStream_t st);
void launch_pack_nodes_genes(const DevBatch &B, const int2 *list, const int *count, int64_t gene_cap, const pgpu_gene *genes,
                             const int64_t *gene_off, const void *mot, pgpu_node *out, cudaStream_t st);
void launch_pack_gene_nodes(int n_contigs, const pgpu_contig_summary *summary, const pgpu_gene *genes,
                            const int64_t *gene_off, const int64_t *gene_out_off, const int64_t *node_out_off,
                            const pgpu_node *nodes, pgpu_node *out, pgpu_gene *genes_out, cudaStream_t st);
void launch_skippable(int n, const int8_t *strand, const uint8_t *type, const int32_t *ndx, int mn, int i,
                      uint8_t *skip, cudaStream_t st);
void launch_skippable_plugin(const uint8_t *strand, const uint8_t *type, const uint8_t *frame, int cnt, uint8_t *skip,
                             cudaStream_t st);
void launch_build_final_chains(const DevBatch &B, int n_contigs, const int32_t *winner_chain, const int64_t *fin_coff,
                               ChainInfo *fin_chains, cudaStream_t st);

// train_kernels.cu -- the training pass works on ONE extraction (the training sequence) and ONE chain at
// chain-node offset 0
struct TrainView {
    train::NodeArrays N;
    int node_off;            // offset of the extraction's nodes (clist is indexed relative to it)
    const int32_t *cb;       // its four class-segment offsets (DevBatch::cbase)
    int8_t *gp;              // GC frame plot, one byte per base (-1 in the tail)
    double *gc_score;        // [3 * nn]
    int8_t *gc_bias;         // [nn]
    double *term;            // [nn] addends of the bias sum
    const double *cscore;    // [nn]
    const uint8_t *rbs;      // [2 * nn]
    const uint64_t *upc, *umot;
    MotifOut *mot;           // [nn] best upstream motif under the current weights (non-SD training)
};
void launch_gc_frame(const uint32_t *gcbits, int slen, int8_t *gp, cudaStream_t st);
void launch_gc_bias(const DevBatch &B, const TrainView &V, double *bias_out, double *gcb, cudaStream_t st);
void launch_training_path(const DevBatch &B, const TrainView &V, int4 *intervals, int cap, int *n_out, cudaStream_t st);
void launch_dicodon(const uint8_t *digits, int slen, const int4 *intervals, const int *n_intervals, int cap,
                    uint32_t *bg_counts, uint32_t *gene_counts, unsigned long long *gene_total, cudaStream_t st);
void launch_type_background(const TrainView &V, uint32_t *cnt, cudaStream_t st);
void launch_sd_iteration(const DevBatch &B, const TrainView &V, const train::SdParams &P, uint32_t *cnt, cudaStream_t st);
void launch_motif_iteration(const DevBatch &B, const TrainView &V, const double *mot_wt, const train::MotParams &P,
                            uint32_t *bg_cells, uint32_t *real_cells, uint32_t *cnt, cudaStream_t st);

}  // namespace pgpu
