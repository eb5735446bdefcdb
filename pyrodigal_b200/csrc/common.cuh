// common.cuh -- shared definitions of the pyrodigal_b200 CUDA library (sm_100a only).
//
// Data model (all device arrays are SoA; see DESIGN.md "Data layout in HBM"):
//   contig      one input sequence.  digits[doff .. doff+slen) (1 B/nt, A0 G1 C2 T3 N6, padded with
//               zeros, doff multiple of 128) and cod[] (1 B/nt: 6-bit code of the codon starting at
//               that base + bit 6 "contains N").
//   extraction  (contig, translation table): the sorted node list  ndx[], stop_val[], cls[], gc_cont[],
//               plus DP index tables (window start, per-class ranks, class lists).
//   chain       (extraction, model): one DP chain.  Per-chain node scores and DP state live at
//               chain-node offset `coff`.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/pyrodigal_b200.h"

namespace pgpu {

constexpr int kMaxNodeDist = 500;  // vendor/Prodigal/dprog.h:29
constexpr int kMaxOppOvlp = 200;   // vendor/Prodigal/dprog.h:30
constexpr int kOperDist = 60;      // src/Prodigal/node.h:33
constexpr int kDcCols = 64;        // row width of the transposed dicodon table (DevBatch::dcT): a power of two >= the 50 built-in models
constexpr int kExtractChunkCodons = 2048;  // codons of one frame scanned by one warp of k_extract_w (64 ballots)
#define PGPU_EDGE_BONUS 0.74       // node.h:34
#define PGPU_EDGE_UPS (-1.00)      // node.h:35
#define PGPU_META_PEN 7.5          // node.h:36

// node kinds: 2*(reverse strand) + (type == STOP)
enum : int { K_FS = 0, K_FE = 1, K_RS = 2, K_RE = 3 };

// cls byte of a node
constexpr int CLS_TYPE = 0x03;   // 0 ATG 1 GTG 2 TTG 3 STOP
constexpr int CLS_REV = 0x04;    // reverse strand
constexpr int CLS_EDGE = 0x08;   // edge flag as produced by extraction
constexpr int CLS_CONV = 0x10;   // start that Nodes._score converts to an edge node (lib.pyx:2424-2434)
constexpr int CLS_FRAME_SHIFT = 5;  // bits 5-6: ndx % 3

__host__ __device__ inline int cls_kind(int c) { return ((c >> 2) & 1) * 2 + ((c & CLS_TYPE) == 3); }
__host__ __device__ inline int cls_frame(int c) { return (c >> CLS_FRAME_SHIFT) & 3; }
__host__ __device__ inline bool cls_is_stop(int c) { return (c & CLS_TYPE) == 3; }
__host__ __device__ inline bool cls_is_rev(int c) { return (c & CLS_REV) != 0; }

// ---- per-model tables derived on the host from the raw `struct _training` blob ----------------
struct DevModel {
    double st_wt, gc, no_mot;
    double bias[3], type_wt[3];
    double rbs_wt[28];
    double uc[32][4];      // 0.4 * st_wt * ups_comp[k][b]            (lib.pyx:1642,1648)
    double lfac[1001];     // log((1-p^g)/p^g) - lfac_min, g = 0..1000 (lib.pyx:2209-2210, host libm)
    double lfac_span;      // lfac_max - lfac_min                      (lib.pyx:2207)
    double igt[61];        // (2.0 - d/60) * 0.15 * st_wt, d = 0..60   (_connection.h:74)
    double ig_neg;         // -0.15 * st_wt                            (_connection.h:48,72)
    int32_t trans_table, uses_sd;
    uint64_t stopmask, startmask;  // bit c: codon code c is a stop / start in trans_table
    uint8_t sd_best[15][64][2];    // best SD bin for (offset, 6-bit match pattern, exact | mismatch): the pair is one 2-byte load
    const double *gene_dc;         // device pointers into the raw blob
    const double *mot_wt;
    const uint32_t *mot_live;      // bitmap over the 4*4*4096 motif cells: weight != -4.0 (the clamped floor, which is
                                   // what all but a few dozen cells of a trained model hold); nullptr = always load
    uint64_t mot_pf[4];            // per motif length: bit (index & 63) set when any cell [len][*][index] is live; a
                                   // register-only pre-filter in front of mot_live (exactness never depends on it)
    double len_neg[250], len_pos[250];   // 250.0 / L and L / 250.0 for ORF lengths L < 250 (lib.pyx:2436-2440): the two FP64
                                   // divisions of every short-ORF start, tabulated (IEEE division: host == device)
    const uint16_t *mot_hit;       // [4096] by six upstream bases: which motifs (length, offset in the window) may be live
                                   // (api.cu: motif_live_bits); with it the search only visits candidate windows
    int32_t col;                   // column of this model in the transposed dicodon table (sorted by tt, gc)
    int32_t pad;
};

struct ContigInfo {
    int64_t doff;   // offset of the contig in digits[] / cod[] (multiple of 128)
    int64_t aoff;   // offset of the contig in the ASCII input
    int32_t slen;
    int32_t mask_off, n_masks;  // masks of this contig in the mask table (pairs)
    int32_t pad;
};

struct ExtractInfo {
    int32_t contig;
    int32_t tt;
    int64_t doff;
    int32_t slen;
    int32_t slot;        // bitmap slot of this extraction within its contig
    int64_t woff;        // first 32-bit word of this extraction's node bitmaps
    int32_t nwords;
    int32_t node_off;    // first node in the extraction-node arrays
    int32_t nn;
    int32_t mask_off, n_masks;
    int32_t n_lo;        // number of nodes with ndx <= 4          (only these and the n_hi last ones can carry an
    int32_t n_hi;        // number of nodes with ndx >= slen - 5    edge flag; written by k_class_index)
    int32_t chunk_off;   // first extraction chunk of this extraction (k_extract_w: one warp per chunk x strand x frame)
    int32_t n_chunks;    // ceil(codons of the longest frame / kExtractChunkCodons), >= 1
    int32_t lut;         // row of this extraction's translation table in DevBatch::codon_lut (or 0)
    int64_t cb_off;      // first word of this extraction's codon bitmaps: [strand*3+frame][n_chunks * kExtractChunkCodons/32]
    uint64_t stopmask, startmask;
};

struct ChainInfo {
    int32_t ext;
    int32_t model;
    int32_t contig;
    int32_t first_pass;  // 1: first model scored after this extraction
    int32_t node_off;    // extraction-node offset
    int32_t nn;
    int64_t coff;        // chain-node offset, chain-major arrays: node i of this chain at coff + i
    int64_t doff;
    int32_t slen;
    int32_t is_meta;
    // The arrays the DP touches (cs, opv, star_ptr, score, traceb, ov_mark) are INTERLEAVED over the chains that share
    // an extraction: node i of this chain sits at ioff + i * istride, istride = number of chains of the extraction,
    // ioff = first element of the extraction's block + this chain's position among them.  k_dp_ml walks those chains
    // together, one lane per chain, so a warp access touches istride consecutive elements (one or two cache lines
    // instead of one line per lane).  A chain that is alone on its extraction has istride 1 and ioff == coff.
    int64_t ioff;
    int32_t istride;
    int32_t lane;        // position among the chains of the extraction
    int64_t coff_in;     // winner pass of meta mode: chain-node offset of this chain in the MAIN pass (raw coding scores)
};

// element i of an interleaved per-chain-node array (see ChainInfo::ioff)
template <typename T>
struct Strided {
    T *p;
    int64_t s;
    __host__ __device__ __forceinline__ T &operator[](int64_t i) const { return p[i * s]; }
};

struct RunOpts {
    int32_t closed, min_gene, min_edge_gene, max_overlap;
};

// ---- device arrays of one (sub-)batch ----------------------------------------------------------
struct DevBatch {
    // sequences
    const uint8_t *ascii;
    uint8_t *digits;
    uint8_t *cod;
    // dicodon (6-mer) indices, per contig three FRAME PLANES of dic_plane(slen) elements each (at doff): a coding-score walk
    // visits every third base, so in a plane its indices are consecutive 2-byte elements (16-byte loads of eight codons)
    uint16_t *dic_f;      // plane p % 3, element p / 3: index of the forward 6-mer starting at p = cod[p] | cod[p+3] << 6
    uint16_t *dic_r;      // plane p % 3, element dic_plane - 1 - p / 3 (REVERSED, so that walking away from a stop means
                          // decreasing elements on both strands): reverse-strand 6-mer whose first codon has its 5' base at
                          // p (N indexes as C)
    uint32_t *gcbits;     // 1 bit per base: not A / not T (unknown bases count as GC, _sequence.h:35-43)
    int32_t *gcpre;       // exclusive prefix of popcount(gcbits) per 32-base word
    ContigInfo *contigs;
    int32_t *gc_count;   // per contig
    int32_t *unknown;    // per contig
    int32_t *masks;      // pairs [begin,end)
    // extraction
    ExtractInfo *exts;
    uint32_t *bits_fwd, *bits_rev;  // node bitmaps (per extraction, word offset woff)
    uint32_t *cb_stop, *cb_start;   // codon bitmaps in scan order (extract_device.cuh), per extraction at cb_off
    const uint8_t *codon_lut;       // optional [rows][128] codon-byte -> flags tables (codon_lut_build); nullptr = mask arithmetic
    int32_t *wordbase;              // exclusive prefix of node counts per bitmap word
    // extraction nodes
    int32_t *ndx, *stop_val;
    uint8_t *cls;
    float *gc_cont;
    uint32_t *sdbits;     // upstream A/G pattern for the SD motif search
    uint64_t *upc;        // 2-bit strand-oriented bases at start-1, start-2, start-15 .. start-44 (composition)
    uint64_t *umot;       // 2-bit strand-oriented bases at start-21 .. start-4 (upstream motif search)
    int32_t *win_min;     // DP window start (lib.pyx:1224-1233)
    int32_t *crank;       // [4 * node]: number of class-c nodes before node
    int32_t *clist;       // class-sorted node indices (local), segments per class
    int32_t *cbase;       // [4 * ext]: start of each class segment in clist (relative to node_off)
    int32_t *cndx;        // ndx in class order: cndx[p] = ndx[clist[p]]
    int32_t *feq;         // class order, +STOP segment only: position in the merged +STOP / -start stream
    int4 *dpx;            // per node: pre-resolved DP candidates / ranges (see k_dp_index)
    int32_t *ig_node;     // merged stream of +STOP and -start nodes ("intergenic sources") in node order:
    int32_t *ig_ndx;      //   node index (bit 31 set for +STOP) and position
    int4 *dqx;            // per node: candidates / ranges in merged-stream positions (k_dp_index)
    int32_t *ext_chain_off;   // [n_ext + 1] chains that use an extraction ...
    int32_t *ext_chains;      // ... chain indices, grouped by extraction
    const double *dcT;        // dicodon table transposed: dcT[index * kDcCols + DevModel.col]
    int32_t n_models;
    // chains
    ChainInfo *chains;
    double *cscore, *sscore, *rscore, *uscore, *tscore;  // per chain-node
    double *cs;           // per chain-node, interleaved: cscore + sscore as one array (main pass of find_genes; nullptr otherwise)
    double *rupen;        // lean main pass of meta mode (interleaved, sparse): what _intergenic_mod_same subtracts for a start
                          // that touches another node; nullptr = read rscore / uscore
    const double *cscore_in;  // winner pass of meta mode: the raw coding scores of the main pass (at ChainInfo::coff_in)
    double *opv;          // [3 * chain-node] operon values (cs[n3] + igm) for STOP nodes
    double *gcb;          // training DP: bias . gc_score per chain-node (only final == 0)
    int32_t *star_ptr;    // [3 * chain-node]
    uint8_t *rbs;         // [2 * chain-node]
    double *score;
    int32_t *traceb;
    int8_t *ov_mark;
    // class-ordered DP state (fast DP): per chain, indexed like clist
    double *dp_sv;        // score of a finalized +STOP / -start node, or -DBL_MAX when nothing leads into it
    int32_t *dp_tbn;      // ndx of the traceback node of a +STOP node
    double *dp_bx;        // per 16-entry block: max of (sv + ig_neg), ties -> later entry
    int32_t *dp_bj;       // node index of that maximum
    double *dp_svig;      // merged-stream order: score of a finalized +STOP / -start, -DBL_MAX if no traceback
    int32_t *dp_tbig;     // merged-stream order: traceback node of a +STOP
    double *dp_fmv;       // k_dp_ml: suffix maxima of the far window (value), interleaved [entry][lane]
    int32_t *dp_fmj;      // k_dp_ml: ... and their nodes
    // block -> owner tables (optional; nullptr = binary search): the chain that contains chain-node index 128*b and
    // the extraction that contains node index 128*b, so that a thread block does not start with a serial search
    const int32_t *blk_chain;
    const int32_t *blk_ext;
    // grouped mapping of k_coding_orf (optional; nullptr = one warp per ORF): thread range of every extraction
    // (orf_toff[0..n_ext], multiples of 32), lanes per ORF (4 / 8 / 16 / 32), owner of every 256-thread block
    const int64_t *orf_toff;   // indexed by the position r of an extraction in the planned order
    const int32_t *orf_ext;    // extraction at position r: extractions with the same model set are neighbours, so that the
                               // warps resident on an SM read the same columns of the dicodon table (L1 locality)
    const uint8_t *orf_w;
    const int32_t *orf_blk;
    int64_t orf_threads;
    // shared-memory mapping of the raw coding score (k_coding_flat; optional, dcS == nullptr => k_coding_orf).  Plan entries
    // r = (extraction, up to four chains on neighbouring table columns), sorted by class = (table set, lanes per ORF); every
    // class ends with a padding entry (cq_ext = -1) that fills the last CTA span of the class.  The host plans the entries,
    // the device counts the STOP nodes and lays out the ORF slots (k_cq_plan): entry r owns [cq_soff[r], cq_soff[r + 1]).
    const double *dcS;         // table sets: dcS[(s * 4096 + index) * 4 + k] = dicodon weight `index` of table column s + k
    const int32_t *cq_ext;     // extraction of entry r (-1: padding)
    const uint8_t *cq_cls;     // class of entry r
    const int32_t *cq_chain;   // [4 * r + k]: chain of lane k, or -1
    const int64_t *cq_cbase;   // [4 * r + k]: chain-node offset of lane k's chain minus the extraction's node offset
                               //   (cscore[cq_cbase + batch node index]), INT64_MIN = no chain
    const int32_t *cq_hs0;     // first ORF descriptor of entry r's extraction (-1: padding entry)
    const int32_t *cq_colmodel;// model on table column c
    int32_t *cq_soff;          // [n_ent + 1], device
    int32_t *cq_cta;           // [max CTAs + 1], device: plan entry that holds the first slot of every CTA span
    int32_t *cq_ncta;          // device: number of CTA spans in use
    int32_t cq_n_ent, cq_span, cq_span_shift, cq_max_cta;
    // ORF links and descriptors (k_orf_links), in batch node indices / element offsets from dic_f (dic_r lies behind it)
    int4 *link;                // start node: x = next in-frame start further away from the stop (-1: none), y = the element
                               //   where the walk towards that start ends, z = previous start towards the stop (or the STOP
                               //   node), w = distance to the stop position (ORF length - 3)
    int4 *orfd;                // per STOP node, at ((node_off + 1) >> 1) + rank among the STOP nodes of its extraction:
                               //   x = element of the first codon of the walk, y = first start (-1: none), z = element
                               //   where the walk to it ends, w = the STOP node
    // per chain results
    int32_t *chain_ipath;
    double *chain_score;
};

// best upstream motif of a start node (struct _motif, src/Prodigal/node.h:39-46)
struct MotifOut {
    double score;
    uint16_t ndx;
    uint8_t len, spacer, spacendx, pad[3];
};

// ---- 1-D bulk copy global -> shared (TMA) completing on an mbarrier; raw PTX, sm_90+ ---------------
#ifndef PGPU_HOST_EMULATION
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// arm the barrier with the byte count of the copy, then start the copy (one elected lane)
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
// the same in two steps, for several copies that complete on one barrier phase
__device__ __forceinline__ void mbar_expect(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MBAR_DONE;\n"
        "bra MBAR_WAIT;\n"
        "MBAR_DONE:\n"
        "}\n" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
#else   // host emulation (tests/emu): the copy happens at once, the barrier is always complete
__device__ __forceinline__ void mbar_init(uint64_t *, int) {}
__device__ __forceinline__ void mbar_init_fence() {}
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *) { memcpy(dst, src, bytes); }
__device__ __forceinline__ void mbar_expect(uint64_t *, uint32_t) {}
__device__ __forceinline__ void bulk_copy(void *dst, const void *src, uint32_t bytes, uint64_t *) { memcpy(dst, src, bytes); }
__device__ __forceinline__ void mbar_wait(uint64_t *, uint32_t) {}
#endif

// elements of one frame plane of the dicodon index arrays: ceil(slen / 3) plus slack for whole 16-byte chunks, a multiple of 8
__host__ __device__ inline int dic_plane(int slen) { return ((slen + 2) / 3 + 8 + 7) & ~7; }

// ---- device helpers ------------------------------------------------------------------------------

// codon code of the reverse-strand codon whose 5' base is at forward position p, from the code of the
// forward codon starting at p-2: complement every base and reverse the base order.
__host__ __device__ inline int rev_code(int c) {
    int r = ((c >> 4) & 3) | (c & 0xC) | ((c & 3) << 4);
    return r ^ 63;
}

}  // namespace pgpu
