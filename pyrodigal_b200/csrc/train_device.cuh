// train_device.cuh -- per-item logic of the training kernels (GeneFinder.train, lib.pyx:5236-5279).
//
// Every function here handles ONE work item (a codon triplet, an ORF, a start node, ...) and is
// __host__ __device__: train_kernels.cu maps items to CUDA threads; tests/emu/train_emu.cu runs the very same
// functions in a host loop so that the logic can be checked against the oracle in a container without a GPU
// (test infrastructure only -- the product never executes them on the host).
//
// Reference semantics: Sequence._max_gc_frame_plot (lib.pyx:724-768), record_gc_bias
// (vendor/Prodigal/node.c:263-317), TrainingInfo._calc_dicodon_gene (lib.pyx:4284-4358),
// _count_upstream_composition (4360-4389), _train_starts_sd (4391-4599), _update_motif_counts (4226-4282),
// Node._find_best_upstream_motif (1557-1616), _train_starts_nonsd (4601-4826).
//
// Why this parallelises exactly: every statistic the training loops accumulate is an integer-valued count
// (the reference adds 1.0 to doubles), so the order of accumulation does not matter and atomics on integers
// reproduce it bit for bit; the per-frame "best start since the last STOP" state of the reference's sweeps is
// reset at every STOP node, so an ORF (one STOP node and the starts it closes) is an independent item.  The
// log-odds conversion of the small count tables stays on the host (train_host.hpp) with the reference's libm.
#pragma once
#include "common.cuh"

#if defined(__CUDA_ARCH__)
#define PGPU_COUNT(p) atomicAdd((p), 1u)
#else
#define PGPU_COUNT(p) (++*(p))
#endif
#define PGPU_HD __host__ __device__ inline

namespace pgpu {
namespace train {

// the sorted node list of the training sequence (one extraction)
struct NodeArrays {
    const int32_t *ndx, *sv;
    const uint8_t *cls;
    int nn, slen;
};

// layout of the small counter block shared by the SD / non-SD iterations
enum : int {
    C_RBG = 0,        // [28] motif bin of every non-edge start (background)
    C_RREAL = 28,     // [28] motif bin of the best start of every accepted ORF
    C_TREAL = 56,     // [3]  start codon type of the best start of every accepted ORF
    C_TBG = 59,       // [3]  start codon type of every start (background)
    C_NGENES = 62,    // accepted ORFs
    C_ZBG = 63,       // non-SD: background starts without a motif
    C_ZREAL = 64,     // non-SD: accepted starts without a motif
    C_UPS = 65,       // [32][4] upstream base composition (last iteration only)
    C_TOTAL = 65 + 128
};
constexpr int kMotCells = 4 * 4 * 4096;

// vendor/Prodigal/sequence.c:559-564: index of the largest of three counts, the later one on ties
PGPU_HD int frame_of_max(int a, int b, int c) {
    if (a > b) return a > c ? 0 : 2;
    return b > c ? 1 : 2;
}

// ---- GC frame plot (lib.pyx:724-768) --------------------------------------------------------------------
// item = codon triplet t (positions 3t .. 3t+2).  The reference's running sums give, for position x,
// tot[x] = number of GC bases among x-57, x-54, .., x+57 inside the sequence; the triplet gets the frame with
// the largest tot.  The three sums are the three residue classes of one 117-base window of the GC bitmap.
PGPU_HD int gc_bit(const uint32_t *gcbits, int slen, int x) {
    return (x >= 0 && x < slen) ? (int)((gcbits[x >> 5] >> (x & 31)) & 1u) : 0;
}
PGPU_HD void gc_frame_triplet(const uint32_t *gcbits, int slen, int t, int8_t *gp) {
    const int i = 3 * t;
    if (i + 2 >= slen) return;  // the tail without a full triplet keeps -1
    int tot[3] = {0, 0, 0};
    int f = 0;
    for (int b = 0; b < 117; b++) {
        tot[f] += gc_bit(gcbits, slen, i - 57 + b);
        f = f == 2 ? 0 : f + 1;
    }
    const int8_t w = (int8_t)frame_of_max(tot[0], tot[1], tot[2]);
    gp[i] = w; gp[i + 1] = w; gp[i + 2] = w;
}

// ---- GC frame bias of every start (node.c:263-317) ------------------------------------------------------
// item = STOP node z.  Walks the starts of its ORF outwards (the reference sweeps the node list once per
// strand and resets its counters at every STOP), counting for every codon which codon position the GC-richest
// frame falls on.  term[i] is the addend of the (order dependent) bias sum, evaluated later in node order.
PGPU_HD void gc_bias_orf(int z, const NodeArrays N, const int8_t *gp, double *gc_score, int8_t *gc_bias, double *term) {
    const int c = N.cls[z];
    const bool rev = (c & CLS_REV) != 0;
    const int f = cls_frame(c);
    int ctr[3] = {0, 0, 0};
    int last = N.ndx[z];
    const int shift = rev ? f : 3 - f;
    ctr[((rev ? 3 - gp[last] : gp[last]) + shift) % 3] = 1;
    const int step = rev ? 1 : -1;
    for (int i = z + step; i >= 0 && i < N.nn; i += step) {
        const int ci = N.cls[i];
        if (((ci & CLS_REV) != 0) != rev || cls_frame(ci) != f) continue;
        if (cls_is_stop(ci)) break;
        const int ni = N.ndx[i];
        if (!rev) for (int j = last - 3; j >= ni; j -= 3) ctr[(gp[j] + shift) % 3]++;
        else for (int j = last + 3; j <= ni; j += 3) ctr[((3 - gp[j]) + shift) % 3]++;
        const int m = frame_of_max(ctr[0], ctr[1], ctr[2]);
        const double span = 1.0 * (rev ? ni - N.sv[i] + 3 : N.sv[i] - ni + 3);
        double gs[3];
        for (int k = 0; k < 3; k++) {
            gs[k] = 3.0 * ctr[k];
            gs[k] /= span;
            gc_score[3 * (int64_t)i + k] = gs[k];
        }
        gc_bias[i] = (int8_t)m;
        const int len = abs(N.sv[i] - ni) + 1;
        term[i] = (gs[m] * len) / 1000.0;
        last = ni;
    }
}

// ---- training path: untangle overlaps, list the genes (lib.pyx:1253-1289, 4317-4343) ----------------------
// single item.  Returns the number of (left, right, strand) intervals written; a gene contributes the 6-mers at
// left, left+3, .. < right-5 in strand coordinates.
PGPU_HD int training_path(int ipath, const NodeArrays N, int32_t *traceb, int8_t *ov_mark, const int32_t *star_ptr,
                          int4 *out, int cap) {
    if (ipath < 0) return 0;
    for (int path = ipath; traceb[path] != -1; path = traceb[path]) {   // triple overlaps
        const int nxt = traceb[path];
        if (cls_kind(N.cls[path]) == K_RE && cls_kind(N.cls[nxt]) == K_FE && ov_mark[path] != -1 &&
            N.ndx[path] > N.ndx[nxt]) {
            const int tmp = star_ptr[3 * (int64_t)path + ov_mark[path]];
            int i = tmp;
            while (N.ndx[i] != N.sv[tmp]) i--;
            traceb[path] = tmp;
            traceb[tmp] = i;
            ov_mark[i] = -1;
            traceb[i] = nxt;
        }
    }
    for (int path = ipath; traceb[path] != -1; path = traceb[path]) {   // simple overlaps
        const int nxt = traceb[path];
        const int kp = cls_kind(N.cls[path]), kn = cls_kind(N.cls[nxt]);
        if (kp == K_RS && kn == K_FE) {
            int i = path;
            while (N.ndx[i] != N.sv[path]) i--;
            traceb[path] = i;
            traceb[i] = nxt;
        }
        if (kp == K_FE && kn == K_FE) {
            traceb[path] = star_ptr[3 * (int64_t)nxt + N.ndx[path] % 3];
            traceb[traceb[path]] = nxt;
        }
        if (kp == K_RE && kn == K_RE) {
            traceb[path] = star_ptr[3 * (int64_t)path + N.ndx[nxt] % 3];
            traceb[traceb[path]] = nxt;
        }
    }
    int n = 0, in_gene = 0, left = -1, right = -1;
    for (int p = ipath; p != -1; p = traceb[p]) {
        const int c = N.cls[p];
        const bool stop = cls_is_stop(c);
        int strand = 0;
        if (!(c & CLS_REV)) {
            if (stop) { in_gene = 1; right = N.ndx[p] + 2; }
            else if (in_gene == 1) { left = N.ndx[p]; strand = 1; in_gene = 0; }
        } else {
            if (!stop) { in_gene = -1; left = N.slen - N.ndx[p] - 1; }
            else if (in_gene == -1) { right = N.slen - N.ndx[p] + 1; strand = -1; in_gene = 0; }
        }
        if (strand != 0) {
            if (n < cap) out[n] = make_int4(left, right, strand, 0);
            n++;
        }
    }
    return n;
}

// ---- dicodon statistics (lib.pyx:4284-4358) -----------------------------------------------------------------
// 6-mer index at strand coordinate i (_sequence.h:207-220: first base in the low bits, N indexes as C)
PGPU_HD int mer6(const uint8_t *d, int slen, int i, bool rev) {
    int r = 0;
    if (!rev) {
        for (int j = 0; j < 6; j++) r |= (d[i + j] & 3) << (2 * j);
    } else {
        for (int j = 0; j < 6; j++) {
            const int b = d[slen - 1 - i - j];
            r |= (b == 6 ? 2 : (b ^ 3)) << (2 * j);
        }
    }
    return r;
}

// ---- upstream base composition of an accepted start (lib.pyx:4360-4389) --------------------------------------
// pc = the packed strand-oriented bases at start-1, start-2, start-15 .. start-44 (k_node_prep); the slots that
// fall off the sequence are exactly the trailing ones
PGPU_HD void count_upstream(uint64_t pc, int start, uint32_t *ups) {
    const int ncomp = (start < 2 ? start : 2) + (start - 14 < 0 ? 0 : (start - 14 > 30 ? 30 : start - 14));
    for (int k = 0; k < ncomp; k++, pc >>= 2) PGPU_COUNT(&ups[4 * k + (int)(pc & 3)]);
}

// ---- Shine-Dalgarno start training, one iteration (lib.pyx:4435-4538) ------------------------------------------
struct SdParams {
    double rbs_wt[28];
    double type_wt[3];
    double wt, sthresh;
    int last;   // 1: final iteration, also tally the upstream composition
};

// which of the two SD bins (exact, one mismatch) a start is credited with (lib.pyx:4442-4449)
PGPU_HD int preferred_rbs(int a, int b, const double *w) {
    if (w[a] > w[b] + 1.0 || b == 0) return a;
    if (w[a] < w[b] - 1.0 || a == 0) return b;
    return a > b ? a : b;
}

// item = STOP node z: background bins of its non-edge starts, and the best start of the ORF if it clears the
// threshold.  The reference scans towards the STOP and keeps the later start on ties (>=); walking outwards from
// the STOP with a strict > selects the same node.
PGPU_HD void sd_orf(int z, const NodeArrays N, const double *cscore, const uint8_t *rbs, const uint64_t *upc,
                    const SdParams &P, uint32_t *cnt) {
    const int c = N.cls[z];
    const bool rev = (c & CLS_REV) != 0;
    const int f = cls_frame(c), step = rev ? 1 : -1;
    double best = 0.0;
    int bi = -1, brb = 0;
    for (int i = z + step; i >= 0 && i < N.nn; i += step) {
        const int ci = N.cls[i];
        if (((ci & CLS_REV) != 0) != rev || cls_frame(ci) != f) continue;
        if (cls_is_stop(ci)) break;
        if (ci & CLS_EDGE) continue;
        const int rb = preferred_rbs(rbs[2 * (int64_t)i], rbs[2 * (int64_t)i + 1], P.rbs_wt);
        PGPU_COUNT(&cnt[C_RBG + rb]);
        const double v = cscore[i] + P.wt * P.rbs_wt[rb] + P.wt * P.type_wt[ci & CLS_TYPE];
        if (v > best) { best = v; bi = i; brb = rb; }
    }
    if (bi >= 0 && best >= P.sthresh) {
        PGPU_COUNT(&cnt[C_RREAL + brb]);
        PGPU_COUNT(&cnt[C_TREAL + (N.cls[bi] & CLS_TYPE)]);
        if (P.last) count_upstream(upc[bi], rev ? N.slen - 1 - N.ndx[bi] : N.ndx[bi], cnt + C_UPS);
    }
}

// ---- upstream motif training (non-SD organisms) ------------------------------------------------------------------
struct MotParams {
    double type_wt[3];
    double wt, sthresh, no_mot;
    int stage;  // 0: count every window, 1: the best motif and its sub-words, 2: the best motif only
    int last;
};

// spacer class of the p-th window (j = start-18-l+p) of a motif length (lib.pyx:1586-1593)
PGPU_HD int spacer_class(int p) { return p <= 2 ? 3 : (p <= 4 ? 2 : (p >= 11 ? 1 : 0)); }

// best upstream motif of one start under the current weights (lib.pyx:1557-1616).  U = packed strand-oriented
// bases start-21 .. start-4 (k_node_prep), so the window starting at j is bit field 2*(j-start+21).
PGPU_HD MotifOut best_motif(uint64_t U, int start, const double *mot_wt, double no_mot, int stage) {
    int max_spacer = 0, max_spacendx = 0, max_len = 0, max_ndx = 0;
    double max_sc = -100.0;
    for (int l = 3; l >= 0; l--) {
        const uint32_t lmask = (1u << (2 * (l + 3))) - 1u;
        for (int p = 0; p < 13; p++) {
            const int j = start - 18 - l + p;
            if (j < 0) continue;
            const int sp = spacer_class(p);
            const int index = (int)((U >> (2 * (3 - l + p))) & lmask);
            const double sc = mot_wt[(l * 4 + sp) * 4096 + index];
            if (sc > max_sc) { max_sc = sc; max_spacendx = sp; max_spacer = start - j - l - 3; max_ndx = index; max_len = l + 3; }
        }
    }
    MotifOut m;
    m.pad[0] = m.pad[1] = m.pad[2] = 0;
    if (stage == 2 && (max_sc == -4.0 || max_sc < no_mot + 0.69)) {
        m.ndx = 0; m.len = 0; m.spacendx = 0; m.spacer = 0; m.score = no_mot;
    } else {
        m.ndx = (uint16_t)max_ndx; m.len = (uint8_t)max_len; m.spacendx = (uint8_t)max_spacendx;
        m.spacer = (uint8_t)max_spacer; m.score = max_sc;
    }
    return m;
}

// lib.pyx:4226-4282 for a non-edge start.  Stage 0 credits a window to all four spacer classes; only class 0 is
// counted here and the host replicates it.
PGPU_HD void count_motif(const MotifOut &m, uint64_t U, int start, int stage, uint32_t *cells, uint32_t *zero) {
    if (m.len == 0) { PGPU_COUNT(zero); return; }
    if (stage == 0) {
        for (int l = 3; l >= 0; l--) {
            const uint32_t lmask = (1u << (2 * (l + 3))) - 1u;
            for (int p = 0; p < 13; p++) {
                if (start - 18 - l + p < 0) continue;
                PGPU_COUNT(&cells[(l * 4) * 4096 + (int)((U >> (2 * (3 - l + p))) & lmask)]);
            }
        }
    } else if (stage == 1) {
        PGPU_COUNT(&cells[((m.len - 3) * 4 + m.spacendx) * 4096 + m.ndx]);
        const int j0 = start - m.spacer - m.len;
        for (int l = 0; l < m.len - 3; l++) {
            const uint32_t lmask = (1u << (2 * (l + 3))) - 1u;
            for (int j = j0; j <= start - m.spacer - l - 3; j++) {
                if (j < 0) continue;
                int sp;
                if (j <= start - 16 - l) sp = 3;
                else if (j <= start - 14 - l) sp = 2;
                else if (j >= start - 7 - l) sp = 1;
                else sp = 0;
                PGPU_COUNT(&cells[(l * 4 + sp) * 4096 + (int)((U >> (2 * (j - start + 21))) & lmask)]);
            }
        }
    } else {
        PGPU_COUNT(&cells[((m.len - 3) * 4 + m.spacendx) * 4096 + m.ndx]);
    }
}

// item = non-edge start node i: re-evaluate its motif and add it to the background (lib.pyx:4658-4664)
PGPU_HD void motif_background(int i, const NodeArrays N, const uint64_t *umot, const double *mot_wt, const MotParams &P,
                              MotifOut *mot, uint32_t *bg_cells, uint32_t *cnt) {
    const int c = N.cls[i];
    if (cls_is_stop(c) || (c & CLS_EDGE)) return;
    const int start = (c & CLS_REV) ? N.slen - 1 - N.ndx[i] : N.ndx[i];
    const MotifOut m = best_motif(umot[i], start, mot_wt, P.no_mot, P.stage);
    mot[i] = m;
    count_motif(m, umot[i], start, P.stage, bg_cells, cnt + C_ZBG);
}

// item = STOP node z: the best start of its ORF, credited to the "real" tables when it clears the threshold
// (lib.pyx:4687-4731)
PGPU_HD void motif_orf(int z, const NodeArrays N, const double *cscore, const uint64_t *umot, const uint64_t *upc,
                       const MotifOut *mot, const MotParams &P, uint32_t *real_cells, uint32_t *cnt) {
    const int c = N.cls[z];
    const bool rev = (c & CLS_REV) != 0;
    const int f = cls_frame(c), step = rev ? 1 : -1;
    double best = 0.0;
    int bi = -1;
    for (int i = z + step; i >= 0 && i < N.nn; i += step) {
        const int ci = N.cls[i];
        if (((ci & CLS_REV) != 0) != rev || cls_frame(ci) != f) continue;
        if (cls_is_stop(ci)) break;
        if (ci & CLS_EDGE) continue;
        const double v = cscore[i] + P.wt * mot[i].score + P.wt * P.type_wt[ci & CLS_TYPE];
        if (v > best) { best = v; bi = i; }
    }
    if (bi >= 0 && best >= P.sthresh) {
        const int start = rev ? N.slen - 1 - N.ndx[bi] : N.ndx[bi];
        PGPU_COUNT(&cnt[C_NGENES]);
        PGPU_COUNT(&cnt[C_TREAL + (N.cls[bi] & CLS_TYPE)]);
        count_motif(mot[bi], umot[bi], start, P.stage, real_cells, cnt + C_ZREAL);
        if (P.last) count_upstream(upc[bi], start, cnt + C_UPS);
    }
}

}  // namespace train
}  // namespace pgpu
