// train_host.hpp -- host half of the training path: the raw `struct _training` layout and the conversion of the
// count tables gathered on the GPU into log-odds weights.
//
// These tables are tiny (3, 28, 128, 4096 and 65536 entries) and need log() of the process' libm to stay
// bit-identical with the reference, so they are finished on the host between kernel launches;
// everything that scales with the sequence or the node list runs on the device (train_kernels.cu).
// Reference: TrainingInfo._calc_dicodon_gene (lib.pyx:4345-4358), _train_starts_sd (4411-4599),
// _train_starts_nonsd (4619-4826), determine_sd_usage / build_coverage_map (vendor/Prodigal/node.c:686-693,
// 1307-1357).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/pyrodigal_b200.h"
#include "train_device.cuh"  // SdParams / MotParams / counter layout shared with the kernels

// raw training struct (vendor/Prodigal/training.h:29-51) -- layout only
struct RawTraining {
    double gc;
    int32_t trans_table, pad0;
    double st_wt;
    double bias[3];
    double type_wt[3];
    int32_t uses_sd, pad1;
    double rbs_wt[28];
    double ups_comp[32][4];
    double mot_wt[4][4][4096];
    double no_mot;
    double gene_dc[4096];
};
static_assert(sizeof(RawTraining) == PGPU_TRAINING_SIZE, "training struct layout");

namespace pgpu {
namespace train_host {

inline double clamp_log(double v, double lim) {
    if (v > lim) return lim;
    if (v < -lim) return -lim;
    return v;
}

// lib.pyx:4345-4358: log-odds of a 6-mer inside training genes against the whole sequence
inline void finish_dicodon(const uint32_t *bg_counts, long long bg_total, const uint32_t *gene_counts, long long gene_total,
                           double *gene_dc) {
    for (int i = 0; i < 4096; i++) {
        const double bg = (double)(int)bg_counts[i] / (double)(int)bg_total;
        const double prob = (double)(int)gene_counts[i] / (double)(int)gene_total;
        double v;
        if (prob == 0 && bg != 0) v = -5.0;
        else if (bg == 0) v = 0.0;
        else v = log(prob / bg);
        if (v > 5.0) v = 5.0;
        else if (v < -5.0) v = -5.0;
        gene_dc[i] = v;
    }
}

// lib.pyx:4423-4433: relative frequency of the three start codons among all starts
inline void type_background(const uint32_t *type_counts, double tbg[3]) {
    for (int i = 0; i < 3; i++) tbg[i] = (double)type_counts[i];
    const double sum = tbg[0] + tbg[1] + tbg[2];
    for (int i = 0; i < 3; i++) tbg[i] /= sum;
}

// lib.pyx:4540-4557 == 4771-4788; returns the number of accepted genes
inline double update_type_weights(const uint32_t *treal_counts, const double tbg[3], RawTraining &T) {
    double treal[3] = {(double)treal_counts[0], (double)treal_counts[1], (double)treal_counts[2]};
    const double sum = treal[0] + treal[1] + treal[2];
    for (int j = 0; j < 3; j++) {
        if (sum == 0.0) { T.type_wt[j] = 0.0; continue; }
        treal[j] /= sum;
        T.type_wt[j] = clamp_log(tbg[j] != 0 ? log(treal[j] / tbg[j]) : -4.0, 4.0);
    }
    return sum;
}

// one SD iteration: lib.pyx:4452-4557.  rbg / rreal = bin counts of the background and of the accepted starts.
inline void sd_update(const uint32_t *rbg_counts, const uint32_t *rreal_counts, const uint32_t *treal_counts,
                      const double tbg[3], int nn, RawTraining &T, double &sthresh) {
    double rbg[28], rreal[28], sum = 0.0;
    for (int j = 0; j < 28; j++) { rbg[j] = (double)rbg_counts[j]; sum += rbg[j]; }
    for (int j = 0; j < 28; j++) rbg[j] /= sum;
    sum = 0.0;
    for (int j = 0; j < 28; j++) { rreal[j] = (double)rreal_counts[j]; sum += rreal[j]; }
    for (int j = 0; j < 28; j++) {
        if (sum == 0.0) { T.rbs_wt[j] = 0.0; continue; }
        rreal[j] /= sum;
        T.rbs_wt[j] = clamp_log(rbg[j] != 0 ? log(rreal[j] / rbg[j]) : -4.0, 4.0);
    }
    sum = update_type_weights(treal_counts, tbg, T);
    if (sum * 2000.0 <= nn) sthresh /= 2.0;
}

// lib.pyx:4561-4599 == 4793-4826: base counts per upstream position -> clamped log-odds against the GC content
inline void upstream_to_log(const uint32_t *ups_counts, RawTraining &T) {
    for (int i = 0; i < 32; i++) {
        double c[4], sum = 0.0;
        for (int j = 0; j < 4; j++) { c[j] = (double)ups_counts[4 * i + j]; sum += c[j]; }
        for (int j = 0; j < 4; j++) {
            if (sum == 0.0) { T.ups_comp[i][j] = 0.0; continue; }
            const bool at = j == 0 || j == 3;
            double den;
            if (T.gc <= 0.1) den = at ? 0.90 : 0.10;
            else if (T.gc >= 0.9) den = at ? 0.10 : 0.90;
            else den = at ? 1.0 - T.gc : T.gc;
            double v = c[j] / sum;
            v = log(v * 2.0 / den);
            T.ups_comp[i][j] = clamp_log(v, 4.0);
        }
    }
}

// vendor/Prodigal/node.c:686-693
inline void determine_sd_usage(RawTraining &T) {
    const double *w = T.rbs_wt;
    T.uses_sd = 1;
    if (w[0] >= 0.0) T.uses_sd = 0;
    if (w[16] < 1.0 && w[13] < 1.0 && w[15] < 1.0 && (w[0] >= -0.5 || (w[22] < 2.0 && w[24] < 2.0 && w[27] < 2.0)))
        T.uses_sd = 0;
}

// vendor/Prodigal/node.c:1307-1357: motifs that are frequent among the accepted starts, or built from such words
inline void coverage_map(const double *real, int *good, double ngenes) {
    auto R = [&](int l, int s, int m) -> double { return real[(l * 4 + s) * 4096 + m]; };
    auto G = [&](int l, int s, int m) -> int & { return good[(l * 4 + s) * 4096 + m]; };
    memset(good, 0, sizeof(int) * 4 * 4 * 4096);
    for (int s = 0; s < 4; s++)
        for (int m = 0; m < 64; m++)
            if (R(0, s, m) / ngenes >= 0.2)
                for (int k = 0; k < 4; k++) G(0, k, m) = 1;
    for (int s = 0; s < 4; s++)
        for (int m = 0; m < 256; m++)
            if (G(0, s, (m & 252) >> 2) && G(0, s, m & 63)) G(1, s, m) = 1;
    for (int s = 0; s < 4; s++)
        for (int m = 0; m < 1024; m++) {
            if (!G(0, s, (m & 1008) >> 4) || !G(0, s, (m & 252) >> 2) || !G(0, s, m & 63)) continue;
            G(2, s, m) = 1;
            int v = m;  // variants of the middle base count as "mismatch" motifs
            for (int a = 0; a <= 16; a += 16) {
                v ^= a;
                for (int b = 0; b <= 32; b += 32) {
                    v ^= b;
                    if (G(2, s, v) == 0) G(2, s, v) = 2;
                }
            }
        }
    for (int s = 0; s < 4; s++)
        for (int m = 0; m < 4096; m++) {
            const int a = G(2, s, (m & 4092) >> 2), b = G(2, s, m & 1023);
            if (a == 0 || b == 0) continue;
            G(3, s, m) = (a == 1 && b == 1) ? 1 : 2;
        }
}

// scratch of the non-SD iterations (3 x 64K doubles + 64K ints)
struct MotifScratch {
    std::vector<double> bg, real;
    std::vector<int> good;
    MotifScratch() : bg(4 * 4 * 4096), real(4 * 4 * 4096), good(4 * 4 * 4096, 0) {}
};

// one non-SD iteration: lib.pyx:4666-4790.  bg_cells / real_cells are the raw counts from the device; in stage 0
// only spacer class 0 was counted and stands for all four classes.
inline void nonsd_update(const uint32_t *bg_cells, const uint32_t *real_cells, double zbg_count, double zreal_count,
                         double ngenes, const uint32_t *treal_counts, const double tbg[3], int stage, int nn,
                         RawTraining &T, double &sthresh, MotifScratch &S) {
    const int N = 4 * 4 * 4096;
    double *mbg = S.bg.data(), *mreal = S.real.data();
    int *mgood = S.good.data();
    for (int x = 0; x < N; x++) {
        const int src = stage == 0 ? (x & ~(3 * 4096)) : x;  // cell (l, k, m) <- (l, 0, m)
        mbg[x] = (double)bg_cells[src];
        mreal[x] = (double)real_cells[src];
    }
    double zbg = zbg_count, zreal = zreal_count, sum = 0.0;
    for (int x = 0; x < N; x++) sum += mbg[x];
    sum += zbg;
    for (int x = 0; x < N; x++) mbg[x] /= sum;
    zbg /= sum;

    if (stage < 2) coverage_map(mreal, mgood, ngenes);
    sum = 0.0;
    for (int x = 0; x < N; x++) sum += mreal[x];
    sum += zreal;
    double *wt = &T.mot_wt[0][0][0];
    if (sum == 0.0) {
        for (int x = 0; x < N; x++) wt[x] = 0.0;
        T.no_mot = 0.0;
    } else {
        for (int x = 0; x < N; x++) {
            if (mgood[x] == 0) {
                zreal += mreal[x];
                zbg += mreal[x];
                mreal[x] = 0.0;
                mbg[x] = 0.0;
            }
            mreal[x] /= sum;
            wt[x] = clamp_log(mbg[x] != 0 ? log(mreal[x] / mbg[x]) : -4.0, 4.0);
        }
    }
    zreal /= sum;
    T.no_mot = clamp_log(zbg != 0 ? log(zreal / zbg) : -4.0, 4.0);
    sum = update_type_weights(treal_counts, tbg, T);
    if (sum * 2000.0 <= nn) sthresh /= 2.0;
}

// The training schedule (GeneFinder._train, lib.pyx:5236-5279) over a backend that performs the counting passes.
// api.cu supplies the CUDA backend (kernel launches + small D2H copies); tests/emu/train_emu.cu supplies a host
// loop over the same per-item functions so that this schedule can be checked without a GPU.  A backend provides
//   int first_gene_set(const RawTraining&, double bias[3], uint32_t dicodon[2*4096], long long *gene_codons)
//       GC frame plot, GC frame bias, training DP, 6-mer counts (background | genes of the training path)
//   int score_starts(const RawTraining&, uint32_t *cnt)      coding score + SD bins of every start; C_TBG counts
//   int sd_iteration(const train::SdParams&, uint32_t *cnt)
//   int motif_iteration(const train::MotParams&, const RawTraining&, uint32_t *cells /* bg | real */, uint32_t *cnt)
// each returning 0 or a PGPU_E* code.
template <class Backend>
int run_training(Backend &be, RawTraining &T, int nn, int slen, int force_nonsd) {
    using namespace pgpu::train;
    double bias[3] = {0, 0, 0};
    std::vector<uint32_t> dicodon(2 * 4096);
    long long gene_codons = 0;
    int rc = be.first_gene_set(T, bias, dicodon.data(), &gene_codons);                        // lib.pyx:5258-5266
    if (rc) return rc;
    if (nn > 0) for (int k = 0; k < 3; k++) T.bias[k] = bias[k];   // record_gc_bias returns early without nodes
    finish_dicodon(dicodon.data(), 2LL * (slen > 5 ? slen - 5 : 0), dicodon.data() + 4096, gene_codons, T.gene_dc);

    std::vector<uint32_t> cnt(C_TOTAL);
    rc = be.score_starts(T, cnt.data());                                                       // lib.pyx:5269-5271
    if (rc) return rc;
    double tbg[3];
    type_background(cnt.data() + C_TBG, tbg);

    double sthresh = 35.0;                                                                     // lib.pyx:4391-4599
    for (int it = 0; it < 10; it++) {
        SdParams P;
        for (int j = 0; j < 28; j++) P.rbs_wt[j] = T.rbs_wt[j];
        for (int j = 0; j < 3; j++) P.type_wt[j] = T.type_wt[j];
        P.wt = T.st_wt; P.sthresh = sthresh; P.last = it == 9;
        rc = be.sd_iteration(P, cnt.data());
        if (rc) return rc;
        sd_update(cnt.data() + C_RBG, cnt.data() + C_RREAL, cnt.data() + C_TREAL, tbg, nn, T, sthresh);
    }
    upstream_to_log(cnt.data() + C_UPS, T);
    if (force_nonsd) T.uses_sd = 0;
    else determine_sd_usage(T);
    if (T.uses_sd) return 0;

    MotifScratch scratch;                                                                      // lib.pyx:4601-4826
    std::vector<uint32_t> cells(2 * (size_t)kMotCells);
    memset(T.ups_comp, 0, sizeof(T.ups_comp));
    memset(T.type_wt, 0, sizeof(T.type_wt));
    sthresh = 35.0;
    for (int it = 0; it < 20; it++) {
        MotParams P;
        for (int j = 0; j < 3; j++) P.type_wt[j] = T.type_wt[j];
        P.wt = T.st_wt; P.sthresh = sthresh; P.no_mot = T.no_mot;
        P.stage = it < 4 ? 0 : (it < 12 ? 1 : 2);
        P.last = it == 19;
        rc = be.motif_iteration(P, T, cells.data(), cnt.data());
        if (rc) return rc;
        nonsd_update(cells.data(), cells.data() + kMotCells, (double)cnt[C_ZBG], (double)cnt[C_ZREAL],
                     (double)cnt[C_NGENES], cnt.data() + C_TREAL, tbg, P.stage, nn, T, sthresh, scratch);
    }
    upstream_to_log(cnt.data() + C_UPS, T);
    return 0;
}

}  // namespace train_host
}  // namespace pgpu
