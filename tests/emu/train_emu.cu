// train_emu.cu -- TEST INFRASTRUCTURE ONLY (never shipped, never on the product path).
//
// Runs the training schedule of the product (train_host::run_training) with a backend that executes the
// per-item functions of pyrodigal_b200/csrc/train_device.cuh in plain host loops, so that the logic of the
// training kernels and of the host-side table math can be checked against the oracle in a container without a
// GPU.  The steps that are *other* CUDA kernels of the product (overlapping starts, the training DP, coding
// score, SD bins -- all covered by their own GPU parity tests) are delegated to the oracle through callbacks.
// Built on demand by tests/test_train_emulation.py with `nvcc -x cu` (host code only is executed).
#include <vector>

#include "../../pyrodigal_b200/csrc/train_host.hpp"

using namespace pgpu;
using namespace pgpu::train;

extern "C" {
// fills traceb / ov_mark / star_ptr (3 per node) and the arg-max node of the training DP (final == 0)
typedef void (*emu_dp_cb)(const double *gc_score, const double *bias, int32_t *traceb, int8_t *ov_mark, int32_t *star_ptr,
                          int32_t *ipath);
// fills cscore and rbs (2 per node) for the model `training`
typedef void (*emu_score_cb)(const void *training, double *cscore, uint8_t *rbs);
}

namespace {

struct Cpu {
    const uint8_t *d; int slen; NodeArrays N;
    emu_dp_cb dp_cb; emu_score_cb score_cb;
    std::vector<uint32_t> gcbits;
    std::vector<int8_t> gp, gc_bias, ov_mark;
    std::vector<double> gc_score, term, cscore;
    std::vector<uint8_t> rbs;
    std::vector<uint64_t> upc, umot;
    std::vector<MotifOut> mot;
    std::vector<int32_t> traceb, star_ptr;
    std::vector<int> stops;
    int8_t *gp_out = nullptr; double *gc_score_out = nullptr; int32_t *n_intervals_out = nullptr;

    static int mer_base(const uint8_t *d, int slen, int x, bool rev) {   // score_kernels.cu:mer_base
        if (!rev) return d[x] & 3;
        const int b = d[slen - 1 - x];
        return b == 6 ? 2 : (b ^ 3);
    }
    void prepare() {
        const int nn = N.nn;
        gcbits.assign(slen / 32 + 2, 0);
        for (int i = 0; i < slen; i++)
            if (d[i] != 0 && d[i] != 3) gcbits[i >> 5] |= 1u << (i & 31);          // seq_kernels.cu:k_encode
        gp.assign(slen + 1, -1); gc_bias.assign(nn + 1, 0); ov_mark.assign(nn + 1, 0);
        gc_score.assign(3 * (size_t)nn + 3, 0.0); term.assign(nn + 1, 0.0); cscore.assign(nn + 1, 0.0);
        rbs.assign(2 * (size_t)nn + 2, 0); upc.assign(nn + 1, 0); umot.assign(nn + 1, 0); mot.resize(nn + 1);
        traceb.assign(nn + 1, -1); star_ptr.assign(3 * (size_t)nn + 3, -1);
        for (int i = 0; i < nn; i++) {
            const int c = N.cls[i];
            if (cls_is_stop(c)) { stops.push_back(i); continue; }
            const bool rev = c & CLS_REV;
            const int start = rev ? slen - 1 - N.ndx[i] : N.ndx[i];
            uint64_t pc = 0, U = 0;                                                 // score_kernels.cu:k_node_prep
            int cnt = 0;
            for (int q = 1; q < 3 && q <= start; q++, cnt++) pc |= (uint64_t)mer_base(d, slen, start - q, rev) << (2 * cnt);
            for (int q = 15; q < 45 && q <= start; q++, cnt++) pc |= (uint64_t)mer_base(d, slen, start - q, rev) << (2 * cnt);
            for (int q = 0; q < 18; q++) {
                const int x = start - 21 + q;
                if (x >= 0 && x < slen) U |= (uint64_t)mer_base(d, slen, x, rev) << (2 * q);
            }
            upc[i] = pc; umot[i] = U;
        }
    }
    int first_gene_set(const RawTraining &T, double *bias, uint32_t *dicodon, long long *gene_codons) {
        for (int t = 0; t < slen / 3; t++) gc_frame_triplet(gcbits.data(), slen, t, gp.data());          // k_gc_frame
        for (int z : stops) gc_bias_orf(z, N, gp.data(), gc_score.data(), gc_bias.data(), term.data()); // k_gc_bias
        double b[3] = {0, 0, 0};                                                                         // k_bias_sum
        for (int i = 0; i < N.nn; i++)
            if (!cls_is_stop(N.cls[i])) b[gc_bias[i]] += term[i];
        const double tot = b[0] + b[1] + b[2];
        for (int k = 0; k < 3; k++) bias[k] = b[k] * (3.0 / tot);
        if (gp_out) memcpy(gp_out, gp.data(), slen);
        if (gc_score_out) memcpy(gc_score_out, gc_score.data(), sizeof(double) * 3 * N.nn);
        int32_t ipath = -1;
        dp_cb(gc_score.data(), bias, traceb.data(), ov_mark.data(), star_ptr.data(), &ipath);
        std::vector<int4> iv(N.nn / 2 + 2);
        const int n = training_path(ipath, N, traceb.data(), ov_mark.data(), star_ptr.data(), iv.data(), (int)iv.size());
        if (n_intervals_out) *n_intervals_out = n;
        memset(dicodon, 0, sizeof(uint32_t) * 2 * 4096);
        for (int i = 0; i < slen - 5; i++) { dicodon[mer6(d, slen, i, false)]++; dicodon[mer6(d, slen, i, true)]++; }  // k_dicodon_bg
        long long total = 0;
        for (int g = 0; g < n; g++)                                                                       // k_dicodon_genes
            for (int i = iv[g].x; i < iv[g].y - 5; i += 3) { dicodon[4096 + mer6(d, slen, i, iv[g].z < 0)]++; total++; }
        *gene_codons = total;
        (void)T;
        return 0;
    }
    int score_starts(const RawTraining &T, uint32_t *cnt) {
        score_cb(&T, cscore.data(), rbs.data());
        memset(cnt, 0, sizeof(uint32_t) * C_TOTAL);
        for (int i = 0; i < N.nn; i++)
            if (!cls_is_stop(N.cls[i])) cnt[C_TBG + (N.cls[i] & CLS_TYPE)]++;                             // k_type_background
        return 0;
    }
    int sd_iteration(const SdParams &P, uint32_t *cnt) {
        memset(cnt, 0, sizeof(uint32_t) * C_TOTAL);
        for (int z : stops) sd_orf(z, N, cscore.data(), rbs.data(), upc.data(), P, cnt);                  // k_sd_iteration
        return 0;
    }
    int motif_iteration(const MotParams &P, const RawTraining &T, uint32_t *cells, uint32_t *cnt) {
        memset(cnt, 0, sizeof(uint32_t) * C_TOTAL);
        memset(cells, 0, sizeof(uint32_t) * 2 * kMotCells);
        const double *w = &T.mot_wt[0][0][0];
        for (int i = 0; i < N.nn; i++) motif_background(i, N, umot.data(), w, P, mot.data(), cells, cnt);   // k_motif_background
        for (int z : stops) motif_orf(z, N, cscore.data(), umot.data(), upc.data(), mot.data(), P, cells + kMotCells, cnt);
        return 0;
    }
};

}  // namespace

extern "C" int emu_train(const uint8_t *digits, int slen, int gc_count, int nn, const int32_t *ndx, const int32_t *sv,
                         const uint8_t *cls, int tt, double st_wt, int force_nonsd, emu_dp_cb dp_cb, emu_score_cb score_cb,
                         void *out_training, int8_t *gp_out, double *gc_score_out, int32_t *n_intervals_out) {
    Cpu be;
    be.d = digits; be.slen = slen; be.N = NodeArrays{ndx, sv, cls, nn, slen};
    be.dp_cb = dp_cb; be.score_cb = score_cb;
    be.gp_out = gp_out; be.gc_score_out = gc_score_out; be.n_intervals_out = n_intervals_out;
    be.prepare();
    std::vector<RawTraining> T(1);
    memset(&T[0], 0, sizeof(RawTraining));
    T[0].gc = slen > 0 ? (double)gc_count / (double)slen : 0.0;
    T[0].trans_table = tt; T[0].st_wt = st_wt; T[0].uses_sd = 1;
    const int rc = train_host::run_training(be, T[0], nn, slen, force_nonsd);
    memcpy(out_training, &T[0], sizeof(RawTraining));
    return rc;
}
