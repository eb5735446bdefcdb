// sd_device.cuh -- Shine-Dalgarno motif of ONE window, as the operator Sequence.shine_dalgarno exposes it
// (src/pyrodigal/lib.pyx:1028-1072 -> _shine_dalgarno_exact 791-890 / _shine_dalgarno_mm 892-979).
//
// The scoring kernel k_start_score does not call this: it searches the 15 windows of a start through per-model
// tables built from the same rules (api.cu: build_sd_masks / prepare_model).  This function evaluates the rules
// directly for an arbitrary (pos, start), which is what the operator-level API and its tests need.
//   window  : up to 6 bases from `pos`, cut off 4 bases before `start`; base i matches A when i % 3 == 0, G otherwise
//   exact   : every sub-window of length >= 3 without a mismatch scores sum(match) - 2 (A = 2, G = 3)
//   mismatch: sub-windows of length >= 5 with exactly one mismatch that is not among the two bases at either end
//   the motif bin follows from (score, spacer class); the bin with the larger rbs weight wins, the larger bin on a tie
#pragma once
#include "common.cuh"

namespace pgpu {

// strand-oriented digit at strand coordinate p (A0 G1 C2 T3; unknown bases never match)
__host__ __device__ inline int sd_base(const uint8_t *d, int slen, int p, int strand) {
    if (strand == 1) return d[p];
    const int b = d[slen - 1 - p];
    return b > 3 ? b : 3 - b;  // complement: A<->T (0<->3), G<->C (1<->2)
}

__host__ __device__ inline int sd_window(const uint8_t *d, int slen, int pos, int start, const double *rbs_wt, int strand,
                                         bool exact) {
    int match[6];
    int limit = start - 4 - pos;
    if (limit > 6) limit = 6;
    for (int i = 0; i < 6; i++) match[i] = exact ? -10 : (i % 3 == 0 ? -3 : -2);
    for (int i = 0; i < limit; i++) {
        const int p = pos + i;
        if (p < 0 || p >= slen) continue;
        const int b = sd_base(d, slen, p, strand);
        if (i % 3 == 0) { if (b == 0) match[i] = 2; }
        else { if (b == 1) match[i] = 3; }
    }
    int best = 0, cur = 0;  // `cur` keeps its value across sub-windows, as in the reference
    for (int len = limit; len > (exact ? 2 : 4); len--) {
        for (int j = 0; j <= limit - len; j++) {
            int ctr = -2, mism = 0;
            for (int k = j; k < j + len; k++) {
                ctr += match[k];
                if (!exact && match[k] < 0) {
                    mism++;
                    if (k <= j + 1 || k >= j + len - 2) ctr -= 10;
                }
            }
            if (ctr < 6 || (!exact && mism != 1)) continue;
            const int rdis = start - (pos + j + len);
            int flag;
            if (rdis < 5) flag = exact ? (len < 5 ? 2 : 1) : 1;
            else if (rdis < 11) flag = 0;
            else if (rdis < 13) flag = exact ? (len < 5 ? 1 : 2) : 2;
            else if (rdis < 16) flag = 3;
            else continue;
            if (exact) {
                // score -> bins by spacer class {5-10, 3-4 / 11-12 (by length), ..., 13-15}
                switch (ctr) {
                case 6: { const int v[4] = {13, 6, 1, 2}; cur = v[flag]; break; }
                case 8: { const int v[4] = {15, 12, 11, 3}; cur = v[flag]; break; }
                case 9: { const int v[4] = {16, 12, 11, 3}; cur = v[flag]; break; }
                case 11: { const int v[4] = {22, 21, 20, 10}; cur = v[flag]; break; }
                case 12: { const int v[4] = {24, 23, 20, 10}; cur = v[flag]; break; }
                case 14: { const int v[4] = {27, 26, 25, 10}; cur = v[flag]; break; }
                default: cur = 0; break;
                }
            } else {
                if (ctr == 6) { const int v[4] = {9, 5, 4, 2}; cur = v[flag]; }
                else if (ctr == 7) { const int v[4] = {14, 8, 7, 2}; cur = v[flag]; }
                else if (ctr == 9) { const int v[4] = {19, 18, 17, 3}; cur = v[flag]; }
            }
            if (rbs_wt[cur] < rbs_wt[best]) continue;
            if (rbs_wt[cur] == rbs_wt[best] && cur < best) continue;
            best = cur;
        }
    }
    return best;
}

}  // namespace pgpu
