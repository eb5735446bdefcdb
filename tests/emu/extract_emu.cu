// extract_emu.cu -- TEST INFRASTRUCTURE ONLY (never shipped, never on the product path).
//
// Runs the per-item functions of pyrodigal_b200/csrc/extract_device.cuh (the bit-parallel node extraction the
// kernels k_codon_bits / k_extract_b are made of) in plain host loops, so that their logic can be checked against
// the oracle's node list in a container without a GPU.  Built on demand by tests/test_extract_emulation.py with
// `nvcc -x cu` (only host code is executed).
#include <vector>

#include "../../pyrodigal_b200/csrc/codon_masks.hpp"
#include "../../pyrodigal_b200/csrc/extract_device.cuh"

using namespace pgpu;

namespace {
struct Collect {
    int slen; bool rev; const uint8_t *cod; int closed;
    std::vector<int32_t> *ndx, *sv, *type, *strand, *edge;
    void put(int pos, int type_, int sv_, int edge_) {
        const int p = rev ? slen - 1 - pos : pos;   // seq_kernels.cu: emit()
        ndx->push_back(p);
        sv->push_back(rev ? slen - 1 - sv_ : sv_);
        type->push_back(type_);
        strand->push_back(rev ? -1 : 1);
        edge->push_back(edge_);
    }
    void start(int i, int last, int edge_) {
        int t = 0;
        if (!edge_) {
            int c = rev ? rev_code(cod[slen - 3 - i] & 63) : (cod[i] & 63);
            const int b0 = c & 3;
            t = b0 == 0 ? 0 : (b0 == 1 ? 1 : 2);
        }
        put(i, t, last, edge_);
    }
    void stop(int last, int sv_, int edge_) { put(last, 3, sv_, edge_); }
};
}  // namespace

extern "C" int emu_extract(const uint8_t *digits, int slen, int tt, int closed, int min_gene, int min_edge_gene,
                           int32_t *out, int cap) {
    // codon codes exactly as k_encode writes them (zero padded past the end)
    std::vector<uint8_t> cod(slen + 16, 0);
    auto dg = [&](int p) { return p < slen ? (int)digits[p] : 0; };
    for (int p = 0; p < slen; p++) {
        const int b0 = dg(p), b1 = dg(p + 1), b2 = dg(p + 2);
        cod[p] = (uint8_t)((b0 & 3) | ((b1 & 3) << 2) | ((b2 & 3) << 4) | (((b0 | b1 | b2) & 4) << 4));
    }
    uint64_t stopmask, startmask;
    codon_masks(tt, &stopmask, &startmask);
    std::vector<int32_t> ndx, sv, type, strand, edge;
    if (slen >= 3) {
        for (int rev = 0; rev < 2; rev++)
            for (int f = 0; f < 3; f++) {
                ExtractFrame F;
                extract_frame_geometry(slen, f, &F.i_top0, &F.n_codons);
                F.n_words = (F.n_codons + 31) / 32;
                F.f = f; F.closed = closed;
                F.d_real = extract_min_codons(min_gene); F.d_virt = extract_min_codons(min_edge_gene);
                F.min_edge_gene = min_edge_gene;
                std::vector<uint32_t> S(F.n_words + 1, 0), C(F.n_words + 1, 0);
                for (int u = 0; u < F.n_codons; u++) {
                    const int fl = codon_flags(cod.data(), slen, rev != 0, F.i_top0 - 3 * u, stopmask, startmask);
                    if (fl & 1) S[u >> 5] |= 1u << (u & 31);
                    if (fl & 2) C[u >> 5] |= 1u << (u & 31);
                }
                F.S = S.data(); F.C = C.data();
                Collect col{slen, rev != 0, cod.data(), closed, &ndx, &sv, &type, &strand, &edge};
                for (int w = 0; w < F.n_words; w++) extract_word(F, w, col);
            }
    }
    const int n = (int)ndx.size();
    if (n > cap) return -n;
    for (int k = 0; k < n; k++) {
        out[5 * k] = ndx[k]; out[5 * k + 1] = sv[k]; out[5 * k + 2] = type[k]; out[5 * k + 3] = strand[k]; out[5 * k + 4] = edge[k];
    }
    return n;
}

// codon_lut_build against codon_flags for every byte value, both strands: number of mismatches
extern "C" int emu_lut_mismatches(int tt) {
    uint64_t stopmask, startmask;
    codon_masks(tt, &stopmask, &startmask);
    uint8_t lut[128];
    codon_lut_build(stopmask, startmask, lut);
    int bad = 0;
    for (int b = 0; b < 128; b++)
        for (int rev = 0; rev < 2; rev++) {
            const uint8_t cod[4] = {(uint8_t)b, 0, 0, 0};
            const int want = codon_flags(cod, 3, rev != 0, 0, stopmask, startmask);
            const int got = (lut[b] >> (rev ? 2 : 0)) & 3;
            bad += want != got;
        }
    return bad;
}
