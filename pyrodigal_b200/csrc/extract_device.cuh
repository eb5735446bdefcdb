// extract_device.cuh -- bit-parallel node extraction (add_nodes): per-item logic shared by the CUDA kernels
// (seq_kernels.cu: k_codon_bits, k_extract_b) and by the host emulation used to check it without a GPU
// (tests/emu/extract_emu.cu).
//
// Reference semantics: Nodes._extract (src/pyrodigal/lib.pyx:1905-2117) = add_nodes (Prodigal node.c:28-172): per
// strand and frame the codons are scanned from the 3' end of the strand towards its 5' end; the state of that
// scan is (last stop seen, was it a real stop, has a start been emitted since) and a codon yields
//   * a start node when it is a start codon at least min_gene (min_edge_gene before the first real stop) away
//     from the last stop, or -- for the last codon of the frame only -- an edge start;
//   * at a stop codon (and at the end of the frame) the STOP node of the ORF that just ended, if it had a start.
//
// Formulation here.  Scan index u = 0, 1, ... numbers the codons of one (strand, frame) in scan order; two
// bitmaps over u hold "is a stop codon" (S) and "is a start codon" (C), 32 codons per word.  Everything a codon
// needs from the scan state is a function of the nearest stop before it (u_g):
//   start at u  <=>  C[u] and u - u_g >= D          (D = ceil((min_gene - 3) / 3): a mask of the bits >= u_g + D)
//   STOP node of u_g, emitted at the next stop u' <=> some start in (u_g, u')
// so one THREAD handles a whole word of 32 codons with a handful of bit operations per stop in the word, after
// looking back for the nearest stop before the word (on average one word; the words in between are only needed
// for "has any start qualified since").  Ownership: a start belongs to the word of its codon, a STOP node to the
// word of the stop that closes its ORF (the trailing one to the last word), so every node is produced exactly once
// and the same function serves the marking pass and the filling pass.  N-run masks (GeneFinder(mask=True)) are not
// handled here: batches with masks use the warp-cooperative kernel k_extract_w.
#pragma once
#include <stdint.h>

#include "common.cuh"

namespace pgpu {

struct ExtractFrame {      // one (strand, frame) of an extraction in scan-index space
    const uint32_t *S;     // stop codons, bit (u & 31) of word (u >> 5); bits past n_codons are zero
    const uint32_t *C;     // start codons
    int n_codons;          // codons of the frame (u < n_codons); 0: nothing to do
    int n_words;           // (n_codons + 31) / 32
    int i_top0;            // strand coordinate of the codon u = 0 (the largest i <= slen - 3 of the frame)
    int f;                 // frame (i % 3)
    int closed;
    int d_real;            // u - u_g must be >= d_real after a real stop        (min_gene)
    int d_virt;            // ... >= d_virt before the first real stop           (min_edge_gene)
    int min_edge_gene;
};

// strand geometry of a frame: lib.pyx:1933-1939 (the initial `last` is i_top0 when the ends are open and
// i_top0 + 3 when they are closed, i.e. the virtual stop sits at u = 0 or u = -1)
__host__ __device__ inline void extract_frame_geometry(int slen, int f, int *i_top0, int *n_codons) {
    int t = slen - 3;
    t -= ((t % 3) - f + 3) % 3;
    *i_top0 = t;
    *n_codons = t >= 0 ? t / 3 + 1 : 0;
}
__host__ __device__ inline int extract_min_codons(int min_dist) { return (min_dist - 3 + 2) / 3; }  // min_dist >= 1

// bit 0: stop codon, bit 1: start codon of the codon at strand coordinate i (0 <= i <= slen - 3)
__host__ __device__ inline int codon_flags(const uint8_t *cod, int slen, bool rev, int i, uint64_t stopmask,
                                           uint64_t startmask) {
    int c = rev ? cod[slen - 3 - i] : cod[i];
    if (c & 64) return 0;  // holds an unknown base: neither (lib.pyx:1943-1990 test has_n first)
    c &= 63;
    if (rev) c = rev_code(c);
    return (int)((stopmask >> c) & 1) | ((int)((startmask >> c) & 1) << 1);
}

// The same flags for every value of a codon byte, both strands at once: bit 0 / 1 = stop / start on the forward
// strand, bit 2 / 3 = stop / start of the reverse-strand codon that the byte encodes (k_codon_bits with
// PGPU_CODON_LUT=1 replaces the mask arithmetic by one table load per codon).
__host__ __device__ inline void codon_lut_build(uint64_t stopmask, uint64_t startmask, uint8_t *lut /*[128]*/) {
    for (int b = 0; b < 128; b++) {
        int v = 0;
        if (!(b & 64)) {
            const int c = b & 63, r = rev_code(c);
            v = (int)((stopmask >> c) & 1) | ((int)((startmask >> c) & 1) << 1) | ((int)((stopmask >> r) & 1) << 2) |
                ((int)((startmask >> r) & 1) << 3);
        }
        lut[b] = (uint8_t)v;
    }
}

__host__ __device__ inline int ex_clz(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __clz((int)x);
#else
    return x ? __builtin_clz(x) : 32;
#endif
}
__host__ __device__ inline int ex_ctz(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __ffs((int)x) - 1;
#else
    return x ? __builtin_ctz(x) : -1;
#endif
}

// Emit: void start(int i, int last, int edge)  -- start node at strand coordinate i, its stop at `last`
//       void stop(int last, int sv, int edge)  -- STOP node at `last`, stop_val sv
template <class Emit>
__host__ __device__ inline void extract_word(const ExtractFrame &F, int w, Emit &emit) {
    const int u0 = 32 * w;
    const uint32_t Sw = F.S[w], Cw = F.C[w];
    // nearest stop before the word, and the highest start codon between it and the word
    int ug = 0, cmax = -1;
    bool real = false;
    for (int ww = w - 1; ww >= 0; ww--) {
        const uint32_t s = F.S[ww];
        uint32_t c = F.C[ww];
        if (s) {
            const int hb = 31 - ex_clz(s);
            ug = 32 * ww + hb;
            real = true;
            c = hb == 31 ? 0u : (c & ~((2u << hb) - 1u));
            if (c && cmax < 0) cmax = 32 * ww + 31 - ex_clz(c);
            break;
        }
        if (c && cmax < 0) cmax = 32 * ww + 31 - ex_clz(c);
    }
    if (!real) ug = F.closed ? -1 : 0;
    // first scan index that can hold a start under the governing stop (ug, real)
    auto first_ok = [&](int g, bool r) { return r ? g + (F.d_real > 1 ? F.d_real : 1) : (g + F.d_virt > 0 ? g + F.d_virt : 0); };
    bool dead = F.closed && !real;  // closed ends: nothing before the first real stop (lib.pyx: `last >= slen`)
    bool saw = !dead && cmax >= first_ok(ug, real);

    uint32_t rem = Sw;
    int lo = 0;
    while (true) {
        const int b = rem ? ex_ctz(rem) : 32;
        if (b > lo && !dead) {
            const uint32_t seg = (b == 32 ? 0xffffffffu : ((1u << b) - 1u)) & ~((1u << lo) - 1u);
            const int thr = first_ok(ug, real) - u0;
            const uint32_t qm = thr <= 0 ? 0xffffffffu : (thr >= 32 ? 0u : ~((1u << thr) - 1u));
            const uint32_t Q = Cw & seg & qm;
            uint32_t E = 0;
            const int ul = F.n_codons - 1 - u0;  // the last codon of the frame (i <= 2): edge start, lib.pyx:1996-2003
            if (!F.closed && ul >= lo && ul < b && !((Q >> ul) & 1u) && (int64_t)3 * (F.n_codons - 1 - ug) > F.min_edge_gene)
                E = 1u << ul;
            uint32_t all = Q | E;
            if (all) saw = true;
            const int last = F.i_top0 - 3 * ug;
            while (all) {
                const int k = ex_ctz(all);
                all &= all - 1;
                emit.start(F.i_top0 - 3 * (u0 + k), last, (int)((E >> k) & 1u));
            }
        }
        if (b == 32) break;
        const int u = u0 + b;
        if (saw) emit.stop(F.i_top0 - 3 * ug, F.i_top0 - 3 * u, real ? 0 : 1);
        ug = u; real = true; dead = false; saw = false;
        lo = b + 1;
        rem &= rem - 1;
        if (lo == 32) break;
    }
    if (w == F.n_words - 1 && saw) emit.stop(F.i_top0 - 3 * ug, F.f - 6, real ? 0 : 1);
}

}  // namespace pgpu
