// codon_masks.hpp -- host-side codon sets per translation table (shared by api.cu and the host emulation tests)
#pragma once
#include <stdint.h>

#include <initializer_list>

namespace pgpu {

// stop / start codon sets per translation table as bit masks over the 6-bit codon code
// (code = b0 | b1<<2 | b2<<4, A0 G1 C2 T3).  Rules: src/pyrodigal/_sequence.h:45-73, 117-157.
inline void codon_masks(int tt, uint64_t *stopmask, uint64_t *startmask) {
    auto in = [&](std::initializer_list<int> l) { for (int v : l) if (v == tt) return true; return false; };
    const bool taa = in({1, 2, 3, 4, 5, 9, 10, 11, 12, 13, 15, 16, 21, 22, 23, 24, 25, 26, 32});
    const bool tag = in({1, 2, 3, 4, 5, 9, 10, 11, 12, 13, 14, 21, 23, 24, 25, 26, 33});
    const bool tga = in({1, 6, 11, 12, 15, 16, 22, 23, 26, 29, 30, 32});
    uint64_t sm = 0, am = 0;
    enum { A = 0, G = 1, C = 2, T = 3 };
    for (int c = 0; c < 64; c++) {
        const int x0 = c & 3, x1 = (c >> 2) & 3, x2 = (c >> 4) & 3;
        bool stop = false;
        if (x0 == T && x1 == A && x2 == G) stop = tag;
        else if (x0 == T && x1 == G && x2 == A) stop = tga;
        else if (x0 == T && x1 == A && x2 == A) stop = taa;
        else if (tt == 2) stop = x0 == A && x1 == G && (x2 == A || x2 == G);
        else if (tt == 22) stop = x0 == T && x1 == C && x2 == A;
        else if (tt == 23) stop = x0 == T && x1 == T && x2 == A;
        bool start = false;
        if (x1 == T && x2 == G) {
            if (x0 == A) start = true;
            else if (in({6, 10, 14, 15, 16, 2})) start = false;
            else if (x0 == G) start = !in({1, 3, 12, 2});
            else if (x0 == T) start = !(tt < 4 || tt == 9 || (tt >= 21 && tt < 25));
        }
        if (stop) sm |= 1ull << c;
        if (start) am |= 1ull << c;
    }
    *stopmask = sm;
    *startmask = am;
}

}  // namespace pgpu
