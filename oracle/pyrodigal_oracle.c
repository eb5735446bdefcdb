/* pyrodigal_oracle.c -- TEST INFRASTRUCTURE ONLY.  See pyrodigal_oracle.h.
 *
 * CPU restatement (plain C, IEEE double, no FMA contraction: build with -ffp-contract=off)
 * of the Pyrodigal hot path.  Each function cites the reference lines it follows.
 * Nothing here is on the product path; the CUDA library never links or loads it.
 */
#include "pyrodigal_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define MAX_NODE_DIST 500 /* vendor/Prodigal/dprog.h:29 */
#define MAX_OPP_OVLP 200  /* vendor/Prodigal/dprog.h:30 */
#define OPER_DIST 60      /* src/Prodigal/node.h:33 */
#define EDGE_BONUS 0.74   /* node.h:34 */
#define EDGE_UPS (-1.00)  /* node.h:35 */
#define META_PEN 7.5      /* node.h:36 */

/* digit alphabet of _sequence.h:8-17: A=0 G=1 C=2 T=3 N=6; complement == xor 3 */
enum { dA = 0, dG = 1, dC = 2, dT = 3, dN = 6 };

/* ------------------------------------------------------------------------------------ */
/* sequence helpers                                                                      */
/* ------------------------------------------------------------------------------------ */

/* lib.pyx:664-697 */
int orc_encode(const uint8_t *ascii, int n, uint8_t *digits, int *gc_count) {
    int unknown = 0, gc = 0;
    for (int i = 0; i < n; i++) {
        switch (ascii[i]) {
        case 'A': case 'a': digits[i] = dA; break;
        case 'T': case 't': digits[i] = dT; break;
        case 'G': case 'g': digits[i] = dG; gc++; break;
        case 'C': case 'c': digits[i] = dC; gc++; break;
        default: digits[i] = dN; unknown++; break;
        }
    }
    if (gc_count) *gc_count = gc;
    return unknown;
}

/* lib.pyx:699-713: runs of N of length >= mask_size; a trailing run is always kept */
int orc_find_masks(const uint8_t *digits, int slen, int mask_size, int32_t *out, int cap) {
    int n = 0, begin = -1;
    for (int i = 0; i < slen; i++) {
        if (digits[i] == dN) {
            if (begin == -1) begin = i;
        } else if (begin != -1) {
            if (i >= mask_size + begin) {
                if (n < cap) { out[2 * n] = begin; out[2 * n + 1] = i; }
                n++;
            }
            begin = -1;
        }
    }
    if (begin != -1) {
        if (n < cap) { out[2 * n] = begin; out[2 * n + 1] = slen; }
        n++;
    }
    return n;
}

/* k-th base (k=0..2) of the codon whose 5' base sits at strand coordinate i.
 * Reverse strand is read in place: _sequence.h:45-57 */
static inline int base_at(const uint8_t *d, int slen, int i, int strand) {
    return strand == 1 ? d[i] : (d[slen - 1 - i] ^ 3);
}

/* _sequence.h:117-157 */
static int is_stop(const uint8_t *d, int slen, int i, int tt, int strand) {
    /* which translation tables read TAA / TAG / TGA as stop */
    static const uint8_t taa[34] = {0,1,1,1,1,1,0,0,0,1,1,1,1,1,0,1,1,0,0,0,0,1,1,1,1,1,1,0,0,0,0,0,1,0};
    static const uint8_t tag[34] = {0,1,1,1,1,1,0,0,0,1,1,1,1,1,1,0,0,0,0,0,0,1,0,1,1,1,1,0,0,0,0,0,0,1};
    static const uint8_t tga[34] = {0,1,0,0,0,0,1,0,0,0,0,1,1,0,0,1,1,0,0,0,0,0,1,1,0,0,1,0,0,1,1,0,1,0};
    int x0 = base_at(d, slen, i, strand), x1 = base_at(d, slen, i + 1, strand), x2 = base_at(d, slen, i + 2, strand);
    if (x0 == dT && x1 == dA && x2 == dG) return tag[tt];
    if (x0 == dT && x1 == dG && x2 == dA) return tga[tt];
    if (x0 == dT && x1 == dA && x2 == dA) return taa[tt];
    if (tt == 2) return x0 == dA && x1 == dG && (x2 == dA || x2 == dG);
    if (tt == 22) return x0 == dT && x1 == dC && x2 == dA;
    if (tt == 23) return x0 == dT && x1 == dT && x2 == dA;
    return 0;
}

/* _sequence.h:45-73 */
static int is_start(const uint8_t *d, int slen, int i, int tt, int strand) {
    int x0 = base_at(d, slen, i, strand), x1 = base_at(d, slen, i + 1, strand), x2 = base_at(d, slen, i + 2, strand);
    if (x1 != dT || x2 != dG) return 0;
    if (x0 == dA) return 1;
    if (tt == 6 || tt == 10 || tt == 14 || tt == 15 || tt == 16 || tt == 2) return 0;
    if (x0 == dG) return !(tt == 1 || tt == 3 || tt == 12 || tt == 2);
    if (x0 == dT) return !(tt < 4 || tt == 9 || (tt >= 21 && tt < 25));
    return 0;
}

/* _sequence.h:35-43: everything that is not A/T counts (N included); forward digits only */
static inline int is_gc_fwd(const uint8_t *d, int i) { return d[i] != dA && d[i] != dT; }

/* _sequence.h:207-220: 2 bits per base, first base in the low bits; N indexes as C */
static int mer_ndx(const uint8_t *d, int slen, int i, int len, int strand) {
    int ndx = 0;
    if (strand == 1) {
        for (int j = 0; j < len; j++) ndx |= (d[i + j] & 3) << (2 * j);
    } else {
        static const uint8_t comp[7] = {dT, dC, dG, dA, dN, dN, dN};
        for (int j = 0; j < len; j++) ndx |= (comp[d[slen - 1 - i - j]] & 3) << (2 * j);
    }
    return ndx;
}

/* ------------------------------------------------------------------------------------ */
/* node extraction (add_nodes)                                                           */
/* ------------------------------------------------------------------------------------ */

/* lib.pyx:337-340 for ONE mask (index m; -1 = the NULL cursor; n_masks = the zero-initialised slot behind the last
 * mask that the reverse-strand cursor can rest on, lib.pyx:2055 compares with &masks[length]) */
static int mask_hits(const orc_opts *o, int m, int begin, int end) {
    if (m < 0) return 0;
    const int mb = m < o->n_masks ? o->masks[2 * m] : 0, me = m < o->n_masks ? o->masks[2 * m + 1] : 0;
    return mb < end && begin < me;
}

typedef struct { orc_node *out; int n, cap; } node_sink;

static void emit(node_sink *s, int slen, int strand, int pos, int type, int stop_val, int edge) {
    if (s->n < s->cap) {
        orc_node *x = &s->out[s->n];
        memset(x, 0, sizeof(*x));
        /* reverse-strand coordinates are stored mirrored: lib.pyx:2039-2042, 2067-2070 */
        x->ndx = strand == 1 ? pos : slen - 1 - pos;
        x->stop_val = strand == 1 ? stop_val : slen - 1 - stop_val;
        x->strand = strand;
        x->type = type;
        x->edge = edge;
    }
    s->n++;
}

/* one strand of lib.pyx:1928-2020 (forward) / 2022-2115 (reverse); `i` runs over strand
 * coordinates, frame = i % 3 */
static void scan_strand(const uint8_t *d, int slen, int tt, const orc_opts *o, int strand, node_sink *s) {
    int last[3], saw[3] = {0, 0, 0}, min_dist[3];
    /* The reference does NOT test a candidate ORF against every mask: each frame keeps one cursor into the sorted mask
     * list (lib.pyx:1929-1932 / 2023-2026) and only the mask under the cursor is tested (1959-1966 / 2053-2061).
     * Forward strand: the cursor starts at the last mask and steps down while the ORF end lies before the mask
     * (-> the last mask with begin <= last); reverse strand: it starts at the first mask and steps up while the
     * ORF's left end lies beyond the mask's end (-> the first mask with end >= left end; it may rest on the zeroed
     * slot behind the list before it becomes NULL).  A second N run further inside the ORF is therefore not seen. */
    int cur[3];
    for (int k = 0; k < 3; k++) cur[k] = o->n_masks > 0 ? (strand == 1 ? o->n_masks - 1 : 0) : -1;
    for (int k = 0; k < 3; k++) {
        int f = (slen + k) % 3;
        last[f] = slen + k;
        min_dist[k] = o->min_edge_gene;
        if (!o->closed)
            while (last[f] + 3 > slen) last[f] -= 3;
    }
    for (int i = slen - 3; i >= 0; i--) {
        int f = i % 3;
        if (is_stop(d, slen, i, tt, strand)) {
            if (saw[f]) emit(s, slen, strand, last[f], ORC_STOP, i, !is_stop(d, slen, last[f], tt, strand));
            min_dist[f] = o->min_gene;
            last[f] = i;
            saw[f] = 0;
            continue;
        }
        if (last[f] >= slen) continue;
        if (o->n_masks) {
            int hit;
            if (strand == 1) {
                while (cur[f] >= 0 && last[f] < o->masks[2 * cur[f]]) cur[f] = cur[f] == 0 ? -1 : cur[f] - 1;
                hit = mask_hits(o, cur[f], i, last[f]);
            } else {
                const int left = slen - last[f] - 1;
                while (cur[f] >= 0 && left > (cur[f] < o->n_masks ? o->masks[2 * cur[f] + 1] : 0))
                    cur[f] = cur[f] == o->n_masks ? -1 : cur[f] + 1;
                hit = mask_hits(o, cur[f], left, slen - i - 1);
            }
            if (hit) continue;
        }
        if (last[f] - i + 3 >= min_dist[f] && is_start(d, slen, i, tt, strand)) {
            int b = base_at(d, slen, i, strand);
            emit(s, slen, strand, i, b == dA ? ORC_ATG : (b == dG ? ORC_GTG : ORC_TTG), last[f], 0);
            saw[f] = 1;
        } else if (i <= 2 && !o->closed && last[f] - i > o->min_edge_gene) {
            emit(s, slen, strand, i, ORC_ATG, last[f], 1);
            saw[f] = 1;
        }
    }
    for (int f = 0; f < 3; f++)
        if (saw[f]) emit(s, slen, strand, last[f], ORC_STOP, f - 6, !is_stop(d, slen, last[f], tt, strand));
}

int orc_extract(const uint8_t *digits, int slen, int tt, const orc_opts *o, orc_node *out, int cap) {
    node_sink s = {out, 0, cap};
    if (slen < 3) return 0;
    scan_strand(digits, slen, tt, o, 1, &s);
    scan_strand(digits, slen, tt, o, -1, &s);
    return s.n <= cap ? s.n : -1;
}

/* node.c:1578-1587 */
static int cmp_nodes(const void *a, const void *b) {
    const orc_node *x = a, *y = b;
    if (x->ndx != y->ndx) return x->ndx < y->ndx ? -1 : 1;
    if (x->strand != y->strand) return x->strand > y->strand ? -1 : 1;
    return 0;
}
void orc_sort(orc_node *nodes, int nn) { qsort(nodes, nn, sizeof(orc_node), cmp_nodes); }

/* node.c:176-197 -- note: `edge`, ndx, stop_val, strand, type, gc_cont survive */
void orc_reset_scores(orc_node *nodes, int nn) {
    for (int i = 0; i < nn; i++) {
        orc_node *x = &nodes[i];
        for (int k = 0; k < 3; k++) { x->star_ptr[k] = 0; x->gc_score[k] = 0.0; }
        x->rbs[0] = x->rbs[1] = 0;
        x->score = x->cscore = x->sscore = x->rscore = x->tscore = x->uscore = 0.0;
        x->traceb = x->tracef = x->ov_mark = -1;
        x->elim = 0; x->gc_bias = 0;
        x->mot_ndx = x->mot_len = x->mot_spacer = x->mot_spacendx = 0; x->mot_score = 0.0;
    }
}

/* ------------------------------------------------------------------------------------ */
/* node scoring                                                                          */
/* ------------------------------------------------------------------------------------ */

static int gc_triplet(const uint8_t *d, int slen, int lo) {
    int c = 0;
    for (int k = lo; k < lo + 3; k++)
        if (k >= 0 && k < slen) c += is_gc_fwd(d, k);
    return c;
}

/* lib.pyx:1846-1896 */
void orc_calc_orf_gc(const uint8_t *d, int slen, orc_node *nodes, int nn) {
    int last[3] = {0, 0, 0};
    double gc[3] = {0, 0, 0};
    for (int i = nn - 1; i >= 0; i--) {
        orc_node *x = &nodes[i];
        if (x->strand != 1) continue;
        int f = x->ndx % 3;
        if (x->type == ORC_STOP) {
            last[f] = x->ndx;
            gc[f] = gc_triplet(d, slen, x->ndx);
        } else {
            for (int j = last[f] - 3; j >= x->ndx; j -= 3) gc[f] += gc_triplet(d, slen, j);
            x->gc_cont = (float)(gc[f] / (abs(x->stop_val - x->ndx) + 3.0));
            last[f] = x->ndx;
        }
    }
    gc[0] = gc[1] = gc[2] = 0.0;
    for (int i = 0; i < nn; i++) {
        orc_node *x = &nodes[i];
        if (x->strand != -1) continue;
        int f = x->ndx % 3;
        if (x->type == ORC_STOP) {
            last[f] = x->ndx;
            gc[f] = gc_triplet(d, slen, x->ndx - 2); /* bases ndx, ndx-1, ndx-2 */
        } else {
            /* window deliberately starts at j (two bases right of the codon): lib.pyx:1887-1890 */
            for (int j = last[f] + 3; j <= x->ndx; j += 3) gc[f] += gc_triplet(d, slen, j);
            x->gc_cont = (float)(gc[f] / (abs(x->stop_val - x->ndx) + 3.0));
            last[f] = x->ndx;
        }
    }
}

/* lib.pyx:2119-2239 */
void orc_raw_coding_score(const uint8_t *d, int slen, orc_node *nodes, int nn, const orc_training *t) {
    double score[3], no_stop, lfac, lfac_min, lfac_max;
    long last[3] = {0, 0, 0};
    double a = 1 - t->gc;
    if (t->trans_table != 11) {
        no_stop = (a * a * t->gc) / 8.0;
        no_stop += (a * a * a) / 8.0;
    } else {
        no_stop = (a * a * t->gc) / 4.0;
        no_stop += (a * a * a) / 8.0;
    }
    no_stop = 1 - no_stop;
    lfac_max = log((1 - pow(no_stop, 1000.0)) / pow(no_stop, 1000.0));
    lfac_min = log((1 - pow(no_stop, 80)) / pow(no_stop, 80));

    /* pass 1: dicodon log-odds summed from the stop towards each start */
    score[0] = score[1] = score[2] = 0.0;
    for (int i = nn - 1; i >= 0; i--) {
        orc_node *x = &nodes[i];
        if (x->strand != 1) continue;
        int f = x->ndx % 3;
        if (x->type == ORC_STOP) { last[f] = x->ndx; score[f] = 0.0; continue; }
        for (long j = last[f] - 3; j >= x->ndx; j -= 3) score[f] += t->gene_dc[mer_ndx(d, slen, (int)j, 6, 1)];
        x->cscore = score[f];
        last[f] = x->ndx;
    }
    score[0] = score[1] = score[2] = 0.0;
    for (int i = 0; i < nn; i++) {
        orc_node *x = &nodes[i];
        if (x->strand != -1) continue;
        int f = x->ndx % 3;
        if (x->type == ORC_STOP) { last[f] = x->ndx; score[f] = 0.0; continue; }
        for (long j = last[f] + 3; j <= x->ndx; j += 3)
            score[f] += t->gene_dc[mer_ndx(d, slen, slen - 1 - (int)j, 6, -1)];
        x->cscore = score[f];
        last[f] = x->ndx;
    }

    /* pass 2: penalise starts with better-coding starts upstream of them */
    for (int dir = 0; dir < 2; dir++) {
        score[0] = score[1] = score[2] = -10000.0;
        for (int k = 0; k < nn; k++) {
            int i = dir == 0 ? k : nn - 1 - k;
            orc_node *x = &nodes[i];
            if (x->strand != (dir == 0 ? 1 : -1)) continue;
            int f = x->ndx % 3;
            if (x->type == ORC_STOP) score[f] = -10000.0;
            else if (x->cscore > score[f]) score[f] = x->cscore;
            else x->cscore -= (score[f] - x->cscore);
        }
    }

    /* pass 3: length factor.  NOTE lib.pyx:2199-2217: the per-frame state is NOT re-initialised
     * between pass 2 and pass 3 (it carries over until the first STOP node resets it). */
    for (int dir = 0; dir < 2; dir++) {
        for (int k = 0; k < nn; k++) {
            int i = dir == 0 ? k : nn - 1 - k;
            orc_node *x = &nodes[i];
            if (x->strand != (dir == 0 ? 1 : -1)) continue;
            int f = x->ndx % 3;
            if (x->type == ORC_STOP) { score[f] = -10000.0; continue; }
            double gsize = dir == 0 ? (((double)x->stop_val - x->ndx) + 3.0) / 3.0
                                    : (((double)x->ndx - x->stop_val) + 3.0) / 3.0;
            if (gsize > 1000.0) {
                lfac = (lfac_max - lfac_min) * (gsize - 80) / 920.0;
            } else {
                double tmp = pow(no_stop, gsize);
                lfac = log((1 - tmp) / tmp) - lfac_min;
            }
            if (lfac > score[f]) score[f] = lfac;
            else lfac -= fmax(fmin(score[f] - lfac, lfac), 0);
            if (lfac > 3.0 && x->cscore < 0.5 * lfac) x->cscore = 0.5 * lfac;
            x->cscore += lfac;
        }
    }
}

/* lib.pyx:791-890 */
int orc_shine_dalgarno_exact(const uint8_t *d, int slen, int pos, int start, const double *rbs_wt, int strand) {
    static const int8_t tab[15][4] = {
        [6] = {13, 6, 1, 2},   [8] = {15, 12, 11, 3},  [9] = {16, 12, 11, 3},
        [11] = {22, 21, 20, 10}, [12] = {24, 23, 20, 10}, [14] = {27, 26, 25, 10}};
    int match[6] = {-10, -10, -10, -10, -10, -10};
    int limit = start - 4 - pos;
    if (limit > 6) limit = 6;
    for (int i = 0; i < limit; i++) {
        int p = pos + i;
        if (p < 0 || p >= slen) continue;
        int b = base_at(d, slen, p, strand);
        if (i % 3 == 0) { if (b == dA) match[i] = 2; }
        else            { if (b == dG) match[i] = 3; }
    }
    int max_val = 0, cur_val = 0;
    for (int len = limit; len > 2; len--) {
        for (int j = 0; j <= limit - len; j++) {
            int ctr = -2;
            for (int k = j; k < j + len; k++) ctr += match[k];
            if (ctr < 6) continue;
            int rdis = start - (pos + j + len), flag;
            if (rdis < 5) flag = len < 5 ? 2 : 1;
            else if (rdis < 11) flag = 0;
            else if (rdis < 13) flag = len < 5 ? 1 : 2;
            else if (rdis < 16) flag = 3;
            else continue;
            cur_val = (ctr <= 14) ? tab[ctr][flag] : 0;
            if (rbs_wt[cur_val] < rbs_wt[max_val]) continue;
            if (rbs_wt[cur_val] == rbs_wt[max_val] && cur_val < max_val) continue;
            max_val = cur_val;
        }
    }
    return max_val;
}

/* lib.pyx:892-979 */
int orc_shine_dalgarno_mm(const uint8_t *d, int slen, int pos, int start, const double *rbs_wt, int strand) {
    int match[6] = {-10, -10, -10, -10, -10, -10};
    int limit = start - 4 - pos;
    if (limit > 6) limit = 6;
    for (int i = 0; i < limit; i++) {
        int p = pos + i;
        if (p >= 0 && p < slen) {
            int b = base_at(d, slen, p, strand);
            if (i % 3 == 0) match[i] = b == dA ? 2 : -3;
            else            match[i] = b == dG ? 3 : -2;
        } else {
            match[i] = i % 3 == 0 ? -3 : -2;
        }
    }
    int max_val = 0, cur_val = 0; /* cur_val is sticky across iterations, as in the reference */
    for (int len = limit; len > 4; len--) {
        for (int j = 0; j <= limit - len; j++) {
            int ctr = -2, mism = 0;
            for (int k = j; k < j + len; k++) {
                ctr += match[k];
                if (match[k] < 0) {
                    mism++;
                    if (k <= j + 1 || k >= j + len - 2) ctr -= 10;
                }
            }
            if (mism != 1 || ctr < 6) continue;
            int rdis = start - (pos + j + len), flag;
            if (rdis < 5) flag = 1;
            else if (rdis < 11) flag = 0;
            else if (rdis < 13) flag = 2;
            else if (rdis < 16) flag = 3;
            else continue;
            if (ctr == 6)      { static const int8_t v[4] = {9, 5, 4, 2};    cur_val = v[flag]; }
            else if (ctr == 7) { static const int8_t v[4] = {14, 8, 7, 2};   cur_val = v[flag]; }
            else if (ctr == 9) { static const int8_t v[4] = {19, 18, 17, 3}; cur_val = v[flag]; }
            if (rbs_wt[cur_val] < rbs_wt[max_val]) continue;
            if (rbs_wt[cur_val] == rbs_wt[max_val] && cur_val < max_val) continue;
            max_val = cur_val;
        }
    }
    return max_val;
}

/* lib.pyx:2241-2277 */
void orc_rbs_score(const uint8_t *d, int slen, orc_node *nodes, int nn, const orc_training *t) {
    for (int i = 0; i < nn; i++) {
        orc_node *x = &nodes[i];
        if (x->type == ORC_STOP || x->edge) continue;
        x->rbs[0] = x->rbs[1] = 0;
        int start = x->strand == 1 ? x->ndx : slen - 1 - x->ndx;
        for (int j = start - 20; j < start - 5; j++) {
            /* forward skips negative offsets; reverse only skips j >= slen (lib.pyx:2257-2269) */
            if (x->strand == 1 ? j < 0 : j >= slen) continue;
            int e = orc_shine_dalgarno_exact(d, slen, j, start, t->rbs_wt, x->strand);
            int m = orc_shine_dalgarno_mm(d, slen, j, start, t->rbs_wt, x->strand);
            if (e > x->rbs[0]) x->rbs[0] = e;
            if (m > x->rbs[1]) x->rbs[1] = m;
        }
    }
}

/* lib.pyx:1557-1616; only stage 2 (scoring, last training rounds) may report "no motif" */
static void best_upstream_motif(const uint8_t *d, int slen, orc_node *x, const orc_training *t, int stage) {
    if (x->type == ORC_STOP || x->edge) return;
    int start = x->strand == 1 ? x->ndx : slen - 1 - x->ndx;
    int max_spacer = 0, max_spacendx = 0, max_len = 0, max_ndx = 0;
    double max_sc = -100.0;
    for (int i = 3; i >= 0; i--) {
        for (int j = start - 18 - i; j <= start - 6 - i; j++) {
            if (j < 0) continue;
            int spacendx;
            if (j <= start - 16 - i) spacendx = 3;
            else if (j <= start - 14 - i) spacendx = 2;
            else if (j >= start - 7 - i) spacendx = 1;
            else spacendx = 0;
            int index = mer_ndx(d, slen, j, i + 3, x->strand);
            double sc = t->mot_wt[i][spacendx][index];
            if (sc > max_sc) {
                max_sc = sc; max_spacendx = spacendx; max_spacer = start - j - i - 3;
                max_ndx = index; max_len = i + 3;
            }
        }
    }
    if (stage == 2 && (max_sc == -4.0 || max_sc < t->no_mot + 0.69)) {
        x->mot_ndx = 0; x->mot_len = 0; x->mot_spacendx = 0; x->mot_spacer = 0; x->mot_score = t->no_mot;
    } else {
        x->mot_ndx = max_ndx; x->mot_len = max_len; x->mot_spacendx = max_spacendx;
        x->mot_spacer = max_spacer; x->mot_score = max_sc;
    }
}

/* lib.pyx:1619-1650 */
static double upstream_composition(const uint8_t *d, int slen, const orc_node *x, const orc_training *t) {
    int start = x->strand == 1 ? x->ndx : slen - 1 - x->ndx;
    int count = 0;
    double u = 0.0;
    for (int i = 1; i < 3 && i <= start; i++, count++)
        u += 0.4 * t->st_wt * t->ups_comp[count][mer_ndx(d, slen, start - i, 1, x->strand)];
    for (int i = 15; i < 45 && i <= start; i++, count++)
        u += 0.4 * t->st_wt * t->ups_comp[count][mer_ndx(d, slen, start - i, 1, x->strand)];
    return u;
}

/* lib.pyx:2331-2487 */
void orc_score(const uint8_t *d, int slen, orc_node *nodes, int nn, const orc_training *t, int closed, int is_meta) {
    orc_calc_orf_gc(d, slen, nodes, nn);
    orc_raw_coding_score(d, slen, nodes, nn, t);
    if (t->uses_sd) orc_rbs_score(d, slen, nodes, nn, t);
    else for (int i = 0; i < nn; i++) best_upstream_motif(d, slen, &nodes[i], t, 2);

    for (int i = 0; i < nn; i++) {
        orc_node *x = &nodes[i];
        if (x->type == ORC_STOP) continue;
        int orf_length = abs(x->ndx - x->stop_val);
        int edge_gene = 0;
        if (x->edge) edge_gene++;
        if ((x->strand == 1 && !is_stop(d, slen, x->stop_val, t->trans_table, 1)) ||
            (x->strand == -1 && !is_stop(d, slen, slen - 1 - x->stop_val, t->trans_table, -1)))
            edge_gene++;

        if (x->edge) {
            x->tscore = EDGE_BONUS * t->st_wt / edge_gene;
            x->uscore = 0.0;
            x->rscore = 0.0;
        } else {
            x->tscore = t->type_wt[x->type] * t->st_wt;
            double rbs1 = t->rbs_wt[x->rbs[0]], rbs2 = t->rbs_wt[x->rbs[1]];
            double sd_score = fmax(rbs1, rbs2) * t->st_wt;
            if (t->uses_sd) {
                x->rscore = sd_score;
            } else {
                x->rscore = t->st_wt * x->mot_score;
                if (x->rscore < sd_score && t->no_mot > -0.5) x->rscore = sd_score;
            }
            x->uscore = upstream_composition(d, slen, x, t);
            /* penalise starts that would stop the gene running off the edge (lib.pyx:2407-2422) */
            if (!closed && x->ndx <= 2 && x->strand == 1) {
                x->uscore += EDGE_UPS * t->st_wt;
            } else if (!closed && x->ndx >= slen - 3 && x->strand == -1) {
                x->uscore += EDGE_UPS * t->st_wt;
            } else if (i < 500 && x->strand == 1) {
                for (int j = i - 1; j >= 0; j--)
                    if (nodes[j].edge && x->stop_val == nodes[j].stop_val) { x->uscore += EDGE_UPS * t->st_wt; break; }
            } else if (i + 500 >= nn && x->strand == -1) {
                for (int j = i + 1; j < nn; j++)
                    if (nodes[j].edge && x->stop_val == nodes[j].stop_val) { x->uscore += EDGE_UPS * t->st_wt; break; }
            }
        }

        /* starts at the very first/last bases become edge nodes (lib.pyx:2424-2434) */
        if (!closed && !x->edge &&
            ((x->ndx <= 2 && x->strand == 1) || (x->ndx >= slen - 3 && x->strand == -1))) {
            edge_gene++;
            x->edge = 1;
            x->tscore = 0.0;
            x->uscore = EDGE_BONUS * t->st_wt / edge_gene;
            x->rscore = 0.0;
        }
        if (!x->edge && edge_gene == 1) x->uscore -= 0.5 * EDGE_BONUS * t->st_wt;

        if (edge_gene == 0 && orf_length < 250) {
            double negf = 250.0 / (float)orf_length, posf = (float)orf_length / 250.0;
            x->rscore *= x->rscore < 0 ? negf : posf;
            x->uscore *= x->uscore < 0 ? negf : posf;
            x->tscore *= x->tscore < 0 ? negf : posf;
        }
        if (is_meta && slen < 3000 && edge_gene == 0 && (x->cscore < 5.0 || orf_length < 120))
            x->cscore -= META_PEN * fmax(0, (3000.0 - slen) / 2700.0);

        x->sscore = x->tscore + x->rscore + x->uscore;

        if (x->cscore < 0.0) {
            if (edge_gene > 0 && !x->edge) {
                if (!is_meta || slen > 1500) x->sscore -= t->st_wt;
                else x->sscore -= 10.31 - 0.004 * slen;
            } else if (is_meta && slen < 3000 && x->edge) {
                double min_meta_len = sqrt((double)slen) * 5.0;
                if (orf_length >= min_meta_len) {
                    if (x->cscore >= 0) x->cscore = -1.0;
                    x->sscore = 0.0;
                    x->uscore = 0.0;
                }
            } else {
                x->sscore -= 0.5;
            }
        } else if (is_meta && x->cscore < 5.0 && orf_length < 120 && x->sscore < 0.0) {
            x->sscore -= t->st_wt;
        }
    }
}

/* ------------------------------------------------------------------------------------ */
/* intergenic modifiers: _connection.h:43-91 (== node.c:1377-1402)                       */
/* ------------------------------------------------------------------------------------ */

static double igm_diff(double st_wt) { return -0.15 * st_wt; }

static double igm_same(const orc_node *n1, const orc_node *n2, double st_wt) {
    int dist = abs(n1->ndx - n2->ndx);
    int overlap = n1->ndx + 2 * n1->strand >= n2->ndx;
    double r = 0.0;
    if (n1->ndx + 2 == n2->ndx || n1->ndx == n2->ndx + 1) {
        const orc_node *s = n1->strand == 1 ? n2 : n1; /* the start whose RBS is waived */
        if (s->rscore < 0) r -= s->rscore;
        if (s->uscore < 0) r -= s->uscore;
    }
    if (dist > 3 * OPER_DIST) r -= 0.15 * st_wt;
    else if ((dist <= OPER_DIST && !overlap) || dist * 4 < OPER_DIST)
        r += (2.0 - ((double)dist / OPER_DIST)) * 0.15 * st_wt;
    return r;
}

static double igm(const orc_node *n1, const orc_node *n2, double st_wt) {
    return n1->strand == n2->strand ? igm_same(n1, n2, st_wt) : igm_diff(st_wt);
}

/* lib.pyx:2279-2329 */
void orc_record_overlapping_starts(orc_node *nodes, int nn, const orc_training *t, int flag, int max_overlap) {
    for (int i = 0; i < nn; i++) {
        orc_node *x = &nodes[i];
        x->star_ptr[0] = x->star_ptr[1] = x->star_ptr[2] = -1;
        if (x->type != ORC_STOP || x->edge == 1) continue;
        double max_sc = -100;
        if (x->strand == 1) {
            for (int j = i + 3; j >= 0; j--) {
                if (j >= nn || nodes[j].ndx > x->ndx + 2) continue;
                if (nodes[j].ndx + max_overlap < x->ndx) break;
                if (nodes[j].strand != 1 || nodes[j].type == ORC_STOP) continue;
                if (nodes[j].stop_val <= x->ndx) continue;
                int f = nodes[j].ndx % 3;
                if (flag == 0) {
                    if (x->star_ptr[f] == -1) x->star_ptr[f] = j;
                } else {
                    double sc = nodes[j].cscore + nodes[j].sscore + igm_same(x, &nodes[j], t->st_wt);
                    if (sc > max_sc) { x->star_ptr[f] = j; max_sc = sc; }
                }
            }
        } else {
            for (int j = i - 3; j < nn; j++) {
                if (j < 0 || nodes[j].ndx < x->ndx - 2) continue;
                if (nodes[j].ndx - max_overlap > x->ndx) break;
                if (nodes[j].strand != -1 || nodes[j].type == ORC_STOP) continue;
                if (nodes[j].stop_val >= x->ndx) continue;
                int f = nodes[j].ndx % 3;
                if (flag == 0) {
                    if (x->star_ptr[f] == -1) x->star_ptr[f] = j;
                } else {
                    double sc = nodes[j].cscore + nodes[j].sscore + igm_same(&nodes[j], x, t->st_wt);
                    if (sc > max_sc) { x->star_ptr[f] = j; max_sc = sc; }
                }
            }
        }
    }
}

/* ------------------------------------------------------------------------------------ */
/* connection scoring DP                                                                 */
/* ------------------------------------------------------------------------------------ */

static inline int kind_of(const orc_node *x) { return 2 * (x->strand != 1) + (x->type == ORC_STOP); }
enum { K_FS = 0 /* +start */, K_FE = 1 /* +STOP */, K_RS = 2 /* -start */, K_RE = 3 /* -STOP */ };

/* impl/generic.h:29-36 restated as "which predecessor kinds a target kind accepts" */
int orc_skippable(const orc_node *nodes, int j, int i) {
    int k1 = kind_of(&nodes[j]), k2 = kind_of(&nodes[i]);
    int same_frame = nodes[j].ndx % 3 == nodes[i].ndx % 3;
    switch (k2) {
    case K_FS: return !(k1 == K_FE || k1 == K_RS);
    case K_FE: return !((k1 == K_FS && same_frame) || k1 == K_FE);
    case K_RS: return !((k1 == K_RE && same_frame) || k1 == K_FE);
    default:   return !(k1 == K_FE || k1 == K_RS || k1 == K_RE);
    }
}

static double gc_bias_sum(const orc_node *x, const orc_training *t) {
    return t->bias[0] * x->gc_score[0] + t->bias[1] * x->gc_score[1] + t->bias[2] * x->gc_score[2];
}

/* One admissible (j -> i) connection: _connection.h:94-367.  Returns 0 when the pair is
 * rejected, else 1 with *val = the increment and *maxfr = the overlap marker. */
static int connection(const orc_node *nodes, int j, int i, const orc_training *t, int final, double *val, int *maxfr_out) {
    const orc_node *n1 = &nodes[j], *n2 = &nodes[i], *n3;
    int k1 = kind_of(n1), k2 = kind_of(n2);
    int left = n1->ndx, right = n2->ndx, ovlp = 0, maxfr = -1;
    double score = 0.0, scr_mod = 0.0;

    /* edge artifacts: a gene end that nothing leads into */
    if (n1->traceb == -1 && (k1 == K_FE || k1 == K_RS)) return 0;

    if (k2 == K_FS) {
        if (k1 == K_FE) { /* 3'fwd -> 5'fwd */
            left += 2;
            if (left >= right) return 0;
            if (final) score = igm_same(n1, n2, t->st_wt);
        } else if (k1 == K_RS) { /* 5'rev -> 5'fwd */
            if (left >= right) return 0;
            if (final) score = igm_diff(t->st_wt);
        }
    } else if (k2 == K_FE) {
        if (k1 == K_FS) { /* gene */
            if (n2->stop_val >= n1->ndx) return 0;
            right += 2;
            if (final) score = n1->cscore + n1->sscore;
            else scr_mod = gc_bias_sum(n1, t);
        } else if (k1 == K_FE) { /* operon: overlapping forward genes */
            if (n2->stop_val >= n1->ndx) return 0;
            int s = n1->star_ptr[n2->ndx % 3];
            if (s == -1) return 0;
            n3 = &nodes[s];
            left = n3->ndx;
            right += 2;
            if (final) score = n3->cscore + n3->sscore + igm(n1, n3, t->st_wt);
            else scr_mod = gc_bias_sum(n3, t);
        }
    } else if (k2 == K_RS) {
        if (k1 == K_RE) { /* reverse gene */
            if (n1->stop_val <= n2->ndx) return 0;
            left -= 2;
            if (final) score = n2->cscore + n2->sscore;
            else scr_mod = gc_bias_sum(n2, t);
        } else if (k1 == K_FE) { /* opposite-strand 3' overlap */
            if (n2->stop_val - 2 >= n1->ndx + 2) return 0;
            ovlp = (n1->ndx + 2) - (n2->stop_val - 2) + 1;
            if (ovlp >= MAX_OPP_OVLP) return 0;
            if ((n1->ndx - n2->stop_val) >= (n2->ndx - n1->ndx + 3)) return 0;
            int bnd = n1->traceb == -1 ? 0 : nodes[n1->traceb].ndx;
            if ((n1->ndx - n2->stop_val) >= (n2->stop_val - 3 - bnd)) return 0;
            left = n2->stop_val - 2;
            if (final) score = n2->cscore + n2->sscore + igm_diff(t->st_wt);
            else scr_mod = gc_bias_sum(n2, t);
        }
    } else { /* K_RE */
        if (k1 == K_FE) { /* 3'fwd -> 3'rev, with the triple-overlap search */
            left += 2;
            right -= 2;
            if (left >= right) return 0;
            double maxval = 0.0;
            for (int k = 0; k < 3; k++) {
                if (n2->star_ptr[k] == -1) continue;
                n3 = &nodes[n2->star_ptr[k]];
                ovlp = left - n3->stop_val + 3; /* keeps the last examined value, as in the reference */
                if (ovlp <= 0 || ovlp >= MAX_OPP_OVLP) continue;
                if (ovlp >= n3->ndx - left) continue;
                if (n1->traceb == -1) continue;
                if (ovlp >= n3->stop_val - nodes[n1->traceb].ndx - 2) continue;
                double cur = n3->cscore + n3->sscore + igm(n3, n2, t->st_wt);
                if ((final && cur > maxval) || (!final && gc_bias_sum(n3, t) > maxval)) { maxfr = k; maxval = cur; }
            }
            if (maxfr != -1) {
                n3 = &nodes[n2->star_ptr[maxfr]];
                if (final) score = n3->cscore + n3->sscore + igm(n3, n2, t->st_wt);
                else scr_mod = gc_bias_sum(n3, t);
            } else if (final) {
                score = igm_diff(t->st_wt);
            }
        } else if (k1 == K_RS) { /* 5'rev -> 3'rev */
            right -= 2;
            if (left >= right) return 0;
            if (final) score = igm_same(n1, n2, t->st_wt);
        } else if (k1 == K_RE) { /* operon: overlapping reverse genes */
            if (n1->stop_val <= n2->ndx) return 0;
            int s = n2->star_ptr[n1->ndx % 3];
            if (s == -1) return 0;
            n3 = &nodes[s];
            left -= 2;
            right = n3->ndx;
            if (final) score = n3->cscore + n3->sscore + igm(n3, n2, t->st_wt);
            else scr_mod = gc_bias_sum(n3, t);
        }
    }
    if (!final) score = ((double)(right - left + 1 - ovlp * 2)) * scr_mod;
    *val = score;
    *maxfr_out = maxfr;
    return 1;
}

/* lib.pyx:1224-1233: start of node i's predecessor window */
static int window_start(const orc_node *nodes, int i) {
    int m = i < MAX_NODE_DIST ? 0 : i - MAX_NODE_DIST;
    int k = kind_of(&nodes[i]);
    if ((k == K_RS || k == K_FE) && nodes[m].ndx > nodes[i].stop_val)
        while (m > 0 && nodes[m].ndx != nodes[i].stop_val) m--;
    return m < MAX_NODE_DIST ? 0 : m - MAX_NODE_DIST;
}

/* lib.pyx:1205-1237 */
void orc_score_connections(orc_node *nodes, int nn, const orc_training *t, int final, int64_t *pairs) {
    int64_t np = 0;
    for (int i = 0; i < nn; i++) { nodes[i].score = 0; nodes[i].traceb = -1; nodes[i].tracef = -1; }
    for (int i = 0; i < nn; i++) {
        int m = window_start(nodes, i);
        np += i - m;
        for (int j = m; j < i; j++) {
            double v; int fr;
            if (orc_skippable(nodes, j, i)) continue;
            if (!connection(nodes, j, i, t, final, &v, &fr)) continue;
            if (nodes[j].score + v >= nodes[i].score) {
                nodes[i].score = nodes[j].score + v;
                nodes[i].traceb = j;
                nodes[i].ov_mark = fr;
            }
        }
    }
    if (pairs) *pairs = np;
}

/* lib.pyx:1239-1311 (== dprog.c:58-108) */
int orc_dynamic_programming(orc_node *nodes, int nn, const orc_training *t, int final) {
    if (nn == 0) return -1;
    orc_score_connections(nodes, nn, t, final, NULL);

    int best = -1;
    double best_sc = -1.0;
    for (int i = nn - 1; i >= 0; i--) {
        int k = kind_of(&nodes[i]);
        if (k == K_FS || k == K_RE) continue;
        if (nodes[i].score > best_sc) { best_sc = nodes[i].score; best = i; }
    }
    if (best < 0) return -1; /* the reference would index nodes[-1] here */

    /* pass 1: triple overlaps */
    for (int path = best; nodes[path].traceb != -1; path = nodes[path].traceb) {
        int nxt = nodes[path].traceb;
        if (kind_of(&nodes[path]) == K_RE && kind_of(&nodes[nxt]) == K_FE && nodes[path].ov_mark != -1 &&
            nodes[path].ndx > nodes[nxt].ndx) {
            int tmp = nodes[path].star_ptr[nodes[path].ov_mark], i = tmp;
            while (nodes[i].ndx != nodes[tmp].stop_val) i--;
            nodes[path].traceb = tmp;
            nodes[tmp].traceb = i;
            nodes[i].ov_mark = -1;
            nodes[i].traceb = nxt;
        }
    }
    /* pass 2: simple overlaps */
    for (int path = best; nodes[path].traceb != -1; path = nodes[path].traceb) {
        int nxt = nodes[path].traceb;
        int kp = kind_of(&nodes[path]), kn = kind_of(&nodes[nxt]);
        if (kp == K_RS && kn == K_FE) {
            int i = path;
            while (nodes[i].ndx != nodes[path].stop_val) i--;
            nodes[path].traceb = i;
            nodes[i].traceb = nxt;
        }
        if (kp == K_FE && kn == K_FE) {
            nodes[path].traceb = nodes[nxt].star_ptr[nodes[path].ndx % 3];
            nodes[nodes[path].traceb].traceb = nxt;
        }
        if (kp == K_RE && kn == K_RE) {
            nodes[path].traceb = nodes[path].star_ptr[nodes[nxt].ndx % 3];
            nodes[nodes[path].traceb].traceb = nxt;
        }
    }
    /* forward pointers */
    for (int path = best; nodes[path].traceb != -1; path = nodes[path].traceb)
        nodes[nodes[path].traceb].tracef = path;

    return nodes[best].traceb == -1 ? -1 : best;
}

/* dprog.c:306-335 */
void orc_eliminate_bad_genes(orc_node *nodes, int ipath, const orc_training *t) {
    if (ipath == -1) return;
    int head = ipath;
    while (nodes[head].traceb != -1) head = nodes[head].traceb;
    for (int p = head; nodes[p].tracef != -1; p = nodes[p].tracef) {
        int nx = nodes[p].tracef, k = kind_of(&nodes[p]);
        if (k == K_FE) nodes[nx].sscore += igm(&nodes[p], &nodes[nx], t->st_wt);
        if (k == K_RS) nodes[p].sscore += igm(&nodes[p], &nodes[nx], t->st_wt);
    }
    for (int p = head; nodes[p].tracef != -1; p = nodes[p].tracef) {
        int nx = nodes[p].tracef, k = kind_of(&nodes[p]);
        if (k == K_FS && nodes[p].cscore + nodes[p].sscore < 0) { nodes[p].elim = 1; nodes[nx].elim = 1; }
        if (k == K_RE && nodes[nx].cscore + nodes[nx].sscore < 0) { nodes[p].elim = 1; nodes[nx].elim = 1; }
    }
}

/* lib.pyx:3231-3270 */
int orc_genes_extract(const orc_node *nodes, int ipath, orc_gene *genes, int cap) {
    int ng = 0, begin = 0, end = 0, start_ndx = 0, stop_ndx = 0;
    if (ipath == -1) return 0;
    int p = ipath;
    while (nodes[p].traceb != -1) p = nodes[p].traceb;
    for (; p != -1; p = nodes[p].tracef) {
        const orc_node *x = &nodes[p];
        if (x->elim == 1) continue;
        int emit_gene = 0;
        if (x->strand == 1) {
            if (x->type != ORC_STOP) { begin = x->ndx + 1; start_ndx = p; }
            else { end = x->ndx + 3; stop_ndx = p; emit_gene = 1; }
        } else {
            if (x->type != ORC_STOP) { end = x->ndx + 1; start_ndx = p; emit_gene = 1; }
            else { begin = x->ndx - 1; stop_ndx = p; }
        }
        if (emit_gene) {
            if (ng < cap) { genes[ng].begin = begin; genes[ng].end = end; genes[ng].start_ndx = start_ndx; genes[ng].stop_ndx = stop_ndx; }
            ng++;
        }
    }
    return ng;
}

/* lib.pyx:3272-3401 */
void orc_tweak_final_starts(orc_gene *genes, int ng, const orc_node *nodes, int nn, const orc_training *t, int max_overlap) {
    for (int i = 0; i < ng; i++) {
        int ndx = genes[i].start_ndx;
        const orc_node *cur = &nodes[ndx];
        double sc = cur->sscore + cur->cscore, igm0 = 0.0;
        const orc_node *pstart = i > 0 ? &nodes[genes[i - 1].start_ndx] : NULL;
        const orc_node *pstop = i > 0 ? &nodes[genes[i - 1].stop_ndx] : NULL;
        const orc_node *nstart = i < ng - 1 ? &nodes[genes[i + 1].start_ndx] : NULL;
        const orc_node *nstop = i < ng - 1 ? &nodes[genes[i + 1].stop_ndx] : NULL;

        if (pstart && cur->strand == 1 && pstart->strand == 1) igm0 = igm_same(pstop, cur, t->st_wt);
        if (pstart && cur->strand == 1 && pstart->strand == -1) igm0 = igm_diff(t->st_wt);
        if (nstart && cur->strand == -1 && nstart->strand == 1) igm0 = igm_diff(t->st_wt);
        if (nstart && cur->strand == -1 && nstart->strand == -1) igm0 = igm_same(cur, nstop, t->st_wt);

        int maxndx[2] = {-1, -1};
        double maxsc[2] = {0, 0}, maxigm[2] = {0, 0};
        for (int j = ndx - 100; j < ndx + 100; j++) {
            if (j < 0 || j >= nn || j == ndx) continue;
            const orc_node *c = &nodes[j];
            if (c->type == ORC_STOP || c->stop_val != cur->stop_val) continue;
            double tigm = 0.0;
            if (pstart && c->strand == 1 && pstart->strand == 1) {
                if (pstop->ndx - c->ndx > max_overlap) continue;
                tigm = igm_same(pstop, c, t->st_wt);
            }
            if (pstart && c->strand == 1 && pstart->strand == -1) {
                if (pstart->ndx - c->ndx >= 0) continue;
                tigm = igm_diff(t->st_wt);
            }
            if (nstart && c->strand == -1 && nstart->strand == 1) {
                if (c->ndx - nstart->ndx >= 0) continue;
                tigm = igm_diff(t->st_wt);
            }
            if (nstart && c->strand == -1 && nstart->strand == -1) {
                if (c->ndx - nstop->ndx > max_overlap) continue;
                tigm = igm_same(c, nstop, t->st_wt);
            }
            double csc = c->cscore + c->sscore;
            if (maxndx[0] == -1) {
                maxndx[0] = j; maxsc[0] = csc; maxigm[0] = tigm;
            } else if (csc + tigm > maxsc[0]) {
                maxndx[1] = maxndx[0]; maxsc[1] = maxsc[0]; maxigm[1] = maxigm[0];
                maxndx[0] = j; maxsc[0] = csc; maxigm[0] = tigm;
            } else if (maxndx[1] == -1 || csc + tigm > maxsc[1]) {
                maxndx[1] = j; maxsc[1] = csc; maxigm[1] = tigm;
            }
        }
        for (int j = 0; j < 2; j++) {
            if (maxndx[j] == -1) continue;
            const orc_node *m = &nodes[maxndx[j]];
            if (m->tscore < cur->tscore && maxsc[j] - m->tscore >= sc - cur->tscore + t->st_wt &&
                m->rscore > cur->rscore && m->uscore > cur->uscore && m->cscore > cur->cscore &&
                abs(m->ndx - cur->ndx) > 15) {
                maxsc[j] += cur->tscore - m->tscore;
            } else if (abs(m->ndx - cur->ndx) <= 15 && m->rscore + m->tscore > cur->rscore + cur->tscore &&
                       cur->edge == 0 && m->edge == 0) {
                if (cur->cscore > m->cscore) maxsc[j] += cur->cscore - m->cscore;
                if (cur->uscore > m->uscore) maxsc[j] += cur->uscore - m->uscore;
                if (igm0 > maxigm[j]) maxsc[j] += igm0 - maxigm[j];
            } else {
                maxsc[j] = -1000.0;
            }
        }
        int pick = -1;
        for (int j = 0; j < 2; j++) {
            if (maxndx[j] == -1) continue;
            if (pick == -1 && maxsc[j] + maxigm[j] > sc + igm0) pick = j;
            else if (pick >= 0 && maxsc[j] + maxigm[j] > maxsc[pick] + maxigm[pick]) pick = j;
        }
        if (pick != -1) {
            const orc_node *m = &nodes[maxndx[pick]];
            genes[i].start_ndx = maxndx[pick];
            if (m->strand == 1) genes[i].begin = m->ndx + 1;
            else genes[i].end = m->ndx + 1;
        }
    }
}

/* ------------------------------------------------------------------------------------ */
/* drivers                                                                               */
/* ------------------------------------------------------------------------------------ */

/* lib.pyx:5281-5315 */
int orc_find_genes_single(const uint8_t *digits, int slen, const orc_training *t, const orc_opts *o, orc_node *nodes,
                          int node_cap, int *nn_out, orc_gene *genes, int gene_cap, int *ipath_out) {
    int nn = orc_extract(digits, slen, t->trans_table, o, nodes, node_cap);
    if (nn < 0) return -1;
    orc_sort(nodes, nn);
    orc_reset_scores(nodes, nn);
    orc_score(digits, slen, nodes, nn, t, o->closed, 0);
    orc_record_overlapping_starts(nodes, nn, t, 1, o->max_overlap);
    int ipath = orc_dynamic_programming(nodes, nn, t, 1);
    if (nn > 0) orc_eliminate_bad_genes(nodes, ipath, t);
    int ng = orc_genes_extract(nodes, ipath, genes, gene_cap);
    if (ng > gene_cap) return -1;
    orc_tweak_final_starts(genes, ng, nodes, nn, t, o->max_overlap);
    if (nn_out) *nn_out = nn;
    if (ipath_out) *ipath_out = ipath;
    return ng;
}

/* lib.pyx:5317-5396 */
int orc_find_genes_meta(const uint8_t *digits, int slen, double gc, const orc_training *bins, int n_bins,
                        const orc_opts *o, orc_node *nodes, int node_cap, int *nn_out, orc_gene *genes, int gene_cap,
                        int *winner, int64_t *pairs) {
    double low = fmin(0.65, 0.88495 * gc - 0.0102337), high = fmax(0.35, 0.86596 * gc + 0.1131991);
    double max_score = -100.0;
    int tt = -1, max_phase = -1, nn = 0, ng = 0;
    int64_t tot_pairs = 0;
    for (int b = 0; b < n_bins; b++) {
        const orc_training *t = &bins[b];
        if (t->gc < low || t->gc > high) continue;
        if (t->trans_table != tt) {
            tt = t->trans_table;
            nn = orc_extract(digits, slen, tt, o, nodes, node_cap);
            if (nn < 0) return -1;
            orc_sort(nodes, nn);
        }
        orc_reset_scores(nodes, nn);
        orc_score(digits, slen, nodes, nn, t, o->closed, 1);
        orc_record_overlapping_starts(nodes, nn, t, 1, o->max_overlap);
        int ipath = -1;
        if (nn > 0) {
            /* same as orc_dynamic_programming but counting pairs */
            int64_t np = 0;
            if (pairs) {
                for (int i = 0; i < nn; i++) np += i - window_start(nodes, i);
                tot_pairs += np;
            }
            ipath = orc_dynamic_programming(nodes, nn, t, 1);
        }
        if (nn > 0 && ipath >= 0 && nodes[ipath].score > max_score) {
            max_phase = b;
            max_score = nodes[ipath].score;
            orc_eliminate_bad_genes(nodes, ipath, t);
            ng = orc_genes_extract(nodes, ipath, genes, gene_cap);
            if (ng > gene_cap) return -1;
            orc_tweak_final_starts(genes, ng, nodes, nn, t, o->max_overlap);
        }
    }
    if (max_phase >= 0) {
        const orc_training *t = &bins[max_phase];
        nn = orc_extract(digits, slen, t->trans_table, o, nodes, node_cap);
        orc_sort(nodes, nn);
        orc_reset_scores(nodes, nn);
        orc_score(digits, slen, nodes, nn, t, o->closed, 1);
    }
    if (nn_out) *nn_out = nn;
    if (winner) *winner = max_phase;
    if (pairs) *pairs = tot_pairs;
    return ng;
}

/* ------------------------------------------------------------------------------------ */
/* training (GeneFinder.train, lib.pyx:5236-5279)                                        */
/* ------------------------------------------------------------------------------------ */

/* vendor/Prodigal/sequence.c:559-564: index of the strictly largest value, later index on ties */
static int frame_of_max(int a, int b, int c) {
    if (a > b) return a > c ? 0 : 2;
    return b > c ? 1 : 2;
}

/* lib.pyx:724-768.  tot[i] = number of GC bases among i-57, i-54, .., i+57 that lie inside the
 * sequence (the reference builds it from two running sums 60 positions apart); every codon
 * triplet then gets the frame with the largest count, the tail that has no full triplet stays -1. */
void orc_gc_frame_plot(const uint8_t *d, int slen, int8_t *gp) {
    for (int i = 0; i < slen; i++) gp[i] = -1;
    for (int i = 0; i + 2 < slen; i += 3) {
        int tot[3];
        for (int f = 0; f < 3; f++) {
            int c = 0;
            for (int k = -19; k <= 19; k++) {
                int x = i + f + 3 * k;
                if (x >= 0 && x < slen) c += is_gc_fwd(d, x);
            }
            tot[f] = c;
        }
        int w = frame_of_max(tot[0], tot[1], tot[2]);
        gp[i] = gp[i + 1] = gp[i + 2] = (int8_t)w;
    }
}

/* vendor/Prodigal/node.c:263-317.  Each ORF is walked from its stop towards its starts, counting in
 * which codon position the GC-richest frame of every codon falls; bias[] is an order-dependent
 * floating point sum over the start nodes in index order. */
void orc_record_gc_bias(const int8_t *gp, orc_node *nodes, int nn, orc_training *t) {
    if (nn == 0) return;
    int ctr[3][3] = {{0}}, last[3] = {0, 0, 0};
    for (int i = nn - 1; i >= 0; i--) {
        orc_node *x = &nodes[i];
        if (x->strand != 1) continue;
        int fr = x->ndx % 3, shift = 3 - fr;
        if (x->type == ORC_STOP) {
            ctr[fr][0] = ctr[fr][1] = ctr[fr][2] = 0;
            last[fr] = x->ndx;
            ctr[fr][(gp[x->ndx] + shift) % 3] = 1;
        } else {
            for (int j = last[fr] - 3; j >= x->ndx; j -= 3) ctr[fr][(gp[j] + shift) % 3]++;
            x->gc_bias = frame_of_max(ctr[fr][0], ctr[fr][1], ctr[fr][2]);
            for (int j = 0; j < 3; j++) {
                x->gc_score[j] = 3.0 * ctr[fr][j];
                x->gc_score[j] /= 1.0 * (x->stop_val - x->ndx + 3);
            }
            last[fr] = x->ndx;
        }
    }
    for (int i = 0; i < nn; i++) {
        orc_node *x = &nodes[i];
        if (x->strand != -1) continue;
        int fr = x->ndx % 3, shift = fr;
        if (x->type == ORC_STOP) {
            ctr[fr][0] = ctr[fr][1] = ctr[fr][2] = 0;
            last[fr] = x->ndx;
            ctr[fr][((3 - gp[x->ndx]) + shift) % 3] = 1;
        } else {
            for (int j = last[fr] + 3; j <= x->ndx; j += 3) ctr[fr][((3 - gp[j]) + shift) % 3]++;
            x->gc_bias = frame_of_max(ctr[fr][0], ctr[fr][1], ctr[fr][2]);
            for (int j = 0; j < 3; j++) {
                x->gc_score[j] = 3.0 * ctr[fr][j];
                x->gc_score[j] /= 1.0 * (x->ndx - x->stop_val + 3);
            }
            last[fr] = x->ndx;
        }
    }
    t->bias[0] = t->bias[1] = t->bias[2] = 0.0;
    for (int i = 0; i < nn; i++) {
        const orc_node *x = &nodes[i];
        if (x->type == ORC_STOP) continue;
        int len = abs(x->stop_val - x->ndx) + 1;
        t->bias[x->gc_bias] += (x->gc_score[x->gc_bias] * len) / 1000.0;
    }
    double tot = t->bias[0] + t->bias[1] + t->bias[2];
    for (int i = 0; i < 3; i++) t->bias[i] *= (3.0 / tot);
}

/* lib.pyx:4284-4358: log-odds of every 6-mer inside the genes of the training path against the
 * whole sequence (both strands), clamped to [-5, 5] */
void orc_calc_dicodon_gene(const uint8_t *d, int slen, const orc_node *nodes, int ipath, orc_training *t) {
    int counts[4096];
    double bg[4096];
    int glob = 0;
    memset(counts, 0, sizeof(counts));
    for (int i = 0; i < slen - 5; i++) {
        counts[mer_ndx(d, slen, i, 6, 1)]++;
        counts[mer_ndx(d, slen, i, 6, -1)]++;
        glob += 2;
    }
    for (int i = 0; i < 4096; i++) bg[i] = (double)counts[i] / (double)glob;

    glob = 0;
    memset(counts, 0, sizeof(counts));
    int in_gene = 0, left = -1, right = -1;
    for (int p = ipath; p != -1; p = nodes[p].traceb) {
        const orc_node *x = &nodes[p];
        if (x->strand == 1) {
            if (x->type == ORC_STOP) {
                in_gene = 1;
                right = x->ndx + 2;
            } else if (in_gene == 1) {
                left = x->ndx;
                for (int i = left; i < right - 5; i += 3) { counts[mer_ndx(d, slen, i, 6, 1)]++; glob++; }
                in_gene = 0;
            }
        } else {
            if (x->type != ORC_STOP) {
                in_gene = -1;
                left = slen - x->ndx - 1;
            } else if (in_gene == -1) {
                right = slen - x->ndx + 1;
                for (int i = left; i < right - 5; i += 3) { counts[mer_ndx(d, slen, i, 6, -1)]++; glob++; }
                in_gene = 0;
            }
        }
    }
    for (int i = 0; i < 4096; i++) {
        double prob = (double)counts[i] / (double)glob, v;
        if (prob == 0 && bg[i] != 0) v = -5.0;
        else if (bg[i] == 0) v = 0.0;
        else v = log(prob / bg[i]);
        if (v > 5.0) v = 5.0;
        else if (v < -5.0) v = -5.0;
        t->gene_dc[i] = v;
    }
}

/* lib.pyx:4360-4389: tally the bases at start-1, start-2, start-15 .. start-44 (strand oriented) */
static void count_upstream(const uint8_t *d, int slen, int pos, int strand, orc_training *t) {
    static const int off[2][2] = {{1, 3}, {15, 45}};
    int slot = 0;
    for (int g = 0; g < 2; g++)
        for (int j = off[g][0]; j < off[g][1]; j++, slot++) {
            if (strand == 1) {
                if (pos >= j) t->ups_comp[slot][d[pos - j] & 3] += 1.0;
            } else {
                static const uint8_t comp[7] = {dT, dC, dG, dA, dN, dN, dN};
                if (pos + j < slen) t->ups_comp[slot][comp[d[pos + j]] & 3] += 1.0;
            }
        }
}

/* shared tail of both start-training loops: counts -> clamped log-odds against the GC content
 * (lib.pyx:4561-4599 == 4793-4826) */
static void upstream_to_log(orc_training *t) {
    for (int i = 0; i < 32; i++) {
        double sum = 0.0;
        for (int j = 0; j < 4; j++) sum += t->ups_comp[i][j];
        if (sum == 0.0) {
            for (int j = 0; j < 4; j++) t->ups_comp[i][j] = 0.0;
            continue;
        }
        for (int j = 0; j < 4; j++) {
            int at = j == 0 || j == 3;
            double den;
            if (t->gc <= 0.1) den = at ? 0.90 : 0.10;
            else if (t->gc >= 0.9) den = at ? 0.10 : 0.90;
            else den = at ? 1.0 - t->gc : t->gc;
            double v = t->ups_comp[i][j] / sum;
            v = log(v * 2.0 / den);
            if (v > 4.0) v = 4.0;
            if (v < -4.0) v = -4.0;
            t->ups_comp[i][j] = v;
        }
    }
}

/* counts -> clamped log-odds of the start codon types (lib.pyx:4540-4557 == 4771-4788);
 * returns the number of genes that contributed */
static double type_weights_from_counts(double treal[3], const double tbg[3], orc_training *t) {
    double sum = treal[0] + treal[1] + treal[2];
    for (int j = 0; j < 3; j++) {
        if (sum == 0.0) { t->type_wt[j] = 0.0; continue; }
        treal[j] /= sum;
        double v = tbg[j] != 0 ? log(treal[j] / tbg[j]) : -4.0;
        if (v > 4.0) v = 4.0;
        else if (v < -4.0) v = -4.0;
        t->type_wt[j] = v;
    }
    return sum;
}

/* which of the two SD motif bins (exact / one mismatch) a start is credited with
 * (lib.pyx:4442-4449 and three more copies) */
static int preferred_rbs(const orc_node *x, const double *rbs_wt) {
    int a = x->rbs[0], b = x->rbs[1];
    if (rbs_wt[a] > rbs_wt[b] + 1.0 || b == 0) return a;
    if (rbs_wt[a] < rbs_wt[b] - 1.0 || a == 0) return b;
    return a > b ? a : b;
}

static void type_background(const orc_node *nodes, int nn, double tbg[3]) {
    tbg[0] = tbg[1] = tbg[2] = 0.0;
    for (int i = 0; i < nn; i++)
        if (nodes[i].type != ORC_STOP) tbg[nodes[i].type] += 1.0;
    double sum = tbg[0] + tbg[1] + tbg[2];
    for (int i = 0; i < 3; i++) tbg[i] /= sum;
}

/* lib.pyx:4391-4599 */
void orc_train_starts_sd(const uint8_t *d, int slen, orc_node *nodes, int nn, orc_training *t) {
    double tbg[3], sthresh = 35.0;
    const double wt = t->st_wt;
    memset(t->type_wt, 0, sizeof(t->type_wt));
    memset(t->rbs_wt, 0, sizeof(t->rbs_wt));
    memset(t->ups_comp, 0, sizeof(t->ups_comp));
    type_background(nodes, nn, tbg);

    for (int it = 0; it < 10; it++) {
        double rbg[28] = {0}, rreal[28] = {0}, treal[3] = {0};
        for (int j = 0; j < nn; j++)
            if (nodes[j].type != ORC_STOP && !nodes[j].edge) rbg[preferred_rbs(&nodes[j], t->rbs_wt)] += 1.0;
        double sum = 0.0;
        for (int j = 0; j < 28; j++) sum += rbg[j];
        for (int j = 0; j < 28; j++) rbg[j] /= sum;

        /* one sweep per strand: forward in index order, reverse against it; per frame the best
         * scoring start since the last STOP is credited when that STOP is reached */
        for (int strand = 1; strand >= -1; strand -= 2) {
            double best[3] = {0, 0, 0};
            int bndx[3] = {-1, -1, -1}, brbs[3] = {0, 0, 0}, btype[3] = {0, 0, 0};
            for (int q = 0; q < nn; q++) {
                int j = strand == 1 ? q : nn - 1 - q;
                const orc_node *x = &nodes[j];
                if ((x->type != ORC_STOP && x->edge) || x->strand != strand) continue;
                int ph = x->ndx % 3;
                if (x->type == ORC_STOP) {
                    if (best[ph] >= sthresh && nodes[bndx[ph]].ndx % 3 == ph) {
                        rreal[brbs[ph]] += 1.0;
                        treal[btype[ph]] += 1.0;
                        if (it == 9) count_upstream(d, slen, nodes[bndx[ph]].ndx, strand, t);
                    }
                    best[ph] = 0.0; bndx[ph] = -1; brbs[ph] = 0; btype[ph] = 0;
                } else {
                    int rb = preferred_rbs(x, t->rbs_wt);
                    double v = x->cscore + wt * t->rbs_wt[rb] + wt * t->type_wt[x->type];
                    if (v >= best[ph]) { best[ph] = v; bndx[ph] = j; btype[ph] = x->type; brbs[ph] = rb; }
                }
            }
        }

        sum = 0.0;
        for (int j = 0; j < 28; j++) sum += rreal[j];
        for (int j = 0; j < 28; j++) {
            if (sum == 0.0) { t->rbs_wt[j] = 0.0; continue; }
            rreal[j] /= sum;
            double v = rbg[j] != 0 ? log(rreal[j] / rbg[j]) : -4.0;
            if (v > 4.0) v = 4.0;
            else if (v < -4.0) v = -4.0;
            t->rbs_wt[j] = v;
        }
        sum = type_weights_from_counts(treal, tbg, t);
        if (sum * 2000.0 <= nn) sthresh /= 2.0;
    }
    upstream_to_log(t);
}

/* vendor/Prodigal/node.c:686-693 */
void orc_determine_sd_usage(orc_training *t) {
    const double *w = t->rbs_wt;
    t->uses_sd = 1;
    if (w[0] >= 0.0) t->uses_sd = 0;
    if (w[16] < 1.0 && w[13] < 1.0 && w[15] < 1.0 && (w[0] >= -0.5 || (w[22] < 2.0 && w[24] < 2.0 && w[27] < 2.0)))
        t->uses_sd = 0;
}

/* lib.pyx:4226-4282 */
static void update_motif_counts(double (*cnt)[4][4096], double *zero, const uint8_t *d, int slen, const orc_node *x,
                                int stage) {
    if (x->type == ORC_STOP || x->edge == 1) return;
    if (x->mot_len == 0) { *zero += 1.0; return; }
    int start = x->strand == 1 ? x->ndx : slen - 1 - x->ndx;
    if (stage == 0) {
        /* every window of every length, credited to all four spacer classes */
        for (int l = 3; l >= 0; l--)
            for (int j = start - 18 - l; j <= start - 6 - l; j++) {
                if (j < 0) continue;
                int mer = mer_ndx(d, slen, j, l + 3, x->strand);
                for (int k = 0; k < 4; k++) cnt[l][k][mer] += 1.0;
            }
    } else if (stage == 1) {
        /* the best motif and every shorter word inside it */
        cnt[x->mot_len - 3][x->mot_spacendx][x->mot_ndx] += 1.0;
        for (int l = 0; l < x->mot_len - 3; l++)
            for (int j = start - x->mot_spacer - x->mot_len; j <= start - x->mot_spacer - l - 3; j++) {
                if (j < 0) continue;
                int sp;
                if (j <= start - 16 - l) sp = 3;
                else if (j <= start - 14 - l) sp = 2;
                else if (j >= start - 7 - l) sp = 1;
                else sp = 0;
                cnt[l][sp][mer_ndx(d, slen, j, l + 3, x->strand)] += 1.0;
            }
    } else {
        cnt[x->mot_len - 3][x->mot_spacendx][x->mot_ndx] += 1.0;
    }
}

/* vendor/Prodigal/node.c:1307-1357: which motifs are frequent enough (or made of frequent words) */
static void coverage_map(double (*real)[4][4096], int (*good)[4][4096], double ng, int stage) {
    (void)stage;
    memset(good, 0, sizeof(int) * 4 * 4 * 4096);
    for (int s = 0; s < 4; s++)
        for (int m = 0; m < 64; m++)
            if (real[0][s][m] / ng >= 0.2)
                for (int k = 0; k < 4; k++) good[0][k][m] = 1;
    for (int s = 0; s < 4; s++)
        for (int m = 0; m < 256; m++)
            if (good[0][s][(m & 252) >> 2] && good[0][s][m & 63]) good[1][s][m] = 1;
    for (int s = 0; s < 4; s++)
        for (int m = 0; m < 1024; m++) {
            if (!good[0][s][(m & 1008) >> 4] || !good[0][s][(m & 252) >> 2] || !good[0][s][m & 63]) continue;
            good[2][s][m] = 1;
            /* the three variants with the middle base changed are allowed as "mismatch" motifs */
            int v = m;
            for (int a = 0; a <= 16; a += 16) {
                v ^= a;
                for (int b = 0; b <= 32; b += 32) {
                    v ^= b;
                    if (good[2][s][v] == 0) good[2][s][v] = 2;
                }
            }
        }
    for (int s = 0; s < 4; s++)
        for (int m = 0; m < 4096; m++) {
            int a = good[2][s][(m & 4092) >> 2], b = good[2][s][m & 1023];
            if (a == 0 || b == 0) continue;
            good[3][s][m] = (a == 1 && b == 1) ? 1 : 2;
        }
}

/* lib.pyx:4601-4826 */
void orc_train_starts_nonsd(const uint8_t *d, int slen, orc_node *nodes, int nn, orc_training *t) {
    double(*mbg)[4][4096] = malloc(sizeof(double) * 4 * 4 * 4096);
    double(*mreal)[4][4096] = malloc(sizeof(double) * 4 * 4 * 4096);
    int(*mgood)[4][4096] = malloc(sizeof(int) * 4 * 4 * 4096);
    double tbg[3], sthresh = 35.0;
    const double wt = t->st_wt;
    memset(t->ups_comp, 0, sizeof(t->ups_comp));
    memset(t->type_wt, 0, sizeof(t->type_wt));
    type_background(nodes, nn, tbg);
    /* the reference never initialises mgood before stage 2 is reached without stages 0/1 having
     * run; iterations 0..11 always fill it, so start from zero */
    memset(mgood, 0, sizeof(int) * 4 * 4 * 4096);

    for (int it = 0; it < 20; it++) {
        int stage = it < 4 ? 0 : (it < 12 ? 1 : 2);
        double zbg = 0.0, zreal = 0.0, treal[3] = {0}, ngenes = 0.0;
        memset(mbg, 0, sizeof(double) * 4 * 4 * 4096);
        memset(mreal, 0, sizeof(double) * 4 * 4 * 4096);
        for (int j = 0; j < nn; j++) {
            if (nodes[j].type == ORC_STOP || nodes[j].edge) continue;
            best_upstream_motif(d, slen, &nodes[j], t, stage);
            update_motif_counts(mbg, &zbg, d, slen, &nodes[j], stage);
        }
        double sum = 0.0;
        for (int a = 0; a < 4; a++) for (int b = 0; b < 4; b++) for (int c = 0; c < 4096; c++) sum += mbg[a][b][c];
        sum += zbg;
        for (int a = 0; a < 4; a++) for (int b = 0; b < 4; b++) for (int c = 0; c < 4096; c++) mbg[a][b][c] /= sum;
        zbg /= sum;

        for (int strand = 1; strand >= -1; strand -= 2) {
            double best[3] = {0, 0, 0};
            int bndx[3] = {-1, -1, -1};
            for (int q = 0; q < nn; q++) {
                int j = strand == 1 ? q : nn - 1 - q;
                const orc_node *x = &nodes[j];
                if ((x->type != ORC_STOP && x->edge) || x->strand != strand) continue;
                int fr = x->ndx % 3;
                if (x->type == ORC_STOP) {
                    if (best[fr] >= sthresh) {
                        ngenes += 1.0;
                        treal[nodes[bndx[fr]].type] += 1.0;
                        update_motif_counts(mreal, &zreal, d, slen, &nodes[bndx[fr]], stage);
                        if (it == 19) count_upstream(d, slen, nodes[bndx[fr]].ndx, strand, t);
                    }
                    best[fr] = 0.0; bndx[fr] = -1;
                } else {
                    double v = x->cscore + wt * x->mot_score + wt * t->type_wt[x->type];
                    if (v >= best[fr]) { best[fr] = v; bndx[fr] = j; }
                }
            }
        }

        if (stage < 2) coverage_map(mreal, mgood, ngenes, stage);
        sum = 0.0;
        for (int a = 0; a < 4; a++) for (int b = 0; b < 4; b++) for (int c = 0; c < 4096; c++) sum += mreal[a][b][c];
        sum += zreal;
        if (sum == 0.0) {
            memset(t->mot_wt, 0, sizeof(t->mot_wt));
            t->no_mot = 0.0;
        } else {
            for (int a = 0; a < 4; a++) for (int b = 0; b < 4; b++) for (int c = 0; c < 4096; c++) {
                if (mgood[a][b][c] == 0) {
                    zreal += mreal[a][b][c];
                    zbg += mreal[a][b][c];
                    mreal[a][b][c] = 0.0;
                    mbg[a][b][c] = 0.0;
                }
                mreal[a][b][c] /= sum;
                double v = mbg[a][b][c] != 0 ? log(mreal[a][b][c] / mbg[a][b][c]) : -4.0;
                if (v > 4.0) v = 4.0;
                else if (v < -4.0) v = -4.0;
                t->mot_wt[a][b][c] = v;
            }
        }
        /* no_mot is recomputed even when sum == 0 (0/0 = NaN compares false everywhere) */
        zreal /= sum;
        {
            double v = zbg != 0 ? log(zreal / zbg) : -4.0;
            if (v > 4.0) v = 4.0;
            else if (v < -4.0) v = -4.0;
            t->no_mot = v;
        }
        sum = type_weights_from_counts(treal, tbg, t);
        if (sum * 2000.0 <= nn) sthresh /= 2.0;
    }
    upstream_to_log(t);
    free(mbg); free(mreal); free(mgood);
}

/* lib.pyx:5236-5279 (+ the TrainingInfo constructor, 3960-4003).  `nodes` is scratch of node_cap
 * entries; returns the node count or -1 when node_cap is too small. */
int orc_train(const uint8_t *digits, int slen, double gc, int tt, double st_wt, int force_nonsd, const orc_opts *o,
              orc_node *nodes, int node_cap, orc_training *t) {
    memset(t, 0, sizeof(*t));
    t->gc = gc; t->trans_table = tt; t->st_wt = st_wt; t->uses_sd = 1;
    int nn = orc_extract(digits, slen, tt, o, nodes, node_cap);
    if (nn < 0) return -1;
    orc_sort(nodes, nn);
    int8_t *gp = malloc(slen > 0 ? slen : 1);
    orc_gc_frame_plot(digits, slen, gp);
    orc_record_gc_bias(gp, nodes, nn, t);
    free(gp);
    orc_record_overlapping_starts(nodes, nn, t, 0, o->max_overlap);
    int ipath = orc_dynamic_programming(nodes, nn, t, 0);
    orc_calc_dicodon_gene(digits, slen, nodes, ipath, t);
    orc_raw_coding_score(digits, slen, nodes, nn, t);
    orc_rbs_score(digits, slen, nodes, nn, t);
    orc_train_starts_sd(digits, slen, nodes, nn, t);
    if (force_nonsd) t->uses_sd = 0;
    else orc_determine_sd_usage(t);
    if (!t->uses_sd) orc_train_starts_nonsd(digits, slen, nodes, nn, t);
    return nn;
}

/* layout probes for the ctypes/numpy binding in oracle/oracle.py */
int orc_sizeof_node(void) { return (int)sizeof(orc_node); }
int orc_sizeof_training(void) { return (int)sizeof(orc_training); }
