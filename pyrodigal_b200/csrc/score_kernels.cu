// score_kernels.cu -- per-node scoring (score_nodes family) and overlapping-start bookkeeping.
//
// Reference semantics: Nodes._calc_orf_gc (lib.pyx:1846-1896), _raw_coding_score (2119-2239),
// _rbs_score + Sequence._shine_dalgarno_exact/_mm (2241-2277, 791-979), Node._find_best_upstream_motif
// (1557-1616), _score_upstream_composition (1619-1650), Nodes._score (2331-2487),
// Nodes._record_overlapping_starts (2279-2329), BaseConnectionScorer._index / window logic (1126-1162,
// 1224-1233).
//
// B200 design notes
//  * every ORF (stop-to-stop segment of one strand/frame) is an independent unit: the reference's three
//    sequential sweeps over all nodes reset their state at every STOP node, so one thread per STOP node
//    reproduces the exact summation order (bit-identical doubles) while exposing 10^4..10^6-way
//    parallelism per batch.
//  * dicodon indices slide over the 1-byte codon-code array (one load per codon).
//  * the Shine-Dalgarno search is table driven: the motif found for a 6-base window depends only on the
//    window's A/G match pattern and its distance to the start, so a per-model 15x64 byte table (built on
//    the host) replaces the nested loops; the per-node 32-bit A/G pattern is model independent.
//  * log/pow of the length factor come from a host table (host libm == the reference's libm).
#include <cstdlib>

#include "kernels.cuh"
#include "sd_device.cuh"

namespace pgpu {

__device__ __forceinline__ int find_chain(const ChainInfo *__restrict__ ch, int n, int64_t g) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (ch[mid].coff <= g) lo = mid; else hi = mid - 1;
    }
    return lo;
}
__device__ __forceinline__ int find_ext(const ExtractInfo *__restrict__ ex, int n, int g) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (ex[mid].node_off <= g) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// First candidate owner of flat index g: from the block table when the batch has one (one broadcast load), else by
// binary search shared through the CTA.  Callers advance linearly from it (`while (next.off <= g) k++`).
__device__ __forceinline__ int chain_hint(const DevBatch &B, int n_chains, int64_t g, int64_t total, int *s_first,
                                          int64_t block_first = -1) {
    if (B.blk_chain) return B.blk_chain[g >> 7];
    if (block_first < 0) block_first = (int64_t)blockIdx.x * blockDim.x;
    if (threadIdx.x == 0) *s_first = find_chain(B.chains, n_chains, min(block_first, total - 1));
    __syncthreads();
    return *s_first;
}
__device__ __forceinline__ int ext_hint(const DevBatch &B, int n_ext, int g, int block_first, int total_nodes, int *s_first) {
    if (B.blk_ext) return B.blk_ext[g >> 7];
    if (threadIdx.x == 0) *s_first = find_ext(B.exts, n_ext, min(block_first, total_nodes - 1));
    __syncthreads();
    return *s_first;
}

// GC count of the three bases starting at p, from the codon code (N counts as GC: _sequence.h:35-43;
// padding beyond the sequence end is 'A').
__device__ __forceinline__ int gc3(int code) {
    int b0 = code & 3, b1 = (code >> 2) & 3, b2 = (code >> 4) & 3;
    return ((b0 ^ (b0 >> 1)) & 1) + ((b1 ^ (b1 >> 1)) & 1) + ((b2 ^ (b2 >> 1)) & 1);
}

// strand-oriented 2-bit base for k-mer indices (_sequence.h:207-220): N -> C on both strands
__device__ __forceinline__ int mer_base(const uint8_t *__restrict__ d, int slen, int x, bool rev) {
    if (!rev) return d[x] & 3;
    int b = d[slen - 1 - x];
    return b == 6 ? 2 : (b ^ 3);
}

// code of the reverse-strand codon whose 5' base is at forward position p (bases p, p-1, p-2)
__device__ __noinline__ int rcode_slow(const uint8_t *__restrict__ d, int p) {
    // a codon with unknown bases: _mer_ndx complements N to N and then keeps its low 2 bits (= C)
    int r = 0;
    for (int k = 0; k < 3; k++) {
        int b = d[p - k];
        r |= (b == 6 ? 2 : (b ^ 3)) << (2 * k);
    }
    return r;
}
__device__ __forceinline__ int rcode_at(const uint8_t *__restrict__ d, const uint8_t *__restrict__ cod, int p) {
    const int c = cod[p - 2];
    if (__builtin_expect((c & 64) != 0, 0)) return rcode_slow(d, p);  // rare: keep it out of the hot loop
    return rev_code(c & 63);
}

// --------------------------------------------------------------------------------------------------
// per-extraction preparation: gc_cont, SD match bits, DP window start
// --------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_node_prep(DevBatch B, int n_ext, int total_nodes, int seq_parts) {
    __shared__ int s_first;
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    int e = ext_hint(B, n_ext, min(g, total_nodes - 1), blockIdx.x * blockDim.x, total_nodes, &s_first);
    if (g >= total_nodes) return;
    while (e + 1 < n_ext && B.exts[e + 1].node_off <= g) e++;
    const ExtractInfo X = B.exts[e];
    const int z = g - X.node_off, nn = X.nn, slen = X.slen;
    const int32_t *__restrict__ ndx = B.ndx + X.node_off;
    const int32_t *__restrict__ sv = B.stop_val + X.node_off;
    const uint8_t *__restrict__ cls = B.cls + X.node_off;
    const uint8_t *__restrict__ d = B.digits + X.doff;
    const uint8_t *__restrict__ cod = B.cod + X.doff;
    const int c = cls[z], kind = cls_kind(c), f = cls_frame(c), my = ndx[z];

    // ---- DP window start (lib.pyx:1224-1233) ----
    {
        int m = z < kMaxNodeDist ? 0 : z - kMaxNodeDist;
        if ((kind == K_RS || kind == K_FE) && ndx[m] > sv[z]) {
            // highest index whose ndx == stop_val, or 0 (binary search instead of the reference's walk)
            const int target = sv[z];
            int lo = 0, hi = m;  // first index with ndx > target in [0, m]
            while (lo < hi) {
                int mid = (lo + hi) >> 1;
                if (ndx[mid] > target) hi = mid; else lo = mid + 1;
            }
            m = (lo > 0 && ndx[lo - 1] == target) ? lo - 1 : 0;
        }
        B.win_min[X.node_off + z] = m < kMaxNodeDist ? 0 : m - kMaxNodeDist;
    }
    if (!seq_parts) return;  // operator-level DP: nodes supplied by the caller, no sequence

    if (kind == K_FS || kind == K_RS) {
        // ---- upstream A/G pattern for the SD search: bit k = A at start-20+k, bit 16+k = G ----
        const bool rev = kind == K_RS;
        const int start = rev ? slen - 1 - my : my;
        uint32_t bits = 0;
#pragma unroll 4
        for (int k = 0; k < 16; k++) {
            int x = start - 20 + k;
            if (x < 0 || x >= slen) continue;
            int b = rev ? d[slen - 1 - x] : d[x];
            if (b == (rev ? 3 : 0)) bits |= 1u << k;
            if (b == (rev ? 2 : 1)) bits |= 1u << (16 + k);
        }
        B.sdbits[X.node_off + z] = bits;
        // packed upstream bases (model independent): composition positions in the reference's loop order
        // (lib.pyx:1638-1649: q = 1, 2 then 15..44, stopping at the first q > start), and the motif window
        uint64_t pc = 0;
        int cnt = 0;
        for (int q = 1; q < 3 && q <= start; q++, cnt++) pc |= (uint64_t)mer_base(d, slen, start - q, rev) << (2 * cnt);
        for (int q = 15; q < 45 && q <= start; q++, cnt++) pc |= (uint64_t)mer_base(d, slen, start - q, rev) << (2 * cnt);
        B.upc[X.node_off + z] = pc;
        uint64_t U = 0;
        for (int q = 0; q < 18; q++) {
            const int x = start - 21 + q;
            if (x >= 0 && x < slen) U |= (uint64_t)mer_base(d, slen, x, rev) << (2 * q);
        }
        B.umot[X.node_off + z] = U;
        // gc_cont (lib.pyx:1846-1896): the reference accumulates GC counts of codon triplets from the stop to
        // the start; that is the GC count of a contiguous range, taken here from the GC bitmap + word prefix:
        //   forward: [ndx, stop+3);  reverse: the stop's bases [stop-2, stop] plus [stop+3, ndx+3) (T14: the
        //   reverse window is shifted by two bases), positions beyond the sequence end count 0
        const uint32_t *__restrict__ gb = B.gcbits + (X.doff >> 5);
        const int32_t *__restrict__ gp = B.gcpre + (X.doff >> 5);
        auto P = [&](int x) {  // number of GC bits in [0, x)
            x = max(0, min(x, slen));
            return gp[x >> 5] + __popc(gb[x >> 5] & ((1u << (x & 31)) - 1u));
        };
        const int stop = sv[z];
        const int gc = rev ? (P(stop + 1) - P(stop - 2)) + (P(my + 3) - P(stop + 3)) : P(stop + 3) - P(my);
        B.gc_cont[X.node_off + z] = (float)((double)gc / (abs(stop - my) + 3.0));
        return;
    }

    // STOP nodes: nothing else to prepare
}

// per-class ranks and class-sorted index lists (SoA replacement of ConnectionScorer._index,
// lib.pyx:1126-1162): one warp per extraction
__global__ void __launch_bounds__(128) k_class_index(DevBatch B, int n_ext) {
    const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (e >= n_ext) return;
    const ExtractInfo X = B.exts[e];
    const uint8_t *__restrict__ cls = B.cls + X.node_off;
    int32_t *crank = B.crank + 4 * (int64_t)X.node_off;
    int32_t *clist = B.clist + X.node_off;
    const uint32_t lt = (1u << lane) - 1u;
    int tot[4] = {0, 0, 0, 0};
    for (int base = 0; base < X.nn; base += 32) {
        int i = base + lane;
        int k = i < X.nn ? cls_kind(cls[i]) : -1;
#pragma unroll
        for (int c = 0; c < 4; c++) tot[c] += __popc(__ballot_sync(0xffffffffu, k == c));
    }
    int cb[4] = {0, tot[0], tot[0] + tot[1], tot[0] + tot[1] + tot[2]};
    if (lane < 4) B.cbase[4 * e + lane] = cb[lane];
    if (lane == 0) {
        // edge-flagged nodes sit within 5 bp of either sequence end: count the nodes there (ndx is sorted)
        const int32_t *__restrict__ ndx = B.ndx + X.node_off;
        int lo = 0, hi = X.nn;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (ndx[mid] <= 4) lo = mid + 1; else hi = mid; }
        B.exts[e].n_lo = lo;
        lo = 0; hi = X.nn;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (ndx[mid] < X.slen - 5) lo = mid + 1; else hi = mid; }
        B.exts[e].n_hi = X.nn - lo;
    }
    int run[4] = {0, 0, 0, 0};
    for (int base = 0; base < X.nn; base += 32) {
        int i = base + lane;
        int k = i < X.nn ? cls_kind(cls[i]) : -1;
        int r[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            uint32_t m = __ballot_sync(0xffffffffu, k == c);
            r[c] = run[c] + __popc(m & lt);
            run[c] += __popc(m);
        }
        if (i < X.nn) {
            *reinterpret_cast<int4 *>(crank + 4 * (int64_t)i) = make_int4(r[0], r[1], r[2], r[3]);
            clist[cb[k] + r[k]] = i;
        }
    }
    // sentinel rank row (index nn) is not needed: the DP reads crank[i] for i < nn only
}

// --------------------------------------------------------------------------------------------------
// raw coding score: one thread per STOP node (= per ORF)
// --------------------------------------------------------------------------------------------------
// the ORF that ends at STOP node z of chain C (all three sweeps), by one thread
__device__ __forceinline__ void coding_orf_thread(const DevBatch &B, const DevModel *__restrict__ models, const ChainInfo &C,
                                                  int z) {
    const uint8_t *__restrict__ cls = B.cls + C.node_off;
    const int c = cls[z];
    const DevModel &M = models[C.model];
    const double *__restrict__ dc = M.gene_dc;
    const int32_t *__restrict__ ndx = B.ndx + C.node_off;
    const int32_t *__restrict__ sv = B.stop_val + C.node_off;
    const uint16_t *__restrict__ dicf = B.dic_f + C.doff;
    const uint16_t *__restrict__ dicr = B.dic_r + C.doff;
    double *__restrict__ cscore = B.cscore + C.coff;
    const int f = cls_frame(c), my = ndx[z], nn = C.nn;
    const bool rev = c & CLS_REV;

    // sweep A: dicodon log-odds accumulated from the stop towards each start (lib.pyx:2149-2173); the 6-mer index of
    // every position is precomputed (k_dicodon_index, frame planes: a walk reads consecutive elements downwards on both
    // strands), so a codon costs one 2-byte load and one weight load
    int far = -1, last = my;
    double acc = 0.0;
    const int P = dic_plane(C.slen);
    const uint16_t *__restrict__ pl = (rev ? dicr : dicf) + (my % 3) * P;
    for (int i = rev ? z + 1 : z - 1; rev ? i < nn : i >= 0; i += rev ? 1 : -1) {
        int ci = cls[i];
        if (((ci & CLS_REV) != 0) != rev || cls_frame(ci) != f) continue;
        if (cls_is_stop(ci)) break;
        const int ni = ndx[i];
        const uint16_t *__restrict__ q = pl + (rev ? P - 2 - last / 3 : last / 3 - 1);
        int rem = rev ? (ni - last) / 3 : (last - ni) / 3;
        for (; rem >= 4; rem -= 4, q -= 4) {   // four independent loads in flight, adds in the reference's order
            const double w0 = __ldg(&dc[q[0]]), w1 = __ldg(&dc[q[-1]]), w2 = __ldg(&dc[q[-2]]), w3 = __ldg(&dc[q[-3]]);
            acc += w0; acc += w1; acc += w2; acc += w3;
        }
        for (; rem > 0; rem--, q--) acc += __ldg(&dc[q[0]]);
        cscore[i] = acc;
        last = ni;
        far = i;
    }
    if (far < 0) return;

    // sweep B: the two penalty passes fused, walking back towards the stop (lib.pyx:2175-2236)
    double s2 = -10000.0, s3 = -10000.0;
    const int step = rev ? -1 : 1;
    for (int i = far; i != z; i += step) {
        int ci = cls[i];
        if (((ci & CLS_REV) != 0) != rev || cls_frame(ci) != f) continue;
        double cs = cscore[i];
        if (cs > s2) s2 = cs; else cs -= (s2 - cs);
        const double gsize = rev ? (((double)ndx[i] - sv[i]) + 3.0) / 3.0 : (((double)sv[i] - ndx[i]) + 3.0) / 3.0;
        double lfac;
        if (gsize > 1000.0) lfac = M.lfac_span * (gsize - 80) / 920.0;
        else lfac = M.lfac[(int)gsize];
        if (lfac > s3) s3 = lfac; else lfac -= fmax(fmin(s3 - lfac, lfac), 0.0);
        if (lfac > 3.0 && cs < 0.5 * lfac) cs = 0.5 * lfac;
        cs += lfac;
        cscore[i] = cs;
    }
}

__global__ void __launch_bounds__(128) k_coding(DevBatch B, const DevModel *__restrict__ models, int n_chains,
                                                 int64_t total) {
    __shared__ int s_first;
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int k = chain_hint(B, n_chains, min(g, total - 1), total, &s_first);
    if (g >= total) return;
    while (k + 1 < n_chains && B.chains[k + 1].coff <= g) k++;
    const ChainInfo C = B.chains[k];
    // the first (#STOP nodes) threads of a chain's index range take one STOP node each, through the
    // class-sorted index list, so that warps are either fully busy or exit at once
    const int t = (int)(g - C.coff);
    if (t >= C.nn) return;
    const int32_t *__restrict__ cbase = B.cbase + 4 * C.ext;
    const int n_fe = cbase[2] - cbase[1], n_re = C.nn - cbase[3];
    if (t >= n_fe + n_re) return;
    coding_orf_thread(B, models, C, (B.clist + C.node_off)[t < n_fe ? cbase[1] + t : cbase[3] + (t - n_fe)]);
}

// --------------------------------------------------------------------------------------------------
// raw coding score, warp per ORF: all models that share an extraction walk the same codons, so the lanes
// of a warp take one model each -- identical control flow (no divergence), codon loads are warp-uniform
// broadcasts and the per-model dicodon weights of one index sit next to each other in the transposed
// table.  Summation order per (ORF, model) is unchanged, so results stay bit-identical.
// --------------------------------------------------------------------------------------------------
constexpr int kOrfWarps = 8;

// The ORF that ends at STOP node z (position my, frame f, strand rev) for the chain of this lane: all three sweeps.
// Wt(index) = the dicodon weight of this lane's model; a lane without a chain (active == false) walks along.
template <class WF>
__device__ __forceinline__ void coding_orf_lane(const uint8_t *__restrict__ cls, const int32_t *__restrict__ ndx,
                                                const int32_t *__restrict__ sv, const uint16_t *__restrict__ dicf,
                                                const uint16_t *__restrict__ dicr, int P, int nn, int z, int f, int my, bool rev,
                                                bool active, double *__restrict__ cscore, const DevModel &M, WF Wt) {
    // sweep A: dicodon log-odds accumulated from the stop towards each start (lib.pyx:2149-2173).  The chain
    // (index load -> weight load -> add) is latency bound: eight codons per round while at least eight remain
    // (their loads are independent and in flight together), then 4 / 2 / 1; adds keep the reference's order.
    // The indices of a walk are consecutive elements of one frame plane, downwards on both strands (DevBatch::dic_f).
    int far = -1, last = my;
    double acc = 0.0;
    const uint16_t *__restrict__ pl = (rev ? dicr : dicf) + (my % 3) * P;
    for (int i = rev ? z + 1 : z - 1; rev ? i < nn : i >= 0; i += rev ? 1 : -1) {
        const int ci = cls[i];
        if (((ci & CLS_REV) != 0) != rev || cls_frame(ci) != f) continue;
        if (cls_is_stop(ci)) break;
        const int ni = ndx[i];
        const uint16_t *__restrict__ q = pl + (rev ? P - 2 - last / 3 : last / 3 - 1);
        int rem = rev ? (ni - last) / 3 : (last - ni) / 3;
#pragma unroll 1
        for (; rem >= 8; rem -= 8, q -= 8) {
            const uint32_t i0 = q[0], i1 = q[-1], i2 = q[-2], i3 = q[-3], i4 = q[-4], i5 = q[-5], i6 = q[-6], i7 = q[-7];
            const double w0 = Wt(i0), w1 = Wt(i1), w2 = Wt(i2), w3 = Wt(i3), w4 = Wt(i4), w5 = Wt(i5), w6 = Wt(i6), w7 = Wt(i7);
            acc += w0; acc += w1; acc += w2; acc += w3; acc += w4; acc += w5; acc += w6; acc += w7;
        }
        if (rem & 4) {
            const uint32_t i0 = q[0], i1 = q[-1], i2 = q[-2], i3 = q[-3];
            const double w0 = Wt(i0), w1 = Wt(i1), w2 = Wt(i2), w3 = Wt(i3);
            acc += w0; acc += w1; acc += w2; acc += w3;
            q -= 4;
        }
        if (rem & 2) {
            const uint32_t i0 = q[0], i1 = q[-1];
            const double w0 = Wt(i0), w1 = Wt(i1);
            acc += w0; acc += w1;
            q -= 2;
        }
        if (rem & 1) acc += Wt(q[0]);
        if (active) cscore[i] = acc;
        last = ni;
        far = i;
    }
    if (far < 0) return;

    // sweep B: the two penalty passes fused, walking back towards the stop (lib.pyx:2175-2236)
    double s2 = -10000.0, s3 = -10000.0;
    const int step = rev ? -1 : 1;
    for (int i = far; i != z; i += step) {
        const int ci = cls[i];
        if (((ci & CLS_REV) != 0) != rev || cls_frame(ci) != f) continue;
        double cs = active ? cscore[i] : 0.0;
        if (cs > s2) s2 = cs; else cs -= (s2 - cs);
        const double gsize = rev ? (((double)ndx[i] - sv[i]) + 3.0) / 3.0 : (((double)sv[i] - ndx[i]) + 3.0) / 3.0;
        double lfac;
        if (gsize > 1000.0) lfac = M.lfac_span * (gsize - 80) / 920.0;
        else lfac = M.lfac[(int)gsize];
        if (lfac > s3) s3 = lfac; else lfac -= fmax(fmin(s3 - lfac, lfac), 0.0);
        if (lfac > 3.0 && cs < 0.5 * lfac) cs = 0.5 * lfac;
        cs += lfac;
        if (active) cscore[i] = cs;
    }
}

// GROUPED: an extraction with L <= 16 models does not need a whole warp per ORF.  Its ORFs are handled by groups of
// W = 4 / 8 / 16 lanes (the next power of two >= L), 32 / W ORFs per warp: the lanes of a group still take one model
// each and share every node / codon load, different groups of a warp simply diverge (there is no warp-level
// primitive in this kernel), and a warp costs the longest of its ORFs instead of their sum.  About two thirds of the
// (extraction, ORF) items of a metagenome batch are in extractions with L <= 16 (a third with L <= 4: the
// translation-table-4 extractions of low-GC contigs).  Thread ranges per extraction (`orf_toff`, multiples of 32) and
// group widths (`orf_w`) are planned on the host.  MINB: min CTAs per SM, 5 = 46 registers / 40 warps, 6 = 40 registers.
template <int MINB, bool GROUPED>
__global__ void __launch_bounds__(32 * kOrfWarps, MINB) k_coding_orf(DevBatch B, const DevModel *__restrict__ models, int n_ext,
                                                                int total_nodes) {
    __shared__ int s_first;
    int e, tl, lane, W;
    if (GROUPED) {
        const int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (gt >= B.orf_toff[n_ext]) return;
        int r = B.orf_blk[gt >> 8];   // position in the planned order of the extractions (orf_ext: grouped by model set)
        while (r + 1 < n_ext && B.orf_toff[r + 1] <= gt) r++;
        W = B.orf_w[r];
        const int local = (int)(gt - B.orf_toff[r]);
        e = B.orf_ext[r];
        tl = local >> (31 - __clz(W));   // W is a power of two
        lane = local & (W - 1);
    } else {
        // Work items are the STOP nodes.  A STOP node exists only if its ORF has a start, so an extraction with nn
        // nodes has at most nn / 2 of them: the warps index a half-size slot space in which extraction e owns the
        // slots from (node_off + 1) / 2 on, and only its first (#STOP nodes) slots have work.
        const int h = blockIdx.x * kOrfWarps + (threadIdx.x >> 5);
        e = ext_hint(B, n_ext, min(2 * h, total_nodes - 1), 2 * blockIdx.x * kOrfWarps, total_nodes, &s_first);
        if (h > (total_nodes + 1) / 2) return;
        while (e + 1 < n_ext && ((B.exts[e + 1].node_off + 1) >> 1) <= h) e++;
        tl = h - ((B.exts[e].node_off + 1) >> 1);
        lane = threadIdx.x & 31;
        W = 32;
    }
    const ExtractInfo X = B.exts[e];
    const int32_t *__restrict__ cbase = B.cbase + 4 * e;
    const int nn = X.nn;
    const int n_fe = cbase[2] - cbase[1], n_re = nn - cbase[3];
    if (tl >= n_fe + n_re) return;
    const int z = (B.clist + X.node_off)[tl < n_fe ? cbase[1] + tl : cbase[3] + (tl - n_fe)];
    const uint8_t *__restrict__ cls = B.cls + X.node_off;
    const int32_t *__restrict__ ndx = B.ndx + X.node_off;
    const int32_t *__restrict__ sv = B.stop_val + X.node_off;
    const uint16_t *__restrict__ dicf = B.dic_f + X.doff;
    const uint16_t *__restrict__ dicr = B.dic_r + X.doff;
    const int c = cls[z], f = cls_frame(c), my = ndx[z];
    const bool rev = c & CLS_REV;
    const int ch0 = B.ext_chain_off[e], nch = B.ext_chain_off[e + 1] - ch0;

    for (int c0 = 0; c0 < nch; c0 += W) {
        const bool active = c0 + lane < nch;
        const int chain = active ? B.ext_chains[ch0 + c0 + lane] : 0;
        const ChainInfo C = B.chains[active ? chain : B.ext_chains[ch0]];
        const DevModel &M = models[C.model];
        // this lane's column of the transposed table; a weight address is one 32x32->64 multiply-add from it
        const char *__restrict__ wcol = (const char *)(B.dcT + M.col);
        coding_orf_lane(cls, ndx, sv, dicf, dicr, dic_plane(X.slen), nn, z, f, my, rev, active, B.cscore + C.coff, M, [&](uint32_t index) -> double {
            return *(const double *)(wcol + (uint64_t)index * (uint64_t)(kDcCols * sizeof(double)));
        });
    }
}

// --------------------------------------------------------------------------------------------------
// raw coding score with the dicodon tables in shared memory: k_orf_links + k_cq_plan + k_coding_flat.
// k_coding_orf gathers its weights from the 2 MB transposed table through L1 / L2 (ncu: 70 % of the stall samples
// long-scoreboard, L1 hit rate 54 %).  Here a CTA owns the weights of kCqCols neighbouring table columns -- the models are
// sorted by (translation table, GC), the chains of an extraction are a contiguous column range, cut into groups of up to
// four -- as one 128 KB table dcS[set][index][4], fetched once with 1-D bulk copies (TMA, cp.async.bulk + mbarrier) and
// then read with LDS.64.  Work items are (extraction, column group, STOP node) = "ORF slots": the host plans the entries
// (api.cu: sorted by class = (table set, lanes per ORF), so that a CTA has ONE table set and ONE group width W = 1 / 2 /
// 4), the device counts the STOP nodes and lays the slots out without gaps (k_cq_plan), classes padded to whole CTA spans.
//
// With four models per table a warp has to walk 8 ... 32 ORFs at once, and ORFs differ in length by orders of magnitude.
// Measured on the way (630 Mbp bench shard, k_coding_orf = 19.5 ms): every group of W lanes running the nested loops of
// coding_orf_lane on its own ORF: 5.7 active lanes per LDS instruction, 21.4 ms; one warp-uniform loop with the node scan
// replaced by ORF links, but six dependent loads to set up an ORF and the penalty sweeps in a kernel of their own: 12.7 +
// 9.0 + 1.6 ms.  k_coding_flat therefore
//   * runs ONE warp-uniform loop: per round every group adds up to eight codons of its current walk (predicated); the
//     transitions -- a start reached: keep the sum, continue with the next start of the ORF; ORF finished: penalty
//     sweeps, next slot from a shared-memory counter -- are short divergent blocks;
//   * reads the starts of an ORF as a linked list from its STOP node (k_orf_links, once per extraction instead of a node
//     scan per model) and everything an ORF needs from ONE 16-byte descriptor, requested one ORF ahead;
//   * reads the indices of a walk as consecutive elements of a frame plane (DevBatch::dic_f): eight codons = one aligned
//     16-byte load, requested one chunk ahead;
//   * keeps the sums of the first eight starts of the ORF in shared memory for the penalty sweeps (lib.pyx:2175-2236),
//     which run back from the last start at the end of the ORF: the raw sums never go to global memory.
// Per (ORF, model) the additions and their order are those of the reference.
// --------------------------------------------------------------------------------------------------
constexpr int kCqThreads = 1024;
constexpr int kCqCols = 4;
constexpr int kCqRing = 8;
constexpr int kCqTableBytes = 4096 * kCqCols * (int)sizeof(double);                 // 128 KB
constexpr int kCqSmemBytes = kCqTableBytes + kCqRing * kCqThreads * (int)sizeof(double);   // + 64 KB

// one thread per STOP node: the in-frame starts of its ORF, nearest first (DevBatch::link, ::orfd).  Threads index the
// half-size slot space of the descriptors (extraction e owns the slots from (node_off + 1) / 2 on, its first
// (#STOP nodes) slots have work: a STOP node exists only if its ORF has a start), so busy lanes are neighbours.
__global__ void __launch_bounds__(128) k_orf_links(DevBatch B, int n_ext, int total_nodes, int roff) {
    __shared__ int s_first;
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    int e = ext_hint(B, n_ext, min(2 * h, total_nodes - 1), 2 * blockIdx.x * blockDim.x, total_nodes, &s_first);
    if (h > (total_nodes + 1) / 2) return;
    while (e + 1 < n_ext && ((B.exts[e + 1].node_off + 1) >> 1) <= h) e++;
    const ExtractInfo *__restrict__ X = B.exts + e;
    const int node_off = X->node_off, nn = X->nn;
    const int32_t *__restrict__ cbase = B.cbase + 4 * e;
    const int tl = h - ((node_off + 1) >> 1), n_fe = cbase[2] - cbase[1], n_re = nn - cbase[3];
    if (tl >= n_fe + n_re) return;
    const int z = (B.clist + node_off)[tl < n_fe ? cbase[1] + tl : cbase[3] + (tl - n_fe)];
    const int g = node_off + z;
    const uint8_t *__restrict__ cls = B.cls + node_off;
    const int c = cls[z];
    if (!cls_is_stop(c)) return;
    const int32_t *__restrict__ ndx = B.ndx + node_off;
    const int32_t *__restrict__ sv = B.stop_val + node_off;
    const bool rev = (c & CLS_REV) != 0;
    const int f = cls_frame(c), P = dic_plane(X->slen), my = ndx[z];
    const int plane = (rev ? roff : 0) + (int)X->doff + (my % 3) * P;
    int4 d = make_int4(plane + (rev ? P - 2 - my / 3 : my / 3 - 1), -1, 0, g);
    int prev = -1, pprev = g, pdiff = 0;   // the last start seen, the node before it, its distance to the stop
    for (int i = rev ? z + 1 : z - 1; rev ? i < nn : i >= 0; i += rev ? 1 : -1) {
        const int ci = cls[i];
        if (((ci & CLS_REV) != 0) != rev || cls_frame(ci) != f) continue;
        if (cls_is_stop(ci)) break;
        const int lo = plane + (rev ? P - 1 - ndx[i] / 3 : ndx[i] / 3);
        if (prev < 0) { d.y = node_off + i; d.z = lo; }
        else { B.link[prev] = make_int4(node_off + i, lo, pprev, pdiff); pprev = prev; }
        prev = node_off + i;
        pdiff = rev ? ndx[i] - sv[i] : sv[i] - ndx[i];
    }
    if (prev >= 0) B.link[prev] = make_int4(-1, 0, pprev, pdiff);
    B.orfd[h] = d;
}

// ORF slots of the plan entries (one CTA): entry r gets as many slots as its extraction has STOP nodes, the padding
// entry at the end of a class fills the class up to a whole number of CTA spans
__global__ void __launch_bounds__(1024) k_cq_plan(DevBatch B) {
    __shared__ int s_sum[256], s_base[256], s_gfirst[256], s_warp[32], s_carry;
    const int n = B.cq_n_ent, tid = threadIdx.x;
    int32_t *__restrict__ soff = B.cq_soff;
    if (tid < 256) s_sum[tid] = 0;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    // exclusive prefix of the slot counts over all entries (padding entries count 0), class totals on the side
    for (int r0 = 0; r0 < n; r0 += 1024) {
        const int r = r0 + tid;
        int v = 0;
        if (r < n) {
            const int e = B.cq_ext[r];
            if (e >= 0) {
                const int32_t *__restrict__ cbase = B.cbase + 4 * e;
                v = (cbase[2] - cbase[1]) + (B.exts[e].nn - cbase[3]);
                atomicAdd(&s_sum[B.cq_cls[r]], v);
            }
        }
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if ((tid & 31) >= o) x += y; }
        if ((tid & 31) == 31) s_warp[tid >> 5] = x;
        __syncthreads();
        if (tid < 32) {
            int w = s_warp[tid];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, w, o); if (tid >= o) w += y; }
            s_warp[tid] = w;
        }
        __syncthreads();
        const int excl = s_carry + (tid >= 32 ? s_warp[(tid >> 5) - 1] : 0) + x - v;
        if (r < n) {
            soff[r] = excl;
            if (r == 0 || B.cq_cls[r - 1] != B.cq_cls[r]) s_gfirst[B.cq_cls[r]] = excl;
        }
        __syncthreads();
        if (tid == 1023) s_carry = excl + v;
        __syncthreads();
    }
    if (tid == 0) {
        int at = 0;
        const int span = B.cq_span;
        for (int c = 0; c < 256; c++) {
            s_base[c] = at;
            at += (s_sum[c] + span - 1) / span * span;
        }
        soff[n] = at;
        *B.cq_ncta = at / span;
    }
    __syncthreads();
    for (int r = tid; r < n; r += 1024) {
        const int c = B.cq_cls[r];
        soff[r] = s_base[c] + (soff[r] - s_gfirst[c]);
    }
}

// first plan entry of every CTA span
__global__ void __launch_bounds__(128) k_cq_owner(DevBatch B) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= B.cq_n_ent) return;
    const int a = B.cq_soff[r], e = B.cq_soff[r + 1], sh = B.cq_span_shift;
    for (int c = (a + (1 << sh) - 1) >> sh; (c << sh) < e; c++) B.cq_cta[c] = r;
    if (r == B.cq_n_ent - 1) B.cq_cta[e >> sh] = r;
}

__global__ void __launch_bounds__(kCqThreads, 1) k_coding_flat(DevBatch B, const DevModel *__restrict__ models) {
#ifdef PGPU_HOST_EMULATION
    static double tab_store[4096 * kCqCols + kCqRing * kCqThreads];
    double *tab = tab_store;
#else
    extern __shared__ __align__(128) unsigned char cq_dyn_smem[];
    double *tab = reinterpret_cast<double *>(cq_dyn_smem);
#endif
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ int s_next;
    if ((int)blockIdx.x >= *B.cq_ncta) return;
    double *ring = tab + 4096 * kCqCols + threadIdx.x;   // sum of the j-th start of the current ORF: ring[j * kCqThreads]
    const int span = B.cq_span;
    const int r0 = B.cq_cta[blockIdx.x], r1 = B.cq_cta[blockIdx.x + 1];   // plan entries of this CTA's span: [r0, r1]
    const int first_chain = B.cq_chain[4 * r0];
    int W = 1;
    if (B.cq_chain[4 * r0 + 2] >= 0) W = 4; else if (B.cq_chain[4 * r0 + 1] >= 0) W = 2;
    const int set = models[B.chains[first_chain].model].col;
    if (threadIdx.x == 0) {
        s_next = 0;
        mbar_init(&s_bar, 1);
        mbar_init_fence();
        const char *src = (const char *)(B.dcS + (size_t)set * (4096 * kCqCols));
        mbar_expect(&s_bar, kCqTableBytes);
        for (int k = 0; k < kCqTableBytes; k += 16384) bulk_copy((char *)tab + k, src + k, 16384, &s_bar);
    }
    __syncthreads();
    mbar_wait(&s_bar, 0);

    const int lane = threadIdx.x & (W - 1);
    const unsigned gmask = W == 4 ? 0xFu << (threadIdx.x & 28) : W == 2 ? 0x3u << (threadIdx.x & 30) : 0u;
    const int slot0 = blockIdx.x * span;
    const double *__restrict__ wcol = tab + lane;
    const int mdl = set + lane < B.n_models ? B.cq_colmodel[set + lane] : 0;
    const DevModel &M = models[mdl];
    const uint16_t *__restrict__ dic = B.dic_f;
    const int4 *__restrict__ link = B.link;
    double *__restrict__ cscore = B.cscore;
    int r = r0;
    // next ORF: its descriptor, requested when the current ORF was set up (y == -2: no more slots)
    int4 D1 = make_int4(0, -2, 0, 0);
    int D1r = r0;
    auto draw = [&]() {
        int slot = 0;
        if (lane == 0) slot = atomicAdd(&s_next, 1);
        if (W > 1) slot = __shfl_sync(gmask, slot, 0, W);
        D1.y = -2;
        if (slot >= span) return;
        const int gs = slot0 + slot;
        while (r < r1 && B.cq_soff[r + 1] <= gs) r++;   // slots are drawn in increasing order
        const int hs0 = B.cq_hs0[r];
        if (hs0 < 0) return;                             // class padding: nothing behind it in this span
        D1 = B.orfd[hs0 + (gs - B.cq_soff[r])];
        D1r = r;
    };
    draw();
    // the walk of this group: elements k_cur, k_cur - 1, ... lo (offsets from dic) lead to start node seg (sum so far:
    // acc), the cnt-th start of the ORF that ends at STOP node zstop; nxt = the start after it.  k_cur < lo: transition.
    int k_cur = -1, lo = 0, seg = -1, cnt = 0, zstop = -1;
    int2 nxt = make_int2(-1, 0);
    int64_t cb = INT64_MIN;
    double acc = 0.0;
    uint4 v = make_uint4(0, 0, 0, 0), vn = make_uint4(0, 0, 0, 0);   // chunk [vbase, vbase + 8) and the one below it
    int vbase = -16;
    bool done = false;
    while (!__all_sync(0xffffffffu, done)) {
        if (!done && k_cur < lo) {
            if (seg >= 0) {   // a start is reached
                if (cnt < kCqRing) ring[cnt * kCqThreads] = acc;
                else if (cb != INT64_MIN) cscore[cb + seg] = acc;
                cnt++;
                if (nxt.x >= 0) {   // on to the next start of the ORF
                    seg = nxt.x; lo = nxt.y;
                    const int4 L = link[seg];
                    nxt = make_int2(L.x, L.y);
                } else {
                    // the ORF is finished: the two penalty sweeps, fused, from the last start back to the stop
                    double s2 = -10000.0, s3 = -10000.0;
                    const bool active = cb != INT64_MIN;
                    // one start: cs = its sum, L = its link record (L.w = gene length - 3, L.z = the start before it).
                    // gsize = (L.w + 3) / 3.0 of coding_orf_lane: its integer part and the comparison with 1000 in integer
                    // arithmetic (exact: L.w + 3 < 2^31), the division only for genes above 3 kbp; fmax(fmin(d, lfac), 0)
                    // as two comparisons (the weights are finite: same value, and subtracting +-0 changes nothing)
                    auto sweep = [&](int i, double cs, const int4 &L) {
                        if (cs > s2) s2 = cs; else cs -= (s2 - cs);
                        double lfac;
                        if (L.w + 3 > 3000) lfac = M.lfac_span * (((double)L.w + 3.0) / 3.0 - 80) / 920.0;
                        else lfac = M.lfac[(unsigned)(L.w + 3) / 3u];
                        if (lfac > s3) {
                            s3 = lfac;
                        } else {
                            const double d = s3 - lfac, m = d < lfac ? d : lfac;
                            if (m > 0.0) lfac -= m;
                        }
                        if (lfac > 3.0 && cs < 0.5 * lfac) cs = 0.5 * lfac;
                        cs += lfac;
                        if (active) cscore[cb + i] = cs;
                    };
                    int i = seg, j = cnt - 1;
#pragma unroll 1
                    for (; j >= kCqRing; j--) {   // (an ORF with more than kCqRing starts: the sums beyond are in global memory)
                        const int4 L = link[i];
                        sweep(i, active ? cscore[cb + i] : 0.0, L);
                        i = L.z;
                    }
#pragma unroll 1
                    for (; j >= 0; j--) {
                        const int4 L = link[i];
                        sweep(i, ring[j * kCqThreads], L);
                        i = L.z;
                    }
                    seg = -1;
                }
            }
            if (seg < 0) {   // next ORF
                const int4 D = D1;
                const int rr = D1r;
                if (D.y == -2) {
                    done = true;
                } else {
                    draw();
                    if (D.y >= 0) {
                        k_cur = D.x; seg = D.y; lo = D.z; zstop = D.w;
                        const int4 L = link[seg];
                        nxt = make_int2(L.x, L.y);
                        cb = B.cq_cbase[4 * rr + lane];
                        acc = 0.0;
                        cnt = 0;
                        vbase = -16;
                    }
                }
            }
        }
        if (!done && k_cur >= lo) {
            const int base = k_cur & ~7;
            if (base != vbase) {
                if (base == vbase - 8) v = vn; else v = *reinterpret_cast<const uint4 *>(dic + base);
                vbase = base;
                if (base >= 8) vn = *reinterpret_cast<const uint4 *>(dic + base - 8);   // needed unless the ORF ends in this chunk
            }
            const int hi_t = k_cur - base, lo_t = max(lo - base, 0);
            // elements lo_t .. hi_t of the chunk are added; every element of a plane is a valid index (k_dicodon_index
            // fills the slack), so all eight weights are loaded and the others dropped.  Byte offset of a weight = index * 32.
            const char *__restrict__ wb = reinterpret_cast<const char *>(wcol);
            const double w7 = *reinterpret_cast<const double *>(wb + ((v.w >> 16) << 5)), w6 = *reinterpret_cast<const double *>(wb + ((v.w & 0xffffu) << 5));
            const double w5 = *reinterpret_cast<const double *>(wb + ((v.z >> 16) << 5)), w4 = *reinterpret_cast<const double *>(wb + ((v.z & 0xffffu) << 5));
            const double w3 = *reinterpret_cast<const double *>(wb + ((v.y >> 16) << 5)), w2 = *reinterpret_cast<const double *>(wb + ((v.y & 0xffffu) << 5));
            const double w1 = *reinterpret_cast<const double *>(wb + ((v.x >> 16) << 5)), w0 = *reinterpret_cast<const double *>(wb + ((v.x & 0xffffu) << 5));
            const unsigned take = ((2u << hi_t) - 1u) & ~((1u << lo_t) - 1u);   // bit t: element t is added
            if (take & 0x80u) acc += w7;
            if (take & 0x40u) acc += w6;
            if (take & 0x20u) acc += w5;
            if (take & 0x10u) acc += w4;
            if (take & 0x08u) acc += w3;
            if (take & 0x04u) acc += w2;
            if (take & 0x02u) acc += w1;
            if (take & 0x01u) acc += w0;
            k_cur = base + lo_t - 1;
        }
    }
}

// --------------------------------------------------------------------------------------------------
// start scoring: RBS / upstream motif / type / upstream composition / penalties (lib.pyx:2331-2487)
// --------------------------------------------------------------------------------------------------

// What scoring a start needs from the node itself (model independent)
struct StartNode {
    int c, ndx, stop_val, cc;   // cls byte, position, stop position, codon byte at the stop
    uint32_t bits;              // upstream A/G pattern (SD search)
    uint64_t U, pc;             // packed upstream bases (motif search, composition)
    bool touch;                 // another node lies within 2 bp: _intergenic_mod may read this start's rbs / upstream
                                // scores (_connection.h:60-67: ndx1 + 2 == ndx2 || ndx1 == ndx2 + 1)
};
__device__ __forceinline__ void load_start_node(const DevBatch &B, int node_off, int64_t doff, int nn, int i, StartNode &N) {
    N.c = B.cls[node_off + i];
    N.ndx = B.ndx[node_off + i];
    N.stop_val = B.stop_val[node_off + i];
    N.bits = B.sdbits[node_off + i];
    N.U = B.umot[node_off + i];
    N.pc = B.upc[node_off + i];
    const uint8_t *__restrict__ cod = B.cod + doff;
    N.cc = cls_is_stop(N.c) ? 0 : ((N.c & CLS_REV) ? cod[N.stop_val - 2] : cod[N.stop_val]);
    N.touch = (i > 0 && B.ndx[node_off + i - 1] >= N.ndx - 2) || (i + 1 < nn && B.ndx[node_off + i + 1] <= N.ndx + 2);
}

// node i of chain C.  FULL pass (LEAN = false): every score of the node is written at chain-node g = C.coff + i (plus cs,
// interleaved, when the batch has it); the raw coding score is read from cs_in[g_in].  LEAN pass (the main pass of meta
// mode, every (contig, model) chain): only what the chains that do NOT win are ever asked for is written -- cs =
// cscore + sscore of the starts (record_overlapping_starts, the DP) and, for starts that touch another node, the
// rbs / upstream penalty _intergenic_mod subtracts (`rupen`).  The winner's chain is scored again in full before the
// traceback (api.cu); that is 8 B instead of 54 B written per chain-node.
template <bool LEAN>
__device__ __forceinline__ void start_score_eval(const DevBatch &B, const DevModel *__restrict__ models, const ChainInfo &C,
                                                 const StartNode &N, int64_t g, int i, const double *__restrict__ cs_in,
                                                 int64_t g_in, RunOpts o, MotifOut *__restrict__ mot_out) {
    const uint8_t *__restrict__ cls = B.cls + C.node_off;
    const int c = N.c;
    if (cls_is_stop(c)) {
        if (LEAN) return;   // nothing reads the scores of a STOP node of a chain that is not the winner
        // STOP nodes keep the reset state (node.c:176-197)
        B.cscore[g] = 0.0; B.sscore[g] = 0.0; B.rscore[g] = 0.0; B.uscore[g] = 0.0; B.tscore[g] = 0.0;
        if (B.cs) B.cs[C.ioff + (int64_t)i * C.istride] = 0.0;
        B.rbs[2 * g] = 0; B.rbs[2 * g + 1] = 0;
        if (mot_out) { MotifOut m = {}; mot_out[g] = m; }
        return;
    }
    const DevModel &M = models[C.model];
    const int32_t *__restrict__ sva = B.stop_val + C.node_off;
    const int slen = C.slen, nn = C.nn, ndx = N.ndx, stop_val = N.stop_val;
    const bool rev = c & CLS_REV;
    const int type = c & CLS_TYPE;
    const int edge_mask_now = C.first_pass ? CLS_EDGE : (CLS_EDGE | CLS_CONV);
    bool edge = (c & edge_mask_now) != 0;
    const int start = rev ? slen - 1 - ndx : ndx;
    const double st_wt = M.st_wt;
    // issue every independent load up front: the kernel is latency bound
    const uint32_t pre_bits = N.bits;
    const uint64_t pre_U = N.U;
    const uint64_t pre_pc = N.pc;
    const double pre_cscore = cs_in[g_in];
    const int pre_cc = N.cc;

    int rbs0 = 0, rbs1 = 0;
    MotifOut mot = {};
    if (!edge) {
        if (M.uses_sd) {
            // table-driven Shine-Dalgarno search (lib.pyx:2241-2277 + 791-979)
            const uint32_t bits = pre_bits;
            const uint32_t A = bits & 0xffffu, G = bits >> 16;
            const int omin = rev ? 0 : max(0, 20 - start);  // forward skips negative offsets only
            for (int off = omin; off < 15; off++) {
                const uint32_t gp = ((A >> off) & 0x09u) | ((G >> off) & 0x36u);
                const uint32_t em = *reinterpret_cast<const uint16_t *>(&M.sd_best[off][gp][0]);   // exact | mismatch << 8
                const int e = em & 0xff, m = em >> 8;
                rbs0 = max(rbs0, e);
                rbs1 = max(rbs1, m);
            }
        } else {
            // best upstream motif, stage 2 (lib.pyx:1557-1616); spacer class of the p-th window of a length is
            // fixed: j <= start-16-l (p <= 2) -> 3, p <= 4 -> 2, j >= start-7-l (p >= 11) -> 1, else 0
            const uint64_t U = pre_U;
            int max_spacer = 0, max_spacendx = 0, max_len = 0, max_ndx = 0;
            double max_sc = -100.0;
            const double *__restrict__ mw = M.mot_wt;
            const uint32_t *__restrict__ live = M.mot_live;  // all but a few dozen cells hold the floor weight -4.0
            // A cell that is not live holds exactly -4.0 and cannot replace a maximum that is already >= -4.0 (the
            // comparison is a strict ">"), so once the first window has been taken only live cells matter.
            auto probe = [&](int l, int p) {   // window p of motif length l + 3 (0 <= start - 18 - l + p)
                const int spacendx = p <= 2 ? 3 : (p <= 4 ? 2 : (p >= 11 ? 1 : 0));
                const int index = (int)((U >> (2 * (3 - l + p))) & ((1u << (2 * (l + 3))) - 1u));
                const int cell = (l * 4 + spacendx) * 4096 + index;
                double sc = -4.0;
                if (!live || ((__ldg(&live[cell >> 5]) >> (cell & 31)) & 1u)) sc = __ldg(&mw[cell]);
                if (sc > max_sc) {
                    max_sc = sc; max_spacendx = spacendx; max_spacer = 15 - p;   // = start - j - l - 3, j = start - 18 - l + p
                    max_ndx = index; max_len = l + 3;
                }
            };
            // Candidate windows from the model's table M.mot_hit: four lookups by six upstream bases each tell, for all
            // 52 windows at once, which can hold a live cell; the window the reference visits first is taken
            // unconditionally.  (A live weight below the floor would break the argument: generic loop then.)
            const uint16_t *__restrict__ hit = M.mot_hit;
            bool generic = !hit || !live;
            if (!generic) {
                int l1 = -1, p1 = 0;
#pragma unroll
                for (int l = 3; l >= 0; l--) {
                    const int pm = max(0, 18 + l - start);
                    if (l1 < 0 && pm <= 12) { l1 = l; p1 = pm; }
                }
                if (l1 >= 0) {
                    probe(l1, p1);
                    if (max_sc < -4.0) {
                        generic = true;
                        max_sc = -100.0;
                    } else {
                        const uint32_t e0 = __ldg(&hit[(uint32_t)U & 0xfffu]), e4 = __ldg(&hit[(uint32_t)(U >> 8) & 0xfffu]),
                                       e8 = __ldg(&hit[(uint32_t)(U >> 16) & 0xfffu]), e12 = __ldg(&hit[(uint32_t)(U >> 24) & 0xfffu]);
#pragma unroll
                        for (int l = 3; l >= 0; l--) {
                            const int pm = max(0, 18 + l - start);
                            if (pm > 12) continue;
                            // bit b: a motif of this length that starts at upstream base b = 3 - l + p
                            uint32_t cand = ((e0 >> (4 * l)) & 15u) | (((e4 >> (4 * l)) & 15u) << 4) | (((e8 >> (4 * l)) & 15u) << 8) |
                                            (((e12 >> (4 * l)) & 15u) << 12);
                            cand &= ((1u << (16 - l)) - 1u) & ~((1u << (3 - l + pm)) - 1u);
                            while (cand) {
                                const int bpos = __ffs(cand) - 1;
                                cand &= cand - 1;
                                probe(l, bpos - 3 + l);
                            }
                        }
                    }
                }
            }
            if (generic) {
#pragma unroll
                for (int l = 3; l >= 0; l--) {
                    const uint64_t pf = M.mot_pf[l];
                    if (pf == 0 && max_sc >= -4.0) continue;
#pragma unroll
                    for (int p = 0; p < 13; p++) {
                        if (start - 18 - l + p < 0) continue;
                        const int index = (int)((U >> (2 * (3 - l + p))) & ((1u << (2 * (l + 3))) - 1u));
                        if (!((pf >> (index & 63)) & 1ull) && max_sc >= -4.0) continue;
                        probe(l, p);
                    }
                }
            }
            if (max_sc == -4.0 || max_sc < M.no_mot + 0.69) {
                mot.score = M.no_mot;
            } else {
                mot.ndx = (uint16_t)max_ndx; mot.len = (uint8_t)max_len; mot.spacendx = (uint8_t)max_spacendx;
                mot.spacer = (uint8_t)max_spacer; mot.score = max_sc;
            }
        }
    }

    const int orf_length = abs(ndx - stop_val);
    int edge_gene = edge ? 1 : 0;
    {
        // does the ORF run off the edge?  (is the "stop" a real stop codon)
        int code;
        bool has_n;
        has_n = pre_cc & 64;
        code = rev ? rev_code(pre_cc & 63) : (pre_cc & 63);
        if (has_n || !((M.stopmask >> code) & 1)) edge_gene++;
    }
    double cscore = pre_cscore, tscore, uscore, rscore;
    if (edge) {
        tscore = PGPU_EDGE_BONUS * st_wt / edge_gene;
        uscore = 0.0;
        rscore = 0.0;
    } else {
        tscore = M.type_wt[type] * st_wt;
        const double rbs1w = M.rbs_wt[rbs0], rbs2w = M.rbs_wt[rbs1];
        const double sd_score = fmax(rbs1w, rbs2w) * st_wt;
        if (M.uses_sd) {
            rscore = sd_score;
        } else {
            rscore = st_wt * mot.score;
            if (rscore < sd_score && M.no_mot > -0.5) rscore = sd_score;
        }
        // upstream composition (lib.pyx:1619-1650)
        uscore = 0.0;
        {
            const int ncomp = min(2, start) + max(0, min(30, start - 14));
            uint64_t pc = pre_pc;
            int k = 0;
            for (; k + 8 <= ncomp; k += 8, pc >>= 16) {  // 8 independent loads, adds in the reference's order
                const double w0 = M.uc[k][pc & 3], w1 = M.uc[k + 1][(pc >> 2) & 3], w2 = M.uc[k + 2][(pc >> 4) & 3],
                             w3 = M.uc[k + 3][(pc >> 6) & 3], w4 = M.uc[k + 4][(pc >> 8) & 3], w5 = M.uc[k + 5][(pc >> 10) & 3],
                             w6 = M.uc[k + 6][(pc >> 12) & 3], w7 = M.uc[k + 7][(pc >> 14) & 3];
                uscore += w0; uscore += w1; uscore += w2; uscore += w3; uscore += w4; uscore += w5; uscore += w6; uscore += w7;
            }
            for (; k + 4 <= ncomp; k += 4, pc >>= 8) {
                const double w0 = M.uc[k][pc & 3], w1 = M.uc[k + 1][(pc >> 2) & 3], w2 = M.uc[k + 2][(pc >> 4) & 3],
                             w3 = M.uc[k + 3][(pc >> 6) & 3];
                uscore += w0; uscore += w1; uscore += w2; uscore += w3;
            }
            for (; k < ncomp; k++, pc >>= 2) uscore += M.uc[k][pc & 3];
        }
        // starts that would stop the gene from running off the edge (lib.pyx:2407-2422)
        if (!o.closed && ndx <= 2 && !rev) {
            uscore += PGPU_EDGE_UPS * st_wt;
        } else if (!o.closed && ndx >= slen - 3 && rev) {
            uscore += PGPU_EDGE_UPS * st_wt;
        } else if (i < 500 && !rev) {
            // lib.pyx:2413-2417: is there an edge node before i with the same stop_val?  Only nodes within 5 bp
            // of a sequence end can carry an edge flag, i.e. the first n_lo / last n_hi nodes.  Nodes before i
            // already carry their converted edge flag in this sweep.
            const int n_lo = B.exts[C.ext].n_lo, n_hi = B.exts[C.ext].n_hi;
            bool hit = false;
            for (int j = min(i, n_lo) - 1; j >= 0 && !hit; j--)
                hit = (cls[j] & (CLS_EDGE | CLS_CONV)) && stop_val == sva[j];
            for (int j = i - 1; j >= max(nn - n_hi, n_lo) && !hit; j--)
                hit = (cls[j] & (CLS_EDGE | CLS_CONV)) && stop_val == sva[j];
            if (hit) uscore += PGPU_EDGE_UPS * st_wt;
        } else if (i + 500 >= nn && rev) {
            // lib.pyx:2418-2422; nodes after i still carry the flag of the previous sweep
            const int n_lo = B.exts[C.ext].n_lo, n_hi = B.exts[C.ext].n_hi;
            bool hit = false;
            for (int j = max(i + 1, nn - n_hi); j < nn && !hit; j++)
                hit = (cls[j] & edge_mask_now) && stop_val == sva[j];
            for (int j = i + 1; j < min(n_lo, nn - n_hi) && !hit; j++)
                hit = (cls[j] & edge_mask_now) && stop_val == sva[j];
            if (hit) uscore += PGPU_EDGE_UPS * st_wt;
        }
    }
    // convert starts at the first/last bases into edge nodes (lib.pyx:2424-2434)
    if (!o.closed && !edge && ((ndx <= 2 && !rev) || (ndx >= slen - 3 && rev))) {
        edge_gene++;
        edge = true;
        tscore = 0.0;
        uscore = PGPU_EDGE_BONUS * st_wt / edge_gene;
        rscore = 0.0;
    }
    if (!edge && edge_gene == 1) uscore -= 0.5 * PGPU_EDGE_BONUS * st_wt;
    if (edge_gene == 0 && orf_length < 250) {
        const double negf = M.len_neg[orf_length], posf = M.len_pos[orf_length];   // 250.0 / L, L / 250.0
        rscore *= rscore < 0 ? negf : posf;
        uscore *= uscore < 0 ? negf : posf;
        tscore *= tscore < 0 ? negf : posf;
    }
    if (C.is_meta && slen < 3000 && edge_gene == 0 && (cscore < 5.0 || orf_length < 120))
        cscore -= PGPU_META_PEN * fmax(0.0, (3000.0 - slen) / 2700.0);
    double sscore = tscore + rscore + uscore;
    if (cscore < 0.0) {
        if (edge_gene > 0 && !edge) {
            if (!C.is_meta || slen > 1500) sscore -= st_wt;
            else sscore -= 10.31 - 0.004 * slen;
        } else if (C.is_meta && slen < 3000 && edge) {
            const double min_meta_len = sqrt((double)slen) * 5.0;
            if (orf_length >= min_meta_len) {
                if (cscore >= 0) cscore = -1.0;
                sscore = 0.0;
                uscore = 0.0;
            }
        } else {
            sscore -= 0.5;
        }
    } else if (C.is_meta && cscore < 5.0 && orf_length < 120 && sscore < 0.0) {
        sscore -= st_wt;
    }
    // what record_overlapping_starts and the DP read of a start: one value, interleaved over the chains of the extraction
    if (LEAN) {
        const int64_t gi = C.ioff + (int64_t)i * C.istride;
        B.cs[gi] = cscore + sscore;
        if (N.touch) {   // _intergenic_mod_same (_connection.h:60-67), the part that depends on this start's scores
            double r = 0.0;
            if (rscore < 0) r -= rscore;
            if (uscore < 0) r -= uscore;
            B.rupen[gi] = r;
        }
        return;
    }
    B.cscore[g] = cscore; B.sscore[g] = sscore; B.rscore[g] = rscore; B.uscore[g] = uscore; B.tscore[g] = tscore;
    if (B.cs) B.cs[C.ioff + (int64_t)i * C.istride] = cscore + sscore;
    B.rbs[2 * g] = (uint8_t)rbs0; B.rbs[2 * g + 1] = (uint8_t)rbs1;
    if (mot_out) mot_out[g] = mot;
}

// FULL pass of one node by one thread: node data loaded here
__device__ __forceinline__ void start_score_node(const DevBatch &B, const DevModel *__restrict__ models, const ChainInfo &C,
                                                 int64_t g, int i, RunOpts o, MotifOut *__restrict__ mot_out) {
    StartNode N;
    load_start_node(B, C.node_off, C.doff, C.nn, i, N);
    // the raw coding score: in place, or (winner pass of meta mode) in the main pass' array at the chain's own offset
    const double *cs_in = B.cscore_in ? B.cscore_in : B.cscore;
    start_score_eval<false>(B, models, C, N, g, i, cs_in, B.cscore_in ? C.coff_in + i : g, o, mot_out);
}

// LEAN main pass of meta mode: one thread per chain-node like the full pass (a warp = 32 consecutive nodes of one chain,
// i.e. one model: the table lookups of a warp stay inside one model's tables)
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_start_score_lean(DevBatch B, const DevModel *__restrict__ models, int n_chains,
                                                                 int64_t total, RunOpts o) {
    __shared__ int s_first;
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int k = chain_hint(B, n_chains, min(g, total - 1), total, &s_first);
    if (g >= total) return;
    while (k + 1 < n_chains && B.chains[k + 1].coff <= g) k++;
    const ChainInfo C = B.chains[k];
    // Only starts are scored: the first (#starts) threads of the chain's index range take one start each through the
    // class-sorted node list (+starts, then -starts), so that a warp is either full of starts or exits at once (a
    // quarter of the nodes are STOP nodes; thread per node left 20 of 32 lanes busy).
    const int t = (int)(g - C.coff);
    const int32_t *__restrict__ cbase = B.cbase + 4 * C.ext;
    const int n_fs = cbase[1] - cbase[0], n_rs = cbase[3] - cbase[2];
    if (t >= n_fs + n_rs) return;
    const int i = (B.clist + C.node_off)[t < n_fs ? cbase[0] + t : cbase[2] + (t - n_fs)];
    StartNode N;
    N.c = B.cls[C.node_off + i];
    load_start_node(B, C.node_off, C.doff, C.nn, i, N);
    start_score_eval<true>(B, models, C, N, C.coff + i, i, B.cscore, C.coff + i, o, nullptr);
}

__global__ void __launch_bounds__(128) k_start_score(DevBatch B, const DevModel *__restrict__ models, int n_chains,
                                                      int64_t total, RunOpts o, MotifOut *__restrict__ mot_out) {
    __shared__ int s_first;
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int k = chain_hint(B, n_chains, min(g, total - 1), total, &s_first);
    if (g >= total) return;
    while (k + 1 < n_chains && B.chains[k + 1].coff <= g) k++;
    const ChainInfo C = B.chains[k];
    const int i = (int)(g - C.coff);
    if (i >= C.nn) return;
    start_score_node(B, models, C, C.coff + i, i, o, mot_out);
}

// --------------------------------------------------------------------------------------------------
// Final scoring pass of meta mode restricted to what the result exposes when node arrays are not requested: the
// start and stop node of every gene.  B is the final-pass batch (one chain per contig: chain index == contig).
//   k_gene_list        compacts (contig, gene) pairs into a work list (order is irrelevant)
//   k_coding_genes     one thread per listed gene: the three coding sweeps over the gene's ORF (they need every
//                      start of that ORF, nothing outside it)
//   k_start_score_genes one thread per listed node (2 per gene)
// --------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_gene_list(int n_contigs, const pgpu_contig_summary *__restrict__ summary,
                                                    int2 *__restrict__ list, int *__restrict__ count) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_contigs) return;
    const int ng = summary[c].n_genes;
    if (ng <= 0) return;
    const int base = atomicAdd(count, ng);
    for (int j = 0; j < ng; j++) list[base + j] = make_int2(c, j);
}
__global__ void __launch_bounds__(128) k_coding_genes(DevBatch B, const DevModel *__restrict__ models,
                                                       const int2 *__restrict__ list, const int *__restrict__ count,
                                                       const pgpu_gene *__restrict__ genes, const int64_t *__restrict__ gene_off) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= *count) return;
    const int2 e = list[t];
    const ChainInfo C = B.chains[e.x];
    coding_orf_thread(B, models, C, genes[gene_off[e.x] + e.y].stop_ndx);
}
__global__ void __launch_bounds__(128) k_start_score_genes(DevBatch B, const DevModel *__restrict__ models,
                                                            const int2 *__restrict__ list, const int *__restrict__ count,
                                                            const pgpu_gene *__restrict__ genes,
                                                            const int64_t *__restrict__ gene_off, RunOpts o,
                                                            MotifOut *__restrict__ mot_out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if ((t >> 1) >= *count) return;
    const int2 e = list[t >> 1];
    const ChainInfo C = B.chains[e.x];
    const pgpu_gene G = genes[gene_off[e.x] + e.y];
    const int i = (t & 1) ? G.stop_ndx : G.start_ndx;
    start_score_node(B, models, C, C.coff + i, i, o, mot_out);
}

// --------------------------------------------------------------------------------------------------
// overlapping starts (lib.pyx:2279-2329) + operon values for the DP
// --------------------------------------------------------------------------------------------------
// _intergenic_mod_same (_connection.h:52-78) for n1 = (ndx1, strand1), n2 = ndx2 on the same strand.  The rbs /
// upstream scores of the start involved (node `start` of this chain -- in every use here the node whose scores the
// reference reads is the start, never the STOP) are needed only when the two nodes touch, so they are loaded on
// demand: that case is rare and the two loads were half of this kernel's score traffic.
__device__ __forceinline__ double igm_same(int ndx1, int strand1, int ndx2, int start, const double *__restrict__ rscore,
                                           const double *__restrict__ uscore, const double *__restrict__ rupen, int64_t S,
                                           const DevModel &M) {
    const int dist = abs(ndx1 - ndx2);
    const bool overlap = ndx1 + 2 * strand1 >= ndx2;
    double r = 0.0;
    if (ndx1 + 2 == ndx2 || ndx1 == ndx2 + 1) {
        if (rupen) {   // lean main pass: the two subtractions were done by the scoring pass (interleaved array)
            r = rupen[start * S];
        } else {
            const double rs = rscore[start], us = uscore[start];
            if (rs < 0) r -= rs;
            if (us < 0) r -= us;
        }
    }
    if (dist > 3 * kOperDist) r -= 0.15 * M.st_wt;
    else if ((dist <= kOperDist && !overlap) || dist * 4 < kOperDist) r += M.igt[dist];   // (2.0 - dist / 60) * 0.15 * st_wt, tabulated (no FP64 division)
    return r;
}

// cscore + sscore of a start: from the combined array when the scoring pass wrote one
__device__ __forceinline__ double cs_of(int j, const double *__restrict__ cs, int64_t S, const double *__restrict__ cscore,
                                        const double *__restrict__ sscore) {
    return cs ? cs[j * S] : cscore[j] + sscore[j];
}

// cs[n3] + intergenic_mod for the start n3 recorded in star_ptr of STOP node z:
// forward STOP: _connection.h:189 (n1 = z, n3);  reverse STOP: _connection.h:320,329,354 (n3, n2 = z)
__device__ __forceinline__ double operon_value(int cz, int z, int s, const uint8_t *__restrict__ cls,
                                               const int32_t *__restrict__ ndx, const double *__restrict__ cs, int64_t S,
                                               const double *__restrict__ cscore, const double *__restrict__ sscore,
                                               const double *__restrict__ rscore, const double *__restrict__ uscore,
                                               const double *__restrict__ rupen, const DevModel &M) {
    const int cs_ = cls[s];
    const double base = cs_of(s, cs, S, cscore, sscore);
    if (((cz ^ cs_) & CLS_REV) != 0) return base + M.ig_neg;
    return (cz & CLS_REV) ? base + igm_same(ndx[s], -1, ndx[z], s, rscore, uscore, rupen, S, M)
                          : base + igm_same(ndx[z], 1, ndx[s], s, rscore, uscore, rupen, S, M);
}

// STOP node i of chain C: star_ptr[3] and the operon values (interleaved arrays, ChainInfo::ioff)
__device__ __forceinline__ void overlap_stop_node(const DevBatch &B, const DevModel &M, const ChainInfo &C, int i, RunOpts o,
                                                  int flag) {
    const int nn = C.nn;
    const uint8_t *__restrict__ cls = B.cls + C.node_off;
    const int32_t *__restrict__ ndx = B.ndx + C.node_off;
    const int32_t *__restrict__ sv = B.stop_val + C.node_off;
    const double *__restrict__ cscore = B.cscore + C.coff;
    const double *__restrict__ sscore = B.sscore + C.coff;
    const double *__restrict__ rscore = B.rscore + C.coff;
    const double *__restrict__ uscore = B.uscore + C.coff;
    const int64_t S = C.istride;
    const double *__restrict__ cs = B.cs ? B.cs + C.ioff : nullptr;
    const double *__restrict__ rupen = B.rupen ? B.rupen + C.ioff : nullptr;
    int sp[3] = {-1, -1, -1};
    const int c = cls[i];
    if (cls_is_stop(c) && !(c & CLS_EDGE)) {
        const int my = ndx[i];
        double max_sc = -100.0;
        if (!(c & CLS_REV)) {
            for (int j = i + 3; j >= 0; j--) {
                if (j >= nn || ndx[j] > my + 2) continue;
                if (ndx[j] + o.max_overlap < my) break;
                const int cj = cls[j];
                if ((cj & CLS_REV) || cls_is_stop(cj)) continue;
                if (sv[j] <= my) continue;
                const int f = ndx[j] % 3;
                if (flag == 0) {
                    if (sp[f] == -1) sp[f] = j;
                } else {
                    const double sc = cs_of(j, cs, S, cscore, sscore) + igm_same(my, 1, ndx[j], j, rscore, uscore, rupen, S, M);
                    if (sc > max_sc) { sp[f] = j; max_sc = sc; }
                }
            }
        } else {
            for (int j = i - 3; j < nn; j++) {
                if (j < 0 || ndx[j] < my - 2) continue;
                if (ndx[j] - o.max_overlap > my) break;
                const int cj = cls[j];
                if (!(cj & CLS_REV) || cls_is_stop(cj)) continue;
                if (sv[j] >= my) continue;
                const int f = ndx[j] % 3;
                if (flag == 0) {
                    if (sp[f] == -1) sp[f] = j;
                } else {
                    const double sc = cs_of(j, cs, S, cscore, sscore) + igm_same(ndx[j], -1, my, j, rscore, uscore, rupen, S, M);
                    if (sc > max_sc) { sp[f] = j; max_sc = sc; }
                }
            }
        }
    }
    const int64_t gi = C.ioff + (int64_t)i * S;
#pragma unroll
    for (int f = 0; f < 3; f++) {
        B.star_ptr[3 * gi + f] = sp[f];
        B.opv[3 * gi + f] = sp[f] == -1 ? 0.0 : operon_value(c, i, sp[f], cls, ndx, cs, S, cscore, sscore, rscore, uscore, rupen, M);
    }
}

// one thread per (chain, STOP node); used when the chains do not share extractions (single mode, training, operators)
__global__ void __launch_bounds__(128, 12) k_overlap(DevBatch B, const DevModel *__restrict__ models, int n_chains,
                                                  int64_t total, RunOpts o, int flag) {
    __shared__ int s_first;
    // half-size slot space, as in k_coding_orf: chain k owns the slots from (coff + 1) / 2 on and its STOP nodes (at
    // most nn / 2) take the first of them; star_ptr was preset to -1 for every node
    const int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int k = chain_hint(B, n_chains, min(2 * h, total - 1), total, &s_first, 2 * (int64_t)blockIdx.x * blockDim.x);
    if (h > (total + 1) / 2) return;
    while (k + 1 < n_chains && ((B.chains[k + 1].coff + 1) >> 1) <= h) k++;
    const ChainInfo C = B.chains[k];
    const int t = (int)(h - ((C.coff + 1) >> 1)), nn = C.nn;
    if (t >= nn) return;
    const int32_t *__restrict__ cbase = B.cbase + 4 * C.ext;
    const int n_fe = cbase[2] - cbase[1], n_re = nn - cbase[3];
    if (t >= n_fe + n_re) return;
    const int i = (B.clist + C.node_off)[t < n_fe ? cbase[1] + t : cbase[3] + (t - n_fe)];
    overlap_stop_node(B, models[C.model], C, i, o, flag);
}

// the same with the lanes of a group over the chains (models) that share the extraction, W = 4 / 8 / 16 / 32 lanes per
// STOP node (the thread layout of the grouped k_coding_orf: orf_toff / orf_w / orf_blk): the geometric loop is
// identical for the lanes of a node, so node loads are broadcasts and the interleaved cs / star_ptr / opv accesses of a
// group are contiguous
__global__ void __launch_bounds__(256, 6) k_overlap_lanes(DevBatch B, const DevModel *__restrict__ models, int n_ext, RunOpts o,
                                                        int flag) {
    const int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gt >= B.orf_toff[n_ext]) return;
    int r = B.orf_blk[gt >> 8];
    while (r + 1 < n_ext && B.orf_toff[r + 1] <= gt) r++;
    const int W = B.orf_w[r];
    const int local = (int)(gt - B.orf_toff[r]);
    const int e = B.orf_ext[r];
    const int tl = local >> (31 - __clz(W)), lane = local & (W - 1);   // W is a power of two
    const int32_t *__restrict__ cbase = B.cbase + 4 * e;
    const ExtractInfo &X = B.exts[e];
    const int nn = X.nn;
    const int n_fe = cbase[2] - cbase[1], n_re = nn - cbase[3];
    if (tl >= n_fe + n_re) return;
    const int i = (B.clist + X.node_off)[tl < n_fe ? cbase[1] + tl : cbase[3] + (tl - n_fe)];
    const int ch0 = B.ext_chain_off[e], nch = B.ext_chain_off[e + 1] - ch0;
    for (int c0 = lane; c0 < nch; c0 += W) {
        const ChainInfo C = B.chains[B.ext_chains[ch0 + c0]];
        overlap_stop_node(B, models[C.model], C, i, o, flag);
    }
}

// operon values for caller-supplied star_ptr (operator-level pgpu_score_connections)
__global__ void __launch_bounds__(128) k_opv(DevBatch B, const DevModel *__restrict__ models, int n_chains, int64_t total) {
    __shared__ int s_first;
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int k = chain_hint(B, n_chains, min(g, total - 1), total, &s_first);
    if (g >= total) return;
    while (k + 1 < n_chains && B.chains[k + 1].coff <= g) k++;
    const ChainInfo C = B.chains[k];
    const int i = (int)(g - C.coff);
    if (i >= C.nn) return;
    const uint8_t *__restrict__ cls = B.cls + C.node_off;
    const int c = cls[i];
#pragma unroll
    for (int f = 0; f < 3; f++) {
        const int64_t gi = C.ioff + (int64_t)i * C.istride;
        const int s = B.star_ptr[3 * gi + f];
        B.opv[3 * gi + f] = (s < 0 || s >= C.nn || !cls_is_stop(c)) ? 0.0
            : operon_value(c, i, s, cls, B.ndx + C.node_off, nullptr, 1, B.cscore + C.coff, B.sscore + C.coff, B.rscore + C.coff,
                           B.uscore + C.coff, nullptr, models[C.model]);
    }
}

// node pairs of an extraction: sum_i (i - min_i)  (implementation independent)
__global__ void __launch_bounds__(256) k_pairs(DevBatch B, int n_ext, int total_nodes, unsigned long long *ext_pairs) {
    __shared__ int s_first;
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int hint = ext_hint(B, n_ext, min(g, total_nodes - 1), blockIdx.x * blockDim.x, total_nodes, &s_first);
    int e = -1;
    unsigned long long v = 0;
    if (g < total_nodes) {
        e = hint;
        while (e + 1 < n_ext && B.exts[e + 1].node_off <= g) e++;
        v = (unsigned long long)((g - B.exts[e].node_off) - B.win_min[g]);
    }
    // warp-level pre-reduction when the whole warp belongs to one extraction
    const int e0 = __shfl_sync(0xffffffffu, e, 0);
    if (__all_sync(0xffffffffu, e == e0)) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
        if ((threadIdx.x & 31) == 0 && e0 >= 0 && v) atomicAdd(&ext_pairs[e0], v);
    } else if (e >= 0 && v) {
        atomicAdd(&ext_pairs[e], v);
    }
}

// --------------------------------------------------------------------------------------------------
// launch wrappers
// --------------------------------------------------------------------------------------------------
// block -> owner tables: owner k covers flat indices [off_k, off_k + nn_k); entry b is the owner of index 128*b
__global__ void __launch_bounds__(128) k_block_owner_chains(const ChainInfo *__restrict__ chains, int n, int32_t *__restrict__ tab) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int64_t a = chains[k].coff, e = a + chains[k].nn;
    for (int64_t b = (a + 127) >> 7; (b << 7) < e; b++) tab[b] = k;
}
__global__ void __launch_bounds__(128) k_block_owner_exts(const ExtractInfo *__restrict__ exts, int n, int32_t *__restrict__ tab) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int64_t a = exts[k].node_off, e = a + exts[k].nn;
    for (int64_t b = (a + 127) >> 7; (b << 7) < e; b++) tab[b] = k;
}
// the same for ranges given by an offset array off[0..n] (entry b = owner of index b << shift)
__global__ void __launch_bounds__(128) k_block_owner_off(const int64_t *__restrict__ off, int n, int shift, int32_t *__restrict__ tab) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int64_t a = off[k], e = off[k + 1], step = (int64_t)1 << shift;
    for (int64_t b = (a + step - 1) >> shift; (b << shift) < e; b++) tab[b] = k;
}
void launch_block_owner_off(const int64_t *off, int n, int shift, int32_t *tab, cudaStream_t st) {
    if (n > 0) k_block_owner_off<<<(n + 127) / 128, 128, 0, st>>>(off, n, shift, tab);
}
void launch_block_owner_chains(const ChainInfo *chains, int n, int32_t *tab, cudaStream_t st) {
    if (n > 0) k_block_owner_chains<<<(n + 127) / 128, 128, 0, st>>>(chains, n, tab);
}
void launch_block_owner_exts(const ExtractInfo *exts, int n, int32_t *tab, cudaStream_t st) {
    if (n > 0) k_block_owner_exts<<<(n + 127) / 128, 128, 0, st>>>(exts, n, tab);
}

void launch_node_prep(const DevBatch &B, int n_ext, int total_nodes, int seq_parts, cudaStream_t st) {
    if (n_ext == 0) return;
    if (total_nodes > 0) k_node_prep<<<(total_nodes + 127) / 128, 128, 0, st>>>(B, n_ext, total_nodes, seq_parts);
    k_class_index<<<(n_ext * 32 + 127) / 128, 128, 0, st>>>(B, n_ext);
}
void launch_coding(const DevBatch &B, const DevModel *models, int n_chains, int64_t total, int n_ext, int total_nodes,
                   cudaStream_t st) {
    if (n_chains == 0 || total == 0) return;
    if (B.ext_chains && B.dcS && B.cq_max_cta > 0 && total_nodes > 0) {   // dicodon tables in shared memory
        k_orf_links<<<((total_nodes + 1) / 2 + 1 + 127) / 128, 128, 0, st>>>(B, n_ext, total_nodes, (int)(B.dic_r - B.dic_f));
        k_cq_plan<<<1, 1024, 0, st>>>(B);
        k_cq_owner<<<(B.cq_n_ent + 127) / 128, 128, 0, st>>>(B);
#ifdef PGPU_HOST_EMULATION
        k_coding_flat<<<(unsigned)B.cq_max_cta, kCqThreads, 0, st>>>(B, models);
#else
        // (per device: set before every launch, a process may drive several devices)
        cudaFuncSetAttribute(k_coding_flat, cudaFuncAttributeMaxDynamicSharedMemorySize, kCqSmemBytes);
        k_coding_flat<<<(unsigned)B.cq_max_cta, kCqThreads, kCqSmemBytes, st>>>(B, models);
#endif
    } else if (B.ext_chains && B.dcT && n_ext > 0 && total_nodes > 0) {
        static const int minb = getenv("PGPU_CODING_MINB") ? atoi(getenv("PGPU_CODING_MINB")) : 5;  // A/B switch
        if (B.orf_toff) {  // grouped mapping planned by the host (api.cu)
            const unsigned nb = (unsigned)((B.orf_threads + 32 * kOrfWarps - 1) / (32 * kOrfWarps));
            if (minb == 6) k_coding_orf<6, true><<<nb, 32 * kOrfWarps, 0, st>>>(B, models, n_ext, total_nodes);
            else k_coding_orf<5, true><<<nb, 32 * kOrfWarps, 0, st>>>(B, models, n_ext, total_nodes);
        } else {
            const unsigned nb = ((total_nodes + 1) / 2 + 1 + kOrfWarps - 1) / kOrfWarps;
            if (minb == 6) k_coding_orf<6, false><<<nb, 32 * kOrfWarps, 0, st>>>(B, models, n_ext, total_nodes);
            else k_coding_orf<5, false><<<nb, 32 * kOrfWarps, 0, st>>>(B, models, n_ext, total_nodes);
        }
    } else {
        k_coding<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(B, models, n_chains, total);
    }
}
void launch_start_score(const DevBatch &B, const DevModel *models, int n_chains, int64_t total, RunOpts o, void *mot_out,
                        cudaStream_t st) {
    if (n_chains == 0 || total == 0) return;
    k_start_score<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(B, models, n_chains, total, o, (MotifOut *)mot_out);
}
void launch_start_score_lean(const DevBatch &B, const DevModel *models, int n_chains, int64_t total, RunOpts o, cudaStream_t st) {
    if (n_chains == 0 || total == 0) return;
    static const int minb = getenv("PGPU_SCORE_MINB") ? atoi(getenv("PGPU_SCORE_MINB")) : 9;   // 9 = 56 registers, no spills, 36 warps / SM; 10 = 48 registers (a few spilled words), 40 warps / SM: measured equal (32.5 vs 32.2 ms scoring phase)
    const unsigned nb = (unsigned)((total + 127) / 128);
    if (minb == 12) k_start_score_lean<12><<<nb, 128, 0, st>>>(B, models, n_chains, total, o);
    else if (minb == 10) k_start_score_lean<10><<<nb, 128, 0, st>>>(B, models, n_chains, total, o);
    else k_start_score_lean<9><<<nb, 128, 0, st>>>(B, models, n_chains, total, o);
}
void launch_score_chains(const DevBatch &B, const DevModel *models, int n_chains, int64_t total, RunOpts o,
                         void *mot_out, int n_ext, int total_nodes, cudaStream_t st) {
    launch_coding(B, models, n_chains, total, n_ext, total_nodes, st);
    launch_start_score(B, models, n_chains, total, o, mot_out, st);
}
void launch_score_genes(const DevBatch &B, const DevModel *models, int n_contigs, const void *summary, const void *genes,
                        const int64_t *gene_off, int64_t gene_cap, int2 *list, int *count, RunOpts o, void *mot_out,
                        cudaStream_t st) {
    if (n_contigs == 0 || gene_cap == 0) return;
    cudaMemsetAsync(count, 0, sizeof(int), st);
    k_gene_list<<<(n_contigs + 127) / 128, 128, 0, st>>>(n_contigs, (const pgpu_contig_summary *)summary, list, count);
    k_coding_genes<<<(unsigned)((gene_cap + 127) / 128), 128, 0, st>>>(B, models, list, count, (const pgpu_gene *)genes, gene_off);
    k_start_score_genes<<<(unsigned)((2 * gene_cap + 127) / 128), 128, 0, st>>>(B, models, list, count, (const pgpu_gene *)genes,
                                                                                   gene_off, o, (MotifOut *)mot_out);
}
// operator Sequence.shine_dalgarno: one window of contig 0, rules evaluated directly (sd_device.cuh)
__global__ void k_shine_dalgarno(DevBatch B, const DevModel *__restrict__ models, int model, int pos, int start, int strand,
                                 int exact, int32_t *__restrict__ out) {
    if (blockIdx.x == 0 && threadIdx.x == 0)
        *out = sd_window(B.digits, B.contigs[0].slen, pos, start, models[model].rbs_wt, strand, exact != 0);
}
void launch_shine_dalgarno(const DevBatch &B, const DevModel *models, int model, int pos, int start, int strand, int exact,
                           int32_t *out, cudaStream_t st) {
    k_shine_dalgarno<<<1, 32, 0, st>>>(B, models, model, pos, start, strand, exact, out);
}
void launch_opv(const DevBatch &B, const DevModel *models, int n_chains, int64_t total, cudaStream_t st) {
    if (n_chains == 0 || total == 0) return;
    k_opv<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(B, models, n_chains, total);
}
void launch_pairs(const DevBatch &B, int n_ext, int total_nodes, unsigned long long *ext_pairs, cudaStream_t st) {
    if (n_ext == 0 || total_nodes == 0) return;
    k_pairs<<<(total_nodes + 255) / 256, 256, 0, st>>>(B, n_ext, total_nodes, ext_pairs);
}
void launch_overlap(const DevBatch &B, const DevModel *models, int n_chains, int64_t total, int64_t total_il, int n_ext,
                    RunOpts o, int flag, cudaStream_t st) {
    if (n_chains == 0 || total == 0) return;
    // star_ptr = -1 for every node (node records of single mode show it for starts too); the lean main pass of meta mode
    // only ever reads the STOP nodes' entries, which the kernel writes itself (all three frames of every STOP node)
    if (!B.rupen) cudaMemsetAsync(B.star_ptr, 0xff, 3 * (size_t)total_il * sizeof(int32_t), st);
    if (B.orf_toff && B.ext_chains && n_ext > 0)
        k_overlap_lanes<<<(unsigned)((B.orf_threads + 255) / 256), 256, 0, st>>>(B, models, n_ext, o, flag);
    else
        k_overlap<<<(unsigned)(((total + 1) / 2 + 1 + 127) / 128), 128, 0, st>>>(B, models, n_chains, total, o, flag);
}

}  // namespace pgpu

// --------------------------------------------------------------------------------------------------
// DP index (model independent, once per extraction): class-ordered ndx and, per node, the predecessor
// candidates / ranges that the connection rules pin down geometrically (see dp_kernels.cu, k_dp_fast)
//   +STOP  i: x = first +STOP (class position) with ndx > stop_val(i)                    [operon range]
//   -start i: x = node index of its own -STOP (ndx == stop_val), y/z = class-position range of the
//             +STOPs with ndx in (stop_val-4, stop_val+195)                              [gene, 3' overlap]
//   -STOP  i: x,y,z = for frame 0,1,2 the nearest previous -STOP node whose ORF spans ndx(i), or -1  [operon]
// --------------------------------------------------------------------------------------------------
namespace pgpu {

__device__ __forceinline__ int lower_bound_ndx(const int32_t *__restrict__ a, int n, int key) {
    // first position with a[p] >= key
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(128) k_dp_index(DevBatch B, int n_ext, int total_nodes) {
    __shared__ int s_first;
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    int e = ext_hint(B, n_ext, min(g, total_nodes - 1), blockIdx.x * blockDim.x, total_nodes, &s_first);
    if (g >= total_nodes) return;
    while (e + 1 < n_ext && B.exts[e + 1].node_off <= g) e++;
    const ExtractInfo X = B.exts[e];
    const int p = g - X.node_off, nn = X.nn;  // p = class-ordered position
    const int32_t *__restrict__ ndx = B.ndx + X.node_off;
    const int32_t *__restrict__ sv = B.stop_val + X.node_off;
    const uint8_t *__restrict__ cls = B.cls + X.node_off;
    const int32_t *__restrict__ clist = B.clist + X.node_off;
    const int32_t *__restrict__ cb = B.cbase + 4 * e;
    // (1) class-ordered ndx
    const int i = clist[p];
    B.cndx[X.node_off + p] = ndx[i];
    // (2) per-node candidates; computed by the thread that owns class position p == node clist[p]
    const int c = cls[i], kind = cls_kind(c), my = ndx[i], msv = sv[i];
    int4 r = make_int4(-1, -1, -1, -1);
    // class-ordered ndx of +STOPs / -STOPs are not complete yet for other threads' positions, so search
    // through clist -> ndx (sorted within a class segment)
    auto lb_class = [&](int cbeg, int cend, int key) {  // first class position with ndx >= key
        int lo = cbeg, hi = cend;
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (ndx[clist[mid]] < key) lo = mid + 1; else hi = mid;
        }
        return lo;
    };
    const int fe0 = cb[1], fe1 = cb[2], re0 = cb[3], re1 = nn;
    if (kind == K_FE) {
        r.x = lb_class(fe0, fe1, msv + 1) - fe0;
    } else if (kind == K_RS) {
        const int q = lb_class(re0, re1, msv);
        r.x = (q < re1 && ndx[clist[q]] == msv) ? clist[q] : -1;
        r.y = lb_class(fe0, fe1, msv - 3) - fe0;
        r.z = lb_class(fe0, fe1, msv + 195) - fe0;
    } else if (kind == K_RE) {
        // nearest previous -STOP of every frame, within the 1000-node window
        int found = 0;
        int cand[3] = {-1, -1, -1};
        for (int q = p - 1; q >= re0 && found != 7; q--) {
            const int j = clist[q];
            if (j < i - 2 * kMaxNodeDist) break;
            const int fj = cls_frame(cls[j]);
            if (found & (1 << fj)) continue;
            found |= 1 << fj;
            if (sv[j] > my) cand[fj] = j;
        }
        r.x = cand[0]; r.y = cand[1]; r.z = cand[2];
    }
    B.dpx[X.node_off + i] = r;

    // ---- merged stream of +STOP / -start nodes (k_dp_dq): rank = (#FE + #RS nodes before the node) ----
    const int32_t *__restrict__ crank = B.crank + 4 * (int64_t)X.node_off;
    auto igrank = [&](int node) { return crank[4 * (int64_t)node + 1] + crank[4 * (int64_t)node + 2]; };
    if (kind == K_FE || kind == K_RS) {
        const int q = igrank(i);
        B.ig_node[X.node_off + q] = i | (kind == K_FE ? (int)0x80000000 : 0);
        B.ig_ndx[X.node_off + q] = my;
        if (kind == K_FE && B.feq) B.feq[X.node_off + p] = q;  // class order -> merged position (k_dp_ml)
    }
    // first merged-stream position whose ndx >= key: binary search over node indices (ndx sorted), then rank
    auto ig_lb = [&](int key) {
        int lo = 0, hi = nn;
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (ndx[mid] < key) lo = mid + 1; else hi = mid;
        }
        return lo < nn ? igrank(lo) : igrank(nn - 1) + ((cls_kind(cls[nn - 1]) == K_FE || cls_kind(cls[nn - 1]) == K_RS) ? 1 : 0);
    };
    int4 q4 = make_int4(-1, -1, -1, -1);
    if (kind == K_FE) {
        q4.x = ig_lb(msv + 1);               // operon range: +STOPs with ndx > stop_val
        q4.w = igrank(B.win_min[X.node_off + i]);
    } else if (kind == K_RS) {
        q4.x = r.x;                          // own -STOP (node index)
        q4.y = ig_lb(msv - 3);               // 3' overlap range: +STOPs with ndx in [stop_val-3, stop_val+195)
        q4.z = ig_lb(msv + 195);
        q4.w = igrank(B.win_min[X.node_off + i]);
    } else if (kind == K_RE) {
        q4.x = r.x; q4.y = r.y; q4.z = r.z;
    }
    B.dqx[X.node_off + i] = q4;
}

void launch_dp_index(const DevBatch &B, int n_ext, int total_nodes, cudaStream_t st) {
    if (n_ext == 0 || total_nodes == 0) return;
    k_dp_index<<<(total_nodes + 127) / 128, 128, 0, st>>>(B, n_ext, total_nodes);
}

}  // namespace pgpu
