// dp_kernels.cu -- the connection-scoring dynamic program and everything downstream of it.
//
// Reference semantics: BaseConnectionScorer._score_connections / _find_max_index / _disentangle_overlaps /
// _max_forward_pointers / _dynamic_programming (src/pyrodigal/lib.pyx:1205-1311), the four specialised
// score_connection functions (src/pyrodigal/_connection.h:94-367), the skip filter
// (src/pyrodigal/impl/generic.h:29-36), eliminate_bad_genes (vendor/Prodigal/dprog.c:306-335),
// Genes._extract / _tweak_final_starts (lib.pyx:3231-3401), the meta-mode winner rule (lib.pyx:5364).
//
// B200 design:
//  * one CTA walks one chain (contig x model); chains are independent, so a batch exposes 10^3..10^6 CTAs.
//  * the SIMD skip filter is not a pass: nodes are pre-bucketed by class (+start,+STOP,-start,-STOP) and a
//    target only visits the classes the six filter clauses admit (about one third of the window), in index
//    order, so tie-breaking ("largest j wins") is unchanged.
//  * the DP-dependent state of the last 1024 nodes (score, ndx of the traceback node) lives in a shared
//    memory ring; model-independent node data comes from L1/L2; the per-target constants of the next 32
//    steps are staged cooperatively (coalesced) into shared memory.
//  * per step: lanes stride over admissible predecessors, warp-shuffle arg-max, one __syncthreads.
//  * doubles everywhere, no FMA contraction (-fmad=false): results are bit-identical to the reference.
#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <cstring>

#include "kernels.cuh"

namespace pgpu {

constexpr int kDpThreads = 128;
constexpr int kDpWarps = kDpThreads / 32;
constexpr int kRing = 1024;
constexpr int kStage = 32;
constexpr int kTbNone = INT32_MIN;  // "traceb == -1"

struct StepConst {
    int32_t ndx, sv, m, cls;
    int32_t cr_i[4];
    int32_t cr_m[4];
    int32_t sp[3];
    int32_t n3ndx[3];
    int32_t n3sv[3];
    int32_t pad;
    double opv[3];
    double gcb3[3];
    double cs;   // cscore + sscore of the target (K_RS)
    double gcb;  // training: bias . gc_score of the target
};

struct Cand {
    double v;
    int32_t j, tn, fr;
};

__device__ __forceinline__ void cand_take(Cand &a, double v, int j, int tn, int fr) {
    // candidates of one thread arrive with increasing j: ">=" keeps the largest j among equal values
    if (v >= a.v) { a.v = v; a.j = j; a.tn = tn; a.fr = fr; }
}
__device__ __forceinline__ void cand_merge(Cand &a, double v, int j, int tn, int fr) {
    if (v > a.v || (v == a.v && j > a.j)) { a.v = v; a.j = j; a.tn = tn; a.fr = fr; }
}

template <int FINAL>
__global__ void __launch_bounds__(kDpThreads) k_dp(DevBatch B, const DevModel *__restrict__ models,
                                                    const int32_t *__restrict__ order, int n_chains) {
    __shared__ double s_score[kRing];
    __shared__ int32_t s_tbn[kRing];
    __shared__ StepConst s_c[kStage];
    __shared__ Cand s_red[2][kDpWarps];

    const int chain = order ? order[blockIdx.x] : blockIdx.x;
    const ChainInfo C = B.chains[chain];
    const int nn = C.nn, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (nn == 0) {
        if (tid == 0) { B.chain_ipath[chain] = -1; B.chain_score[chain] = 0.0; }
        return;
    }
    const DevModel &M = models[C.model];
    const int32_t *__restrict__ ndx = B.ndx + C.node_off;
    const int32_t *__restrict__ sv = B.stop_val + C.node_off;
    const uint8_t *__restrict__ cls = B.cls + C.node_off;
    const int32_t *__restrict__ win_min = B.win_min + C.node_off;
    const int32_t *__restrict__ crank = B.crank + 4 * (int64_t)C.node_off;
    const int32_t *__restrict__ clist = B.clist + C.node_off;
    const int32_t *__restrict__ cbase = B.cbase + 4 * C.ext;
    const double *__restrict__ cscore = B.cscore + C.coff;
    const double *__restrict__ sscore = B.sscore + C.coff;
    // interleaved arrays (ChainInfo::ioff): node j at j * S (3-vectors: S3 * j + f)
    const int64_t S = C.istride, S3 = 3 * S;
    const double *__restrict__ csum = B.cs ? B.cs + C.ioff : nullptr;   // cscore + sscore as one (interleaved) array
    const double *__restrict__ opv = B.opv + 3 * C.ioff;
    const double *__restrict__ gcb = FINAL ? nullptr : B.gcb + C.coff;
    const int32_t *__restrict__ star_ptr = B.star_ptr + 3 * C.ioff;
    const Strided<double> score{B.score + C.ioff, S};
    const Strided<int32_t> traceb{B.traceb + C.ioff, S};
    const Strided<int8_t> ov_mark{B.ov_mark + C.ioff, S};
    const int cb0 = cbase[0], cb1 = cbase[1], cb2 = cbase[2], cb3 = cbase[3];
    const double ig_neg = M.ig_neg;

    double prev_score = 0.0;
    int prev_tbn = kTbNone;
    double best_sc = -1.0;
    int best_i = -1, best_tb = -1;

    for (int i0 = 0; i0 < nn; i0 += kStage) {
        // ---- stage the constants of the next kStage targets (one lane per target) ----
        __syncthreads();
        if (tid < kStage && i0 + tid < nn) {
            const int i = i0 + tid;
            StepConst sc;
            sc.ndx = ndx[i]; sc.sv = sv[i]; sc.cls = cls[i]; sc.m = win_min[i];
            const int4 a = *reinterpret_cast<const int4 *>(crank + 4 * (int64_t)i);
            const int4 b = *reinterpret_cast<const int4 *>(crank + 4 * (int64_t)sc.m);
            sc.cr_i[0] = a.x; sc.cr_i[1] = a.y; sc.cr_i[2] = a.z; sc.cr_i[3] = a.w;
            sc.cr_m[0] = b.x; sc.cr_m[1] = b.y; sc.cr_m[2] = b.z; sc.cr_m[3] = b.w;
            const int kind = cls_kind(sc.cls);
            sc.cs = 0.0; sc.gcb = 0.0;
#pragma unroll
            for (int k = 0; k < 3; k++) { sc.sp[k] = -1; sc.n3ndx[k] = 0; sc.n3sv[k] = 0; sc.opv[k] = 0.0; sc.gcb3[k] = 0.0; }
            if (kind == K_RE) {
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const int s = star_ptr[S3 * (int64_t)i + k];
                    sc.sp[k] = s;
                    if (s != -1) {
                        sc.n3ndx[k] = ndx[s]; sc.n3sv[k] = sv[s];
                        sc.opv[k] = opv[S3 * (int64_t)i + k];
                        if (!FINAL) sc.gcb3[k] = gcb[s];
                    }
                }
            } else if (kind == K_RS) {
                sc.cs = csum ? csum[i * S] : cscore[i] + sscore[i];
                if (!FINAL) sc.gcb = gcb[i];
            }
            s_c[tid] = sc;
        }
        __syncthreads();

        const int iend = min(i0 + kStage, nn);
        for (int i = i0; i < iend; i++) {
            const StepConst &K = s_c[i - i0];
            const int ndx_i = K.ndx, sv_i = K.sv, kind2 = cls_kind(K.cls), f2 = cls_frame(K.cls);
            Cand best = {-DBL_MAX, -1, kTbNone, -1};

            // DP state of predecessor j: registers (j == i-1), shared ring, or global (giant-ORF windows)
            auto state = [&](int j, double &sj, int &tj) {
                if (j == i - 1) { sj = prev_score; tj = prev_tbn; }
                else if (i - j < kRing) { sj = s_score[j & (kRing - 1)]; tj = s_tbn[j & (kRing - 1)]; }
                else { sj = score[j]; const int tb = traceb[j]; tj = tb == -1 ? kTbNone : ndx[tb]; }
            };

            if (kind2 == K_FS) {
                // <- +STOP (intergenic, same strand)
                for (int p = cb1 + K.cr_m[1] + tid, pe = cb1 + K.cr_i[1]; p < pe; p += kDpThreads) {
                    const int j = clist[p], nj = ndx[j];
                    double sj; int tj; state(j, sj, tj);
                    if (tj == kTbNone || nj + 2 >= ndx_i) continue;
                    double term = 0.0;
                    if (FINAL) { const int dist = ndx_i - nj; term = dist > 3 * kOperDist ? ig_neg : (dist <= kOperDist ? M.igt[dist] : 0.0); }
                    cand_take(best, sj + term, j, nj, -1);
                }
                Cand b2 = {-DBL_MAX, -1, kTbNone, -1};
                // <- -start (intergenic, strand switch)
                for (int p = cb2 + K.cr_m[2] + tid, pe = cb2 + K.cr_i[2]; p < pe; p += kDpThreads) {
                    const int j = clist[p], nj = ndx[j];
                    double sj; int tj; state(j, sj, tj);
                    if (tj == kTbNone || nj >= ndx_i) continue;
                    cand_take(b2, sj + (FINAL ? ig_neg : 0.0), j, nj, -1);
                }
                cand_merge(best, b2.v, b2.j, b2.tn, b2.fr);
            } else if (kind2 == K_FE) {
                // <- +start of the same frame (gene)
                for (int p = cb0 + K.cr_m[0] + tid, pe = cb0 + K.cr_i[0]; p < pe; p += kDpThreads) {
                    const int j = clist[p];
                    if (cls_frame(cls[j]) != f2) continue;
                    const int nj = ndx[j];
                    if (sv_i >= nj) continue;
                    double sj; int tj; state(j, sj, tj);
                    double term;
                    if (FINAL) term = csum ? csum[j * S] : cscore[j] + sscore[j];
                    else term = ((double)(ndx_i + 2 - nj + 1)) * gcb[j];
                    cand_take(best, sj + term, j, nj, -1);
                }
                Cand b2 = {-DBL_MAX, -1, kTbNone, -1};
                // <- +STOP (operon: overlapping forward genes)
                for (int p = cb1 + K.cr_m[1] + tid, pe = cb1 + K.cr_i[1]; p < pe; p += kDpThreads) {
                    const int j = clist[p], nj = ndx[j];
                    if (sv_i >= nj) continue;
                    double sj; int tj; state(j, sj, tj);
                    if (tj == kTbNone) continue;
                    const int s = star_ptr[S3 * (int64_t)j + f2];
                    if (s == -1) continue;
                    double term;
                    if (FINAL) term = opv[S3 * (int64_t)j + f2];
                    else term = ((double)(ndx_i + 2 - ndx[s] + 1)) * gcb[s];
                    cand_take(b2, sj + term, j, nj, -1);
                }
                cand_merge(best, b2.v, b2.j, b2.tn, b2.fr);
            } else if (kind2 == K_RS) {
                // <- -STOP of the same frame (reverse gene)
                for (int p = cb3 + K.cr_m[3] + tid, pe = cb3 + K.cr_i[3]; p < pe; p += kDpThreads) {
                    const int j = clist[p];
                    if (cls_frame(cls[j]) != f2) continue;
                    if (sv[j] <= ndx_i) continue;
                    const int nj = ndx[j];
                    double sj; int tj; state(j, sj, tj);
                    double term;
                    if (FINAL) term = K.cs;
                    else term = ((double)(ndx_i - (nj - 2) + 1)) * K.gcb;
                    cand_take(best, sj + term, j, nj, -1);
                }
                Cand b2 = {-DBL_MAX, -1, kTbNone, -1};
                // <- +STOP (overlapping opposite-strand 3' ends)
                const double cs_diff = K.cs + ig_neg;
                for (int p = cb1 + K.cr_m[1] + tid, pe = cb1 + K.cr_i[1]; p < pe; p += kDpThreads) {
                    const int j = clist[p], nj = ndx[j];
                    if (sv_i - 2 >= nj + 2) continue;
                    const int ovlp = (nj + 2) - (sv_i - 2) + 1;
                    if (ovlp >= kMaxOppOvlp) continue;
                    if ((nj - sv_i) >= (ndx_i - nj + 3)) continue;
                    double sj; int tj; state(j, sj, tj);
                    if (tj == kTbNone) continue;
                    if ((nj - sv_i) >= (sv_i - 3 - tj)) continue;
                    double term;
                    if (FINAL) term = cs_diff;
                    else term = ((double)(ndx_i - (sv_i - 2) + 1 - ovlp * 2)) * K.gcb;
                    cand_take(b2, sj + term, j, nj, -1);
                }
                cand_merge(best, b2.v, b2.j, b2.tn, b2.fr);
            } else {  // K_RE
                // <- +STOP (intergenic with the triple-overlap search)
                for (int p = cb1 + K.cr_m[1] + tid, pe = cb1 + K.cr_i[1]; p < pe; p += kDpThreads) {
                    const int j = clist[p], nj = ndx[j];
                    const int left = nj + 2, right = ndx_i - 2;
                    if (left >= right) continue;
                    double sj; int tj; state(j, sj, tj);
                    if (tj == kTbNone) continue;
                    int maxfr = -1, ovlp = 0;
                    double maxval = 0.0;
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        if (K.sp[k] == -1) continue;
                        ovlp = left - K.n3sv[k] + 3;
                        if (ovlp <= 0 || ovlp >= kMaxOppOvlp) continue;
                        if (ovlp >= K.n3ndx[k] - left) continue;
                        if (ovlp >= K.n3sv[k] - tj - 2) continue;
                        const double cur = K.opv[k];
                        if (FINAL ? (cur > maxval) : (K.gcb3[k] > maxval)) { maxfr = k; maxval = cur; }
                    }
                    double term;
                    if (FINAL) term = maxfr != -1 ? K.opv[maxfr] : ig_neg;
                    else term = ((double)(right - left + 1 - ovlp * 2)) * (maxfr != -1 ? K.gcb3[maxfr] : 0.0);
                    cand_take(best, sj + term, j, nj, maxfr);
                }
                Cand b2 = {-DBL_MAX, -1, kTbNone, -1};
                // <- -start (intergenic, same strand)
                for (int p = cb2 + K.cr_m[2] + tid, pe = cb2 + K.cr_i[2]; p < pe; p += kDpThreads) {
                    const int j = clist[p], nj = ndx[j];
                    if (nj >= ndx_i - 2) continue;
                    double sj; int tj; state(j, sj, tj);
                    if (tj == kTbNone) continue;
                    double term = 0.0;
                    if (FINAL) { const int dist = ndx_i - nj; term = dist > 3 * kOperDist ? ig_neg : (dist <= kOperDist ? M.igt[dist] : 0.0); }
                    cand_take(b2, sj + term, j, nj, -1);
                }
                cand_merge(best, b2.v, b2.j, b2.tn, b2.fr);
                Cand b3 = {-DBL_MAX, -1, kTbNone, -1};
                // <- -STOP (operon: overlapping reverse genes)
                for (int p = cb3 + K.cr_m[3] + tid, pe = cb3 + K.cr_i[3]; p < pe; p += kDpThreads) {
                    const int j = clist[p];
                    if (sv[j] <= ndx_i) continue;
                    const int fj = cls_frame(cls[j]);
                    if (K.sp[fj] == -1) continue;
                    const int nj = ndx[j];
                    double sj; int tj; state(j, sj, tj);
                    double term;
                    if (FINAL) term = K.opv[fj];
                    else term = ((double)(K.n3ndx[fj] - (nj - 2) + 1)) * K.gcb3[fj];
                    cand_take(b3, sj + term, j, nj, -1);
                }
                cand_merge(best, b3.v, b3.j, b3.tn, b3.fr);
            }

            // ---- arg-max over the CTA: equal values -> larger j ----
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double v = __shfl_xor_sync(0xffffffffu, best.v, off);
                const int j = __shfl_xor_sync(0xffffffffu, best.j, off);
                const int tn = __shfl_xor_sync(0xffffffffu, best.tn, off);
                const int fr = __shfl_xor_sync(0xffffffffu, best.fr, off);
                cand_merge(best, v, j, tn, fr);
            }
            const int par = i & 1;
            if (lane == 0) s_red[par][wid] = best;
            __syncthreads();
            Cand fin = s_red[par][0];
#pragma unroll
            for (int w = 1; w < kDpWarps; w++) { const Cand c = s_red[par][w]; cand_merge(fin, c.v, c.j, c.tn, c.fr); }

            // "if (n1->score + score >= n2->score)" starting from score 0, traceb -1
            double sc_i = 0.0;
            int tb_i = -1, tbn_i = kTbNone, fr_i = -1;
            if (fin.j >= 0 && fin.v >= 0.0) { sc_i = fin.v; tb_i = fin.j; tbn_i = fin.tn; fr_i = fin.fr; }
            if (tid == 0) { score[i] = sc_i; traceb[i] = tb_i; ov_mark[i] = (int8_t)fr_i; }
            if (tid == 32) { s_score[i & (kRing - 1)] = sc_i; s_tbn[i & (kRing - 1)] = tbn_i; }
            prev_score = sc_i;
            prev_tbn = tbn_i;
            if ((kind2 == K_FE || kind2 == K_RS) && sc_i >= best_sc) { best_sc = sc_i; best_i = i; best_tb = tb_i; }
        }
    }
    if (tid == 0) {
        // lib.pyx:1239-1251 + 1311: largest index among equal maxima; -1 when nothing leads into it
        const bool ok = best_i >= 0 && best_tb != -1;
        B.chain_ipath[chain] = ok ? best_i : -1;
        B.chain_score[chain] = ok ? best_sc : 0.0;
    }
}

// --------------------------------------------------------------------------------------------------
// k_dp_fast: the final-scoring DP (final == 1) restructured around what the connection rules actually
// depend on.  One WARP per chain.  For every predecessor class the value "score[j] + term" is either
//   * a constant-term connection (intergenic, more than 180 bp away: term = -0.15*st_wt), answered by a
//     range maximum over per-16-entry block summaries of fl(score + term) -- exact, ties -> later j;
//   * one of a handful of geometrically pinned candidates (own stop of a reverse gene, the +STOPs around a
//     3' overlap, the <=3 reverse STOPs whose ORF spans the target, the +STOPs inside the target's ORF);
//   * or a running per-frame maximum (best start of the current forward ORF).
// Every candidate value is computed with exactly the reference's operations, so score / traceb / ov_mark
// are bit-identical to _connection.h:94-367; only pairs that cannot win are never materialised.
// --------------------------------------------------------------------------------------------------
constexpr int kFastWarps = 4;  // chains per CTA (independent warps)
constexpr double kNeg = -DBL_MAX;

struct WCand {
    double v;
    int32_t j, fr;
};
struct FastK {  // per-target constants staged in shared memory, 32 targets per warp at a time
    int32_t ndx, sv, cls, leave;
    int32_t wlo, wmin;
    int32_t dx, dy, dz;
    int32_t sp0, sp1, sp2;
    int32_t n3n0, n3n1, n3n2, n3s0, n3s1, n3s2;
    double op0, op1, op2;
    double cs;
};
__device__ __forceinline__ void wc_merge(WCand &a, double v, int j, int fr) {
    if (v > a.v || (v == a.v && j > a.j)) { a.v = v; a.j = j; a.fr = fr; }
}

// --------------------------------------------------------------------------------------------------
// k_dp_dq: like k_dp_fast, but the constant-term (far intergenic) maximum is a sliding-window maximum kept
// in a monotone deque (shared memory, per warp), so a DP step costs O(1) instead of a range query:
//   * +STOP and -start nodes form one merged stream ("intergenic sources"); both add the same constant
//     (-0.15*st_wt) once they are more than 180 bp behind the target, for both target kinds that use them
//     (+start and -STOP), and those targets always have the regular window [i-1000, i);
//   * entries enter the deque when they cross the 180-bp boundary (values fl(score + const); an entry pops
//     every older entry with a value <= its own, so the front is the maximum with ties -> latest j) and leave
//     it when they drop out of the window;
//   * everything within 180 bp, and every geometrically pinned candidate, is evaluated individually.
// The warp arg-max uses redux.sync on an order-preserving integer image of the doubles.
// --------------------------------------------------------------------------------------------------
constexpr int kDqCap = 64;
constexpr int kDqRing = 128;   // merged-stream entries kept in shared memory per chain (power of two)

struct DqK {  // per-target constants, staged 32 targets at a time
    int32_t ndx, sv, cls, leave;
    int32_t x, y, z, w;  // dqx
    int32_t sp0, sp1, sp2;
    int32_t n3n0, n3n1, n3n2, n3s0, n3s1, n3s2;
    int32_t pad;
    double op0, op1, op2;
    double cs;          // final DP: cscore + sscore of a start; training DP: gene term of a +start / bias sum of a -start
    double g0, g1, g2;  // training DP: bias . gc_score of the recorded starts
};

// arg-max over the warp of (v, j): larger v wins, equal v -> larger j; returns the winner in every lane
__device__ __forceinline__ void warp_argmax(double &v, int &j, int &fr) {
    if (v == 0.0) v = 0.0;  // -0.0 and +0.0 compare equal as doubles: give them one integer image
    long long b = __double_as_longlong(v);
    b = b >= 0 ? b : (b ^ 0x7fffffffffffffffLL);  // order-preserving map double -> int64
    const int hi = (int)(b >> 32);
    const unsigned lo = (unsigned)b;
    const int mhi = __reduce_max_sync(0xffffffffu, hi);
    const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
    const bool top = hi == mhi && lo == mlo;
    const int mj = __reduce_max_sync(0xffffffffu, top ? j : -1);
    const int mfr = __reduce_max_sync(0xffffffffu, (top && j == mj) ? fr + 1 : 0) - 1;
    long long r = ((long long)mhi << 32) | (long long)mlo;
    r = r >= 0 ? r : (r ^ 0x7fffffffffffffffLL);
    v = __longlong_as_double(r);
    j = mj;
    fr = mfr;
}

template <int MINB, int FINAL>
__global__ void __launch_bounds__(32 * kFastWarps, MINB) k_dp_dq(DevBatch B, const DevModel *__restrict__ models,
                                                                   const int32_t *__restrict__ order, int n_chains) {
    __shared__ DqK s_k[kFastWarps][32];
    __shared__ double s_dqv[kFastWarps][kDqCap];
    __shared__ int32_t s_dqj[kFastWarps][kDqCap];
    // the most recent merged-stream entries (source value, traceback node) of the chain: what a step reads from within
    // 200 bp was written a few steps ago -- from shared memory instead of an L2 round trip (a single chain is pure latency)
    __shared__ double s_rsv[MINB == 4 ? kFastWarps : 1][MINB == 4 ? kDqRing : 1];
    __shared__ int32_t s_rtb[MINB == 4 ? kFastWarps : 1][MINB == 4 ? kDqRing : 1];
    const int lane = threadIdx.x & 31, wslot = threadIdx.x >> 5;
    const int slot = blockIdx.x * kFastWarps + wslot;
    if (slot >= n_chains) return;
    const int chain = order ? order[slot] : slot;
    const ChainInfo C = B.chains[chain];
    const int nn = C.nn;
    if (nn == 0) {
        if (lane == 0) { B.chain_ipath[chain] = -1; B.chain_score[chain] = 0.0; }
        return;
    }
    const DevModel &M = models[C.model];
    const int32_t *__restrict__ ndx = B.ndx + C.node_off;
    const int32_t *__restrict__ sv = B.stop_val + C.node_off;
    const uint8_t *__restrict__ cls = B.cls + C.node_off;
    const int32_t *__restrict__ ig_node = B.ig_node + C.node_off;
    const int32_t *__restrict__ ig_ndx = B.ig_ndx + C.node_off;
    const int4 *__restrict__ dqx = B.dqx + C.node_off;
    const double *__restrict__ cscore = B.cscore + C.coff;
    const double *__restrict__ sscore = B.sscore + C.coff;
    // interleaved arrays (ChainInfo::ioff): node j at j * S (3-vectors: S3 * j + f)
    const int64_t S = C.istride, S3 = 3 * S;
    const double *__restrict__ csum = B.cs ? B.cs + C.ioff : nullptr;   // cscore + sscore as one (interleaved) array
    const double *__restrict__ opv = B.opv + 3 * C.ioff;
    const int32_t *__restrict__ star_ptr = B.star_ptr + 3 * C.ioff;
    const Strided<double> score{B.score + C.ioff, S};       // written and re-read by this warp: no read-only path
    const Strided<int32_t> traceb{B.traceb + C.ioff, S};
    const Strided<int8_t> ov_mark{B.ov_mark + C.ioff, S};
    double *svig = B.dp_svig + C.coff;
    int32_t *tbig = B.dp_tbig + C.coff;
    DqK *sk = s_k[wslot];
    double *dqv = s_dqv[wslot];
    int32_t *dqj = s_dqj[wslot];
    double *rsv = s_rsv[MINB == 4 ? wslot : 0];
    int32_t *rtb = s_rtb[MINB == 4 ? wslot : 0];
    const double ig_neg = FINAL ? M.ig_neg : 0.0;   // training DP: every intergenic connection scores 0
    const double *__restrict__ gcb = FINAL ? nullptr : B.gcb + C.coff;   // training DP: bias . gc_score of a start

    // merged-stream cursors: cur = finalized entries, lo = first entry inside [i-1000, i), far = first entry
    // that is NOT more than 180 bp behind the target
    int cur = 0, lo = 0, far = 0;
    // entry q: from the ring while it is among the last kDqRing - 1 entries (slot q & mask is not reused before entry
    // q + kDqRing is written, and entries are written in order: index <= cur)
    // (only in the instantiation for few chains, MINB == 4: with thousands of chains the kernel is throughput bound and
    // the 64-register instantiation has no room for it)
    constexpr bool kRing = MINB == 4;
    auto SV = [&](int q) -> double { return kRing && cur - q < kDqRing ? rsv[q & (kDqRing - 1)] : svig[q]; };
    auto TB = [&](int q) -> int { return kRing && cur - q < kDqRing ? rtb[q & (kDqRing - 1)] : tbig[q]; };
    int dq_head = 0, dq_cnt = 0;
    bool dq_ok = true;
    double rc_v0 = kNeg, rc_v1 = kNeg, rc_v2 = kNeg;
    // score of the last -STOP of every frame: a reverse gene's own stop and the operon predecessors of a -STOP are that
    // node (no in-frame stop lies inside an ORF), so their score is at hand instead of a global read of what lane 0 stored;
    // any other node falls back to the array
    double re_v0 = 0.0, re_v1 = 0.0, re_v2 = 0.0;
    int re_j0 = -1, re_j1 = -1, re_j2 = -1;
    constexpr bool kFew = MINB == 4;   // (the instantiation for few chains; the 64-register one has no room for this)
    auto SC = [&](int j) -> double {
        if (!kFew) return score[j];
        return j == re_j0 ? re_v0 : (j == re_j1 ? re_v1 : (j == re_j2 ? re_v2 : score[j]));
    };
    int rc_j0 = -1, rc_j1 = -1, rc_j2 = -1;
    // -start nodes depend only on STOP nodes (their own -STOP, +STOPs around a 3' overlap): they are parked and
    // evaluated lane-parallel just before the next target that reads them (a +start or -STOP) or at the end of
    // the staged block
    int pend_cnt = 0, pend_i = 0, pend_q = 0;
    auto flush = [&](int i0) {
        if (lane < pend_cnt) {
            const DqK &P = sk[pend_i - i0];
            double bv = kNeg;
            int bj = -1;
            if (P.x >= P.pad && P.x >= 0 && P.x < pend_i) {   // own -STOP (gene)
                bv = SC(P.x) + (FINAL ? P.cs : ((double)(P.ndx - (ndx[P.x] - 2) + 1)) * P.cs);
                bj = P.x;
            }
            const double cs_diff = P.cs + ig_neg;
            for (int q = max(P.y, P.w); q < min(P.z, pend_q); q++) {  // +STOPs overlapping the 3' end
                const int nd = ig_node[q];
                if (nd >= 0) continue;
                const double s = SV(q);
                if (s == kNeg) continue;
                const int nj = ig_ndx[q];
                if (P.sv - 2 >= nj + 2) continue;
                const int ovlp = (nj + 2) - (P.sv - 2) + 1;
                if (ovlp >= kMaxOppOvlp) continue;
                if ((nj - P.sv) >= (P.ndx - nj + 3)) continue;
                if ((nj - P.sv) >= (P.sv - 3 - ndx[TB(q)])) continue;
                const double v = s + (FINAL ? cs_diff : ((double)(P.ndx - (P.sv - 2) + 1 - ovlp * 2)) * P.cs);
                const int j = nd & 0x7fffffff;
                if (v > bv || (v == bv && j > bj)) { bv = v; bj = j; }
            }
            double sc = 0.0;
            int tb = -1;
            if (bj >= 0 && bv >= 0.0) { sc = bv; tb = bj; }
            score[pend_i] = sc; traceb[pend_i] = tb; ov_mark[pend_i] = -1;
            svig[pend_q] = tb == -1 ? kNeg : sc;
            if (kRing) rsv[pend_q & (kDqRing - 1)] = tb == -1 ? kNeg : sc;   // (a -start entry has no traceback record)
        }
        pend_cnt = 0;
        __syncwarp();
    };

    for (int i0 = 0; i0 < nn; i0 += 32) {
      __syncwarp();
      if (i0 + lane < nn) {
          const int i = i0 + lane;
          DqK k;
          k.ndx = ndx[i]; k.sv = sv[i]; k.cls = cls[i];
          k.leave = i > 2 * kMaxNodeDist ? cls_kind(cls[i - 2 * kMaxNodeDist - 1]) : -1;
          const int kind = cls_kind(k.cls);
          const int4 dx = dqx[i];
          k.x = dx.x; k.y = dx.y; k.z = dx.z; k.w = dx.w;
          if (FINAL) k.cs = (kind == K_FS || kind == K_RS) ? (csum ? csum[i * S] : cscore[i] + sscore[i]) : 0.0;
          else k.cs = kind == K_FS ? ((double)(k.sv + 2 - k.ndx + 1)) * gcb[i] : (kind == K_RS ? gcb[i] : 0.0);   // gene term / bias sum
          k.g0 = k.g1 = k.g2 = 0.0;
          k.sp0 = k.sp1 = k.sp2 = -1;
          k.n3n0 = k.n3n1 = k.n3n2 = k.n3s0 = k.n3s1 = k.n3s2 = 0;
          k.op0 = k.op1 = k.op2 = 0.0;
          k.pad = kind == K_RS ? B.win_min[C.node_off + i] : 0;  // node index of the window start
          if (kind == K_RE) {
              k.sp0 = star_ptr[S3 * (int64_t)i]; k.sp1 = star_ptr[S3 * (int64_t)i + 1]; k.sp2 = star_ptr[S3 * (int64_t)i + 2];
              if (k.sp0 != -1) { k.n3n0 = ndx[k.sp0]; k.n3s0 = sv[k.sp0]; k.op0 = opv[S3 * (int64_t)i]; if (!FINAL) k.g0 = gcb[k.sp0]; }
              if (k.sp1 != -1) { k.n3n1 = ndx[k.sp1]; k.n3s1 = sv[k.sp1]; k.op1 = opv[S3 * (int64_t)i + 1]; if (!FINAL) k.g1 = gcb[k.sp1]; }
              if (k.sp2 != -1) { k.n3n2 = ndx[k.sp2]; k.n3s2 = sv[k.sp2]; k.op2 = opv[S3 * (int64_t)i + 2]; if (!FINAL) k.g2 = gcb[k.sp2]; }
          }
          sk[lane] = k;
      }
      __syncwarp();
      const int iend = min(i0 + 32, nn);
      for (int i = i0; i < iend; i++) {
        const DqK &K = sk[i - i0];
        const int ci = K.cls, kind = cls_kind(ci), f2 = cls_frame(ci), ndx_i = K.ndx, sv_i = K.sv;
        lo += (K.leave == K_FE) | (K.leave == K_RS);
        if (kind == K_RS) {  // park it (merged-stream slot reserved now)
            if (lane == pend_cnt) { pend_i = i; pend_q = cur; }
            pend_cnt++;
            cur++;
            continue;
        }
        if (kind != K_FE && pend_cnt) flush(i0);
        double wv = kNeg;
        int wj = -1, wfr = -1;
        auto cand = [&](double v, int j, int fr) { if (v > wv || (v == wv && j > wj)) { wv = v; wj = j; wfr = fr; } };
        const double cs_i = K.cs;

        if (kind == K_FS || kind == K_RE) {
            // ---- entries that fall more than 180 bp behind move into the deque ----
            const int thr = ndx_i - 3 * kOperDist;
            if (far < cur && ig_ndx[far] < thr)  // usually nothing crosses the boundary at this step
            for (;;) {
                const int q = far + lane;
                const bool in = q < cur;
                const int nq = in ? ig_ndx[q] : 0x7fffffff;
                const double sq = in ? SV(q) : kNeg;
                const int jq = in ? (ig_node[q] & 0x7fffffff) : 0;
                const int c = __popc(__ballot_sync(0xffffffffu, nq < thr));  // ndx sorted: a prefix of the lanes
                for (int t = 0; t < c; t++) {
                    const double s = __shfl_sync(0xffffffffu, sq, t);
                    const int j = __shfl_sync(0xffffffffu, jq, t);
                    if (s == kNeg || !dq_ok) continue;
                    const double x = s + ig_neg;
                    while (dq_cnt > 0 && dqv[(dq_head + dq_cnt - 1) & (kDqCap - 1)] <= x) dq_cnt--;
                    if (dq_cnt == kDqCap) { dq_ok = false; continue; }
                    if (lane == 0) { dqv[(dq_head + dq_cnt) & (kDqCap - 1)] = x; dqj[(dq_head + dq_cnt) & (kDqCap - 1)] = j; }
                    dq_cnt++;
                    __syncwarp();
                }
                far += c;
                if (c < 32) break;
            }
            const int flo = max(far, lo);
            bool far_scan = !dq_ok;   // deque overflowed once (scores fell monotonically over > 64 entries): scan the far range
            if (dq_ok) {
                while (dq_cnt > 0 && dqj[dq_head] < i - 2 * kMaxNodeDist) { dq_head = (dq_head + 1) & (kDqCap - 1); dq_cnt--; }
                if (!FINAL && kind == K_RE && dq_cnt > 0) {
                    // Training DP: a far +STOP that triggers the triple overlap does not score "at least the plain value"
                    // as in the final DP (its term replaces the 0), so the far maximum is only usable when its node
                    // cannot trigger: not a +STOP within [stop - 4, stop + 194] of a recorded start.  Otherwise every far
                    // source is evaluated (rare).
                    const int fj = dqj[dq_head];
                    if (cls_kind(cls[fj]) == K_FE) {
                        const int nf = ndx[fj];
                        far_scan = (K.sp0 != -1 && nf >= K.n3s0 - 4 && nf < K.n3s0 + 195) || (K.sp1 != -1 && nf >= K.n3s1 - 4 && nf < K.n3s1 + 195) ||
                                   (K.sp2 != -1 && nf >= K.n3s2 - 4 && nf < K.n3s2 + 195);
                    }
                }
                if (!far_scan && dq_cnt > 0 && lane == 0) cand(dqv[dq_head], dqj[dq_head], -1);
            }
            if (far_scan && (FINAL || kind == K_FS)) {
                for (int q = lo + lane; q < flo; q += 32) {
                    const double s = SV(q);
                    if (s != kNeg) cand(s + ig_neg, ig_node[q] & 0x7fffffff, -1);
                }
            }
            if (kind == K_FS) {
                // near sources (_connection.h:116-129): +STOP distance-dependent term, -start strand switch
                for (int q = flo + lane; q < cur; q += 32) {
                    const double s = SV(q);
                    if (s == kNeg) continue;
                    const int nd = ig_node[q], nj = ig_ndx[q];
                    if (nd < 0) {
                        if (nj + 2 >= ndx_i) continue;
                        const int dist = ndx_i - nj;
                        cand(s + (!FINAL || dist > 3 * kOperDist ? ig_neg : (dist <= kOperDist ? M.igt[dist] : 0.0)), nd & 0x7fffffff, -1);
                    } else {
                        if (nj >= ndx_i) continue;
                        cand(s + ig_neg, nd, -1);
                    }
                }
            } else {
                // +STOP with the triple-overlap search (_connection.h:297-334), for one merged-stream position
                auto eval_fe = [&](int q, double s, int nj, int nd) {
                    const int left = nj + 2, right = ndx_i - 2;
                    if (left >= right) return;
                    int maxfr = -1, tj = kTbNone, ovlp = 0;
                    double maxval = 0.0, maxg = 0.0;
                    // (training DP: the comparison is between the bias sum and the value of the final DP, the overlap
                    // that enters the score is the one examined last: _connection.h:320-324)
                    auto probe = [&](int k, int spk, int n3n, int n3s, double op, double g) {
                        if (spk == -1) return;
                        ovlp = left - n3s + 3;
                        if (ovlp <= 0 || ovlp >= kMaxOppOvlp) return;
                        if (ovlp >= n3n - left) return;
                        if (tj == kTbNone) tj = ndx[TB(q)];
                        if (ovlp >= n3s - tj - 2) return;
                        if (FINAL ? (op > maxval) : (g > maxval)) { maxfr = k; maxval = op; maxg = g; }
                    };
                    probe(0, K.sp0, K.n3n0, K.n3s0, K.op0, K.g0);
                    probe(1, K.sp1, K.n3n1, K.n3s1, K.op1, K.g1);
                    probe(2, K.sp2, K.n3n2, K.n3s2, K.op2, K.g2);
                    if (FINAL) cand(s + (maxfr != -1 ? maxval : ig_neg), nd & 0x7fffffff, maxfr);
                    else cand(s + ((double)(right - left + 1 - ovlp * 2)) * (maxfr != -1 ? maxg : 0.0), nd & 0x7fffffff, maxfr);
                };
                for (int q = flo + lane; q < cur; q += 32) {
                    const double s = SV(q);
                    if (s == kNeg) continue;
                    const int nd = ig_node[q], nj = ig_ndx[q];
                    if (nd < 0) {
                        eval_fe(q, s, nj, nd);
                    } else {  // -start (_connection.h:335-341)
                        if (nj >= ndx_i - 2) continue;
                        const int dist = ndx_i - nj;
                        cand(s + (!FINAL || dist > 3 * kOperDist ? ig_neg : (dist <= kOperDist ? M.igt[dist] : 0.0)), nd, -1);
                    }
                }
                if (!FINAL && far_scan) {   // every far source, exactly
                    for (int q = lo + lane; q < flo; q += 32) {
                        const double s = SV(q);
                        if (s == kNeg) continue;
                        const int nd = ig_node[q];
                        if (nd < 0) eval_fe(q, s, ig_ndx[q], nd); else cand(s + ig_neg, nd, -1);
                    }
                }
                // far +STOPs whose position can trigger the triple overlap (200 bp after the stop of a recorded
                // start): their plain value is covered by the far maximum, the overlap value can only be larger
                auto special = [&](int spk, int n3s) {
                    if (spk == -1) return;
                    int a_ = lo, b_ = flo;  // ndx in [n3s-4, n3s+194]
                    while (a_ < b_) { const int mid = (a_ + b_) >> 1; if (ig_ndx[mid] < n3s - 4) a_ = mid + 1; else b_ = mid; }
                    const int a = a_;
                    b_ = flo;
                    while (a_ < b_) { const int mid = (a_ + b_) >> 1; if (ig_ndx[mid] < n3s + 195) a_ = mid + 1; else b_ = mid; }
                    for (int q = a + lane; q < a_; q += 32) {
                        const int nd = ig_node[q];
                        const double s = SV(q);
                        if (nd < 0 && s != kNeg) eval_fe(q, s, ig_ndx[q], nd);
                    }
                };
                special(K.sp0, K.n3s0);
                special(K.sp1, K.n3s1);
                special(K.sp2, K.n3s2);
                // -STOPs whose ORF spans this stop: operon (_connection.h:343-356); at most one per frame
                if (lane < 3) {
                    const int j = lane == 0 ? K.x : (lane == 1 ? K.y : K.z);
                    const int spl = lane == 0 ? K.sp0 : (lane == 1 ? K.sp1 : K.sp2);
                    const double opl = lane == 0 ? K.op0 : (lane == 1 ? K.op1 : K.op2);
                    const int n3l = lane == 0 ? K.n3n0 : (lane == 1 ? K.n3n1 : K.n3n2);
                    const double gl = lane == 0 ? K.g0 : (lane == 1 ? K.g1 : K.g2);
                    if (j >= 0 && j >= i - 2 * kMaxNodeDist && spl != -1)
                        cand(SC(j) + (FINAL ? opl : ((double)(n3l - (ndx[j] - 2) + 1)) * gl), j, -1);
                }
            }
        } else if (kind == K_FE) {
            {   // best +start of this ORF (gene): running maximum of fl(score + cscore + sscore)
                const double rv = f2 == 0 ? rc_v0 : (f2 == 1 ? rc_v1 : rc_v2);
                const int rj = f2 == 0 ? rc_j0 : (f2 == 1 ? rc_j1 : rc_j2);
                if (lane == 0 && rj >= 0) cand(rv, rj, -1);
            }
            // +STOPs inside the ORF (operon, _connection.h:178-191)
            for (int q = max(K.x, K.w) + lane; q < cur; q += 32) {
                const int nd = ig_node[q];
                if (nd >= 0) continue;
                const double s = SV(q);
                if (s == kNeg) continue;
                const int j = nd & 0x7fffffff;
                const int sp = star_ptr[S3 * (int64_t)j + f2];
                if (sp == -1) continue;
                cand(s + (FINAL ? opv[S3 * (int64_t)j + f2] : ((double)(ndx_i + 2 - ndx[sp] + 1)) * gcb[sp]), j, -1);
            }
        }

        {
            const unsigned have = __ballot_sync(0xffffffffu, wj >= 0);
            if ((have & (have - 1)) == 0) {  // at most one lane holds a candidate: broadcast it
                const int src = have ? __ffs(have) - 1 : 0;
                wv = __shfl_sync(0xffffffffu, wv, src);
                wj = __shfl_sync(0xffffffffu, wj, src);
                wfr = __shfl_sync(0xffffffffu, wfr, src);
            } else {
                warp_argmax(wv, wj, wfr);
            }
        }
        double sc_i = 0.0;
        int tb_i = -1, fr_i = -1;
        if (wj >= 0 && wv >= 0.0) { sc_i = wv; tb_i = wj; fr_i = wfr; }
        if (lane == 0) {
            score[i] = sc_i; traceb[i] = tb_i; ov_mark[i] = (int8_t)fr_i;
            if (kind == K_FE) {
                svig[cur] = tb_i == -1 ? kNeg : sc_i;  // edge-artifact rule: nothing leads into it
                tbig[cur] = tb_i;
                if (kRing) { rsv[cur & (kDqRing - 1)] = tb_i == -1 ? kNeg : sc_i; rtb[cur & (kDqRing - 1)] = tb_i; }
            }
        }
        if (kFew && kind == K_RE) {
            if (f2 == 0) { re_v0 = sc_i; re_j0 = i; } else if (f2 == 1) { re_v1 = sc_i; re_j1 = i; } else { re_v2 = sc_i; re_j2 = i; }
        }
        if (kind == K_FE) {
            cur++;
            if (f2 == 0) { rc_v0 = kNeg; rc_j0 = -1; } else if (f2 == 1) { rc_v1 = kNeg; rc_j1 = -1; } else { rc_v2 = kNeg; rc_j2 = -1; }
        } else if (kind == K_FS) {
            const double g = sc_i + cs_i;
            if (f2 == 0) { if (g >= rc_v0) { rc_v0 = g; rc_j0 = i; } }
            else if (f2 == 1) { if (g >= rc_v1) { rc_v1 = g; rc_j1 = i; } }
            else { if (g >= rc_v2) { rc_v2 = g; rc_j2 = i; } }
        }
        __syncwarp();
      }
      if (pend_cnt) flush(i0);  // the staged constants are about to be replaced
    }
    // the best terminal node (lib.pyx:1239-1251) is found by k_chain_best
}

// --------------------------------------------------------------------------------------------------
// k_dp_ml ("model lanes"): one warp walks ALL chains of one extraction (contig x translation table) at once,
// one lane per model.  The models of a contig share the node set, so everything geometric -- node kinds,
// windows, the 180-bp boundary, which predecessors the six filter clauses admit, the overlap tests that do not
// involve a traceback -- is warp-uniform and evaluated once; only the scores differ per lane.  A step is a
// short serial loop over the (few) individually evaluated candidates, with no cross-lane reduction at all.
//   * far intergenic maximum: a two-stack sliding-window maximum (exact, O(1) amortised, no capacity limit):
//     the back stack is a running (value, node) maximum in registers, the front stack a suffix maximum written
//     once per entry to an interleaved array when the window start passes the split point.  Window and split
//     positions are geometric, hence uniform: the rebuild loop does not diverge.
//   * every per-chain array is interleaved [node or entry][lane] (ChainInfo::ioff), so a warp access touches S
//     consecutive elements.
//   * the DP window in shared memory.  A step is a chain of dependent loads, so what bounds the kernel is the
//     latency of each load, not bandwidth; the state a target reads most is therefore kept on chip, per warp:
//       - cs = cscore + sscore of the next targets: the rows [i, i + rows) x S lanes of the interleaved array are
//         contiguous, so they are fetched with 1-D bulk copies (TMA, cp.async.bulk + mbarrier complete_tx) into a
//         double-buffered tile one chunk ahead of the walk;
//       - per frame the score of the last -STOP (own stop of a reverse gene, operon predecessors) and the best
//         +start of the open forward ORF.
//   * rules that only admit +STOPs (operon, 3' overlap, triple overlap) walk the +STOP class list (node, ndx,
//     merged position) instead of the merged stream; their ranges are geometric and come from k_dp_index.
// Candidate order does not matter: the arg-max is "larger value, then larger node index", as everywhere else.
// --------------------------------------------------------------------------------------------------
constexpr int kMlWarps = 4;

struct MlK {  // geometric constants of a target, staged 32 targets at a time (model independent)
    int32_t ndx, sv, cls, leave;
    // +STOP : a = first +STOP (class position) inside the ORF and the window            [operon]
    // -start: a = node of its own -STOP (or -1), [b, c) = +STOP class positions of the 3' overlap range
    // -STOP : a, b, c = per frame the previous -STOP whose ORF spans this stop (or -1)   [operon]
    int32_t a, b, c, pad;
};

// (mbarrier / 1-D bulk copy helpers: common.cuh)

// (2.0 - dist / 60) * 0.15 of _connection.h:74 for dist = 0..60: the model-independent part of the distance term; the
// term itself is this times the start weight of the lane's model (same operations, same order as the reference)
__constant__ double c_igA[61] = {
    (2.0 - ((double)0 / 60)) * 0.15,
    (2.0 - ((double)1 / 60)) * 0.15,
    (2.0 - ((double)2 / 60)) * 0.15,
    (2.0 - ((double)3 / 60)) * 0.15,
    (2.0 - ((double)4 / 60)) * 0.15,
    (2.0 - ((double)5 / 60)) * 0.15,
    (2.0 - ((double)6 / 60)) * 0.15,
    (2.0 - ((double)7 / 60)) * 0.15,
    (2.0 - ((double)8 / 60)) * 0.15,
    (2.0 - ((double)9 / 60)) * 0.15,
    (2.0 - ((double)10 / 60)) * 0.15,
    (2.0 - ((double)11 / 60)) * 0.15,
    (2.0 - ((double)12 / 60)) * 0.15,
    (2.0 - ((double)13 / 60)) * 0.15,
    (2.0 - ((double)14 / 60)) * 0.15,
    (2.0 - ((double)15 / 60)) * 0.15,
    (2.0 - ((double)16 / 60)) * 0.15,
    (2.0 - ((double)17 / 60)) * 0.15,
    (2.0 - ((double)18 / 60)) * 0.15,
    (2.0 - ((double)19 / 60)) * 0.15,
    (2.0 - ((double)20 / 60)) * 0.15,
    (2.0 - ((double)21 / 60)) * 0.15,
    (2.0 - ((double)22 / 60)) * 0.15,
    (2.0 - ((double)23 / 60)) * 0.15,
    (2.0 - ((double)24 / 60)) * 0.15,
    (2.0 - ((double)25 / 60)) * 0.15,
    (2.0 - ((double)26 / 60)) * 0.15,
    (2.0 - ((double)27 / 60)) * 0.15,
    (2.0 - ((double)28 / 60)) * 0.15,
    (2.0 - ((double)29 / 60)) * 0.15,
    (2.0 - ((double)30 / 60)) * 0.15,
    (2.0 - ((double)31 / 60)) * 0.15,
    (2.0 - ((double)32 / 60)) * 0.15,
    (2.0 - ((double)33 / 60)) * 0.15,
    (2.0 - ((double)34 / 60)) * 0.15,
    (2.0 - ((double)35 / 60)) * 0.15,
    (2.0 - ((double)36 / 60)) * 0.15,
    (2.0 - ((double)37 / 60)) * 0.15,
    (2.0 - ((double)38 / 60)) * 0.15,
    (2.0 - ((double)39 / 60)) * 0.15,
    (2.0 - ((double)40 / 60)) * 0.15,
    (2.0 - ((double)41 / 60)) * 0.15,
    (2.0 - ((double)42 / 60)) * 0.15,
    (2.0 - ((double)43 / 60)) * 0.15,
    (2.0 - ((double)44 / 60)) * 0.15,
    (2.0 - ((double)45 / 60)) * 0.15,
    (2.0 - ((double)46 / 60)) * 0.15,
    (2.0 - ((double)47 / 60)) * 0.15,
    (2.0 - ((double)48 / 60)) * 0.15,
    (2.0 - ((double)49 / 60)) * 0.15,
    (2.0 - ((double)50 / 60)) * 0.15,
    (2.0 - ((double)51 / 60)) * 0.15,
    (2.0 - ((double)52 / 60)) * 0.15,
    (2.0 - ((double)53 / 60)) * 0.15,
    (2.0 - ((double)54 / 60)) * 0.15,
    (2.0 - ((double)55 / 60)) * 0.15,
    (2.0 - ((double)56 / 60)) * 0.15,
    (2.0 - ((double)57 / 60)) * 0.15,
    (2.0 - ((double)58 / 60)) * 0.15,
    (2.0 - ((double)59 / 60)) * 0.15,
    (2.0 - ((double)60 / 60)) * 0.15
};

#ifdef PGPU_HOST_EMULATION
#define ML_ASSERT(x) do { if (!(x)) { fprintf(stderr, "k_dp_ml invariant violated: %s\n", #x); abort(); } } while (0)
#else
#define ML_ASSERT(x) do { } while (0)
#endif

// CSH: doubles of one half of the cs tile (rows per half = largest even number <= CSH / S)
template <int MINB, int CSH>
__global__ void __launch_bounds__(32 * kMlWarps, MINB) k_dp_ml(DevBatch B, const DevModel *__restrict__ models,
                                                               const int4 *__restrict__ groups, int n_groups, int prefetch) {
    struct __align__(16) WarpMem {
        double cs[2][CSH];       // bulk-copy destinations: 16-byte aligned
        MlK k[32];
        double rcv[3][32];       // per frame: best +start of the open forward ORF (value ...
        double rev[3][32];       // per frame: score of the last -STOP
        int32_t rcj[3][32];      // ... and node)
        int32_t rej[4];          // node of the last -STOP of every frame
        // -STOP target, per frame k and lane: the recorded overlapping -start (node sp, position n3n, its -STOP n3s), the
        // class range [sa, sb) of the +STOPs that can trigger the triple overlap with it, and the operon value -- kept
        // here instead of in 21 registers (only every eighth target is a -STOP; registers decide the occupancy)
        int32_t t_sp[3][32], t_n3n[3][32], t_n3s[3][32], t_sa[3][32], t_sb[3][32];
        double t_op[3][32];
        uint64_t bar[2];
    };
    __shared__ WarpMem s_w[kMlWarps];
    const int lane = threadIdx.x & 31, wslot = threadIdx.x >> 5;
    const int slot = blockIdx.x * kMlWarps + wslot;
    if (slot >= n_groups) return;
    constexpr int W = 32;
    const int4 G = groups[slot];  // x: first entry in ext_chains, y: number of chains (<= 32), z: extraction; longest first
    const int L = G.y;
    const bool act = lane < L;
    const int ll = act ? lane : 0;  // idle lanes shadow lane 0 (loads stay in bounds, stores are suppressed)
    const int chain = B.ext_chains[G.x + ll];
    const ChainInfo C = B.chains[chain];
    const int nn = C.nn;  // the same for every lane: one extraction
    if (nn == 0) {
        if (act) { B.chain_ipath[chain] = -1; B.chain_score[chain] = 0.0; }
        return;
    }
    const DevModel &M = models[C.model];
    // Arrays are addressed from the kernel parameters (constant bank) + two offsets instead of through eighteen
    // precomputed pointers: the pointers alone would take 36 registers, and occupancy is what hides the load latency
    // of the walk.  no: first node of the extraction; io: this lane's element of the extraction's interleaved block
    // (ChainInfo::ioff) -- element (x, lane) at io + x * S, where x is a node index (cs, opv, star_ptr, score, traceb,
    // ov_mark) or a merged-stream position (the DP-private source values / traceback positions / suffix maxima).
    const int no = C.node_off;
    const int64_t io = C.ioff;
    const int S = C.istride;
    const int fe0 = no + B.cbase[4 * G.z + 1];   // +STOPs in class order: node (clist), ndx (cndx), merged-stream position (feq)
    const int n_fe = B.cbase[4 * G.z + 2] - B.cbase[4 * G.z + 1];
    const int64_t rowb = io - C.lane;             // first element of row 0 of the extraction's interleaved block
    auto pf_row = [](const void *p, int bytes) {
        for (int o = 0; o < bytes; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"((const char *)p + o));
    };
#define ndx(j) B.ndx[no + (j)]
#define sv(j) B.stop_val[no + (j)]
#define cls(j) B.cls[no + (j)]
#define dpx(j) B.dpx[no + (j)]
#define ig_node(q) B.ig_node[no + (q)]
#define ig_ndx(q) B.ig_ndx[no + (q)]
#define fe_node(r) B.clist[fe0 + (r)]
#define fe_ndx(r) B.cndx[fe0 + (r)]
#define fe_q(r) B.feq[fe0 + (r)]
#define IL(x) (io + (int64_t)(x) * S)
#define opv(j, f) B.opv[3 * IL(j) + (f)]
#define star_ptr(j, f) B.star_ptr[3 * IL(j) + (f)]
#define score(j) B.score[IL(j)]
#define svig(q) B.dp_svig[IL(q)]   /* source value of a merged-stream entry (-DBL_MAX: nothing leads into it) */
#define tbx(q) B.dp_tbig[IL(q)]    /* +STOP entries: POSITION (ndx) of the node their traceback points to */
#define fmv(q) B.dp_fmv[IL(q)]
#define fmj(q) B.dp_fmj[IL(q)]
    WarpMem &wm = s_w[wslot];
    MlK *sk = wm.k;
    const double ig_neg = M.ig_neg, st_wt = M.st_wt;

    // ---- cs tile: chunks of `rows` target rows, chunk c in half c & 1, fetched one chunk ahead ----
    const int rows = max(2, (CSH / S) & ~1);
    const double *cs_block = B.cs + (io - C.lane);   // row r of the extraction at cs_block + r * S (16-byte aligned for even r)
    const uint32_t chunk_bytes = (uint32_t)rows * S * 8u;
    int chunk = 0, chunk_begin = 0;
    if (lane == 0) { mbar_init(&wm.bar[0], 1); mbar_init(&wm.bar[1], 1); mbar_init_fence(); }
    __syncwarp();
    if (lane == 0) {
        bulk_load(wm.cs[0], cs_block, chunk_bytes, &wm.bar[0]);
        if (rows < nn) bulk_load(wm.cs[1], cs_block + (int64_t)rows * S, chunk_bytes, &wm.bar[1]);
    }
    mbar_wait(&wm.bar[0], 0);

    // merged-stream cursors (uniform): cur = finalized entries, lo = first entry inside [i-1000, i), far = first
    // entry that is NOT more than 180 bp behind the target, split = boundary between front and back stack;
    // *_fe: the same positions counted in +STOPs only
    int cur = 0, lo = 0, far = 0, split = 0, cur_fe = 0, lo_fe = 0, far_fe = 0;
    double bk_v = kNeg;  // back stack: running maximum over the entries [split, far)
    int bk_j = -1;
    double front_top = kNeg;   // maximum of the front stack at the last flip
    // best terminal node of this lane's chain (lib.pyx:1239-1251: +STOP / -start nodes, largest index among equal maxima)
    double best_v = -1.0;
    int best_i = -1, best_tb = -1;
#pragma unroll
    for (int f = 0; f < 3; f++) { wm.rcv[f][lane] = kNeg; wm.rcj[f][lane] = -1; wm.rev[f][lane] = 0.0; }
    if (lane < 4) wm.rej[lane] = -1;

    auto ENT_ND = [&](int q) -> int { return ig_node(q); };
    auto ENT_NX = [&](int q) -> int { return ig_ndx(q); };
    auto ENT_SV = [&](int q) -> double { return svig(q); };

    for (int i0 = 0; i0 < nn; i0 += W) {
      __syncwarp();
      if (i0 + lane < nn) {
          const int i = i0 + lane;
          MlK k;
          k.ndx = ndx(i); k.sv = sv(i); k.cls = cls(i);
          k.leave = i > 2 * kMaxNodeDist ? cls_kind(cls(i - 2 * kMaxNodeDist - 1)) : -1;
          const int kind = cls_kind(k.cls);
          const int4 dx = dpx(i);
          k.a = k.b = k.c = -1; k.pad = 0;
          if (kind == K_FE || kind == K_RS) {
              const int wmin = B.win_min[no + i];                       // first node of the window
              const int wlo = B.crank[4 * (int64_t)(no + wmin) + 1];    // +STOPs before it
              if (kind == K_FE) { k.a = max(dx.x, wlo); }
              else { k.a = (dx.x >= wmin && dx.x >= 0 && dx.x < i) ? dx.x : -1; k.b = max(dx.y, wlo); k.c = dx.z; }
          } else if (kind == K_RE) {
              const int w0 = i - 2 * kMaxNodeDist;
              k.a = dx.x >= max(w0, 0) ? dx.x : -1; k.b = dx.y >= max(w0, 0) ? dx.y : -1; k.c = dx.z >= max(w0, 0) ? dx.z : -1;
          }
          sk[lane] = k;
          // Software prefetch into L2 of what the walk will read for this target and cannot have in cache: rows written
          // (or last touched) hundreds of steps ago are in DRAM by now, and a dependent DRAM round trip on the walk costs
          // about as much as a whole step.  Everything here is addressed geometrically, 1 .. 32 steps ahead of its use.
          // Only when the launch does not oversubscribe the SMs (`prefetch`): under full load the lines are evicted again
          // before they are used and the extra requests cost more than they save (measured: 18.3 -> 19.3 ms on the 630 Mbp
          // shard, but 10.5 -> 8.6 ms on the 1 000-contig batch, where the longest walk is the critical path).
          if (!prefetch) {
          } else if (kind == K_RE) {   // the recorded overlapping starts / operon values of this -STOP (per lane: one row)
              pf_row(&B.star_ptr[3 * (rowb + (int64_t)i * S)], S * 12);
              pf_row(&B.opv[3 * (rowb + (int64_t)i * S)], S * 24);
          } else if (kind == K_RS) {   // +STOPs of the 3' overlap range: source value, traceback position
              const int r1 = min(min(k.c, k.b + 4), n_fe);
              for (int r = k.b; r < r1; r++) {
                  const int64_t q = rowb + (int64_t)fe_q(r) * S;
                  pf_row(&B.dp_svig[q], S * 8);
                  pf_row(&B.dp_tbig[q], S * 4);
              }
          } else if (kind == K_FE) {   // +STOPs inside the ORF (operon): source value, recorded start, operon value
              const int r1 = min(k.a + 4, n_fe);
              for (int r = k.a; r < r1; r++) {
                  const int64_t q = rowb + (int64_t)fe_q(r) * S, j = rowb + (int64_t)fe_node(r) * S;
                  pf_row(&B.dp_svig[q], S * 8);
                  pf_row(&B.star_ptr[3 * j], S * 12);
                  pf_row(&B.opv[3 * j], S * 24);
              }
          }
      }
      __syncwarp();
      const int iend = min(i0 + W, nn);
      for (int i = i0; i < iend; i++) {
        const MlK &K = sk[i - i0];
        const int ci = K.cls, kind = cls_kind(ci), f2 = cls_frame(ci), ndx_i = K.ndx, sv_i = K.sv;
        lo += (K.leave == K_FE) | (K.leave == K_RS);
        lo_fe += K.leave == K_FE;
        // ---- cs tile: entering the next chunk (uniform) ----
        if (i >= chunk_begin + rows) {
            chunk++; chunk_begin += rows;
            mbar_wait(&wm.bar[chunk & 1], (chunk >> 1) & 1);   // every phase is observed exactly once
            // the half of the chunk just left is free (every lane is past its last read: __syncwarp at the end of a step)
            const int nxt = chunk_begin + rows;
            if (lane == 0 && nxt < nn) bulk_load(wm.cs[(chunk + 1) & 1], cs_block + (int64_t)nxt * S, chunk_bytes, &wm.bar[(chunk + 1) & 1]);
        }
        // cscore + sscore of a start target
        double cs_i = 0.0;
        if (kind == K_FS || kind == K_RS) cs_i = wm.cs[chunk & 1][(i - chunk_begin) * S + ll];
        double wv = kNeg;
        int wkey = -1;  // (node << 2) | (overlap frame + 1)
        // larger value, then larger node; the same node seen twice (far maximum + overlap re-evaluation) keeps the
        // overlap frame, as the reference evaluates it only once with it
        auto cand = [&](double v, int j, int fr) {
            const int key = (j << 2) | (fr + 1);
            if (v > wv || (v == wv && key > wkey)) { wv = v; wkey = key; }
        };

        if (kind == K_RS) {
            // own -STOP (gene, _connection.h:228-237): the last -STOP of this frame
            // (no stop codon of this frame lies inside the ORF, so the last -STOP node of the frame IS the own stop)
            if (K.a >= 0) { ML_ASSERT(wm.rej[f2] == K.a); cand(wm.rev[f2][lane] + cs_i, K.a, -1); }
            // +STOPs overlapping the 3' end (_connection.h:239-256)
            const double cs_diff = cs_i + ig_neg;
            const int re = min(K.c, cur_fe);
#pragma unroll 1
            for (int r = K.b; r < re; r++) {
                const int nj = fe_ndx(r), q = fe_q(r), j = fe_node(r);
                if (sv_i - 2 >= nj + 2) continue;
                const int ovlp = (nj + 2) - (sv_i - 2) + 1;
                if (ovlp >= kMaxOppOvlp) continue;
                if ((nj - sv_i) >= (ndx_i - nj + 3)) continue;
                const double s = ENT_SV(q);
                const int tx = tbx(q);
                if (s == kNeg) continue;
                if ((nj - sv_i) >= (sv_i - 3 - tx)) continue;
                cand(s + cs_diff, j, -1);
            }
        } else if (kind == K_FE) {
            {   // best +start of this ORF (gene): running maximum of fl(score + cscore + sscore)
                const int rj = wm.rcj[f2][lane];
                if (rj >= 0) cand(wm.rcv[f2][lane], rj, -1);
            }
            // +STOPs inside the ORF (operon, _connection.h:178-191)
#pragma unroll 1
            for (int r = K.a; r < cur_fe; r++) {
                const int q = fe_q(r), j = fe_node(r);
                const double s = ENT_SV(q);
                const int spj = star_ptr(j, f2);
                const double opj = opv(j, f2);
                if (s != kNeg && spj != -1) cand(s + opj, j, -1);
            }
        } else {  // K_FS, K_RE: intergenic sources
            double fm_v = kNeg;
            int fm_j = -1;
            const bool flip = lo > split;
            // -STOP only: the recorded overlapping -starts of this model (lane): node, its -STOP, and the class
            // range of the +STOPs that can trigger the triple overlap with it
            if (kind == K_RE) {
#pragma unroll
                for (int f = 0; f < 3; f++) {
                    const int sp = star_ptr(i, f);
                    wm.t_sp[f][lane] = sp;
                    if (sp != -1) {
                        const int4 d = dpx(sp);
                        wm.t_n3n[f][lane] = ndx(sp); wm.t_n3s[f][lane] = sv(sp); wm.t_op[f][lane] = opv(i, f);
                        wm.t_sa[f][lane] = d.y; wm.t_sb[f][lane] = d.z;
                    }
                }
            }
            // ---- entries that fall more than 180 bp behind move onto the back stack ----
            const int thr = ndx_i - 3 * kOperDist;
#pragma unroll 1
            while (far < cur && ENT_NX(far) < thr) {
                const int nd = ENT_ND(far);
                if (far >= lo) {
                    const double s = ENT_SV(far);
                    if (s != kNeg) {
                        const double x = s + ig_neg;
                        if (x >= bk_v) { bk_v = x; bk_j = nd & 0x7fffffff; }  // later entry wins a tie
                    }
                }
                far_fe += nd < 0;
                far++;
            }
            // ---- the window start passed the split: flip the back stack into suffix maxima ----
            if (flip) {
                double fv = kNeg;
                int fj = -1;
                // the entries are read once, here: four independent loads per round (values, then nodes)
                int q = far - 1;
#pragma unroll 1
                for (; q - 3 >= lo; q -= 4) {
                    double s8[4];
                    int n8[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) { s8[u] = ENT_SV(q - u); n8[u] = ENT_ND(q - u); }
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        if (s8[u] != kNeg) {
                            const double x = s8[u] + ig_neg;
                            if (x > fv) { fv = x; fj = n8[u] & 0x7fffffff; }  // earlier entry loses a tie
                        }
                        if (act) { fmv(q - u) = fv; fmj(q - u) = fj; }
                    }
                }
#pragma unroll 1
                for (; q >= lo; q--) {
                    const double s = ENT_SV(q);
                    if (s != kNeg) {
                        const double x = s + ig_neg;
                        if (x > fv) { fv = x; fj = ENT_ND(q) & 0x7fffffff; }
                    }
                    if (act) { fmv(q) = fv; fmj(q) = fj; }
                }
                split = far;
                bk_v = kNeg; bk_j = -1;
                front_top = fv;                             // the maximum of the whole front stack
                if (lo < split) { fm_v = fv; fm_j = fj; }   // what the loop stored last is the entry of `lo`
            } else if (lo < split && !(bk_j >= 0 && bk_v >= front_top)) {
                // The front stack's answer for this window start.  DP scores grow along a chain, so soon after a flip the
                // running maximum of the back stack exceeds everything the front stack holds (front_top bounds every
                // later suffix maximum; on a tie the back stack's later node wins anyway): the load -- a DRAM round trip,
                // the suffix maxima were written hundreds of steps ago -- is then skipped.
                fm_j = fmj(lo); fm_v = fmv(lo);
            }
            if (fm_j >= 0) cand(fm_v, fm_j, -1);
            if (bk_j >= 0) cand(bk_v, bk_j, -1);
            const int flo = max(far, lo);
            if (kind == K_FS) {
                // near sources (_connection.h:116-129): +STOP distance-dependent term, -start strand switch
#pragma unroll 1
                for (int q0 = flo; q0 < cur; q0 += 2) {   // two entries per round: their loads are in flight together
                    const int q1 = min(q0 + 1, cur - 1);   // a clamped slot repeats the last entry: harmless
                    const int nd0 = ENT_ND(q0), nj0 = ENT_NX(q0), nd1 = ENT_ND(q1), nj1 = ENT_NX(q1);
                    const double s0 = ENT_SV(q0), s1 = ENT_SV(q1);
                    if (s0 != kNeg && !(nd0 < 0 ? (nj0 + 2 >= ndx_i) : (nj0 >= ndx_i))) {
                        const int dist = ndx_i - nj0;
                        const double term = (nd0 >= 0 || dist > 3 * kOperDist) ? ig_neg : (dist <= kOperDist ? c_igA[dist] * st_wt : 0.0);
                        cand(s0 + term, nd0 & 0x7fffffff, -1);
                    }
                    if (s1 != kNeg && !(nd1 < 0 ? (nj1 + 2 >= ndx_i) : (nj1 >= ndx_i))) {
                        const int dist = ndx_i - nj1;
                        const double term = (nd1 >= 0 || dist > 3 * kOperDist) ? ig_neg : (dist <= kOperDist ? c_igA[dist] * st_wt : 0.0);
                        cand(s1 + term, nd1 & 0x7fffffff, -1);
                    }
                }
            } else {
                // +STOP with the triple-overlap search (_connection.h:297-334), for one merged-stream position
                auto eval_fe = [&](int q, double s, int nj, int j) {
                    const int left = nj + 2, right = ndx_i - 2;
                    if (left >= right) return;
                    int maxfr = -1, tj = kTbNone;
                    double maxval = 0.0;
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        if (wm.t_sp[k][lane] == -1) continue;
                        const int n3s = wm.t_n3s[k][lane];
                        const int ovlp = left - n3s + 3;
                        if (ovlp <= 0 || ovlp >= kMaxOppOvlp) continue;
                        if (ovlp >= wm.t_n3n[k][lane] - left) continue;
                        if (tj == kTbNone) tj = tbx(q);
                        if (ovlp >= n3s - tj - 2) continue;
                        const double op = wm.t_op[k][lane];
                        if (op > maxval) { maxfr = k; maxval = op; }
                    }
                    cand(s + (maxfr != -1 ? maxval : ig_neg), j, maxfr);
                };
#pragma unroll 1
                for (int q = flo; q < cur; q++) {
                    const int nd = ENT_ND(q), nj = ENT_NX(q);
                    const double s = ENT_SV(q);
                    if (s == kNeg) continue;
                    if (nd < 0) {
                        eval_fe(q, s, nj, nd & 0x7fffffff);
                    } else {  // -start (_connection.h:335-341)
                        if (nj >= ndx_i - 2) continue;
                        const int dist = ndx_i - nj;
                        cand(s + (dist > 3 * kOperDist ? ig_neg : (dist <= kOperDist ? c_igA[dist] * st_wt : 0.0)), nd, -1);
                    }
                }
                // far +STOPs whose position can trigger the triple overlap (ndx in [n3s-4, n3s+194], n3s = the -STOP
                // of a recorded start): their plain value is covered by the far maximum, the overlap value can only
                // be larger.  The recorded start differs per model, so the class range is per lane (the 3' overlap
                // range of that -start, k_dp_index, widened by the entries with ndx == n3s-4 and clamped to the far
                // zone); the loop runs over the union of the lanes' ranges.
                const int ffe = max(far_fe, lo_fe);  // +STOPs below the far / near boundary
                if (lo_fe < ffe) {
#pragma unroll 1
                    for (int k = 0; k < 3; k++) {
                        const int spk = wm.t_sp[k][lane], n3s = wm.t_n3s[k][lane];
                        int sa = wm.t_sa[k][lane];
                        const int sb = wm.t_sb[k][lane];
                        int a_ = 0, b_ = 0;
                        if (spk != -1) {
                            while (sa > lo_fe && fe_ndx(sa - 1) >= n3s - 4) sa--;
                            a_ = max(sa, lo_fe); b_ = min(sb, ffe);
                        }
                        const bool any = a_ < b_;
                        if (!__any_sync(0xffffffffu, any)) continue;
                        const int ua = __reduce_min_sync(0xffffffffu, any ? a_ : 0x7fffffff);
                        const int ub = __reduce_max_sync(0xffffffffu, any ? b_ : -1);
#pragma unroll 1
                        for (int r = ua; r < ub; r++) {
                            if (r < a_ || r >= b_) continue;
                            const int q = fe_q(r);
                            const double s = ENT_SV(q);
                            if (s != kNeg) eval_fe(q, s, fe_ndx(r), fe_node(r));
                        }
                    }
                }
                // -STOPs whose ORF spans this stop: operon (_connection.h:343-356); at most one per frame = the last
                // -STOP of that frame
                ML_ASSERT((K.a < 0 || wm.rej[0] == K.a) && (K.b < 0 || wm.rej[1] == K.b) && (K.c < 0 || wm.rej[2] == K.c));
                if (K.a >= 0 && wm.t_sp[0][lane] != -1) cand(wm.rev[0][lane] + wm.t_op[0][lane], K.a, -1);
                if (K.b >= 0 && wm.t_sp[1][lane] != -1) cand(wm.rev[1][lane] + wm.t_op[1][lane], K.b, -1);
                if (K.c >= 0 && wm.t_sp[2][lane] != -1) cand(wm.rev[2][lane] + wm.t_op[2][lane], K.c, -1);
            }
        }

        double sc_i = 0.0;
        int tb_i = -1, fr_i = -1;
        if (wkey >= 0 && wv >= 0.0) { sc_i = wv; tb_i = wkey >> 2; fr_i = (wkey & 3) - 1; }
        if (act) {
            // the scores themselves are only stored when somebody reads them (single mode: node records; self-check)
            if (B.score) score(i) = sc_i;
            B.traceb[IL(i)] = tb_i; B.ov_mark[IL(i)] = (int8_t)fr_i;
            if (kind == K_FE || kind == K_RS) {
                svig(cur) = tb_i == -1 ? kNeg : sc_i;  // edge-artifact rule: nothing leads into it
                if (kind == K_FE) tbx(cur) = tb_i >= 0 ? ndx(tb_i) : 0;
            }
        }
        if ((kind == K_FE || kind == K_RS) && sc_i >= best_v) { best_v = sc_i; best_i = i; best_tb = tb_i; }
        if (kind == K_FE) {
            cur++; cur_fe++;
            wm.rcv[f2][lane] = kNeg; wm.rcj[f2][lane] = -1;
        } else if (kind == K_RS) {
            cur++;
        } else if (kind == K_FS) {
            const double g = sc_i + cs_i;
            if (g >= wm.rcv[f2][lane]) { wm.rcv[f2][lane] = g; wm.rcj[f2][lane] = i; }
        } else {
            wm.rev[f2][lane] = sc_i;
            if (lane == 0) wm.rej[f2] = i;
        }
        __syncwarp();  // the per-frame slots written above are visible to every lane of the next step
      }
    }
    if (act) {   // lib.pyx:1311: no path when nothing leads into the best terminal node
        const bool ok = best_i >= 0 && best_tb != -1;
        B.chain_ipath[chain] = ok ? best_i : -1;
        B.chain_score[chain] = ok ? best_v : 0.0;
    }
}
#undef ndx
#undef sv
#undef cls
#undef dpx
#undef ig_node
#undef ig_ndx
#undef fe_node
#undef fe_ndx
#undef fe_q
#undef IL
#undef opv
#undef star_ptr
#undef score
#undef svig
#undef tbx
#undef fmv
#undef fmj

// arg-max of the DP score over the terminal node kinds (+STOP, -start), largest index among equal maxima
// (lib.pyx:1239-1251 scans from the end with a strict ">"); -1 when nothing leads into it (lib.pyx:1311)
__global__ void __launch_bounds__(128) k_chain_best(DevBatch B, int n_chains) {
    const int chain = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (chain >= n_chains) return;
    const ChainInfo C = B.chains[chain];
    const uint8_t *__restrict__ cls = B.cls + C.node_off;
    const Strided<double> score{B.score + C.ioff, C.istride};
    double bv = -1.0;
    int bi = -1;
    for (int i = lane; i < C.nn; i += 32) {
        const int k = cls_kind(cls[i]);
        if (k != K_FE && k != K_RS) continue;
        const double v = score[i];
        if (v >= bv) { bv = v; bi = i; }   // i increases within a lane
    }
    int fr = 0;
    if (bi < 0) bv = -DBL_MAX;
    warp_argmax(bv, bi, fr);
    if (lane == 0) {
        const bool ok = bi >= 0 && B.traceb[C.ioff + (int64_t)bi * C.istride] != -1;
        B.chain_ipath[chain] = ok ? bi : -1;
        B.chain_score[chain] = ok ? bv : 0.0;
    }
}

// --------------------------------------------------------------------------------------------------
// winner selection + traceback + gene extraction: one thread per contig
// --------------------------------------------------------------------------------------------------
struct NodeRef {
    const int32_t *ndx, *sv;
    const uint8_t *cls;
    double *cscore, *sscore, *rscore, *uscore, *tscore;   // chain-major (ChainInfo::coff)
    Strided<int32_t> traceb;                              // interleaved (ChainInfo::ioff)
    int32_t *tracef;
    int32_t *star_ptr;                                    // interleaved 3-vectors: node j, frame f at s3 * j + f
    int64_t s3;
    Strided<int8_t> ov_mark;
    uint8_t *elim;
    __device__ __forceinline__ int32_t sp(int node, int f) const { return star_ptr[s3 * node + f]; }
};

__device__ __forceinline__ double igm_nodes(const NodeRef &N, int a, int b, const DevModel &M) {
    // _intergenic_mod (_connection.h:81-91)
    const int ca = N.cls[a], cb = N.cls[b];
    if (((ca ^ cb) & CLS_REV) != 0) return M.ig_neg;
    const int sa = (ca & CLS_REV) ? -1 : 1;
    const int na = N.ndx[a], nb = N.ndx[b];
    const int dist = abs(na - nb);
    const bool overlap = na + 2 * sa >= nb;
    double r = 0.0;
    if (na + 2 == nb || na == nb + 1) {
        const int s = sa == 1 ? b : a;
        if (N.rscore[s] < 0) r -= N.rscore[s];
        if (N.uscore[s] < 0) r -= N.uscore[s];
    }
    if (dist > 3 * kOperDist) r -= 0.15 * M.st_wt;
    else if ((dist <= kOperDist && !overlap) || dist * 4 < kOperDist) r += M.igt[dist];   // (2.0 - dist / 60) * 0.15 * st_wt, tabulated (no FP64 division)
    return r;
}

struct TraceArgs {
    const int32_t *contig_chain_begin;  // [n_contigs + 1] chains of contig c
    int32_t *tracef;                    // per chain-node
    uint8_t *elim;                      // per chain-node
    pgpu_gene *genes;                   // gene buffers (final, after k_tweak)
    pgpu_gene *genes_raw;               // gene buffers as extracted (before k_tweak)
    const int64_t *gene_off;            // [n_contigs] offset of each contig's gene buffer
    pgpu_contig_summary *summary;       // [n_contigs]
    int32_t *winner_chain;              // [n_contigs] chain index of the winner or -1
    int meta;
    int max_overlap;
    // Lean main pass of meta mode: only the winner's chain has been scored in full, into arrays of its own (one slot per
    // contig).  wchains[c] is the winner's chain with coff = its slot (ChainInfo::coff_in = its offset in the main
    // pass); the score arrays below, tracef and elim are indexed through it.  nullptr: everything lives in B at C.coff.
    const ChainInfo *wchains;
    double *cscore, *sscore, *rscore, *uscore, *tscore;
};

// meta mode: the winner = strict ">" from -100 in bin order, so the lowest bin wins ties (lib.pyx:5331,5364)
__device__ __forceinline__ int pick_winner(const DevBatch &B, const TraceArgs &A, int c) {
    int win = -1;
    double max_score = -100.0;
    for (int k = A.contig_chain_begin[c]; k < A.contig_chain_begin[c + 1]; k++) {
        const int ip = B.chain_ipath[k];
        if (A.meta) {
            if (B.chains[k].nn > 0 && ip >= 0 && B.chain_score[k] > max_score) { max_score = B.chain_score[k]; win = k; }
        } else {
            win = k;  // single mode: exactly one chain, kept even when no path exists
        }
    }
    return win;
}

// lean main pass: pick the winners first and list their chains for the full scoring pass that precedes the traceback
__global__ void __launch_bounds__(128) k_winner(DevBatch B, int n_contigs, TraceArgs A, const int64_t *__restrict__ slot_off,
                                                 ChainInfo *__restrict__ wchains) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_contigs) return;
    const int win = pick_winner(B, A, c);
    A.winner_chain[c] = win;
    ChainInfo W;
    if (win >= 0) { W = B.chains[win]; W.coff_in = W.coff; }
    else { W = ChainInfo{}; W.contig = c; W.nn = 0; W.istride = 1; }
    W.coff = slot_off[c];
    wchains[c] = W;
}

__global__ void __launch_bounds__(64) k_trace(DevBatch B, const DevModel *__restrict__ models, int n_contigs,
                                               TraceArgs A) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_contigs) return;
    const int win = A.wchains ? A.winner_chain[c] : pick_winner(B, A, c);
    pgpu_contig_summary S = A.summary[c];
    S.n_genes = 0; S.n_nodes = 0; S.winner = -1; S.ipath = -1; S.score = 0.0;
    A.winner_chain[c] = win;
    if (win < 0) { A.summary[c] = S; return; }
    const ChainInfo C = B.chains[win];
    const DevModel &M = models[C.model];
    const int nn = C.nn;
    const int ipath = B.chain_ipath[win];
    S.winner = C.model; S.n_nodes = nn; S.ipath = ipath; S.score = ipath >= 0 ? B.chain_score[win] : 0.0;
    NodeRef N;
    N.ndx = B.ndx + C.node_off; N.sv = B.stop_val + C.node_off; N.cls = B.cls + C.node_off;
    const int64_t so = A.wchains ? A.wchains[c].coff : C.coff;   // where this chain's full scores / tracef / elim live
    N.cscore = (A.wchains ? A.cscore : B.cscore) + so; N.sscore = (A.wchains ? A.sscore : B.sscore) + so;
    N.rscore = (A.wchains ? A.rscore : B.rscore) + so; N.uscore = (A.wchains ? A.uscore : B.uscore) + so;
    N.tscore = (A.wchains ? A.tscore : B.tscore) + so;
    N.traceb = Strided<int32_t>{B.traceb + C.ioff, C.istride}; N.tracef = A.tracef + so;
    N.star_ptr = B.star_ptr + 3 * C.ioff; N.s3 = 3 * (int64_t)C.istride;
    N.ov_mark = Strided<int8_t>{B.ov_mark + C.ioff, C.istride}; N.elim = A.elim + so;
    if (nn == 0) { A.summary[c] = S; return; }

    // the reference untangles overlaps from the arg-max node even when that node has no traceback
    // (then both loops fall through); we only have work when a path exists
    if (ipath >= 0) {
        // pass 1: triple overlaps (lib.pyx:1258-1271)
        for (int path = ipath; N.traceb[path] != -1; path = N.traceb[path]) {
            const int nxt = N.traceb[path];
            if (cls_kind(N.cls[path]) == K_RE && cls_kind(N.cls[nxt]) == K_FE && N.ov_mark[path] != -1 &&
                N.ndx[path] > N.ndx[nxt]) {
                const int tmp = N.sp(path, N.ov_mark[path]);
                int i = tmp;
                while (N.ndx[i] != N.sv[tmp]) i--;
                N.traceb[path] = tmp;
                N.traceb[tmp] = i;
                N.ov_mark[i] = -1;
                N.traceb[i] = nxt;
            }
        }
        // pass 2: simple overlaps (lib.pyx:1273-1289)
        for (int path = ipath; N.traceb[path] != -1; path = N.traceb[path]) {
            const int nxt = N.traceb[path];
            const int kp = cls_kind(N.cls[path]), kn = cls_kind(N.cls[nxt]);
            if (kp == K_RS && kn == K_FE) {
                int i = path;
                while (N.ndx[i] != N.sv[path]) i--;
                N.traceb[path] = i;
                N.traceb[i] = nxt;
            }
            if (kp == K_FE && kn == K_FE) {
                N.traceb[path] = N.sp(nxt, N.ndx[path] % 3);
                N.traceb[N.traceb[path]] = nxt;
            }
            if (kp == K_RE && kn == K_RE) {
                N.traceb[path] = N.sp(path, N.ndx[nxt] % 3);
                N.traceb[N.traceb[path]] = nxt;
            }
        }
        // forward pointers (lib.pyx:1291-1295)
        for (int path = ipath; N.traceb[path] != -1; path = N.traceb[path]) N.tracef[N.traceb[path]] = path;

        // eliminate_bad_genes (dprog.c:306-335)
        int head = ipath;
        while (N.traceb[head] != -1) head = N.traceb[head];
        for (int p = head; N.tracef[p] != -1; p = N.tracef[p]) {
            const int nx = N.tracef[p], k = cls_kind(N.cls[p]);
            if (k == K_FE) N.sscore[nx] += igm_nodes(N, p, nx, M);
            if (k == K_RS) N.sscore[p] += igm_nodes(N, p, nx, M);
        }
        for (int p = head; N.tracef[p] != -1; p = N.tracef[p]) {
            const int nx = N.tracef[p], k = cls_kind(N.cls[p]);
            if (k == K_FS && N.cscore[p] + N.sscore[p] < 0) { N.elim[p] = 1; N.elim[nx] = 1; }
            if (k == K_RE && N.cscore[nx] + N.sscore[nx] < 0) { N.elim[p] = 1; N.elim[nx] = 1; }
        }

        // Genes._extract (lib.pyx:3231-3270)
        pgpu_gene *genes = A.genes_raw + A.gene_off[c];
        int ng = 0, begin = 0, end = 0, start_ndx = 0, stop_ndx = 0;
        for (int p = head; p != -1; p = N.tracef[p]) {
            if (N.elim[p] == 1) continue;
            const int cp = N.cls[p];
            bool emit = false;
            if (!(cp & CLS_REV)) {
                if (!cls_is_stop(cp)) { begin = N.ndx[p] + 1; start_ndx = p; }
                else { end = N.ndx[p] + 3; stop_ndx = p; emit = true; }
            } else {
                if (!cls_is_stop(cp)) { end = N.ndx[p] + 1; start_ndx = p; emit = true; }
                else { begin = N.ndx[p] - 1; stop_ndx = p; }
            }
            if (emit) { genes[ng].begin = begin; genes[ng].end = end; genes[ng].start_ndx = start_ndx; genes[ng].stop_ndx = stop_ndx; ng++; }
        }
        S.n_genes = ng;

        // Genes._tweak_final_starts runs gene-parallel in k_tweak
    }
    A.summary[c] = S;
}

// --------------------------------------------------------------------------------------------------
// Genes._tweak_final_starts (lib.pyx:3272-3401), one thread per gene.  In the reference the genes are tweaked in
// order and in place; gene i only ever reads (a) its predecessor's stop node and strand (never changed by a
// tweak), (b) its predecessor's *start* node position when the predecessor is on the reverse strand and gene i on
// the forward strand, (c) its successor's untouched start/stop.  So two passes reproduce the sequential result:
// pass 0 tweaks the reverse-strand genes (reading the original array), pass 1 the forward-strand genes (reading
// reverse predecessors from the pass-0 output).
// --------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_tweak(DevBatch B, const DevModel *__restrict__ models, int n_contigs, TraceArgs A,
                                                const pgpu_gene *__restrict__ orig, int64_t total_slots, int pass) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total_slots) return;
    // contig of this gene slot
    int lo = 0, hi = n_contigs - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (A.gene_off[mid] <= t) lo = mid; else hi = mid - 1;
    }
    const int c = lo;
    const int gi = (int)(t - A.gene_off[c]);
    const int ng = A.summary[c].n_genes;
    if (gi >= ng) return;
    const int win = A.winner_chain[c];
    const ChainInfo C = B.chains[win];
    const DevModel &M = models[C.model];
    const int nn = C.nn;
    NodeRef N;
    N.ndx = B.ndx + C.node_off; N.sv = B.stop_val + C.node_off; N.cls = B.cls + C.node_off;
    const int64_t so = A.wchains ? A.wchains[c].coff : C.coff;
    N.cscore = (A.wchains ? A.cscore : B.cscore) + so; N.sscore = (A.wchains ? A.sscore : B.sscore) + so;
    N.rscore = (A.wchains ? A.rscore : B.rscore) + so; N.uscore = (A.wchains ? A.uscore : B.uscore) + so;
    N.tscore = (A.wchains ? A.tscore : B.tscore) + so;
    N.traceb = Strided<int32_t>{nullptr, 1}; N.tracef = nullptr; N.star_ptr = nullptr; N.s3 = 3;
    N.ov_mark = Strided<int8_t>{nullptr, 1}; N.elim = nullptr;
    const pgpu_gene *og = orig + A.gene_off[c];   // untouched genes (Genes._extract)
    pgpu_gene *genes = A.genes + A.gene_off[c];    // output (pass 0: reverse genes, pass 1: forward genes)
    auto is_edge = [&](int x) { return (N.cls[x] & (CLS_EDGE | CLS_CONV)) != 0; };
    auto strand_of = [&](int x) { return (N.cls[x] & CLS_REV) ? -1 : 1; };
    const int cur = og[gi].start_ndx;
    const int scur = strand_of(cur);
    if ((pass == 0) != (scur == -1)) return;
    // neighbours as the sequential loop sees them
    int pstart = gi > 0 ? og[gi - 1].start_ndx : -1;
    // a reverse-strand predecessor was tweaked in pass 0 (a tweak keeps stop and strand, only the start moves)
    if (pass == 1 && pstart >= 0 && strand_of(pstart) == -1) pstart = genes[gi - 1].start_ndx;
    const int pstop = gi > 0 ? og[gi - 1].stop_ndx : -1;
    const int nstart = gi < ng - 1 ? og[gi + 1].start_ndx : -1, nstop = gi < ng - 1 ? og[gi + 1].stop_ndx : -1;
    pgpu_gene G = og[gi];
    {
        const double sc = N.sscore[cur] + N.cscore[cur];
        double igm0 = 0.0;
        if (pstart >= 0 && scur == 1 && strand_of(pstart) == 1) igm0 = igm_nodes(N, pstop, cur, M);
        if (pstart >= 0 && scur == 1 && strand_of(pstart) == -1) igm0 = M.ig_neg;
        if (nstart >= 0 && scur == -1 && strand_of(nstart) == 1) igm0 = M.ig_neg;
        if (nstart >= 0 && scur == -1 && strand_of(nstart) == -1) igm0 = igm_nodes(N, cur, nstop, M);

        int maxndx[2] = {-1, -1};
        double maxsc[2] = {0, 0}, maxigm[2] = {0, 0};
        for (int j = cur - 100; j < cur + 100; j++) {
            if (j < 0 || j >= nn || j == cur) continue;
            if (cls_is_stop(N.cls[j]) || N.sv[j] != N.sv[cur]) continue;
            const int sj = strand_of(j);
            double tigm = 0.0;
            if (pstart >= 0 && sj == 1 && strand_of(pstart) == 1) {
                if (N.ndx[pstop] - N.ndx[j] > A.max_overlap) continue;
                tigm = igm_nodes(N, pstop, j, M);
            }
            if (pstart >= 0 && sj == 1 && strand_of(pstart) == -1) {
                if (N.ndx[pstart] - N.ndx[j] >= 0) continue;
                tigm = M.ig_neg;
            }
            if (nstart >= 0 && sj == -1 && strand_of(nstart) == 1) {
                if (N.ndx[j] - N.ndx[nstart] >= 0) continue;
                tigm = M.ig_neg;
            }
            if (nstart >= 0 && sj == -1 && strand_of(nstart) == -1) {
                if (N.ndx[j] - N.ndx[nstop] > A.max_overlap) continue;
                tigm = igm_nodes(N, j, nstop, M);
            }
            const double csc = N.cscore[j] + N.sscore[j];
            if (maxndx[0] == -1) {
                maxndx[0] = j; maxsc[0] = csc; maxigm[0] = tigm;
            } else if (csc + tigm > maxsc[0]) {
                maxndx[1] = maxndx[0]; maxsc[1] = maxsc[0]; maxigm[1] = maxigm[0];
                maxndx[0] = j; maxsc[0] = csc; maxigm[0] = tigm;
            } else if (maxndx[1] == -1 || csc + tigm > maxsc[1]) {
                maxndx[1] = j; maxsc[1] = csc; maxigm[1] = tigm;
            }
        }
        for (int q = 0; q < 2; q++) {
            const int m = maxndx[q];
            if (m == -1) continue;
            if (N.tscore[m] < N.tscore[cur] && maxsc[q] - N.tscore[m] >= sc - N.tscore[cur] + M.st_wt &&
                N.rscore[m] > N.rscore[cur] && N.uscore[m] > N.uscore[cur] && N.cscore[m] > N.cscore[cur] &&
                abs(N.ndx[m] - N.ndx[cur]) > 15) {
                maxsc[q] += N.tscore[cur] - N.tscore[m];
            } else if (abs(N.ndx[m] - N.ndx[cur]) <= 15 &&
                       N.rscore[m] + N.tscore[m] > N.rscore[cur] + N.tscore[cur] && !is_edge(cur) && !is_edge(m)) {
                if (N.cscore[cur] > N.cscore[m]) maxsc[q] += N.cscore[cur] - N.cscore[m];
                if (N.uscore[cur] > N.uscore[m]) maxsc[q] += N.uscore[cur] - N.uscore[m];
                if (igm0 > maxigm[q]) maxsc[q] += igm0 - maxigm[q];
            } else {
                maxsc[q] = -1000.0;
            }
        }
        int pick = -1;
        for (int q = 0; q < 2; q++) {
            if (maxndx[q] == -1) continue;
            if (pick == -1 && maxsc[q] + maxigm[q] > sc + igm0) pick = q;
            else if (pick >= 0 && maxsc[q] + maxigm[q] > maxsc[pick] + maxigm[pick]) pick = q;
        }
        if (pick != -1) {
            const int m = maxndx[pick];
            G.start_ndx = m;
            if (strand_of(m) == 1) G.begin = N.ndx[m] + 1;
            else G.end = N.ndx[m] + 1;
        }
    }
    genes[gi] = G;
}

// --------------------------------------------------------------------------------------------------
// output packing
// --------------------------------------------------------------------------------------------------

// nodes of a chain list -> pgpu_node records.  `dp_state`: 1 = single mode (keep the DP state of the
// winning chain), 0 = meta mode (nodes were re-extracted and re-scored: score 0, traceb/tracef -1,
// star_ptr 0)
__device__ __forceinline__ void pack_node(const DevBatch &B, const ChainInfo &C, int64_t g, int i,
                                          const MotifOut *__restrict__ mot, const int32_t *__restrict__ tracef,
                                          const uint8_t *__restrict__ elim, int dp_state, pgpu_node *__restrict__ out) {
    const int c = B.cls[C.node_off + i];
    pgpu_node n;
    memset(&n, 0, sizeof(n));  // deterministic padding bytes: records are compared / hashed as raw bytes
    n.ndx = B.ndx[C.node_off + i];
    n.stop_val = B.stop_val[C.node_off + i];
    n.strand = (c & CLS_REV) ? -1 : 1;
    n.type = c & CLS_TYPE;
    n.edge = (c & (CLS_EDGE | CLS_CONV)) ? 1 : 0;  // flags after Nodes._score (lib.pyx:2431)
    n.gc_cont = cls_is_stop(c) ? 0.0f : B.gc_cont[C.node_off + i];
    n.rbs[0] = B.rbs[2 * g]; n.rbs[1] = B.rbs[2 * g + 1];
    const MotifOut m = mot ? mot[g] : MotifOut{};
    n.mot_score = m.score; n.mot_ndx = m.ndx; n.mot_len = m.len; n.mot_spacer = m.spacer; n.mot_spacendx = m.spacendx;
    n.cscore = B.cscore[g]; n.uscore = B.uscore[g]; n.tscore = B.tscore[g]; n.rscore = B.rscore[g]; n.sscore = B.sscore[g];
    const int64_t gi = C.ioff + (int64_t)i * C.istride;   // interleaved arrays (ChainInfo::ioff)
    if (dp_state == 1) {
        n.score = B.score[gi]; n.traceb = B.traceb[gi]; n.tracef = tracef[g]; n.ov_mark = B.ov_mark[gi];
        n.elim = elim[g];
        n.star_ptr[0] = B.star_ptr[3 * gi]; n.star_ptr[1] = B.star_ptr[3 * gi + 1]; n.star_ptr[2] = B.star_ptr[3 * gi + 2];
    } else {
        n.score = 0.0; n.traceb = -1; n.tracef = -1; n.ov_mark = -1; n.elim = 0;
        n.star_ptr[0] = n.star_ptr[1] = n.star_ptr[2] = 0;
        if (dp_state == 2) {  // scored + record_overlapping_starts, no DP (pgpu_score_nodes)
            n.star_ptr[0] = B.star_ptr[3 * gi]; n.star_ptr[1] = B.star_ptr[3 * gi + 1]; n.star_ptr[2] = B.star_ptr[3 * gi + 2];
        }
    }
    const uint64_t *src = reinterpret_cast<const uint64_t *>(&n);
    uint64_t *dst = reinterpret_cast<uint64_t *>(out + g);
#pragma unroll
    for (int q = 0; q < (int)(sizeof(pgpu_node) / 8); q++) dst[q] = src[q];
}

__global__ void __launch_bounds__(128) k_pack_nodes(DevBatch B, int n_chains, int64_t total,
                                                     const MotifOut *__restrict__ mot, const int32_t *__restrict__ tracef,
                                                     const uint8_t *__restrict__ elim, int dp_state,
                                                     pgpu_node *__restrict__ out) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total) return;
    int lo = 0, hi = n_chains - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (B.chains[mid].coff <= g) lo = mid; else hi = mid - 1;
    }
    const ChainInfo C = B.chains[lo];
    const int i = (int)(g - C.coff);
    if (i >= C.nn) return;
    pack_node(B, C, g, i, mot, tracef, elim, dp_state, out);
}

// records of the start and stop node of every listed gene only (meta mode without node arrays; B = final-pass
// batch, chain index == contig), written at their usual place out[coff + node]
__global__ void __launch_bounds__(128) k_pack_nodes_genes(DevBatch B, const int2 *__restrict__ list, const int *__restrict__ count,
                                                           const pgpu_gene *__restrict__ genes,
                                                           const int64_t *__restrict__ gene_off,
                                                           const MotifOut *__restrict__ mot, pgpu_node *__restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if ((t >> 1) >= *count) return;
    const int2 e = list[t >> 1];
    const ChainInfo C = B.chains[e.x];
    const pgpu_gene G = genes[gene_off[e.x] + e.y];
    const int i = (t & 1) ? G.stop_ndx : G.start_ndx;
    pack_node(B, C, C.coff + i, i, mot, nullptr, nullptr, 0, out);
}

// meta mode: the winner's nodes are re-extracted and re-scored from scratch (lib.pyx:5380-5394); this
// builds the chain list of that final scoring pass (first_pass = 1: fresh extraction)
__global__ void k_build_final_chains(DevBatch B, int n_contigs, const int32_t *__restrict__ winner_chain,
                                     const int64_t *__restrict__ fin_coff, ChainInfo *__restrict__ fin) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_contigs) return;
    const int w = winner_chain[c];
    ChainInfo F;
    if (w >= 0) { F = B.chains[w]; F.first_pass = 1; }
    else { F = ChainInfo{}; F.contig = c; F.nn = 0; }
    F.coff = fin_coff[c];
    F.ioff = F.coff; F.istride = 1; F.lane = 0;   // the final pass scores one chain per contig: nothing to interleave
    fin[c] = F;
}

// start/stop node records of every gene (what Gene accessors read), 2 per gene
__global__ void __launch_bounds__(128) k_pack_gene_nodes(int n_contigs, const pgpu_contig_summary *__restrict__ summary,
                                                          const pgpu_gene *__restrict__ genes,
                                                          const int64_t *__restrict__ gene_off,
                                                          const int64_t *__restrict__ gene_out_off,
                                                          const int64_t *__restrict__ node_out_off,
                                                          const pgpu_node *__restrict__ nodes, pgpu_node *__restrict__ out,
                                                          pgpu_gene *__restrict__ genes_out) {
    const int c = blockIdx.x;
    const int ng = summary[c].n_genes;
    for (int g = threadIdx.x; g < ng; g += blockDim.x) {
        const pgpu_gene G = genes[gene_off[c] + g];
        genes_out[gene_out_off[c] + g] = G;
        // raw 8-byte copies: a struct assignment may skip padding bytes, records must be byte-identical
        static_assert(sizeof(pgpu_node) % 8 == 0, "pgpu_node size");
        const uint64_t *s0 = reinterpret_cast<const uint64_t *>(nodes + node_out_off[c] + G.start_ndx);
        const uint64_t *s1 = reinterpret_cast<const uint64_t *>(nodes + node_out_off[c] + G.stop_ndx);
        uint64_t *d0 = reinterpret_cast<uint64_t *>(out + 2 * (gene_out_off[c] + g));
#pragma unroll
        for (int q = 0; q < (int)(sizeof(pgpu_node) / 8); q++) { d0[q] = s0[q]; d0[sizeof(pgpu_node) / 8 + q] = s1[q]; }
    }
}

// the skip filter as a stand-alone operator (impl/generic.h:29-36), via the class table
__global__ void k_skippable(int n, const int8_t *__restrict__ strand, const uint8_t *__restrict__ type,
                            const int32_t *__restrict__ ndx, int mn, int i, uint8_t *__restrict__ skip) {
    const int j = mn + blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= i) return;
    const int k1 = 2 * (strand[j] != 1) + (type[j] == 3), k2 = 2 * (strand[i] != 1) + (type[i] == 3);
    const bool same_frame = ndx[j] % 3 == ndx[i] % 3;
    bool ok;
    switch (k2) {
    case K_FS: ok = (k1 == K_FE || k1 == K_RS); break;
    case K_FE: ok = ((k1 == K_FS && same_frame) || k1 == K_FE); break;
    case K_RS: ok = ((k1 == K_RE && same_frame) || k1 == K_FE); break;
    default: ok = (k1 == K_FE || k1 == K_RS || k1 == K_RE); break;
    }
    skip[j] = ok ? 0 : 1;
}

// the same with the arrays of the reference's plug-in ABI (skippable_t, lib.pxd:120): strand (1 / 255), type, frame of the
// window [mn, i]; element 0 = node mn
__global__ void k_skippable_plugin(const uint8_t *__restrict__ strand, const uint8_t *__restrict__ type,
                                   const uint8_t *__restrict__ frame, int cnt, uint8_t *__restrict__ skip) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;   // relative to mn; the target is element cnt
    if (j >= cnt) return;
    const int k1 = 2 * (strand[j] != 1) + (type[j] == 3), k2 = 2 * (strand[cnt] != 1) + (type[cnt] == 3);
    const bool same_frame = frame[j] == frame[cnt];
    bool ok;
    switch (k2) {
    case K_FS: ok = (k1 == K_FE || k1 == K_RS); break;
    case K_FE: ok = ((k1 == K_FS && same_frame) || k1 == K_FE); break;
    case K_RS: ok = ((k1 == K_RE && same_frame) || k1 == K_FE); break;
    default: ok = (k1 == K_FE || k1 == K_RS || k1 == K_RE); break;
    }
    skip[j] = ok ? 0 : 1;
}

// --------------------------------------------------------------------------------------------------
// launch wrappers
// --------------------------------------------------------------------------------------------------
void launch_dp(const DevBatch &B, const DevModel *models, const int32_t *order, int n_chains, int final, int algo,
               cudaStream_t st) {
    if (n_chains == 0) return;
    // final scoring: k_dp_dq (one warp per chain) unless algo 0 asks for the all-pairs kernel, which is also the training DP
    if (final && algo >= 1 && B.dp_svig) {
        const int nb = (n_chains + kFastWarps - 1) / kFastWarps;
        // <4>: 93 registers, no spills -- also whenever the chains fit into one wave at that occupancy (single chains are
        // pure latency: spills only cost); <8>: 64 registers for batches of many chains
        if (algo == 4 || n_chains <= 148 * 4 * kFastWarps) k_dp_dq<4, 1><<<nb, 32 * kFastWarps, 0, st>>>(B, models, order, n_chains);
        else k_dp_dq<8, 1><<<nb, 32 * kFastWarps, 0, st>>>(B, models, order, n_chains);
        k_chain_best<<<(n_chains * 32 + 127) / 128, 128, 0, st>>>(B, n_chains);
    } else if (!final && algo >= 1 && B.dp_svig) {   // training DP on the deque formulation
        k_dp_dq<4, 0><<<(n_chains + kFastWarps - 1) / kFastWarps, 32 * kFastWarps, 0, st>>>(B, models, order, n_chains);
        k_chain_best<<<(n_chains * 32 + 127) / 128, 128, 0, st>>>(B, n_chains);
    }
    else if (final) k_dp<1><<<n_chains, kDpThreads, 0, st>>>(B, models, order, n_chains);
    else k_dp<0><<<n_chains, kDpThreads, 0, st>>>(B, models, order, n_chains);
}
void launch_dp_ml(const DevBatch &B, const DevModel *models, const int4 *groups, int n_groups, int n_chains, int minb,
                  cudaStream_t st) {
    if (n_groups == 0 || n_chains == 0) return;
    const int nb = (n_groups + kMlWarps - 1) / kMlWarps;
    // software prefetch only while every walk has an SM slot from the start (148 SMs x 24 resident warps)
    const int prefetch = n_groups <= 148 * 24 ? 1 : 0;
    if (minb == 8) k_dp_ml<8, 64><<<nb, 32 * kMlWarps, 0, st>>>(B, models, groups, n_groups, prefetch);   // 64 registers, 32 warps / SM
    else k_dp_ml<6, 64><<<nb, 32 * kMlWarps, 0, st>>>(B, models, groups, n_groups, prefetch);             // 80 registers, 24 warps / SM
    // (the best terminal node of every chain is tracked by the walk itself: no k_chain_best pass)
}
// PGPU_DP_VERIFY: element-wise comparison of two DP results (score, traceback, overlap frame)
__global__ void k_dp_compare(const double *__restrict__ sa, const double *__restrict__ sb, const int32_t *__restrict__ ta,
                             const int32_t *__restrict__ tb, const int8_t *__restrict__ oa, const int8_t *__restrict__ ob,
                             int64_t n, unsigned long long *bad) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    if (!(sa[g] == sb[g]) || ta[g] != tb[g] || oa[g] != ob[g]) { atomicAdd(bad, 1ULL); atomicMin(bad + 1, (unsigned long long)g); }
}
void launch_dp_compare(const double *sa, const double *sb, const int32_t *ta, const int32_t *tb, const int8_t *oa,
                       const int8_t *ob, int64_t n, unsigned long long *bad, cudaStream_t st) {
    if (n > 0) k_dp_compare<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(sa, sb, ta, tb, oa, ob, n, bad);
}
void launch_trace(const DevBatch &B, const DevModel *models, int n_contigs, const int32_t *contig_chain_begin,
                  int32_t *tracef, uint8_t *elim, pgpu_gene *genes, pgpu_gene *genes_raw, const int64_t *gene_off,
                  int64_t total_gene_slots, pgpu_contig_summary *summary, int32_t *winner_chain, int meta, int max_overlap,
                  const DevBatch *W, const std::function<void()> &score_winners, const int64_t *slot_off, cudaStream_t st) {
    if (n_contigs == 0) return;
    TraceArgs A = {contig_chain_begin, tracef, elim, genes, genes_raw, gene_off, summary, winner_chain, meta, max_overlap,
                   nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    if (W) {
        // lean main pass: winners first, then their chains are scored in full (W: the batch view of that pass, its chain
        // list is filled here), then the traceback reads those scores
        k_winner<<<(n_contigs + 127) / 128, 128, 0, st>>>(B, n_contigs, A, slot_off, W->chains);
        score_winners();
        A.wchains = W->chains;
        A.cscore = W->cscore; A.sscore = W->sscore; A.rscore = W->rscore; A.uscore = W->uscore; A.tscore = W->tscore;
    }
    k_trace<<<(n_contigs + 63) / 64, 64, 0, st>>>(B, models, n_contigs, A);
    if (total_gene_slots > 0) {
        const unsigned nb = (unsigned)((total_gene_slots + 127) / 128);
        k_tweak<<<nb, 128, 0, st>>>(B, models, n_contigs, A, genes_raw, total_gene_slots, 0);
        k_tweak<<<nb, 128, 0, st>>>(B, models, n_contigs, A, genes_raw, total_gene_slots, 1);
    }
}
void launch_pack_nodes(const DevBatch &B, int n_chains, int64_t total, const void *mot, const int32_t *tracef,
                       const uint8_t *elim, int dp_state, pgpu_node *out, cudaStream_t st) {
    if (n_chains == 0 || total == 0) return;
    k_pack_nodes<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(B, n_chains, total, (const MotifOut *)mot, tracef, elim,
                                                                   dp_state, out);
}
void launch_pack_nodes_genes(const DevBatch &B, const int2 *list, const int *count, int64_t gene_cap, const pgpu_gene *genes,
                             const int64_t *gene_off, const void *mot, pgpu_node *out, cudaStream_t st) {
    if (gene_cap == 0) return;
    k_pack_nodes_genes<<<(unsigned)((2 * gene_cap + 127) / 128), 128, 0, st>>>(B, list, count, genes, gene_off,
                                                                              (const MotifOut *)mot, out);
}
void launch_pack_gene_nodes(int n_contigs, const pgpu_contig_summary *summary, const pgpu_gene *genes,
                            const int64_t *gene_off, const int64_t *gene_out_off, const int64_t *node_out_off,
                            const pgpu_node *nodes, pgpu_node *out, pgpu_gene *genes_out, cudaStream_t st) {
    if (n_contigs == 0) return;
    k_pack_gene_nodes<<<n_contigs, 64, 0, st>>>(n_contigs, summary, genes, gene_off, gene_out_off, node_out_off, nodes, out,
                                                genes_out);
}
void launch_build_final_chains(const DevBatch &B, int n_contigs, const int32_t *winner_chain, const int64_t *fin_coff,
                               ChainInfo *fin_chains, cudaStream_t st) {
    if (n_contigs > 0) k_build_final_chains<<<(n_contigs + 127) / 128, 128, 0, st>>>(B, n_contigs, winner_chain, fin_coff, fin_chains);
}
void launch_skippable(int n, const int8_t *strand, const uint8_t *type, const int32_t *ndx, int mn, int i, uint8_t *skip,
                      cudaStream_t st) {
    if (i > mn) k_skippable<<<(i - mn + 127) / 128, 128, 0, st>>>(n, strand, type, ndx, mn, i, skip);
}
void launch_skippable_plugin(const uint8_t *strand, const uint8_t *type, const uint8_t *frame, int cnt, uint8_t *skip,
                             cudaStream_t st) {
    if (cnt > 0) k_skippable_plugin<<<(cnt + 127) / 128, 128, 0, st>>>(strand, type, frame, cnt, skip);
}

}  // namespace pgpu
