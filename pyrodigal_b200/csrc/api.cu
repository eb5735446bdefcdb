// api.cu -- context, host orchestration and the C ABI of libpyrodigal_b200.so (see include/pyrodigal_b200.h).
//
// Host-side control flow of one batch mirrors GeneFinder._find_genes_meta / _find_genes_single
// (src/pyrodigal/lib.pyx:5281-5396) but evaluates every (contig, model) chain of the batch in the same
// kernel launches instead of looping over bins per contig:
//   H2D -> encode -> [sync: gc] -> plan chains -> mark -> scan -> [sync: node counts] -> fill -> prep ->
//   coding score -> start score -> overlapping starts -> DP -> winner/traceback/genes -> final re-score
//   (meta) -> pack -> [sync: gene counts] -> D2H.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <numeric>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "kernels.cuh"
#include "codon_masks.hpp"
#include "extract_device.cuh"
#include "train_host.hpp"  // RawTraining + host half of the training path

using namespace pgpu;

// pinned host buffers for results: recycled through a small free list so that steady-state calls neither
// allocate pinned memory nor zero-fill / re-copy result arrays.  Reference counted: results keep the pool alive
// even when their context is destroyed first.
struct PinnedBuf { void *p = nullptr; size_t cap = 0; };
struct PinnedPool {
    std::mutex mu;
    std::vector<PinnedBuf> free_list;
    PinnedBuf acquire(size_t bytes) {
        {
            std::lock_guard<std::mutex> g(mu);
            int best = -1;
            for (int i = 0; i < (int)free_list.size(); i++)
                if (free_list[i].cap >= bytes && (best < 0 || free_list[i].cap < free_list[best].cap)) best = i;
            if (best >= 0) {
                PinnedBuf b = free_list[best];
                free_list.erase(free_list.begin() + best);
                return b;
            }
        }
        PinnedBuf b;
        const size_t want = bytes + bytes / 4 + (1 << 16);
        if (cudaHostAlloc(&b.p, want, cudaHostAllocDefault) != cudaSuccess) { b.p = nullptr; return b; }
        b.cap = want;
        return b;
    }
    void release(PinnedBuf b) {
        if (!b.p) return;
        std::lock_guard<std::mutex> g(mu);
        if (free_list.size() >= 16) {  // bound the pinned footprint: keep the larger buffers (a caller that pipelines
                                       // several results -- distributed.gather_result -- keeps three or four in rotation)
            int s = 0;
            for (int i = 1; i < (int)free_list.size(); i++) if (free_list[i].cap < free_list[s].cap) s = i;
            if (free_list[s].cap < b.cap) { cudaFreeHost(free_list[s].p); free_list[s] = b; }
            else cudaFreeHost(b.p);
            return;
        }
        free_list.push_back(b);
    }
    ~PinnedPool() { for (auto &b : free_list) cudaFreeHost(b.p); }
};

struct pgpu_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    int n_models = 0;
    std::vector<DevModel> h_models;
    std::vector<double> model_gc;     // compact copies for the per-contig planning loop
    std::vector<int> model_tt;
    RawTraining *d_raw = nullptr;
    DevModel *d_models = nullptr;
    uint32_t *d_live = nullptr; // [n_models][kMotifWords]: motif cells with a weight other than the floor (DevModel::mot_live) + window table (mot_hit)
    double *d_dcT = nullptr;   // dicodon weights transposed: [4096][kDcCols], columns sorted by (tt, gc); null if n_models > kDcCols
    double *d_dcS = nullptr;   // the same weights as table sets of four neighbouring columns: [n_models][4096][4] (k_coding_flat)
    int coding_smem = 1;       // PGPU_CODING_SMEM: 1 = k_coding_flat for multi-model batches, 0 = k_coding_orf
    size_t ws_limit = 0;
    cudaEvent_t ev[16];
    int64_t launches = 0;
    int dp_ml_minb = 6;        // k_dp_ml register budget: min CTAs/SM 6, 8 or 10 (PGPU_DP_ML_MINB); 6 = 78 registers, no spills, 24 warps / SM
    int extract_algo = 2;      // 2: bit-parallel extraction (k_codon_bits + k_extract_b), 1: warp-cooperative k_extract_w
                               // (PGPU_EXTRACT_ALGO; batches with N-run masks always use 1)
    int final_algo = 2;        // meta mode without node arrays: 2 = final scoring pass over the genes' ORFs only,
                               // 1 = over every node as with want_nodes (PGPU_FINAL_ALGO)
    bool codon_lut = false;    // PGPU_CODON_LUT=1: k_codon_bits reads codon flags from a per-table byte table (written
                               // after the last GPU run of round 1: logic checked by the host emulation only, so off)
    bool coding_verify = false;   // PGPU_CODING_VERIFY=1: run k_coding_orf after k_coding_flat and fail on any differing raw coding score
    bool dp_verify = false;    // PGPU_DP_VERIFY=1: run k_dp_dq after k_dp_ml and fail on any difference (self-check)
    int dp_algo = 5;           // 5: k_dp_ml for multi-model batches, k_dp_dq otherwise (default); 6: k_dp_ml always;
                               // 1-4: k_dp_dq, 0: all-pairs k_dp (PGPU_DP_ALGO=n)
    std::shared_ptr<struct PinnedPool> pinned;  // shared with the results it backs (they may outlive the context)
    // Optional: host-input calls (pgpu_find_genes_batch) on large batches run as sub-batches on two worker threads /
    // streams ("lanes"), so that the H2D copy, the host planning gaps and the D2H of one hide under the kernels of the other.
    // Device-resident batches (pgpu_batch_run) stay on the single stream, so per-kernel timings remain well defined.
    int lanes = 2;                         // PGPU_LANES=1 disables.  Measured on the 630 Mbp bench shard (round 2): 98.8 ms per
                                           // end-to-end step with two lanes against 105 ms with one (the 11 ms input copy of
                                           // the second half hides under the kernels of the first)
    int64_t lane_min_bp = int64_t(64) << 20;   // smaller batches are not split (PGPU_LANE_MIN_BP)
    cudaStream_t lane_stream[2] = {nullptr, nullptr};
    cudaEvent_t lane_ev[2][16];
    cudaEvent_t fork_ev = nullptr, join_ev[2] = {nullptr, nullptr};
};

static thread_local std::string g_create_err;

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(_e);                         \
            return PGPU_ECUDA;                                                                     \
        }                                                                                          \
    } while (0)

static int fail(pgpu_ctx *ctx, int code, const std::string &msg) {
    if (ctx) ctx->err = msg;
    return code;
}

// ------------------------------------------------------------------------------------------------
// stream-ordered device buffers
// ------------------------------------------------------------------------------------------------
struct DevPool {
    pgpu_ctx *ctx;
    std::vector<void *> ptrs;
    size_t bytes = 0;
    bool failed = false;
    explicit DevPool(pgpu_ctx *c) : ctx(c) {}
    template <typename T>
    T *alloc(size_t n, bool zero = false) {
        size_t sz = std::max<size_t>(n, 1) * sizeof(T);
        sz = (sz + 255) & ~size_t(255);
        void *p = nullptr;
        cudaError_t e = cudaMallocAsync(&p, sz, ctx->stream);
        if (e != cudaSuccess) {
            failed = true;
            ctx->err = std::string("cudaMallocAsync(") + std::to_string(sz) + "): " + cudaGetErrorString(e);
            return nullptr;
        }
        ptrs.push_back(p);
        bytes += sz;
        if (zero) cudaMemsetAsync(p, 0, sz, ctx->stream);
        return (T *)p;
    }
    // Host-to-device uploads are staged through recycled pinned memory: a copy from pageable memory makes the host
    // wait until the stream has drained up to it and then runs at bounce-buffer speed, which left the GPU idle between
    // the planning steps; from pinned memory the copy is queued and the host keeps issuing work.
    std::vector<PinnedBuf> stage;
    size_t stage_used = 0;
    void *stage_alloc(size_t bytes) {
        bytes = (bytes + 63) & ~size_t(63);
        if (stage.empty() || stage_used + bytes > stage.back().cap) {
            PinnedBuf b = ctx->pinned->acquire(std::max<size_t>(bytes, size_t(16) << 20));
            if (!b.p) return nullptr;
            stage.push_back(b);
            stage_used = 0;
        }
        void *p = (char *)stage.back().p + stage_used;
        stage_used += bytes;
        return p;
    }
    void copy_in(void *dst, const void *src, size_t bytes) {
        if (!bytes) return;
        void *h = stage_alloc(bytes);
        if (h) { memcpy(h, src, bytes); src = h; }
        cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream);
    }
    template <typename T>
    T *upload(const std::vector<T> &v) {
        T *p = alloc<T>(v.size());
        if (p) copy_in(p, v.data(), v.size() * sizeof(T));
        return p;
    }
    // frees are stream ordered; the staging buffers go back to the pinned pool, so the caller must have synchronised
    // the stream after its last upload (every exit of run_range does, or has failed)
    void release() {
        for (void *p : ptrs) cudaFreeAsync(p, ctx->stream);
        ptrs.clear();
        if (!stage.empty()) {
            cudaStreamSynchronize(ctx->stream);
            for (auto &b : stage) ctx->pinned->release(b);
            stage.clear();
        }
    }
    ~DevPool() { release(); }
};

// ------------------------------------------------------------------------------------------------
// model preparation (host): everything that needs libm or is model-only
// ------------------------------------------------------------------------------------------------

// Model-independent part of the Shine-Dalgarno search: the set of motif bins the reference's two
// enumerations (lib.pyx:827-888 exact, 928-977 one mismatch) can report for a window at offset `off`
// (pos = start-20+off) whose in-range A/G matches are the bits of `gp`.
static uint32_t g_sd_masks[2][15][64];
static bool g_sd_ready = false;
static void build_sd_masks() {
    if (g_sd_ready) return;
    static const int8_t tab[15][4] = {{0}, {0}, {0}, {0}, {0}, {0}, {13, 6, 1, 2}, {0}, {15, 12, 11, 3},
                                      {16, 12, 11, 3}, {0}, {22, 21, 20, 10}, {24, 23, 20, 10}, {0}, {27, 26, 25, 10}};
    for (int off = 0; off < 15; off++) {
        const int d = 20 - off;  // start - pos
        const int limit = std::min(6, d - 4);
        for (int gp = 0; gp < 64; gp++) {
            int match[6];
            // exact
            for (int i = 0; i < 6; i++) match[i] = (i < limit && ((gp >> i) & 1)) ? (i % 3 == 0 ? 2 : 3) : -10;
            uint32_t mask = 1;
            for (int len = limit; len > 2; len--)
                for (int j = 0; j <= limit - len; j++) {
                    int ctr = -2;
                    for (int k = j; k < j + len; k++) ctr += match[k];
                    if (ctr < 6) continue;
                    const int rdis = d - j - len;
                    int flag;
                    if (rdis < 5) flag = len < 5 ? 2 : 1;
                    else if (rdis < 11) flag = 0;
                    else if (rdis < 13) flag = len < 5 ? 1 : 2;
                    else if (rdis < 16) flag = 3;
                    else continue;
                    const bool known = ctr == 6 || ctr == 8 || ctr == 9 || ctr == 11 || ctr == 12 || ctr == 14;
                    mask |= 1u << (known ? tab[ctr][flag] : 0);
                }
            g_sd_masks[0][off][gp] = mask;
            // one mismatch
            for (int i = 0; i < 6; i++) {
                if (i < limit) match[i] = ((gp >> i) & 1) ? (i % 3 == 0 ? 2 : 3) : (i % 3 == 0 ? -3 : -2);
                else match[i] = -10;
            }
            mask = 1;
            int cur_val = 0;  // sticky across iterations, as in the reference
            for (int len = limit; len > 4; len--)
                for (int j = 0; j <= limit - len; j++) {
                    int ctr = -2, mism = 0;
                    for (int k = j; k < j + len; k++) {
                        ctr += match[k];
                        if (match[k] < 0) { mism++; if (k <= j + 1 || k >= j + len - 2) ctr -= 10; }
                    }
                    if (mism != 1 || ctr < 6) continue;
                    const int rdis = d - j - len;
                    int flag;
                    if (rdis < 5) flag = 1;
                    else if (rdis < 11) flag = 0;
                    else if (rdis < 13) flag = 2;
                    else if (rdis < 16) flag = 3;
                    else continue;
                    static const int8_t v6[4] = {9, 5, 4, 2}, v7[4] = {14, 8, 7, 2}, v9[4] = {19, 18, 17, 3};
                    if (ctr == 6) cur_val = v6[flag];
                    else if (ctr == 7) cur_val = v7[flag];
                    else if (ctr == 9) cur_val = v9[flag];
                    mask |= 1u << cur_val;
                }
            g_sd_masks[1][off][gp] = mask;
        }
    }
    g_sd_ready = true;
}

// Host side of DevModel::mot_live / mot_hit, kMotifWords words per model.
// bits[0 .. 2047]: 1 bit per motif cell: weight != -4.0.
// bits[2048 ..]  : 4096 uint16, indexed by six upstream bases x (2 bits each, the first base in the low bits): bit
//   4 l + o is set when a motif of length l + 3 that starts at base o of the window MAY be live -- some live cell
//   [l][*][index] agrees with the window on the min(l + 3, 6 - o) bases of the motif that lie inside the window.
constexpr int kMotifWords = 4096;
static void motif_live_bits(const RawTraining &r, uint32_t *bits) {
    const double *w = &r.mot_wt[0][0][0];
    for (int k = 0; k < kMotifWords; k++) bits[k] = 0;
    for (int c = 0; c < 4 * 4 * 4096; c++)
        if (!(w[c] == -4.0)) bits[c >> 5] |= 1u << (c & 31);
    uint16_t *hit = reinterpret_cast<uint16_t *>(bits + 2048);
    for (int l = 0; l < 4; l++)
        for (int sp = 0; sp < 4; sp++)
            for (int idx = 0; idx < (1 << (2 * (l + 3))); idx++) {
                if (r.mot_wt[l][sp][idx] == -4.0) continue;
                for (int o = 0; o < 4; o++) {
                    const int nb = std::min(l + 3, 6 - o);          // bases of the motif inside the window
                    const int known = idx & ((1 << (2 * nb)) - 1);
                    // every window whose bases o .. o + nb - 1 equal `known`: the other 6 - nb bases are free
                    for (int rest = 0; rest < (1 << (2 * (6 - nb))); rest++) {
                        const int lowb = rest & ((1 << (2 * o)) - 1), highb = rest >> (2 * o);
                        const int x = lowb | (known << (2 * o)) | (highb << (2 * (o + nb)));
                        hit[x] |= (uint16_t)(1u << (4 * l + o));
                    }
                }
            }
}

static void prepare_model(const RawTraining &r, DevModel &m, const RawTraining *d_raw_k, const uint32_t *d_live_k = nullptr) {
    memset(&m, 0, sizeof(m));
    m.st_wt = r.st_wt; m.gc = r.gc; m.no_mot = r.no_mot;
    for (int i = 0; i < 3; i++) { m.bias[i] = r.bias[i]; m.type_wt[i] = r.type_wt[i]; }
    for (int i = 0; i < 28; i++) m.rbs_wt[i] = r.rbs_wt[i];
    for (int k = 0; k < 32; k++)
        for (int b = 0; b < 4; b++) m.uc[k][b] = 0.4 * r.st_wt * r.ups_comp[k][b];
    // length factor table (lib.pyx:2137-2147, 2209-2210): host libm, same as the reference process
    double no_stop;
    const double a = 1 - r.gc;
    if (r.trans_table != 11) { no_stop = (a * a * r.gc) / 8.0; no_stop += (a * a * a) / 8.0; }
    else { no_stop = (a * a * r.gc) / 4.0; no_stop += (a * a * a) / 8.0; }
    no_stop = 1 - no_stop;
    const double lfac_max = log((1 - pow(no_stop, 1000.0)) / pow(no_stop, 1000.0));
    const double lfac_min = log((1 - pow(no_stop, 80)) / pow(no_stop, 80));
    for (int g = 0; g <= 1000; g++) {
        const double tmp = pow(no_stop, (double)g);
        m.lfac[g] = log((1 - tmp) / tmp) - lfac_min;
    }
    m.lfac_span = lfac_max - lfac_min;
    for (int d = 0; d <= 60; d++) m.igt[d] = (2.0 - ((double)d / kOperDist)) * 0.15 * r.st_wt;
    m.ig_neg = -0.15 * r.st_wt;
    m.trans_table = r.trans_table;
    m.uses_sd = r.uses_sd;
    codon_masks(r.trans_table, &m.stopmask, &m.startmask);
    build_sd_masks();
    for (int x = 0; x < 2; x++)
        for (int off = 0; off < 15; off++)
            for (int gp = 0; gp < 64; gp++) {
                // the reference keeps the candidate with the largest (rbs_wt, bin): lib.pyx:884-888
                int best = 0;
                const uint32_t mask = g_sd_masks[x][off][gp];
                for (int v = 1; v < 28; v++) {
                    if (!((mask >> v) & 1)) continue;
                    if (r.rbs_wt[v] < r.rbs_wt[best]) continue;
                    if (r.rbs_wt[v] == r.rbs_wt[best] && v < best) continue;
                    best = v;
                }
                m.sd_best[off][gp][x] = (uint8_t)best;
            }
    for (int L = 0; L < 250; L++) { m.len_neg[L] = 250.0 / (float)L; m.len_pos[L] = (float)L / 250.0; }
    m.gene_dc = d_raw_k->gene_dc;
    m.mot_wt = &d_raw_k->mot_wt[0][0][0];
    m.mot_live = d_live_k;
    m.mot_hit = d_live_k ? reinterpret_cast<const uint16_t *>(d_live_k + 2048) : nullptr;
    for (int l = 0; l < 4; l++) {
        uint64_t f = 0;
        for (int sp = 0; sp < 4; sp++)
            for (int x = 0; x < 4096; x++)
                if (!(r.mot_wt[l][sp][x] == -4.0)) f |= 1ull << (x & 63);
        m.mot_pf[l] = d_live_k ? f : ~0ull;
    }
}

// ------------------------------------------------------------------------------------------------
// results
// ------------------------------------------------------------------------------------------------
struct ResSeg {  // genes of one sub-batch, in a pinned buffer filled directly by the D2H copy
    int lo = 0, hi = 0;
    int64_t g0 = 0, ng = 0;
    PinnedBuf buf;
    pgpu_gene *genes = nullptr;
    pgpu_node *gnodes = nullptr;
};

struct pgpu_result {
    std::shared_ptr<PinnedPool> pinned;
    int n_contigs = 0;
    std::vector<pgpu_contig_summary> summary;
    std::vector<int64_t> gene_off;   // [n+1]
    std::vector<ResSeg> segs;
    int64_t total_genes = 0;
    bool have_nodes = false;
    std::vector<int64_t> node_off;   // [n+1]
    std::vector<pgpu_node> nodes;
    pgpu_stats stats;
    ~pgpu_result() {
        if (pinned) for (auto &s : segs) pinned->release(s.buf);
    }
};

struct pgpu_batch {
    int device = 0;                // the context may be destroyed before its batches: keep the device, not the context
    bool owned = true;             // false: d_ascii is caller-owned device memory (pgpu_batch_wrap_device)
    int n_contigs = 0;
    std::vector<int64_t> offsets;  // [n+1]
    uint8_t *d_ascii = nullptr;    // device copy of the whole concatenated input
    int64_t total = 0;
};

// ------------------------------------------------------------------------------------------------
// one pipeline run over contigs [lo, hi)
// ------------------------------------------------------------------------------------------------
struct TrainRequest {
    int tt = 11, force_nonsd = 0;
    double st_wt = 4.35;
    RawTraining *out = nullptr;
};

struct RunPlan {
    // what to run
    int stage = 3;            // 1: extraction only, 2: + scoring/overlap, 3: everything
    int forced_tt = -1;       // stage 1/2 operator entry points
    int forced_model = -1;
    int forced_first_pass = 1;
    int forced_is_meta = 0;
    TrainRequest *train = nullptr;  // pgpu_train: run the training pass on the (single) extraction
    // lanes: called right before / after the input copy is queued on the sub-batch's stream (run_lanes chains the
    // copies of consecutive sub-batches, so that one half computes while the other half is still being copied)
    std::function<void(cudaStream_t)> before_h2d, after_h2d;
    // operator-level queries on the encoded sequence of a single contig (stage 0: nothing else runs)
    int8_t *gc_frame_out = nullptr;      // pgpu_max_gc_frame_plot: host buffer [slen]
    struct { int pos, start, model, strand, exact; int32_t *out; } sd = {0, 0, 0, 1, 1, nullptr};  // pgpu_shine_dalgarno
};

static int train_stage(pgpu_ctx *ctx, DevPool &pool, DevBatch &B, const ExtractInfo &X, RunOpts ro, int gc_count,
                       const TrainRequest &R, pgpu_stats &S);

struct OperatorOut {  // operator-level outputs (single contig)
    std::vector<int32_t> ndx, stop_val;
    std::vector<uint8_t> cls;
    std::vector<pgpu_node> nodes;
};

static double window_low(double gc) { return fmin(0.65, 0.88495 * gc - 0.0102337); }
static double window_high(double gc) { return fmax(0.35, 0.86596 * gc + 0.1131991); }

// Plan of k_coding_flat (score_kernels.cu): the chains of every extraction in table-column order, cut into groups of up
// to four neighbouring columns = plan entries; entries sorted by class = (first column = table set, lanes per ORF), every
// class followed by a padding entry.  The ORF slots of the entries are laid out on the device (k_cq_plan: the number of
// STOP nodes of an extraction is only known there); the host bounds the number of CTA spans.  Leaves B.dcS null
// (=> k_coding_orf) when the chains of an extraction are not on neighbouring columns (cannot happen with a GC window, but
// nothing here depends on it) or the index arrays are too large for 32-bit element offsets.
static void plan_coding_smem(pgpu_ctx *ctx, DevPool &pool, DevBatch &B, int total_nodes, const std::vector<ExtractInfo> &exts,
                             const std::vector<ChainInfo> &chains, const std::vector<int32_t> &eoff,
                             const std::vector<int32_t> &elist) {
    const int n_ext = (int)exts.size();
    if (B.dic_r - B.dic_f >= (int64_t)1 << 30 || total_nodes <= 0) return;
    struct Ent { int32_t ext, key; int32_t chain[4]; };
    std::vector<Ent> ents;
    ents.reserve((size_t)n_ext * 4);
    int64_t slots = 0;
    int32_t tmp[kDcCols];
    for (int e = 0; e < n_ext; e++) {
        const int L = eoff[e + 1] - eoff[e];
        if (L == 0 || exts[e].nn == 0) continue;
        if (L > kDcCols) return;
        for (int k = 0; k < L; k++) tmp[k] = elist[eoff[e] + k];
        std::sort(tmp, tmp + L, [&](int a, int b) { return ctx->h_models[chains[a].model].col < ctx->h_models[chains[b].model].col; });
        const int c0 = ctx->h_models[chains[tmp[0]].model].col;
        for (int k = 1; k < L; k++) if (ctx->h_models[chains[tmp[k]].model].col != c0 + k) return;
        for (int k = 0; k < L; k += 4) {
            const int w = std::min(4, L - k);
            Ent t;
            t.ext = e;
            t.key = (c0 + k) * 4 + (w >= 3 ? 2 : w - 1);   // (table set, log2 of the lanes per ORF)
            for (int q = 0; q < 4; q++) t.chain[q] = q < w ? tmp[k + q] : -1;
            ents.push_back(t);
            slots += exts[e].nn / 2;
        }
    }
    if (ents.empty()) return;
    // CTA span: about eight spans per SM when the batch is large enough, between 1024 and 16384 ORF slots (the bound
    // nn / 2 counts about twice the STOP nodes there are)
    int span = 8192, shift = 13;
    while (span > 1024 && slots / 2 / span < 148 * 8) { span >>= 1; shift--; }
    // counting sort by class, a padding entry behind every class
    const int n_keys = kDcCols * 4;
    std::vector<int32_t> first(n_keys + 1, 0);
    for (const Ent &t : ents) first[t.key + 1]++;
    for (int k = 0; k < n_keys; k++) first[k + 1] += first[k];
    std::vector<int32_t> order(ents.size());
    {
        std::vector<int32_t> fill(first.begin(), first.end() - 1);
        for (size_t i = 0; i < ents.size(); i++) order[fill[ents[i].key]++] = (int32_t)i;
    }
    std::vector<int32_t> pext, pchain, phs0;
    std::vector<int64_t> pcbase;
    std::vector<uint8_t> pcls;
    const size_t cap = ents.size() + n_keys;
    pext.reserve(cap); pcls.reserve(cap); phs0.reserve(cap); pchain.reserve(4 * cap); pcbase.reserve(4 * cap);
    int64_t max_cta = 0;
    for (int k = 0; k < n_keys; k++) {
        if (first[k + 1] == first[k]) continue;
        int64_t bound = 0;
        for (int i = first[k]; i < first[k + 1]; i++) {
            const Ent &t = ents[order[i]];
            pext.push_back(t.ext); pcls.push_back((uint8_t)k); phs0.push_back((exts[t.ext].node_off + 1) >> 1);
            for (int q = 0; q < 4; q++) {
                pchain.push_back(t.chain[q]);
                pcbase.push_back(t.chain[q] >= 0 ? chains[t.chain[q]].coff - exts[t.ext].node_off : INT64_MIN);
            }
            bound += exts[t.ext].nn / 2;
        }
        pext.push_back(-1); pcls.push_back((uint8_t)k); phs0.push_back(-1);
        for (int q = 0; q < 4; q++) { pchain.push_back(-1); pcbase.push_back(INT64_MIN); }
        max_cta += (bound + span - 1) / span;
    }
    std::vector<int32_t> colmodel(kDcCols, 0);
    for (int m = 0; m < ctx->n_models; m++) colmodel[ctx->h_models[m].col] = m;
    B.cq_n_ent = (int)pext.size();
    B.cq_span = span;
    B.cq_span_shift = shift;
    B.cq_max_cta = (int)max_cta;
    B.cq_ext = pool.upload(pext);
    B.cq_cls = pool.upload(pcls);
    B.cq_chain = pool.upload(pchain);
    B.cq_cbase = pool.upload(pcbase);
    B.cq_hs0 = pool.upload(phs0);
    B.cq_colmodel = pool.upload(colmodel);
    B.cq_soff = pool.alloc<int32_t>(pext.size() + 1);
    B.cq_cta = pool.alloc<int32_t>((size_t)max_cta + 2);
    B.cq_ncta = pool.alloc<int32_t>(1);
    B.link = pool.alloc<int4>(total_nodes);
    B.orfd = pool.alloc<int4>((size_t)(total_nodes + 1) / 2 + 2);
    if (!pool.failed) B.dcS = ctx->d_dcS;
}

static int run_range(pgpu_ctx *ctx, const uint8_t *h_seq, const uint8_t *d_seq, const int64_t *offsets, int lo, int hi,
                     const pgpu_opts &opts, const RunPlan &plan, pgpu_result *res, OperatorOut *op) {
    cudaStream_t st = ctx->stream;
    const int n = hi - lo;
    if (n <= 0) return PGPU_OK;
    const bool meta = opts.meta != 0 && plan.forced_model < 0 && plan.forced_tt < 0;
    if (plan.stage >= 2 && plan.forced_model < 0 && ctx->n_models == 0) return fail(ctx, PGPU_ESTATE, "no models loaded");
    DevPool pool(ctx);
    DevBatch B;
    memset(&B, 0, sizeof(B));
    RunOpts ro = {opts.closed, opts.min_gene, opts.min_edge_gene, opts.max_overlap};
    pgpu_stats &S = res->stats;
    int evi = 0;
    auto mark = [&]() { cudaEventRecord(ctx->ev[evi], st); return evi++; };
    auto now = []() { return std::chrono::steady_clock::now(); };
    auto since = [](std::chrono::steady_clock::time_point t0) {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    };
    auto t_host = now();
    const bool trace = getenv("PGPU_TRACE") != nullptr;
    const auto t_begin = now();
    auto tr = [&](const char *what) { if (trace) fprintf(stderr, "[pgpu %8.2f ms] %s\n", since(t_begin), what); };
    std::vector<std::pair<const char *, cudaEvent_t>> tevs;
    auto tev = [&](const char *name) {   // device-time trace point (PGPU_TRACE only)
        if (!trace) return;
        cudaEvent_t ev;
        cudaEventCreate(&ev);
        cudaEventRecord(ev, st);
        tevs.push_back({name, ev});
    };
    tev("begin");

    // ---- stage A: sequences ----------------------------------------------------------------------
    std::vector<ContigInfo> contigs(n);
    std::vector<int2> tiles;
    int64_t dtot = 0;
    const int64_t abase = offsets[lo];
    for (int c = 0; c < n; c++) {
        const int64_t len = offsets[lo + c + 1] - offsets[lo + c];
        if (len < 0 || len > 0x7ffffff0) return fail(ctx, PGPU_EINVAL, "contig length out of range");
        contigs[c].doff = dtot;
        contigs[c].aoff = offsets[lo + c] - abase;
        contigs[c].slen = (int)len;
        contigs[c].mask_off = 0; contigs[c].n_masks = 0; contigs[c].pad = 0;
        dtot += ((len + 64 + 127) / 128) * 128;   // >= 3 * dic_plane(len) (common.cuh) + the 16 padding bytes of digits / cod
        for (int64_t s = 0; s < len; s += 4096) tiles.push_back(make_int2(c, (int)s));
    }
    const int64_t atot = offsets[hi] - abase;
    int e_start = mark();
    uint8_t *d_ascii_local = nullptr;
    if (d_seq) {
        B.ascii = d_seq + abase;
    } else {
        d_ascii_local = pool.alloc<uint8_t>(atot + 16);
        if (pool.failed) return PGPU_ENOMEM;
        if (plan.before_h2d) plan.before_h2d(st);
        if (atot) CK(cudaMemcpyAsync(d_ascii_local, h_seq + abase, atot, cudaMemcpyHostToDevice, st));
        if (plan.after_h2d) plan.after_h2d(st);
        B.ascii = d_ascii_local;
        S.h2d_bytes += atot;
    }
    int e_h2d = mark();
    // no memset: k_encode writes every byte that is ever read (whole 16-byte groups, zero padded past the end)
    B.digits = pool.alloc<uint8_t>(dtot + 256);
    B.cod = pool.alloc<uint8_t>(dtot + 256);
    B.dic_f = pool.alloc<uint16_t>(2 * (dtot + 256));   // one buffer: k_coding_flat addresses both by offsets from dic_f
    B.dic_r = B.dic_f ? B.dic_f + dtot + 256 : nullptr;
    B.contigs = pool.upload(contigs);
    const int64_t gcwords = dtot / 32 + 8;
    B.gcbits = pool.alloc<uint32_t>(gcwords);
    B.gcpre = pool.alloc<int32_t>(gcwords + 1);
    int *d_gc_bs = pool.alloc<int>(scan_num_blocks(gcwords) + 1);
    int *d_gc_tot = pool.alloc<int>(1);
    B.gc_count = pool.alloc<int32_t>(n, true);
    B.unknown = pool.alloc<int32_t>(n, true);
    int2 *d_tiles = pool.upload(tiles);
    if (pool.failed) return PGPU_ENOMEM;
    tev("setup/alloc/memset");
    launch_encode(B, d_tiles, (int)tiles.size(), st);
    ctx->launches++;
    if (plan.gc_frame_out || plan.sd.out) {
        // stage 0: the query only needs what k_encode wrote (digits, GC bitmap of contig 0 at offset 0)
        const int slen = contigs[0].slen;
        if (plan.gc_frame_out && slen > 0) {
            int8_t *d_gp = pool.alloc<int8_t>((size_t)slen + 16);
            if (pool.failed) return PGPU_ENOMEM;
            launch_gc_frame(B.gcbits, slen, d_gp, st);
            ctx->launches++;
            CK(cudaMemcpyAsync(plan.gc_frame_out, d_gp, slen, cudaMemcpyDeviceToHost, st));
        }
        if (plan.sd.out) {
            int32_t *d_out = pool.alloc<int32_t>(1);
            if (pool.failed) return PGPU_ENOMEM;
            launch_shine_dalgarno(B, ctx->d_models, plan.sd.model, plan.sd.pos, plan.sd.start, plan.sd.strand, plan.sd.exact,
                                  d_out, st);
            ctx->launches++;
            CK(cudaMemcpyAsync(plan.sd.out, d_out, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        }
        CK(cudaStreamSynchronize(st));
        return PGPU_OK;
    }
    std::vector<int4> h_masks;
    if (opts.mask) {
        const int cap = (int)std::min<int64_t>(std::max<int64_t>(1024, atot / std::max(1, opts.min_mask) + n + 16), 1 << 28);
        int4 *d_mk = pool.alloc<int4>(cap);
        int *d_cnt = pool.alloc<int>(1, true);
        if (pool.failed) return PGPU_ENOMEM;
        launch_find_masks(B, d_tiles, (int)tiles.size(), opts.min_mask, d_mk, cap, d_cnt, st);
        ctx->launches++;
        int cnt = 0;
        CK(cudaMemcpyAsync(&cnt, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        cnt = std::min(cnt, cap);
        h_masks.resize(cnt);
        if (cnt) CK(cudaMemcpyAsync(h_masks.data(), d_mk, cnt * sizeof(int4), cudaMemcpyDeviceToHost, st));
    }
    std::vector<int32_t> h_gc(n), h_unk(n);
    CK(cudaMemcpyAsync(h_gc.data(), B.gc_count, n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_unk.data(), B.unknown, n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    // the host only needs the G/C counts to plan the chains: the GC prefix scan and the dicodon index run behind
    // that copy, underneath the planning loop
    CK(cudaEventRecord(ctx->ev[13], st));
    launch_gc_scan(B, dtot / 32, d_gc_bs, d_gc_tot, st);
    ctx->launches += 3;
    if (plan.stage >= 2 || plan.train) { launch_dicodon_index(B, d_tiles, (int)tiles.size(), st); ctx->launches++; }
    tev("k_encode+gc scan");
    int e_enc = mark();
    S.reserved[0] += since(t_host);  // host: stage A issue
    CK(cudaEventSynchronize(ctx->ev[13]));
    t_host = now();
    S.d2h_bytes += 8 * (int64_t)n;

    // masks: sort by (contig, begin) and build the per-contig table
    std::vector<int32_t> mask_tab;
    if (!h_masks.empty()) {
        std::sort(h_masks.begin(), h_masks.end(), [](const int4 &a, const int4 &b) { return a.x != b.x ? a.x < b.x : a.y < b.y; });
        for (size_t k = 0; k < h_masks.size(); k++) {
            ContigInfo &ci = contigs[h_masks[k].x];
            if (ci.n_masks == 0) ci.mask_off = (int)k;
            ci.n_masks++;
            mask_tab.push_back(h_masks[k].y);
            mask_tab.push_back(h_masks[k].z);
        }
    }
    B.masks = pool.upload(mask_tab);

    tr("sync1 done (gc counts)");
    // ---- stage B: plan extractions and chains ------------------------------------------------------
    // (per-thread scratch: a 630 Mbp sub-batch plans 185 k chains = 15 MB, and fresh vectors would be page-faulted in on
    // every call while the GPU waits for this loop)
    static thread_local std::vector<ExtractInfo> tl_exts;
    static thread_local std::vector<ChainInfo> tl_chains;
    std::vector<ExtractInfo> &exts = tl_exts;
    std::vector<ChainInfo> &chains = tl_chains;
    exts.clear();
    chains.clear();
    std::vector<int32_t> contig_chain_begin(n + 1, 0);
    std::vector<int32_t> contig_ext_begin(n + 1, 0);
    int64_t nwords = 0;
    int total_chunks = 0;
    int64_t cb_words = 0;   // codon bitmap words (bit-parallel extraction)
    std::vector<int> lut_tt;            // translation tables met in this sub-batch: one codon table row each
    std::vector<uint8_t> lut_rows;
    exts.reserve((size_t)n * 2);
    chains.reserve((size_t)n * (meta ? 16 : 1));
    for (int c = 0; c < n; c++) {
        contig_chain_begin[c] = (int)chains.size();
        contig_ext_begin[c] = (int)exts.size();
        const ContigInfo &ci = contigs[c];
        auto get_ext = [&](int tt) {
            for (int e = contig_ext_begin[c]; e < (int)exts.size(); e++)
                if (exts[e].tt == tt) return e;
            ExtractInfo X;
            memset(&X, 0, sizeof(X));
            X.contig = c; X.tt = tt; X.doff = ci.doff; X.slen = ci.slen;
            X.slot = (int)exts.size() - contig_ext_begin[c];
            X.woff = nwords; X.nwords = ci.slen / 32 + 1;
            X.chunk_off = total_chunks;
            X.n_chunks = std::max(1, (ci.slen / 3 + kExtractChunkCodons - 1) / kExtractChunkCodons);
            total_chunks += X.n_chunks;
            X.cb_off = cb_words;
            cb_words += (int64_t)6 * X.n_chunks * (kExtractChunkCodons / 32);
            X.mask_off = ci.mask_off; X.n_masks = ci.n_masks;
            {   // codon masks per translation table, computed once
                static thread_local uint64_t cache[34][2];
                static thread_local bool have[34] = {false};
                if (tt >= 0 && tt < 34) {
                    if (!have[tt]) { codon_masks(tt, &cache[tt][0], &cache[tt][1]); have[tt] = true; }
                    X.stopmask = cache[tt][0]; X.startmask = cache[tt][1];
                } else {
                    codon_masks(tt, &X.stopmask, &X.startmask);
                }
            }
            {
                size_t row = std::find(lut_tt.begin(), lut_tt.end(), tt) - lut_tt.begin();
                if (row == lut_tt.size()) {
                    lut_tt.push_back(tt);
                    lut_rows.resize(lut_rows.size() + 128);
                    codon_lut_build(X.stopmask, X.startmask, lut_rows.data() + 128 * row);
                }
                X.lut = (int32_t)row;
            }
            nwords += X.nwords;
            exts.push_back(X);
            return (int)exts.size() - 1;
        };
        auto add_chain = [&](int model, int tt, int first_pass, int is_meta) {
            const int e = get_ext(tt);
            chains.emplace_back();   // filled in place: this loop runs while the GPU waits for the plan
            ChainInfo &K = chains.back();
            memset(&K, 0, sizeof(K));
            K.ext = e; K.model = model; K.contig = c; K.first_pass = first_pass;
            K.doff = ci.doff; K.slen = ci.slen; K.is_meta = is_meta;
        };
        if (plan.forced_tt >= 0) {
            get_ext(plan.forced_tt);
        } else if (plan.forced_model >= 0) {
            add_chain(plan.forced_model, ctx->h_models[plan.forced_model].trans_table, plan.forced_first_pass, plan.forced_is_meta);
        } else if (!meta) {
            add_chain(opts.single_model, ctx->h_models[opts.single_model].trans_table, 1, 0);
        } else {
            // lib.pyx:5335-5357: bins inside the GC window, in order; re-extraction whenever the table changes
            const double gc = ci.slen > 0 ? (double)h_gc[c] / (double)ci.slen : 0.0;
            const double low = window_low(gc), high = window_high(gc);
            int tt = -1;
            const double *mgc = ctx->model_gc.data();
            const int *mtt = ctx->model_tt.data();
            for (int m = 0; m < ctx->n_models; m++) {
                if (mgc[m] < low || mgc[m] > high) continue;
                const int first = mtt[m] != tt;
                tt = mtt[m];
                add_chain(m, tt, first, 1);
            }
        }
    }
    contig_chain_begin[n] = (int)chains.size();
    contig_ext_begin[n] = (int)exts.size();
    const int n_ext = (int)exts.size(), n_chains = (int)chains.size();
    // chains grouped by extraction (the lanes of k_dp_ml / k_coding_orf / k_overlap_lanes); node counts are not needed yet
    std::vector<int32_t> h_eoff(n_ext + 1, 0);   // chains of an extraction: [h_eoff[e], h_eoff[e+1]) in B.ext_chains
    std::vector<int32_t> h_elist(n_chains);
    {
        for (const auto &K : chains) h_eoff[K.ext + 1]++;
        for (int e = 0; e < n_ext; e++) h_eoff[e + 1] += h_eoff[e];
        std::vector<int32_t> fillp(h_eoff.begin(), h_eoff.end() - 1);
        for (int k = 0; k < n_chains; k++) { chains[k].lane = fillp[chains[k].ext] - h_eoff[chains[k].ext]; h_elist[fillp[chains[k].ext]++] = k; }
    }
    int64_t total_il = 0;   // elements of an interleaved per-chain-node array (>= total chain-nodes: blocks are padded)

    tr("planned chains");
    // ---- extraction pass 1: mark + scan ------------------------------------------------------------
    B.bits_fwd = pool.alloc<uint32_t>(nwords + 8, true);
    B.bits_rev = pool.alloc<uint32_t>(nwords + 8, true);
    B.wordbase = pool.alloc<int32_t>(nwords + 8);
    B.exts = pool.upload(exts);
    int *d_block_sums = pool.alloc<int>(scan_num_blocks(nwords) + 1);
    int *d_total = pool.alloc<int>(1, true);
    if (pool.failed) return PGPU_ENOMEM;
    // bit-parallel extraction unless the batch has N-run masks (GeneFinder(mask=True)) or PGPU_EXTRACT_ALGO=1
    bool extract_bits = ctx->extract_algo != 1;
    for (const auto &X : exts) if (X.n_masks) { extract_bits = false; break; }
    if (extract_bits) {
        B.cb_stop = pool.alloc<uint32_t>(cb_words + 8);
        B.cb_start = pool.alloc<uint32_t>(cb_words + 8);
        if (ctx->codon_lut) B.codon_lut = pool.upload(lut_rows);
        if (pool.failed) return PGPU_ENOMEM;
    }
    tev("plan+alloc");
    if (extract_bits) {
        launch_codon_bits(B, n_ext, total_chunks, st);
        tev("k_codon_bits");
        launch_extract_bits(B, n_ext, total_chunks, ro, false, st);
        ctx->launches++;
    } else {
        launch_extract_mark(B, n_ext, total_chunks, ro, st);
    }
    tev("k_extract mark");
    launch_word_scan(B, nwords, d_block_sums, d_total, st);
    tev("word scan");
    ctx->launches += 4;
    // node offsets of every extraction = wordbase[woff]
    std::vector<int32_t> h_base(n_ext + 1, 0);
    {
        // gather with strided copies would be n_ext memcpys; copy the few words we need via a 2D copy is not
        // possible (irregular), so read wordbase[woff_e] with one small kernel-free trick: cudaMemcpy2D is
        // regular only.  Use a pinned gather list instead: n_ext is small compared to the data.
        std::vector<int64_t> idx(n_ext + 1);
        for (int e = 0; e < n_ext; e++) idx[e] = exts[e].woff;
        idx[n_ext] = nwords;
        // device-side gather
        int64_t *d_idx = pool.upload(idx);
        int32_t *d_out = pool.alloc<int32_t>(n_ext + 1);
        if (pool.failed) return PGPU_ENOMEM;
        extern void launch_gather_i32(const int32_t *src, const int64_t *idx, int n, int32_t *dst, cudaStream_t st);
        launch_gather_i32(B.wordbase, d_idx, n_ext + 1, d_out, st);
        ctx->launches++;
        CK(cudaMemcpyAsync(h_base.data(), d_out, (n_ext + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    }
    int e_mark = mark();
    S.reserved[1] += since(t_host);  // host: chain planning + mark/scan issue
    CK(cudaStreamSynchronize(st));
    t_host = now();
    S.d2h_bytes += 4 * (int64_t)(n_ext + 1);
    tr("sync2 done (node counts)");
    const int total_nodes = h_base[n_ext];
    if (total_nodes < 0) return fail(ctx, PGPU_EINVAL, "too many nodes in one sub-batch");
    for (int e = 0; e < n_ext; e++) { exts[e].node_off = h_base[e]; exts[e].nn = h_base[e + 1] - h_base[e]; }
    int64_t total_cn = 0;
    for (auto &K : chains) { K.node_off = exts[K.ext].node_off; K.nn = exts[K.ext].nn; K.coff = total_cn; total_cn += K.nn; }
    {   // interleaved layout of the arrays the DP touches (ChainInfo::ioff): the chains of an extraction share one block
        // of nn x L elements, element (node, chain) at block + node * L + position of the chain among them
        std::vector<int64_t> iblock(n_ext + 1, 0);
        for (int e = 0; e < n_ext; e++) iblock[e + 1] = iblock[e] + (((int64_t)exts[e].nn * (h_eoff[e + 1] - h_eoff[e]) + 1) & ~int64_t(1));
        for (auto &K : chains) { K.istride = h_eoff[K.ext + 1] - h_eoff[K.ext]; K.ioff = iblock[K.ext] + K.lane; }
        total_il = iblock[n_ext];   // blocks start at even elements (16-byte aligned doubles: bulk copies)
    }
    pool.copy_in(B.exts, exts.data(), n_ext * sizeof(ExtractInfo));

    // ---- extraction pass 2: fill, then per-extraction preparation ----------------------------------
    B.ndx = pool.alloc<int32_t>(total_nodes);
    B.stop_val = pool.alloc<int32_t>(total_nodes);
    B.cls = pool.alloc<uint8_t>(total_nodes + 16);
    B.gc_cont = pool.alloc<float>(total_nodes, true);
    B.sdbits = pool.alloc<uint32_t>(total_nodes);
    B.upc = pool.alloc<uint64_t>(total_nodes);
    B.umot = pool.alloc<uint64_t>(total_nodes);
    B.win_min = pool.alloc<int32_t>(total_nodes);
    B.crank = pool.alloc<int32_t>(4 * (size_t)total_nodes + 4);
    B.clist = pool.alloc<int32_t>(total_nodes);
    B.cbase = pool.alloc<int32_t>(4 * (size_t)n_ext + 4);
    B.cndx = pool.alloc<int32_t>(total_nodes);
    B.dpx = pool.alloc<int4>(total_nodes);
    B.ig_node = pool.alloc<int32_t>(total_nodes + 1);
    B.ig_ndx = pool.alloc<int32_t>(total_nodes + 1);
    B.dqx = pool.alloc<int4>(total_nodes);
    B.feq = pool.alloc<int32_t>(total_nodes);
    unsigned long long *d_ext_pairs = pool.alloc<unsigned long long>(n_ext, true);
    if (pool.failed) return PGPU_ENOMEM;
    {
        int32_t *tab = pool.alloc<int32_t>((size_t)total_nodes / 128 + 2);
        if (pool.failed) return PGPU_ENOMEM;
        launch_block_owner_exts(B.exts, n_ext, tab, st);
        B.blk_ext = tab;
        ctx->launches++;
    }
    tev("sync2+alloc");
    if (extract_bits) launch_extract_bits(B, n_ext, total_chunks, ro, true, st);
    else launch_extract_fill(B, n_ext, total_chunks, ro, st);
    tev("k_extract fill");
    launch_node_prep(B, n_ext, total_nodes, 1, st);
    tev("k_node_prep+class_index");
    launch_dp_index(B, n_ext, total_nodes, st);
    tev("k_dp_index");
    launch_pairs(B, n_ext, total_nodes, d_ext_pairs, st);
    tev("k_pairs");
    ctx->launches += 5;
    int e_ext = mark();

    S.n_contigs += n; S.total_bp += atot; S.total_nodes += total_nodes; S.total_chain_nodes += total_cn;
    S.n_chains += n_chains; S.dp_steps += total_cn;

    if (plan.train) return train_stage(ctx, pool, B, exts[0], ro, h_gc[0], *plan.train, S);

    if (plan.stage == 1) {
        if (op) {
            op->ndx.resize(total_nodes); op->stop_val.resize(total_nodes); op->cls.resize(total_nodes);
            if (total_nodes) {
                CK(cudaMemcpyAsync(op->ndx.data(), B.ndx, total_nodes * 4, cudaMemcpyDeviceToHost, st));
                CK(cudaMemcpyAsync(op->stop_val.data(), B.stop_val, total_nodes * 4, cudaMemcpyDeviceToHost, st));
                CK(cudaMemcpyAsync(op->cls.data(), B.cls, total_nodes, cudaMemcpyDeviceToHost, st));
            }
        }
        CK(cudaStreamSynchronize(st));
        return PGPU_OK;
    }

    tr("issued fill/prep");
    // ---- per-chain scoring ---------------------------------------------------------------------------
    B.chains = pool.upload(chains);
    {
        const std::vector<int32_t> &eoff = h_eoff;
        B.ext_chain_off = pool.upload(h_eoff);
        B.ext_chains = pool.upload(h_elist);
        B.dcT = ctx->d_dcT;
        B.n_models = ctx->n_models;
        if (n_chains > n_ext) {
            // thread layout of the kernels that put the lanes of a group over the chains (models) of an extraction
            // (k_coding_orf, k_overlap_lanes): W lanes per STOP node for an extraction with L chains, nn / 2 + 1 slots
            std::vector<int64_t> toff(n_ext + 1, 0);
            std::vector<uint8_t> w(n_ext + 1, 32);
            // order: by the first model of the extraction's chains (= by translation table and GC window: the models are
            // sorted that way), so that concurrently resident warps share dicodon-table columns
            std::vector<int32_t> perm(n_ext);
            std::iota(perm.begin(), perm.end(), 0);
            std::vector<int32_t> key(n_ext, 0);
            for (int e = 0; e < n_ext; e++) if (eoff[e + 1] > eoff[e]) key[e] = ctx->h_models[chains[h_elist[eoff[e]]].model].col;
            std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return key[a] < key[b]; });
            for (int r = 0; r < n_ext; r++) {
                const int e = perm[r];
                const int L = eoff[e + 1] - eoff[e];
                w[r] = (uint8_t)(L <= 4 ? 4 : L <= 8 ? 8 : L <= 16 ? 16 : 32);
                const int64_t threads = L > 0 ? ((int64_t)exts[e].nn / 2 + 1) * w[r] : 0;
                toff[r + 1] = toff[r] + ((threads + 31) & ~int64_t(31));
            }
            B.orf_ext = pool.upload(perm);
            B.orf_threads = toff[n_ext];
            B.orf_toff = pool.upload(toff);
            B.orf_w = pool.upload(w);
            int32_t *tab = pool.alloc<int32_t>((size_t)(B.orf_threads >> 8) + 2);
            if (pool.failed) return PGPU_ENOMEM;
            launch_block_owner_off(B.orf_toff, n_ext, 8, tab, st);
            B.orf_blk = tab;
            ctx->launches++;
            if (ctx->coding_smem && ctx->d_dcS) plan_coding_smem(ctx, pool, B, total_nodes, exts, chains, h_eoff, h_elist);
            if (trace) fprintf(stderr, "[pgpu] coding: %s, at most %d CTA spans of %d ORF slots\n", B.dcS ? "k_coding_flat" : "k_coding_orf", B.cq_max_cta, B.cq_span);
        }
    }
    // Meta mode scores every (contig, model) chain "lean": per chain-node only the raw coding score, cs = cscore + sscore
    // and (sparse) the penalty of touching starts are kept; the winner of every contig is scored in full, into one
    // slot per contig, right before the traceback (launch_trace).  Single mode / operators keep every score.
    const bool lean = meta && plan.stage >= 3;
    // slot per contig of the passes that score one chain per contig (winner pass, final pass): the largest extraction
    std::vector<int64_t> gene_off(n + 1, 0), fin_coff(n + 1, 0);
    for (int c = 0; c < n; c++) {
        int mx = 0;
        for (int e = contig_ext_begin[c]; e < contig_ext_begin[c + 1]; e++) mx = std::max(mx, exts[e].nn);
        gene_off[c + 1] = gene_off[c] + mx / 2 + 1;   // a gene ends at a distinct STOP node: nn / 2 + 1 bounds the genes
        fin_coff[c + 1] = fin_coff[c] + mx;
    }
    const int64_t ftot = fin_coff[n];
    B.cscore = pool.alloc<double>(total_cn);
    if (!lean) {
        B.sscore = pool.alloc<double>(total_cn);
        B.rscore = pool.alloc<double>(total_cn); B.uscore = pool.alloc<double>(total_cn);
        B.tscore = pool.alloc<double>(total_cn);
        B.rbs = pool.alloc<uint8_t>(2 * (size_t)total_cn + 16);
    } else {
        B.rupen = pool.alloc<double>(total_il + 512);
    }
    // interleaved arrays: total_il elements (blocks padded to even sizes) + slack for the bulk copies of the last rows
    B.cs = pool.alloc<double>(total_il + 512);
    B.opv = pool.alloc<double>(3 * (size_t)total_il + 512);
    B.star_ptr = pool.alloc<int32_t>(3 * (size_t)total_il + 512);
    const bool dp_ml = ctx->dp_algo >= 6 || (ctx->dp_algo == 5 && n_chains > n_ext);
    // the DP score of every chain-node is only kept when somebody reads it: node records of single mode, the per-chain
    // kernels (their k_chain_best pass), the self-check.  k_dp_ml tracks the best terminal node of a chain itself.
    if (!lean || !dp_ml || ctx->dp_verify) B.score = pool.alloc<double>(total_il + 512);
    B.traceb = pool.alloc<int32_t>(total_il + 512);
    B.ov_mark = pool.alloc<int8_t>(total_il + 512);
    B.chain_ipath = pool.alloc<int32_t>(n_chains);
    B.chain_score = pool.alloc<double>(n_chains);
    if (ctx->dp_algo >= 1) {
        B.dp_svig = pool.alloc<double>(total_il + 512);
        B.dp_tbig = pool.alloc<int32_t>(total_il + 512);
        if (dp_ml) {
            B.dp_fmv = pool.alloc<double>(total_il + 512);
            B.dp_fmj = pool.alloc<int32_t>(total_il + 512);
        }
    }
    const int64_t trace_cn = lean ? ftot : total_cn;   // forward pointers / elimination flags: winner slots or every chain
    int32_t *d_tracef = pool.alloc<int32_t>(trace_cn);
    uint8_t *d_elim = pool.alloc<uint8_t>(trace_cn + 16, true);
    MotifOut *d_mot_main = (!meta) ? pool.alloc<MotifOut>(total_cn) : nullptr;
    if (pool.failed) return PGPU_ENOMEM;
    if (trace_cn) CK(cudaMemsetAsync(d_tracef, 0xff, trace_cn * sizeof(int32_t), st));
    {
        int32_t *tab = pool.alloc<int32_t>((size_t)total_cn / 128 + 2);
        if (pool.failed) return PGPU_ENOMEM;
        launch_block_owner_chains(B.chains, n_chains, tab, st);
        B.blk_chain = tab;
        ctx->launches++;
    }
    tev("chain alloc/upload");
    const bool coding_verify = ctx->coding_verify && B.dcS && total_cn > 0;
    if (coding_verify) cudaMemsetAsync(B.cscore, 0, total_cn * sizeof(double), st);   // STOP nodes are never written
    launch_coding(B, ctx->d_models, n_chains, total_cn, n_ext, total_nodes, st);
    tev("k_coding_orf");
    if (coding_verify) {
        // self-check: the lanes-over-models kernel (weights through L1 / L2) must reproduce every raw coding score
        DevBatch V = B;
        V.dcS = nullptr;
        V.cscore = pool.alloc<double>(total_cn);
        unsigned long long *d_bad = pool.alloc<unsigned long long>(2);
        if (pool.failed) return PGPU_ENOMEM;
        cudaMemsetAsync(V.cscore, 0, total_cn * sizeof(double), st);
        const unsigned long long init[2] = {0ULL, ~0ULL};
        CK(cudaMemcpyAsync(d_bad, init, sizeof(init), cudaMemcpyHostToDevice, st));
        launch_coding(V, ctx->d_models, n_chains, total_cn, n_ext, total_nodes, st);
        launch_dp_compare(B.cscore, V.cscore, (const int32_t *)B.cscore, (const int32_t *)B.cscore, (const int8_t *)B.cscore,
                          (const int8_t *)B.cscore, total_cn, d_bad, st);
        unsigned long long bad[2] = {0, 0};
        CK(cudaMemcpyAsync(bad, d_bad, sizeof(bad), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (bad[0]) {
            char msg[160];
            snprintf(msg, sizeof msg, "PGPU_CODING_VERIFY: %llu of %lld raw coding scores differ between k_coding_flat and k_coding_orf, first at chain-node %llu",
                     bad[0], (long long)total_cn, bad[1]);
            return fail(ctx, PGPU_ESTATE, msg);
        }
        ctx->launches += 2;   // (tests read this: the comparison did run)
    }
    if (lean) launch_start_score_lean(B, ctx->d_models, n_chains, total_cn, ro, st);
    else launch_start_score(B, ctx->d_models, n_chains, total_cn, ro, d_mot_main, st);
    tev("k_start_score");
    ctx->launches += B.dcS ? 5 : 2;   // k_orf_links + k_cq_plan + k_cq_owner + k_coding_flat, or k_coding_orf; the start scoring
    int e_score = mark();
    launch_overlap(B, ctx->d_models, n_chains, total_cn, total_il, n_ext, ro, 1, st);
    tev("k_overlap");
    ctx->launches++;
    int e_ovl = mark();

    if (plan.stage == 2) {
        if (op) {
            pgpu_node *d_nodes = pool.alloc<pgpu_node>(total_cn);
            if (pool.failed) return PGPU_ENOMEM;
            launch_pack_nodes(B, n_chains, total_cn, d_mot_main, d_tracef, d_elim, 2, d_nodes, st);
            op->nodes.resize(total_cn);
            if (total_cn) CK(cudaMemcpyAsync(op->nodes.data(), d_nodes, total_cn * sizeof(pgpu_node), cudaMemcpyDeviceToHost, st));
        }
        CK(cudaStreamSynchronize(st));
        return PGPU_OK;
    }

    tr("issued scoring/overlap");
    // ---- DP: largest chains first ---------------------------------------------------------------------
    std::vector<int32_t> order(n_chains);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return chains[a].nn > chains[b].nn; });
    int32_t *d_order = pool.upload(order);
    int32_t *d_ccb = pool.upload(contig_chain_begin);
    pgpu_gene *d_genes = pool.alloc<pgpu_gene>(gene_off[n]);
    pgpu_gene *d_genes_raw = pool.alloc<pgpu_gene>(gene_off[n]);
    int64_t *d_gene_off = pool.upload(gene_off);
    std::vector<pgpu_contig_summary> summ(n);
    for (int c = 0; c < n; c++) { memset(&summ[c], 0, sizeof(summ[c])); summ[c].unknown = h_unk[c]; summ[c].gc_count = h_gc[c]; }
    pgpu_contig_summary *d_summ = pool.upload(summ);
    int32_t *d_winner_chain = pool.alloc<int32_t>(n);
    if (pool.failed) return PGPU_ENOMEM;
    tr("uploaded dp tables");
    tev("dp uploads");
    if (dp_ml) {
        // a group = an extraction (contig x translation table) and <= 32 of its chains, one lane per chain = what one warp
        // walks; longest first, so that the tail of the launch consists of short walks
        std::vector<int4> groups;
        for (int e = 0; e < n_ext; e++)
            for (int c = h_eoff[e]; c < h_eoff[e + 1]; c += 32) groups.push_back(make_int4(c, std::min(32, h_eoff[e + 1] - c), e, exts[e].nn));
        std::stable_sort(groups.begin(), groups.end(), [](const int4 &a, const int4 &b) { return a.w > b.w; });
        int4 *d_groups = pool.upload(groups);
        if (pool.failed) return PGPU_ENOMEM;
        if (ctx->dp_verify) {   // the self-check compares whole arrays: define the padding between the blocks
            cudaMemsetAsync(B.score, 0, (total_il + 512) * sizeof(double), st);
            cudaMemsetAsync(B.traceb, 0, (total_il + 512) * sizeof(int32_t), st);
            cudaMemsetAsync(B.ov_mark, 0, (total_il + 512) * sizeof(int8_t), st);
        }
        launch_dp_ml(B, ctx->d_models, d_groups, (int)groups.size(), n_chains, ctx->dp_ml_minb, st);
        ctx->launches++;
        if (ctx->dp_verify && total_cn > 0) {
            // self-check: the per-chain kernel must reproduce every score / traceback / overlap frame
            DevBatch V = B;
            V.score = pool.alloc<double>(total_il + 512);
            V.traceb = pool.alloc<int32_t>(total_il + 512);
            V.ov_mark = pool.alloc<int8_t>(total_il + 512);
            unsigned long long *d_bad = pool.alloc<unsigned long long>(2);
            if (pool.failed) return PGPU_ENOMEM;
            cudaMemsetAsync(V.score, 0, (total_il + 512) * sizeof(double), st);
            cudaMemsetAsync(V.traceb, 0, (total_il + 512) * sizeof(int32_t), st);
            cudaMemsetAsync(V.ov_mark, 0, (total_il + 512) * sizeof(int8_t), st);
            const unsigned long long init[2] = {0ULL, ~0ULL};
            CK(cudaMemcpyAsync(d_bad, init, sizeof(init), cudaMemcpyHostToDevice, st));
            launch_dp(V, ctx->d_models, d_order, n_chains, 1, 3, st);
            // the padding elements between blocks are never written: compare zeros there
            if (pool.failed) return PGPU_ENOMEM;
            launch_dp_compare(B.score, V.score, B.traceb, V.traceb, B.ov_mark, V.ov_mark, total_il, d_bad, st);
            unsigned long long bad[2] = {0, 0};
            CK(cudaMemcpyAsync(bad, d_bad, sizeof(bad), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            if (bad[0]) {
                // describe the first differing element (interleaved index, ChainInfo::ioff)
                const int64_t g = (int64_t)bad[1];
                int k = 0;
                for (int q = 0; q < n_chains; q++) {
                    const int64_t d = g - chains[q].ioff;
                    if (d >= 0 && d % chains[q].istride == 0 && d / chains[q].istride < chains[q].nn) { k = q; break; }
                }
                const int node = (int)((g - chains[k].ioff) / chains[k].istride);
                double sc[2]; int32_t tb[2]; int8_t ov[2]; int32_t nx = 0, svv = 0; uint8_t cl = 0;
                cudaMemcpy(&sc[0], B.score + g, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&sc[1], V.score + g, 8, cudaMemcpyDeviceToHost);
                cudaMemcpy(&tb[0], B.traceb + g, 4, cudaMemcpyDeviceToHost); cudaMemcpy(&tb[1], V.traceb + g, 4, cudaMemcpyDeviceToHost);
                cudaMemcpy(&ov[0], B.ov_mark + g, 1, cudaMemcpyDeviceToHost); cudaMemcpy(&ov[1], V.ov_mark + g, 1, cudaMemcpyDeviceToHost);
                cudaMemcpy(&nx, B.ndx + chains[k].node_off + node, 4, cudaMemcpyDeviceToHost);
                cudaMemcpy(&svv, B.stop_val + chains[k].node_off + node, 4, cudaMemcpyDeviceToHost);
                cudaMemcpy(&cl, B.cls + chains[k].node_off + node, 1, cudaMemcpyDeviceToHost);
                char msg[320];
                snprintf(msg, sizeof(msg), "PGPU_DP_VERIFY: %llu of %lld chain-nodes differ; first: chain %d (contig %d model %d nn %d) "
                         "node %d kind %d ndx %d sv %d: score %.17g/%.17g traceb %d/%d ov %d/%d", bad[0], (long long)total_cn, k,
                         chains[k].contig, chains[k].model, chains[k].nn, node, cls_kind(cl), nx, svv, sc[0], sc[1], tb[0], tb[1],
                         (int)ov[0], (int)ov[1]);
                return fail(ctx, PGPU_ESTATE, msg);
            }
        }
    } else {
        launch_dp(B, ctx->d_models, d_order, n_chains, 1, ctx->dp_algo >= 5 ? 3 : ctx->dp_algo, st);
    }
    tev("k_dp");
    ctx->launches++;
    int e_dp = mark();
    // one-chain-per-contig passes of meta mode (winner pass before the traceback, final pass after it): their own
    // chain-node index space (one slot per contig) and score arrays, shared by the two passes
    DevBatch F = B;
    ChainInfo *d_fin = nullptr, *d_win = nullptr;
    int64_t *d_fin_coff = nullptr;
    MotifOut *d_mot = nullptr;
    if (meta) {
        d_fin = pool.alloc<ChainInfo>(n);
        d_win = pool.alloc<ChainInfo>(n);
        d_fin_coff = pool.upload(fin_coff);
        F.chains = d_fin;
        F.blk_chain = nullptr;   // its own chain-node index space
        F.cs = nullptr; F.rupen = nullptr;
        F.cscore = pool.alloc<double>(ftot); F.sscore = pool.alloc<double>(ftot); F.rscore = pool.alloc<double>(ftot);
        F.uscore = pool.alloc<double>(ftot); F.tscore = pool.alloc<double>(ftot);
        F.rbs = pool.alloc<uint8_t>(2 * (size_t)ftot + 16);
        F.ext_chains = nullptr;  // one chain per contig: the per-chain kernels
        F.orf_toff = nullptr;
        F.dcS = nullptr;
        d_mot = pool.alloc<MotifOut>(ftot);
        if (pool.failed) return PGPU_ENOMEM;
    }
    if (lean) {
        // winner pass: the winner's chain with its own first_pass flag, raw coding scores taken from the main pass
        DevBatch Wb = F;
        Wb.chains = d_win;
        Wb.cscore_in = B.cscore;
        launch_trace(B, ctx->d_models, n, d_ccb, d_tracef, d_elim, d_genes, d_genes_raw, d_gene_off, gene_off[n], d_summ,
                     d_winner_chain, 1, opts.max_overlap, &Wb,
                     [&]() { launch_start_score(Wb, ctx->d_models, n, ftot, ro, nullptr, st); }, d_fin_coff, st);
        ctx->launches += 2;
    } else {
        launch_trace(B, ctx->d_models, n, d_ccb, d_tracef, d_elim, d_genes, d_genes_raw, d_gene_off, gene_off[n], d_summ,
                     d_winner_chain, meta ? 1 : 0, opts.max_overlap, nullptr, nullptr, nullptr, st);
    }
    ctx->launches += 3;
    tev("k_trace");
    int e_trace = mark();

    tr("issued dp/trace");
    // ---- final node records ------------------------------------------------------------------------------
    pgpu_node *d_nodes = nullptr;
    int64_t *d_node_out_off = nullptr;
    std::vector<int64_t> node_out_off(n + 1, 0);
    if (meta) {
        // re-score the winner's nodes as a fresh first pass (lib.pyx:5380-5394)
        d_nodes = pool.alloc<pgpu_node>(ftot);
        if (pool.failed) return PGPU_ENOMEM;
        launch_build_final_chains(B, n, d_winner_chain, d_fin_coff, d_fin, st);
        // chains with nn == 0 keep coff monotone, so the chain search inside the kernels stays valid; the
        // unused tail of every contig's slot is never touched because kernels index by chain
        if (opts.want_nodes || ctx->final_algo == 1) {
            launch_score_chains(F, ctx->d_models, n, ftot, ro, d_mot, n_ext, total_nodes, st);
            launch_pack_nodes(F, n, ftot, d_mot, nullptr, nullptr, 0, d_nodes, st);
            ctx->launches += 4;
        } else {
            // only the start / stop node records of the genes leave the device: re-score just their ORFs
            int2 *d_list = pool.alloc<int2>(gene_off[n]);
            int *d_count = pool.alloc<int>(1);
            if (pool.failed) return PGPU_ENOMEM;
            launch_score_genes(F, ctx->d_models, n, d_summ, d_genes, d_gene_off, gene_off[n], d_list, d_count, ro, d_mot, st);
            launch_pack_nodes_genes(F, d_list, d_count, gene_off[n], d_genes, d_gene_off, d_mot, d_nodes, st);
            ctx->launches += 5;
        }
        node_out_off = fin_coff;
        d_node_out_off = d_fin_coff;
    } else {
        d_nodes = pool.alloc<pgpu_node>(total_cn);
        if (pool.failed) return PGPU_ENOMEM;
        launch_pack_nodes(B, n_chains, total_cn, d_mot_main, d_tracef, d_elim, 1, d_nodes, st);
        ctx->launches++;
        for (int c = 0; c < n; c++) node_out_off[c] = chains[contig_chain_begin[c]].coff;
        node_out_off[n] = total_cn;
        d_node_out_off = pool.upload(node_out_off);
    }
    tev("final rescoring + pack");
    int e_final = mark();
    std::vector<unsigned long long> h_ext_pairs(n_ext);
    if (n_ext) CK(cudaMemcpyAsync(h_ext_pairs.data(), d_ext_pairs, n_ext * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(summ.data(), d_summ, n * sizeof(pgpu_contig_summary), cudaMemcpyDeviceToHost, st));
    S.reserved[2] += since(t_host);  // host: everything issued between the node-count sync and the summary sync
    CK(cudaStreamSynchronize(st));
    t_host = now();
    S.d2h_bytes += n * sizeof(pgpu_contig_summary);
    for (const auto &K : chains) S.pairs += (int64_t)h_ext_pairs[K.ext];

    tr("sync3 done (summaries)");
    // ---- compact genes + their start/stop node records, copy back ---------------------------------------
    std::vector<int64_t> gene_out_off(n + 1, 0);
    for (int c = 0; c < n; c++) gene_out_off[c + 1] = gene_out_off[c] + summ[c].n_genes;
    const int64_t ng = gene_out_off[n];
    int64_t *d_gene_out_off = pool.upload(gene_out_off);
    pgpu_gene *d_genes_out = pool.alloc<pgpu_gene>(ng);
    pgpu_node *d_gene_nodes = pool.alloc<pgpu_node>(2 * ng);
    if (pool.failed) return PGPU_ENOMEM;
    launch_pack_gene_nodes(n, d_summ, d_genes, d_gene_off, d_gene_out_off, d_node_out_off, d_nodes, d_gene_nodes, d_genes_out, st);
    ctx->launches++;
    ResSeg seg;
    seg.lo = lo; seg.hi = hi; seg.g0 = res->total_genes; seg.ng = ng;
    const size_t gbytes = ng * sizeof(pgpu_gene), nbytes = 2 * ng * sizeof(pgpu_node);
    if (ng) {
        seg.buf = res->pinned->acquire(gbytes + nbytes);
        if (!seg.buf.p) return fail(ctx, PGPU_ENOMEM, "pinned result allocation failed");
        seg.genes = (pgpu_gene *)seg.buf.p;
        seg.gnodes = (pgpu_node *)((char *)seg.buf.p + gbytes);
        CK(cudaMemcpyAsync(seg.genes, d_genes_out, gbytes, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(seg.gnodes, d_gene_nodes, nbytes, cudaMemcpyDeviceToHost, st));
        S.d2h_bytes += gbytes + nbytes;
    }
    const int64_t g0 = res->total_genes;
    res->segs.push_back(seg);
    res->total_genes += ng;
    if (opts.want_nodes) {
        const size_t n0 = res->nodes.size();
        int64_t tot = 0;
        for (int c = 0; c < n; c++) tot += summ[c].n_nodes;
        res->nodes.resize(n0 + tot);
        int64_t w = n0;
        for (int c = 0; c < n; c++) {
            res->node_off[lo + c] = w;
            if (summ[c].n_nodes)
                CK(cudaMemcpyAsync(res->nodes.data() + w, d_nodes + node_out_off[c], summ[c].n_nodes * sizeof(pgpu_node),
                                   cudaMemcpyDeviceToHost, st));
            w += summ[c].n_nodes;
        }
        res->node_off[hi] = w;
        S.d2h_bytes += tot * sizeof(pgpu_node);
        res->have_nodes = true;
    }
    tr("issued pack + d2h");
    tev("compaction + d2h issue");
    int e_d2h = mark();
    S.reserved[3] += since(t_host);  // host: result compaction issue
    CK(cudaStreamSynchronize(st));
    for (int c = 0; c < n; c++) {
        res->summary[lo + c] = summ[c];
        res->gene_off[lo + c + 1] = g0 + gene_out_off[c + 1];
    }
    tr("sync4 done");
    if (trace) {
        for (size_t k = 1; k < tevs.size(); k++) {
            float t = 0;
            cudaEventElapsedTime(&t, tevs[k - 1].second, tevs[k].second);
            fprintf(stderr, "[pgpu dev] %-32s %8.3f ms\n", tevs[k].first, t);
        }
        for (auto &p : tevs) cudaEventDestroy(p.second);
    }
    S.total_genes += ng;
    auto ms = [&](int a, int b) { float t = 0; cudaEventElapsedTime(&t, ctx->ev[a], ctx->ev[b]); return (double)t; };
    S.ms_h2d += ms(e_start, e_h2d);
    S.ms_encode += ms(e_h2d, e_enc);
    S.ms_extract += ms(e_enc, e_mark) + ms(e_mark, e_ext);
    S.ms_score += ms(e_ext, e_score);
    S.ms_overlap += ms(e_score, e_ovl);
    S.ms_dp += ms(e_ovl, e_dp);
    S.ms_trace += ms(e_dp, e_trace);
    S.ms_final += ms(e_trace, e_final);
    S.ms_d2h += ms(e_final, e_d2h);
    S.ms_total_device += ms(e_h2d, e_final);
    return PGPU_OK;
}

// ------------------------------------------------------------------------------------------------
// training (GeneFinder._train, lib.pyx:5236-5279) on the extraction run_range has just prepared
// ------------------------------------------------------------------------------------------------
static int train_stage(pgpu_ctx *ctx, DevPool &pool, DevBatch &B, const ExtractInfo &X, RunOpts ro, int gc_count,
                       const TrainRequest &R, pgpu_stats &S) {
    namespace th = pgpu::train_host;
    using namespace pgpu::train;
    cudaStream_t st = ctx->stream;
    const int nn = X.nn, slen = X.slen;
    RawTraining &T = *R.out;
    memset(&T, 0, sizeof(T));   // TrainingInfo.__init__ (lib.pyx:3960-4003)
    T.gc = slen > 0 ? (double)gc_count / (double)slen : 0.0;
    T.trans_table = R.tt; T.st_wt = R.st_wt; T.uses_sd = 1;
    cudaEventRecord(ctx->ev[8], st);

    // one model slot and one chain (chain-node offset 0) on the device
    RawTraining *d_raw = pool.alloc<RawTraining>(1);
    DevModel *d_model = pool.alloc<DevModel>(1);
    std::vector<ChainInfo> chain(1);
    memset(&chain[0], 0, sizeof(ChainInfo));
    chain[0].first_pass = 1; chain[0].node_off = X.node_off; chain[0].nn = nn; chain[0].doff = X.doff; chain[0].slen = slen;
    chain[0].istride = 1;   // one chain: the interleaved arrays are plain arrays (ioff = coff = 0)
    B.chains = pool.upload(chain);
    B.ext_chains = nullptr; B.dcT = nullptr; B.n_models = 1;
    const size_t n1 = std::max(nn, 1);
    B.cscore = pool.alloc<double>(n1, true); B.sscore = pool.alloc<double>(n1, true);
    B.rscore = pool.alloc<double>(n1, true); B.uscore = pool.alloc<double>(n1, true); B.tscore = pool.alloc<double>(n1, true);
    B.opv = pool.alloc<double>(3 * n1, true); B.gcb = pool.alloc<double>(n1, true);
    B.star_ptr = pool.alloc<int32_t>(3 * n1); B.rbs = pool.alloc<uint8_t>(2 * n1 + 16, true);
    B.score = pool.alloc<double>(n1, true); B.traceb = pool.alloc<int32_t>(n1); B.ov_mark = pool.alloc<int8_t>(n1 + 16, true);
    B.chain_ipath = pool.alloc<int32_t>(1); B.chain_score = pool.alloc<double>(1);
    if (ctx->dp_algo >= 1) { B.dp_svig = pool.alloc<double>(n1); B.dp_tbig = pool.alloc<int32_t>(n1); }   // k_dp_dq<.., 0>
    TrainView V;
    memset(&V, 0, sizeof(V));
    V.N.ndx = B.ndx + X.node_off; V.N.sv = B.stop_val + X.node_off; V.N.cls = B.cls + X.node_off; V.N.nn = nn; V.N.slen = slen;
    V.node_off = X.node_off; V.cb = B.cbase;   // extraction 0
    V.gp = pool.alloc<int8_t>((size_t)slen + 16);
    V.gc_score = pool.alloc<double>(3 * n1, true); V.gc_bias = pool.alloc<int8_t>(n1 + 16, true);
    V.term = pool.alloc<double>(n1, true);
    V.cscore = B.cscore; V.rbs = B.rbs; V.upc = B.upc + X.node_off; V.umot = B.umot + X.node_off;
    V.mot = pool.alloc<MotifOut>(n1, true);
    const int icap = nn / 2 + 2;
    int4 *d_iv = pool.alloc<int4>(icap);
    int *d_niv = pool.alloc<int>(1, true);
    double *d_bias = pool.alloc<double>(4, true);
    uint32_t *d_dc = pool.alloc<uint32_t>(2 * 4096, true);             // background | genes
    unsigned long long *d_gene_total = pool.alloc<unsigned long long>(1, true);
    uint32_t *d_cnt = pool.alloc<uint32_t>(C_TOTAL, true);
    uint32_t *d_cells = pool.alloc<uint32_t>(2 * (size_t)kMotCells);    // background | accepted
    if (pool.failed) return PGPU_ENOMEM;

    // the counting passes as a backend of train_host::run_training (which owns the schedule and the host math)
    struct Gpu {
        pgpu_ctx *ctx; cudaStream_t st; DevBatch &B; TrainView &V; const ExtractInfo &X; RunOpts ro;
        RawTraining *d_raw; DevModel *d_model; DevModel hm;
        int4 *d_iv; int *d_niv; int icap; double *d_bias; uint32_t *d_dc; unsigned long long *d_gene_total;
        uint32_t *d_cnt, *d_cells;
        // PGPU_TRAIN_DUMP=<prefix>: write intermediate device arrays to <prefix>.<name>.bin (diagnostics for the
        // GPU parity tests: tells which stage diverges from the oracle first)
        void dump(const char *name, const void *dptr, size_t bytes) {
            const char *prefix = getenv("PGPU_TRAIN_DUMP");
            if (!prefix || !bytes) return;
            std::vector<char> h(bytes);
            cudaMemcpyAsync(h.data(), dptr, bytes, cudaMemcpyDeviceToHost, st);
            cudaStreamSynchronize(st);
            FILE *f = fopen((std::string(prefix) + "." + name + ".bin").c_str(), "wb");
            if (f) { fwrite(h.data(), 1, bytes, f); fclose(f); }
        }
        int err(cudaError_t e, const char *what) {
            if (e == cudaSuccess) return 0;
            ctx->err = std::string(what) + ": " + cudaGetErrorString(e);
            return PGPU_ECUDA;
        }
        void upload_model(const RawTraining &T) {
            prepare_model(T, hm, d_raw);
            cudaMemcpyAsync(d_raw, &T, sizeof(T), cudaMemcpyHostToDevice, st);
            cudaMemcpyAsync(d_model, &hm, sizeof(hm), cudaMemcpyHostToDevice, st);
        }
        int first_gene_set(const RawTraining &T, double *bias, uint32_t *dicodon, long long *gene_codons) {
            const int nn = X.nn, slen = X.slen;
            launch_gc_frame(B.gcbits + (X.doff >> 5), slen, V.gp, st);
            launch_gc_bias(B, V, d_bias, B.gcb, st);
            upload_model(T);
            launch_overlap(B, d_model, 1, nn, nn, 1, ro, 0, st); // first start of each frame, no scores yet
            launch_dp(B, d_model, nullptr, 1, 0, ctx->dp_algo, st);   // final == 0: GC frame bias is the only score
            launch_training_path(B, V, d_iv, icap, d_niv, st);
            launch_dicodon(B.digits + X.doff, slen, d_iv, d_niv, icap, d_dc, d_dc + 4096, d_gene_total, st);
            ctx->launches += 10;
            unsigned long long total = 0;
            cudaMemcpyAsync(dicodon, d_dc, 2 * 4096 * 4, cudaMemcpyDeviceToHost, st);
            cudaMemcpyAsync(&total, d_gene_total, 8, cudaMemcpyDeviceToHost, st);
            cudaMemcpyAsync(bias, d_bias, 3 * 8, cudaMemcpyDeviceToHost, st);
            cudaEventRecord(ctx->ev[9], st);
            const int rc = err(cudaStreamSynchronize(st), "training: first gene set");
            *gene_codons = (long long)total;
            if (!rc && getenv("PGPU_TRAIN_DUMP")) {
                dump("gp", V.gp, (size_t)slen); dump("gc_score", V.gc_score, 24 * (size_t)nn);
                dump("gc_bias", V.gc_bias, (size_t)nn); dump("gcb", B.gcb, 8 * (size_t)nn);
                dump("star_ptr", B.star_ptr, 12 * (size_t)nn); dump("score", B.score, 8 * (size_t)nn);
                dump("traceb", B.traceb, 4 * (size_t)nn); dump("ipath", B.chain_ipath, 4);
                dump("n_intervals", d_niv, 4); dump("dicodon", d_dc, 2 * 4096 * 4);
                dump("ndx", V.N.ndx, 4 * (size_t)nn); dump("cls", V.N.cls, (size_t)nn);
            }
            return rc;
        }
        int score_starts(const RawTraining &T, uint32_t *cnt) {
            upload_model(T);
            launch_coding(B, d_model, 1, X.nn, 1, X.nn, st);
            launch_start_score(B, d_model, 1, X.nn, ro, nullptr, st);   // rbs[2] of every non-edge start
            cudaMemsetAsync(d_cnt, 0, C_TOTAL * 4, st);
            launch_type_background(V, d_cnt, st);
            ctx->launches += 3;
            cudaMemcpyAsync(cnt, d_cnt, C_TOTAL * 4, cudaMemcpyDeviceToHost, st);
            const int rc = err(cudaStreamSynchronize(st), "training: start scores");
            if (!rc && getenv("PGPU_TRAIN_DUMP")) { dump("cscore", B.cscore, 8 * (size_t)X.nn); dump("rbs", B.rbs, 2 * (size_t)X.nn); }
            return rc;
        }
        int sd_iteration(const SdParams &P, uint32_t *cnt) {
            cudaMemsetAsync(d_cnt, 0, C_TOTAL * 4, st);
            launch_sd_iteration(B, V, P, d_cnt, st);
            ctx->launches++;
            cudaMemcpyAsync(cnt, d_cnt, C_TOTAL * 4, cudaMemcpyDeviceToHost, st);
            return err(cudaStreamSynchronize(st), "training: SD iteration");
        }
        int motif_iteration(const MotParams &P, const RawTraining &T, uint32_t *cells, uint32_t *cnt) {
            double *d_mot_wt = &d_raw->mot_wt[0][0][0];
            cudaMemcpyAsync(d_mot_wt, &T.mot_wt[0][0][0], sizeof(T.mot_wt), cudaMemcpyHostToDevice, st);
            cudaMemsetAsync(d_cnt, 0, C_TOTAL * 4, st);
            cudaMemsetAsync(d_cells, 0, 2 * (size_t)kMotCells * 4, st);
            launch_motif_iteration(B, V, d_mot_wt, P, d_cells, d_cells + kMotCells, d_cnt, st);
            ctx->launches += 2;
            cudaMemcpyAsync(cells, d_cells, 2 * (size_t)kMotCells * 4, cudaMemcpyDeviceToHost, st);
            cudaMemcpyAsync(cnt, d_cnt, C_TOTAL * 4, cudaMemcpyDeviceToHost, st);
            return err(cudaStreamSynchronize(st), "training: motif iteration");
        }
    } gpu{ctx, st, B, V, X, ro, d_raw, d_model, DevModel(), d_iv, d_niv, icap, d_bias, d_dc, d_gene_total, d_cnt, d_cells};
    cudaEventRecord(ctx->ev[9], st);  // re-recorded after the first gene set
    const int rc = th::run_training(gpu, T, nn, slen, R.force_nonsd);
    if (rc) return rc;
    cudaEventRecord(ctx->ev[10], st);
    CK(cudaStreamSynchronize(st));
    float t1 = 0, t2 = 0;
    cudaEventElapsedTime(&t1, ctx->ev[8], ctx->ev[9]);
    cudaEventElapsedTime(&t2, ctx->ev[9], ctx->ev[10]);
    S.ms_dp += t1;       // frame plot + bias + training DP + dicodon counts
    S.ms_score += t2;    // coding / SD scores + the start-training rounds
    S.ms_total_device += t1 + t2;
    S.n_chains = 1; S.total_chain_nodes = nn; S.dp_steps = nn;
    return PGPU_OK;
}

// small gather kernel used by run_range (kept here: it is plumbing, not part of the hot path)
__global__ void k_gather_i32(const int32_t *__restrict__ src, const int64_t *__restrict__ idx, int n, int32_t *__restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}
void launch_gather_i32(const int32_t *src, const int64_t *idx, int n, int32_t *dst, cudaStream_t st) {
    if (n > 0) k_gather_i32<<<(n + 255) / 256, 256, 0, st>>>(src, idx, n, dst);
}

// ------------------------------------------------------------------------------------------------
// batching: split the contigs into sub-batches that fit the workspace
// ------------------------------------------------------------------------------------------------
static int check_opts(pgpu_ctx *ctx, const pgpu_opts *o) {
    if (!o) return fail(ctx, PGPU_EINVAL, "opts is NULL");
    if (o->min_gene <= 0) return fail(ctx, PGPU_EINVAL, "`min_gene` must be strictly positive");
    if (o->min_edge_gene <= 0) return fail(ctx, PGPU_EINVAL, "`min_edge_gene` must be strictly positive");
    if (o->min_mask < 0) return fail(ctx, PGPU_EINVAL, "`min_mask` must be positive");
    if (o->max_overlap < 0) return fail(ctx, PGPU_EINVAL, "`max_overlap` must be positive");
    if (o->max_overlap > o->min_gene) return fail(ctx, PGPU_EINVAL, "`max_overlap` must be lower than `min_gene`");
    if (!o->meta && (o->single_model < 0 || o->single_model >= ctx->n_models))
        return fail(ctx, PGPU_ESTATE, "cannot find genes without having trained in single mode");
    return PGPU_OK;
}

// Two worker threads take the sub-batches in order, each on its own stream with its own copy of the context's host
// state; the partial results are stitched together in sub-batch order afterwards.  The lane streams are forked
// from / joined to the context's stream with events, so work queued on that stream (e.g. the stopwatch events of
// pgpu_timer_*) still brackets the whole call.
static int run_lanes(pgpu_ctx *ctx, const uint8_t *h_seq, const uint8_t *d_seq, const int64_t *offsets, int n, const pgpu_opts &opts,
                     RunPlan &plan, const std::vector<std::pair<int, int>> &ranges, pgpu_result *res) {
    if (!ctx->lane_stream[0]) {
        for (int k = 0; k < 2; k++) {
            if (cudaStreamCreateWithFlags(&ctx->lane_stream[k], cudaStreamNonBlocking) != cudaSuccess)
                return fail(ctx, PGPU_ECUDA, "cudaStreamCreate (lane) failed");
            for (auto &ev : ctx->lane_ev[k]) cudaEventCreate(&ev);
            cudaEventCreateWithFlags(&ctx->join_ev[k], cudaEventDisableTiming);
        }
        cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming);
    }
    CK(cudaEventRecord(ctx->fork_ev, ctx->stream));
    const int P = (int)ranges.size();
    std::vector<std::unique_ptr<pgpu_result>> parts(P);
    std::vector<int> prc(P, PGPU_OK);
    std::vector<std::string> perr(P);
    std::atomic<int> next{0};
    std::atomic<int64_t> launches{0};
    // the input copies are chained in sub-batch order: two copies issued together would share the link and finish
    // together, and neither lane could start computing before both were done
    std::vector<cudaEvent_t> h2d_ev(P, nullptr);
    std::unique_ptr<std::atomic<int>[]> h2d_issued(new std::atomic<int>[P]);
    for (int p = 0; p < P; p++) { cudaEventCreateWithFlags(&h2d_ev[p], cudaEventDisableTiming); h2d_issued[p].store(0); }
    auto worker = [&](int k) {
        cudaSetDevice(ctx->device);
        pgpu_ctx L = *ctx;          // host-side state is small (prepared model tables); device pointers are shared
        L.stream = ctx->lane_stream[k];
        memcpy(L.ev, ctx->lane_ev[k], sizeof(L.ev));
        L.launches = 0;
        cudaStreamWaitEvent(L.stream, ctx->fork_ev, 0);
        for (;;) {
            const int p = next.fetch_add(1);
            if (p >= P) break;
            parts[p].reset(new pgpu_result());
            pgpu_result *r = parts[p].get();
            r->pinned = ctx->pinned;
            r->n_contigs = n;
            r->summary.resize(n);
            r->gene_off.assign(n + 1, 0);
            r->node_off.assign(n + 1, 0);
            memset(&r->stats, 0, sizeof(r->stats));
            L.err.clear();
            RunPlan pl = plan;
            pl.before_h2d = [&, p](cudaStream_t st) {
                if (p == 0) return;
                while (!h2d_issued[p - 1].load()) std::this_thread::yield();   // its record must be queued first
                cudaStreamWaitEvent(st, h2d_ev[p - 1], 0);
            };
            pl.after_h2d = [&, p](cudaStream_t st) {
                cudaEventRecord(h2d_ev[p], st);
                h2d_issued[p].store(1);
            };
            prc[p] = run_range(&L, h_seq, d_seq, offsets, ranges[p].first, ranges[p].second, opts, pl, r, nullptr);
            h2d_issued[p].store(1);  // also when the sub-batch failed before its copy: nobody may wait forever
            if (prc[p]) { perr[p] = L.err; break; }
        }
        cudaEventRecord(ctx->join_ev[k], L.stream);
        launches += L.launches;
    };
    std::thread t1(worker, 1);
    worker(0);
    t1.join();
    for (int k = 0; k < 2; k++) cudaStreamWaitEvent(ctx->stream, ctx->join_ev[k], 0);
    for (auto &ev : h2d_ev) cudaEventDestroy(ev);
    ctx->launches += launches.load();
    for (int p = 0; p < P; p++)
        if (prc[p]) return fail(ctx, prc[p], perr[p]);
    // stitch: genes of part p follow those of parts 0..p-1
    for (int p = 0; p < P; p++) {
        pgpu_result *r = parts[p].get();
        if (!r) return fail(ctx, PGPU_ECUDA, "lane left a sub-batch unprocessed");
        const int64_t shift = res->total_genes;
        for (auto &sg : r->segs) { sg.g0 += shift; res->segs.push_back(sg); }
        r->segs.clear();  // the pinned buffers now belong to `res`
        const size_t n0 = res->nodes.size();
        if (r->have_nodes) {
            res->nodes.insert(res->nodes.end(), r->nodes.begin(), r->nodes.end());
            res->have_nodes = true;
        }
        for (int c = ranges[p].first; c < ranges[p].second; c++) {
            res->summary[c] = r->summary[c];
            res->gene_off[c + 1] = r->gene_off[c + 1] + shift;
            if (r->have_nodes) res->node_off[c] = r->node_off[c] + (int64_t)n0;
        }
        if (r->have_nodes) res->node_off[ranges[p].second] = r->node_off[ranges[p].second] + (int64_t)n0;
        res->total_genes += r->total_genes;
        pgpu_stats &S = res->stats;
        const pgpu_stats &T = r->stats;
        S.n_contigs += T.n_contigs; S.total_bp += T.total_bp; S.total_nodes += T.total_nodes;
        S.total_chain_nodes += T.total_chain_nodes; S.n_chains += T.n_chains; S.total_genes += T.total_genes;
        S.pairs += T.pairs; S.dp_steps += T.dp_steps; S.h2d_bytes += T.h2d_bytes; S.d2h_bytes += T.d2h_bytes;
        // phase times are per-lane stream times and overlap between lanes: sums, not wall time
        S.ms_total_device += T.ms_total_device; S.ms_encode += T.ms_encode; S.ms_extract += T.ms_extract;
        S.ms_score += T.ms_score; S.ms_overlap += T.ms_overlap; S.ms_dp += T.ms_dp; S.ms_trace += T.ms_trace;
        S.ms_final += T.ms_final; S.ms_h2d += T.ms_h2d; S.ms_d2h += T.ms_d2h;
        for (int q = 0; q < 4; q++) S.reserved[q] += T.reserved[q];
    }
    return PGPU_OK;
}

static int run_all(pgpu_ctx *ctx, const uint8_t *h_seq, const uint8_t *d_seq, const int64_t *offsets, int n,
                   const pgpu_opts *opts, pgpu_result **out) {
    if (!ctx) return PGPU_EINVAL;
    if (!out || !offsets || n < 0 || (!h_seq && !d_seq && n > 0 && offsets[n] > offsets[0]))
        return fail(ctx, PGPU_EINVAL, "bad arguments");
    int rc = check_opts(ctx, opts);
    if (rc) return rc;
    if (ctx->n_models == 0) return fail(ctx, PGPU_ESTATE, "no models loaded");
    cudaSetDevice(ctx->device);
    const bool trace = getenv("PGPU_TRACE") != nullptr;
    const auto t_entry = std::chrono::steady_clock::now();
    auto tr = [&](const char *what) {
        if (trace) fprintf(stderr, "[pgpu api %8.2f ms] %s\n",
                           std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_entry).count(), what);
    };
    tr("entry");
    pgpu_result *res = new pgpu_result();
    res->pinned = ctx->pinned;
    res->n_contigs = n;
    res->summary.resize(n);
    res->gene_off.assign(n + 1, 0);
    res->node_off.assign(n + 1, 0);
    memset(&res->stats, 0, sizeof(res->stats));
    const int64_t launches0 = ctx->launches;
    // sub-batch size: bytes per base pair is dominated by per-chain node arrays (~100 B per chain-node,
    // ~0.06 nodes/bp, up to ~26 chains): budget 160 B/bp against the workspace limit
    size_t freeb = 0, totalb = 0;
    cudaMemGetInfo(&freeb, &totalb);
    {   // blocks cached by the stream-ordered pool are reusable: count them as free
        cudaMemPool_t mp;
        uint64_t reserved = 0, used = 0;
        if (cudaDeviceGetDefaultMemPool(&mp, ctx->device) == cudaSuccess &&
            cudaMemPoolGetAttribute(mp, cudaMemPoolAttrReservedMemCurrent, &reserved) == cudaSuccess &&
            cudaMemPoolGetAttribute(mp, cudaMemPoolAttrUsedMemCurrent, &used) == cudaSuccess && reserved > used)
            freeb += reserved - used;
    }
    size_t limit = ctx->ws_limit ? ctx->ws_limit : (size_t)(0.6 * (double)freeb);
    const int64_t total_bp = n > 0 ? offsets[n] - offsets[0] : 0;
    // (device-resident batches stay on one stream: two lanes measured 6.43 against 6.48 Gbp/s on the bench shard)
    const int lanes = (h_seq && !d_seq && ctx->lanes > 1 && n >= 2 && total_bp >= ctx->lane_min_bp) ? 2 : 1;
    const int64_t bp_budget = std::max<int64_t>((int64_t)(limit / 160) / lanes, 1 << 20);
    RunPlan plan;
    std::vector<std::pair<int, int>> ranges;
    for (int lo = 0; lo < n;) {
        int hi = lo;
        int64_t bp = 0;
        while (hi < n && (hi == lo || bp + (offsets[hi + 1] - offsets[hi]) <= bp_budget) && hi - lo < (1 << 22)) {
            bp += offsets[hi + 1] - offsets[hi];
            hi++;
        }
        ranges.push_back({lo, hi});
        lo = hi;
    }
    if (lanes == 2 && ranges.size() == 1) {  // one sub-batch: cut it where the bases split evenly
        int mid = 1;
        while (mid < n - 1 && offsets[mid] - offsets[0] < total_bp / 2) mid++;
        ranges = {{0, mid}, {mid, n}};
    }
    if (lanes == 1 || ranges.size() < 2) {
        for (auto &r : ranges) {
            tr("sub-batch begin");
            rc = run_range(ctx, h_seq, d_seq, offsets, r.first, r.second, *opts, plan, res, nullptr);
            tr("sub-batch end (buffers released)");
            if (rc) { delete res; return rc; }
        }
    } else {
        rc = run_lanes(ctx, h_seq, d_seq, offsets, n, *opts, plan, ranges, res);
        tr("lanes joined");
        if (rc) { delete res; return rc; }
    }
    res->stats.kernel_launches = ctx->launches - launches0;
    *out = res;
    tr("exit");
    return PGPU_OK;
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int pgpu_create(int device, pgpu_ctx **out) {
    if (!out) return PGPU_EINVAL;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_create_err = std::string("no CUDA device: ") + cudaGetErrorString(e);
        return PGPU_ENODEV;
    }
    if (device < 0 || device >= count) { g_create_err = "device index out of range"; return PGPU_ENODEV; }
    if (cudaSetDevice(device) != cudaSuccess) { g_create_err = "cudaSetDevice failed"; return PGPU_ENODEV; }
    pgpu_ctx *ctx = new pgpu_ctx();
    ctx->device = device;
    ctx->pinned = std::make_shared<PinnedPool>();
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        g_create_err = "cudaStreamCreate failed";
        delete ctx;
        return PGPU_ECUDA;
    }
    for (auto &ev : ctx->ev) cudaEventCreate(&ev);
    if (const char *a = getenv("PGPU_DP_ALGO")) ctx->dp_algo = atoi(a);
    if (const char *a = getenv("PGPU_DP_VERIFY")) ctx->dp_verify = atoi(a) != 0;
    if (const char *a = getenv("PGPU_CODING_VERIFY")) ctx->coding_verify = atoi(a) != 0;
    if (const char *a = getenv("PGPU_DP_ML_MINB")) ctx->dp_ml_minb = atoi(a);
    if (const char *a = getenv("PGPU_EXTRACT_ALGO")) ctx->extract_algo = atoi(a);
    if (const char *a = getenv("PGPU_FINAL_ALGO")) ctx->final_algo = atoi(a);
    if (const char *a = getenv("PGPU_CODON_LUT")) ctx->codon_lut = atoi(a) != 0;
    if (const char *a = getenv("PGPU_CODING_SMEM")) ctx->coding_smem = atoi(a);
    if (const char *a = getenv("PGPU_LANES")) ctx->lanes = atoi(a);
    if (const char *a = getenv("PGPU_LANE_MIN_BP")) ctx->lane_min_bp = atoll(a);
    if (const char *a = getenv("PGPU_WS_LIMIT_MB")) ctx->ws_limit = (size_t)atoll(a) << 20;   // = pgpu_set_workspace_limit
    // keep freed blocks cached in the stream-ordered pool: sub-batches reuse them without going to the driver
    cudaMemPool_t mp;
    if (cudaDeviceGetDefaultMemPool(&mp, device) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    *out = ctx;
    return PGPU_OK;
}

void pgpu_destroy(pgpu_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->d_raw) cudaFree(ctx->d_raw);
    if (ctx->d_models) cudaFree(ctx->d_models);
    if (ctx->d_dcT) cudaFree(ctx->d_dcT);
    if (ctx->d_dcS) cudaFree(ctx->d_dcS);
    if (ctx->d_live) cudaFree(ctx->d_live);
    for (auto &ev : ctx->ev) cudaEventDestroy(ev);
    if (ctx->lane_stream[0]) {
        for (int k = 0; k < 2; k++) {
            cudaStreamSynchronize(ctx->lane_stream[k]);
            for (auto &ev : ctx->lane_ev[k]) cudaEventDestroy(ev);
            cudaEventDestroy(ctx->join_ev[k]);
            cudaStreamDestroy(ctx->lane_stream[k]);
        }
        cudaEventDestroy(ctx->fork_ev);
    }
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *pgpu_last_error(const pgpu_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int pgpu_set_models(pgpu_ctx *ctx, const void *blobs, int n, size_t stride) {
    if (!ctx) return PGPU_EINVAL;
    if (!blobs || n <= 0 || stride < sizeof(RawTraining)) return fail(ctx, PGPU_EINVAL, "bad model blobs");
    std::vector<RawTraining> h_raw_v(n);   // host copy of the raw structs, only needed to prepare the device tables
    for (int k = 0; k < n; k++) memcpy(&h_raw_v[k], (const char *)blobs + k * stride, sizeof(RawTraining));
    for (int k = 0; k < n; k++) {
        // the translation tables the reference accepts (lib.pyx TRANSLATION_TABLES: 1-6, 9-16, 21-33); the value
        // selects codon sets in every kernel, so a corrupt blob is refused here
        const int tt = h_raw_v[k].trans_table;
        const bool ok = (tt >= 1 && tt <= 6) || (tt >= 9 && tt <= 16) || (tt >= 21 && tt <= 33);
        if (!ok) return fail(ctx, PGPU_EINVAL, "model " + std::to_string(k) + ": translation table " + std::to_string(tt) + " is not valid");
    }
    cudaSetDevice(ctx->device);
    CK(cudaStreamSynchronize(ctx->stream));
    // Build the new device tables next to the old ones and swap them in only when every step has succeeded: a failed
    // call leaves the context with the model set it had.
    RawTraining *d_raw = nullptr;
    DevModel *d_models = nullptr;
    uint32_t *d_live = nullptr;
    double *d_dcT = nullptr, *d_dcS = nullptr;
    std::vector<DevModel> h_models(n);
    auto build = [&]() -> cudaError_t {
        cudaError_t e;
        if ((e = cudaMalloc(&d_raw, n * sizeof(RawTraining))) != cudaSuccess) return e;
        if ((e = cudaMalloc(&d_models, n * sizeof(DevModel))) != cudaSuccess) return e;
        if ((e = cudaMalloc(&d_live, (size_t)n * kMotifWords * sizeof(uint32_t))) != cudaSuccess) return e;
        std::vector<uint32_t> live((size_t)n * kMotifWords);
        for (int k = 0; k < n; k++) motif_live_bits(h_raw_v[k], live.data() + (size_t)k * kMotifWords);
        if ((e = cudaMemcpy(d_live, live.data(), live.size() * sizeof(uint32_t), cudaMemcpyHostToDevice)) != cudaSuccess) return e;
        for (int k = 0; k < n; k++) prepare_model(h_raw_v[k], h_models[k], d_raw + k, d_live + (size_t)k * kMotifWords);
        // transposed dicodon table: models that are evaluated together (same table, neighbouring GC) get
        // neighbouring columns, so the lanes of k_coding_orf read one or two cache lines per codon
        std::vector<int> ord(n);
        std::iota(ord.begin(), ord.end(), 0);
        std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) {
            const RawTraining &x = h_raw_v[a], &y = h_raw_v[b];
            return x.trans_table != y.trans_table ? x.trans_table < y.trans_table : x.gc < y.gc;
        });
        // rows are kDcCols wide (a compile-time stride: one multiply-add per weight address in the kernel); a model
        // set with more columns than that runs the per-chain kernel k_coding instead (d_dcT stays null)
        for (int c = 0; c < n; c++) h_models[ord[c]].col = c;
        if (n <= kDcCols) {
            std::vector<double> t((size_t)4096 * kDcCols, 0.0);
            for (int c = 0; c < n; c++)
                for (int i = 0; i < 4096; i++) t[(size_t)i * kDcCols + c] = h_raw_v[ord[c]].gene_dc[i];
            if ((e = cudaMalloc(&d_dcT, t.size() * sizeof(double))) != cudaSuccess) return e;
            if ((e = cudaMemcpy(d_dcT, t.data(), t.size() * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess) return e;
            // table sets of k_coding_flat: set s = columns s .. s + 3 (zero past the last one), 128 KB each, contiguous
            // so that a CTA fetches its set with a few bulk copies
            std::vector<double> ts((size_t)n * 4096 * 4, 0.0);
            for (int c = 0; c < n; c++)
                for (int k = 0; k < 4 && c + k < n; k++) {
                    const double *src = h_raw_v[ord[c + k]].gene_dc;
                    double *dst = ts.data() + (size_t)c * 4096 * 4 + k;
                    for (int i = 0; i < 4096; i++) dst[(size_t)i * 4] = src[i];
                }
            if ((e = cudaMalloc(&d_dcS, ts.size() * sizeof(double))) != cudaSuccess) return e;
            if ((e = cudaMemcpy(d_dcS, ts.data(), ts.size() * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess) return e;
        }
        if ((e = cudaMemcpy(d_raw, h_raw_v.data(), n * sizeof(RawTraining), cudaMemcpyHostToDevice)) != cudaSuccess) return e;
        return cudaMemcpy(d_models, h_models.data(), n * sizeof(DevModel), cudaMemcpyHostToDevice);
    };
    const cudaError_t e = build();
    if (e != cudaSuccess) {
        cudaFree(d_raw); cudaFree(d_models); cudaFree(d_live); cudaFree(d_dcT); cudaFree(d_dcS);
        return fail(ctx, e == cudaErrorMemoryAllocation ? PGPU_ENOMEM : PGPU_ECUDA, std::string("pgpu_set_models: ") + cudaGetErrorString(e));
    }
    cudaFree(ctx->d_raw); cudaFree(ctx->d_models); cudaFree(ctx->d_live); cudaFree(ctx->d_dcT); cudaFree(ctx->d_dcS);
    ctx->d_raw = d_raw; ctx->d_models = d_models; ctx->d_live = d_live; ctx->d_dcT = d_dcT; ctx->d_dcS = d_dcS;
    ctx->h_models.swap(h_models);
    ctx->n_models = n;
    ctx->model_gc.resize(n);
    ctx->model_tt.resize(n);
    for (int k = 0; k < n; k++) { ctx->model_gc[k] = h_raw_v[k].gc; ctx->model_tt[k] = h_raw_v[k].trans_table; }
    return PGPU_OK;
}

int pgpu_num_models(const pgpu_ctx *ctx) { return ctx ? ctx->n_models : 0; }

void *pgpu_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void pgpu_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

int pgpu_set_workspace_limit(pgpu_ctx *ctx, size_t bytes) {
    if (!ctx) return PGPU_EINVAL;
    ctx->ws_limit = bytes;
    return PGPU_OK;
}

int pgpu_timer_start(pgpu_ctx *ctx) {
    if (!ctx) return PGPU_EINVAL;
    cudaSetDevice(ctx->device);
    CK(cudaEventRecord(ctx->ev[14], ctx->stream));
    return PGPU_OK;
}

int pgpu_timer_stop(pgpu_ctx *ctx, double *ms) {
    if (!ctx || !ms) return PGPU_EINVAL;
    cudaSetDevice(ctx->device);
    CK(cudaEventRecord(ctx->ev[15], ctx->stream));
    CK(cudaEventSynchronize(ctx->ev[15]));
    float t = 0;
    CK(cudaEventElapsedTime(&t, ctx->ev[14], ctx->ev[15]));
    *ms = t;
    return PGPU_OK;
}

int pgpu_find_genes_batch(pgpu_ctx *ctx, const uint8_t *seq, const int64_t *offsets, int n_contigs,
                          const pgpu_opts *opts, pgpu_result **out) {
    return run_all(ctx, seq, nullptr, offsets, n_contigs, opts, out);
}

int pgpu_batch_upload(pgpu_ctx *ctx, const uint8_t *seq, const int64_t *offsets, int n_contigs, pgpu_batch **out) {
    if (!ctx) return PGPU_EINVAL;
    if (!out || !offsets || n_contigs < 0) return fail(ctx, PGPU_EINVAL, "bad arguments");
    cudaSetDevice(ctx->device);
    pgpu_batch *b = new pgpu_batch();
    b->device = ctx->device;
    b->n_contigs = n_contigs;
    b->offsets.assign(offsets, offsets + n_contigs + 1);
    const int64_t base = offsets[0];
    for (auto &o : b->offsets) o -= base;
    b->total = b->offsets[n_contigs];
    cudaError_t e = cudaMalloc(&b->d_ascii, std::max<int64_t>(b->total, 1) + 16);
    if (e != cudaSuccess) { delete b; return fail(ctx, PGPU_ENOMEM, cudaGetErrorString(e)); }
    if (b->total) {
        e = cudaMemcpyAsync(b->d_ascii, seq + base, b->total, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { cudaFree(b->d_ascii); delete b; return fail(ctx, PGPU_ECUDA, cudaGetErrorString(e)); }
    }
    *out = b;
    return PGPU_OK;
}

int pgpu_batch_run(pgpu_ctx *ctx, pgpu_batch *batch, const pgpu_opts *opts, pgpu_result **out) {
    if (!ctx || !batch) return PGPU_EINVAL;
    return run_all(ctx, nullptr, batch->d_ascii, batch->offsets.data(), batch->n_contigs, opts, out);
}

int pgpu_batch_wrap_device(pgpu_ctx *ctx, const uint8_t *d_seq, const int64_t *offsets, int n_contigs, pgpu_batch **out) {
    if (!ctx) return PGPU_EINVAL;
    if (!out || !offsets || n_contigs < 0 || (!d_seq && n_contigs > 0 && offsets[n_contigs] > offsets[0]))
        return fail(ctx, PGPU_EINVAL, "bad arguments");
    pgpu_batch *b = new pgpu_batch();
    b->device = ctx->device;
    b->owned = false;
    b->n_contigs = n_contigs;
    b->offsets.assign(offsets, offsets + n_contigs + 1);
    const int64_t base = offsets[0];
    for (auto &o : b->offsets) o -= base;
    b->total = b->offsets[n_contigs];
    b->d_ascii = const_cast<uint8_t *>(d_seq) + base;
    *out = b;
    return PGPU_OK;
}

void pgpu_batch_free(pgpu_batch *b) {
    if (!b) return;
    if (b->d_ascii && b->owned) { cudaSetDevice(b->device); cudaFree(b->d_ascii); }
    delete b;
}

int pgpu_result_num_contigs(const pgpu_result *res) { return res ? res->n_contigs : PGPU_EINVAL; }

int pgpu_result_summaries(const pgpu_result *res, pgpu_contig_summary *dst) {
    if (!res || !dst) return PGPU_EINVAL;
    if (res->n_contigs) memcpy(dst, res->summary.data(), res->n_contigs * sizeof(pgpu_contig_summary));
    return PGPU_OK;
}

int pgpu_result_genes(const pgpu_result *res, int contig, pgpu_gene *dst) {
    if (!res || contig < 0 || contig >= res->n_contigs) return PGPU_EINVAL;
    const int64_t a = res->gene_off[contig], b = res->gene_off[contig + 1];
    if (b > a) {
        if (!dst) return PGPU_EINVAL;
        for (const auto &s : res->segs)
            if (contig >= s.lo && contig < s.hi) { memcpy(dst, s.genes + (a - s.g0), (b - a) * sizeof(pgpu_gene)); break; }
    }
    return PGPU_OK;
}

int pgpu_result_all_genes(const pgpu_result *res, pgpu_gene *dst) {
    if (!res) return PGPU_EINVAL;
    if (res->total_genes && !dst) return PGPU_EINVAL;
    for (const auto &s : res->segs)
        if (s.ng) memcpy(dst + s.g0, s.genes, s.ng * sizeof(pgpu_gene));
    return PGPU_OK;
}

int pgpu_result_gene_nodes(const pgpu_result *res, pgpu_node *dst) {
    if (!res) return PGPU_EINVAL;
    if (res->total_genes && !dst) return PGPU_EINVAL;
    for (const auto &s : res->segs)
        if (s.ng) memcpy(dst + 2 * s.g0, s.gnodes, 2 * s.ng * sizeof(pgpu_node));
    return PGPU_OK;
}

int pgpu_result_num_segments(const pgpu_result *res) { return res ? (int)res->segs.size() : PGPU_EINVAL; }

long long pgpu_result_segment(const pgpu_result *res, int k, long long *first_gene, const pgpu_gene **genes,
                              const pgpu_node **gene_nodes) {
    if (!res || k < 0 || k >= (int)res->segs.size()) return PGPU_EINVAL;
    const auto &s = res->segs[k];
    if (first_gene) *first_gene = s.g0;
    if (genes) *genes = s.ng ? s.genes : nullptr;
    if (gene_nodes) *gene_nodes = s.ng ? s.gnodes : nullptr;
    return s.ng;
}

int pgpu_result_nodes(const pgpu_result *res, int contig, pgpu_node *dst) {
    if (!res || contig < 0 || contig >= res->n_contigs) return PGPU_EINVAL;
    if (!res->have_nodes) return PGPU_ESTATE;
    const int64_t a = res->node_off[contig], b = res->node_off[contig + 1];
    if (b > a) { if (!dst) return PGPU_EINVAL; memcpy(dst, res->nodes.data() + a, (b - a) * sizeof(pgpu_node)); }
    return PGPU_OK;
}

// `struct _node` of the reference (src/Prodigal/node.h:41-76 as packed by Pyrodigal: 128 bytes): the motif record first
// (its bit fields share one 32-bit unit: ndx 12 bits, spacer 4, len 3, spacendx 2), then the doubles, gc_cont, the ints
// and the eight byte-sized fields.  gc_score / gc_bias are training state and are zero in a find_genes result.
int pgpu_result_nodes_struct(const pgpu_result *res, int contig, void *dst) {
    if (!res || contig < 0 || contig >= res->n_contigs) return PGPU_EINVAL;
    if (!res->have_nodes) return PGPU_ESTATE;
    const int64_t a = res->node_off[contig], b = res->node_off[contig + 1];
    if (b > a && !dst) return PGPU_EINVAL;
    unsigned char *out = static_cast<unsigned char *>(dst);
    for (int64_t k = a; k < b; k++, out += PGPU_NODE_STRUCT_SIZE) {
        const pgpu_node &n = res->nodes[(size_t)k];
        memset(out, 0, PGPU_NODE_STRUCT_SIZE);
        const uint32_t bits = (uint32_t)(n.mot_ndx & 0xfffu) | ((uint32_t)(n.mot_spacer & 0xfu) << 12) |
                              ((uint32_t)(n.mot_len & 0x7u) << 16) | ((uint32_t)(n.mot_spacendx & 0x3u) << 19);
        memcpy(out + 0, &n.mot_score, 8);
        memcpy(out + 8, &bits, 4);
        // out + 16: gc_score[3] = 0
        memcpy(out + 40, &n.cscore, 8); memcpy(out + 48, &n.uscore, 8); memcpy(out + 56, &n.tscore, 8);
        memcpy(out + 64, &n.rscore, 8); memcpy(out + 72, &n.sscore, 8); memcpy(out + 80, &n.score, 8);
        memcpy(out + 88, &n.gc_cont, 4);
        memcpy(out + 92, n.star_ptr, 12);
        memcpy(out + 104, &n.traceb, 4); memcpy(out + 108, &n.tracef, 4);
        memcpy(out + 112, &n.ndx, 4); memcpy(out + 116, &n.stop_val, 4);
        out[120] = (unsigned char)n.ov_mark; out[121] = (unsigned char)n.strand;
        out[122] = n.rbs[0]; out[123] = n.rbs[1];
        out[124] = n.edge; out[125] = n.elim; out[126] = 0; out[127] = n.type;
    }
    return PGPU_OK;
}

int pgpu_result_stats(const pgpu_result *res, pgpu_stats *dst) {
    if (!res || !dst) return PGPU_EINVAL;
    *dst = res->stats;
    return PGPU_OK;
}

void pgpu_result_free(pgpu_result *res) {
    const bool trace = res && getenv("PGPU_TRACE") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    delete res;
    if (trace) fprintf(stderr, "[pgpu api] result_free %.2f ms\n",
                       std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
}

// ---- operator-level twins ---------------------------------------------------------------------------

int pgpu_extract_nodes(pgpu_ctx *ctx, const uint8_t *seq, int slen, int translation_table, const pgpu_opts *opts,
                       int cap, int32_t *ndx, int32_t *stop_val, int8_t *strand, uint8_t *type, uint8_t *edge) {
    if (!ctx) return PGPU_EINVAL;
    if (!opts || slen < 0 || (!seq && slen > 0) || translation_table < 1 || translation_table > 33)
        return fail(ctx, PGPU_EINVAL, "bad arguments");
    cudaSetDevice(ctx->device);
    pgpu_result tmp;
    tmp.pinned = ctx->pinned;
    memset(&tmp.stats, 0, sizeof(tmp.stats));
    tmp.summary.resize(1); tmp.gene_off.assign(2, 0); tmp.node_off.assign(2, 0);
    OperatorOut op;
    RunPlan plan;
    plan.stage = 1;
    plan.forced_tt = translation_table;
    const int64_t offs[2] = {0, slen};
    int rc = run_range(ctx, seq, nullptr, offs, 0, 1, *opts, plan, &tmp, &op);
    if (rc) return rc;
    const int nn = (int)op.ndx.size();
    if (ndx || stop_val || strand || type || edge) {
        if (nn > cap) return fail(ctx, PGPU_EINVAL, "node capacity too small");
        for (int i = 0; i < nn; i++) {
            if (ndx) ndx[i] = op.ndx[i];
            if (stop_val) stop_val[i] = op.stop_val[i];
            if (strand) strand[i] = (op.cls[i] & CLS_REV) ? -1 : 1;
            if (type) type[i] = op.cls[i] & CLS_TYPE;
            if (edge) edge[i] = (op.cls[i] & CLS_EDGE) ? 1 : 0;
        }
    }
    return nn;
}

static int run_query(pgpu_ctx *ctx, const uint8_t *seq, int slen, RunPlan &plan) {
    cudaSetDevice(ctx->device);
    pgpu_result tmp;
    tmp.pinned = ctx->pinned;
    memset(&tmp.stats, 0, sizeof(tmp.stats));
    tmp.summary.resize(1); tmp.gene_off.assign(2, 0); tmp.node_off.assign(2, 0);
    pgpu_opts opts;
    memset(&opts, 0, sizeof(opts));
    opts.min_gene = 90; opts.min_edge_gene = 60; opts.max_overlap = 60; opts.min_mask = 50;
    plan.stage = 1;
    const int64_t offs[2] = {0, slen};
    return run_range(ctx, seq, nullptr, offs, 0, 1, opts, plan, &tmp, nullptr);
}

int pgpu_max_gc_frame_plot(pgpu_ctx *ctx, const uint8_t *seq, int slen, int8_t *out) {
    if (!ctx) return PGPU_EINVAL;
    if (slen < 0 || (slen > 0 && (!seq || !out))) return fail(ctx, PGPU_EINVAL, "bad arguments");
    if (slen == 0) return PGPU_OK;
    RunPlan plan;
    plan.gc_frame_out = out;
    return run_query(ctx, seq, slen, plan);
}

int pgpu_shine_dalgarno(pgpu_ctx *ctx, const uint8_t *seq, int slen, int pos, int start, int model, int strand, int exact,
                        int32_t *out) {
    if (!ctx) return PGPU_EINVAL;
    if (!out || slen < 0 || (slen > 0 && !seq)) return fail(ctx, PGPU_EINVAL, "bad arguments");
    if (strand != 1 && strand != -1) return fail(ctx, PGPU_EINVAL, "Invalid strand (must be +1 or -1)");
    if (pos < 0) return fail(ctx, PGPU_EINVAL, "`pos` must be positive");
    if (start < 0) return fail(ctx, PGPU_EINVAL, "`start` must be positive");
    if (model < 0 || model >= ctx->n_models) return fail(ctx, PGPU_ESTATE, "model index out of range");
    *out = 0;
    if (slen == 0) return PGPU_OK;
    RunPlan plan;
    plan.sd = {pos, start, model, strand, exact, out};
    return run_query(ctx, seq, slen, plan);
}

int pgpu_score_nodes(pgpu_ctx *ctx, const uint8_t *seq, int slen, int model, const pgpu_opts *opts, int is_meta,
                     int first_pass, int cap, pgpu_node *dst) {
    if (!ctx) return PGPU_EINVAL;
    if (!opts || slen < 0 || (!seq && slen > 0)) return fail(ctx, PGPU_EINVAL, "bad arguments");
    if (model < 0 || model >= ctx->n_models) return fail(ctx, PGPU_ESTATE, "model index out of range");
    cudaSetDevice(ctx->device);
    pgpu_result tmp;
    tmp.pinned = ctx->pinned;
    memset(&tmp.stats, 0, sizeof(tmp.stats));
    tmp.summary.resize(1); tmp.gene_off.assign(2, 0); tmp.node_off.assign(2, 0);
    OperatorOut op;
    RunPlan plan;
    plan.stage = 2;
    plan.forced_model = model;
    plan.forced_first_pass = first_pass;
    plan.forced_is_meta = is_meta;
    const int64_t offs[2] = {0, slen};
    int rc = run_range(ctx, seq, nullptr, offs, 0, 1, *opts, plan, &tmp, &op);
    if (rc) return rc;
    const int nn = (int)op.nodes.size();
    if (dst) {
        if (nn > cap) return fail(ctx, PGPU_EINVAL, "node capacity too small");
        if (nn) memcpy(dst, op.nodes.data(), nn * sizeof(pgpu_node));
    }
    return nn;
}

int pgpu_score_connections(pgpu_ctx *ctx, int n, const int32_t *ndx, const int32_t *stop_val, const int8_t *strand,
                           const uint8_t *type, const double *cscore, const double *sscore, const double *rscore,
                           const double *uscore, const double *gc_score, const int32_t *star_ptr, int model, int final,
                           double *out_score, int32_t *out_traceb, int8_t *out_ov_mark, int64_t *out_pairs,
                           double *out_ms) {
    if (!ctx) return PGPU_EINVAL;
    if (n < 0 || (n > 0 && (!ndx || !stop_val || !strand || !type || !cscore || !sscore || !rscore || !uscore ||
                            !star_ptr || !out_score || !out_traceb || !out_ov_mark)))
        return fail(ctx, PGPU_EINVAL, "bad arguments");
    if (!final && n > 0 && !gc_score) return fail(ctx, PGPU_EINVAL, "gc_score required when final == 0");
    if (model < 0 || model >= ctx->n_models) return fail(ctx, PGPU_ESTATE, "model index out of range");
    if (out_pairs) *out_pairs = 0;
    if (out_ms) *out_ms = 0.0;
    if (n == 0) return PGPU_OK;
    cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->stream;
    DevPool pool(ctx);
    DevBatch B;
    memset(&B, 0, sizeof(B));
    std::vector<uint8_t> cls(n);
    for (int i = 0; i < n; i++)
        cls[i] = (uint8_t)((type[i] & 3) | (strand[i] != 1 ? CLS_REV : 0) | ((((ndx[i] % 3) + 3) % 3) << CLS_FRAME_SHIFT));
    std::vector<ExtractInfo> exts(1);
    memset(&exts[0], 0, sizeof(ExtractInfo));
    exts[0].nn = n;
    std::vector<ChainInfo> chains(1);
    memset(&chains[0], 0, sizeof(ChainInfo));
    chains[0].model = model; chains[0].nn = n; chains[0].first_pass = 1; chains[0].istride = 1;
    const DevModel &M = ctx->h_models[model];
    std::vector<double> gcb(n, 0.0);
    if (!final)
        for (int i = 0; i < n; i++)
            gcb[i] = M.bias[0] * gc_score[3 * i] + M.bias[1] * gc_score[3 * i + 1] + M.bias[2] * gc_score[3 * i + 2];
    B.exts = pool.upload(exts);
    B.chains = pool.upload(chains);
    B.ndx = pool.alloc<int32_t>(n); B.stop_val = pool.alloc<int32_t>(n); B.cls = pool.upload(cls);
    B.win_min = pool.alloc<int32_t>(n); B.crank = pool.alloc<int32_t>(4 * (size_t)n + 4); B.clist = pool.alloc<int32_t>(n);
    B.cbase = pool.alloc<int32_t>(8);
    B.cscore = pool.alloc<double>(n); B.sscore = pool.alloc<double>(n); B.rscore = pool.alloc<double>(n);
    B.uscore = pool.alloc<double>(n); B.opv = pool.alloc<double>(3 * (size_t)n); B.gcb = pool.upload(gcb);
    B.star_ptr = pool.alloc<int32_t>(3 * (size_t)n);
    B.score = pool.alloc<double>(n); B.traceb = pool.alloc<int32_t>(n); B.ov_mark = pool.alloc<int8_t>(n + 16);
    B.chain_ipath = pool.alloc<int32_t>(1); B.chain_score = pool.alloc<double>(1);
    B.cndx = pool.alloc<int32_t>(n);
    B.dpx = pool.alloc<int4>(n);
    B.ig_node = pool.alloc<int32_t>(n + 1); B.ig_ndx = pool.alloc<int32_t>(n + 1); B.dqx = pool.alloc<int4>(n);
    if (ctx->dp_algo >= 1) {
        B.dp_svig = pool.alloc<double>(n); B.dp_tbig = pool.alloc<int32_t>(n);
    }
    unsigned long long *d_pairs = pool.alloc<unsigned long long>(1, true);
    if (pool.failed) return PGPU_ENOMEM;
    CK(cudaMemcpyAsync(B.ndx, ndx, n * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(B.stop_val, stop_val, n * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(B.cscore, cscore, n * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(B.sscore, sscore, n * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(B.rscore, rscore, n * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(B.uscore, uscore, n * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(B.star_ptr, star_ptr, 3 * (size_t)n * 4, cudaMemcpyHostToDevice, st));
    launch_node_prep(B, 1, n, 0, st);
    launch_dp_index(B, 1, n, st);
    launch_pairs(B, 1, n, d_pairs, st);
    launch_opv(B, ctx->d_models, 1, n, st);
    cudaEventRecord(ctx->ev[0], st);
    launch_dp(B, ctx->d_models, nullptr, 1, final ? 1 : 0, ctx->dp_algo, st);
    cudaEventRecord(ctx->ev[1], st);
    ctx->launches += 5;
    unsigned long long pairs = 0;
    CK(cudaMemcpyAsync(out_score, B.score, n * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(out_traceb, B.traceb, n * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(out_ov_mark, B.ov_mark, n, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&pairs, d_pairs, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (out_pairs) *out_pairs = (int64_t)pairs;
    if (out_ms) { float t = 0; cudaEventElapsedTime(&t, ctx->ev[0], ctx->ev[1]); *out_ms = t; }
    return PGPU_OK;
}

/* GeneFinder.train (lib.pyx:5471-5575 / _train 5236-5279) */
int pgpu_train(pgpu_ctx *ctx, const uint8_t *seq, int64_t slen, const pgpu_opts *opts, const pgpu_train_opts *topts,
               void *out_training, pgpu_stats *stats) {
    if (!ctx) return PGPU_EINVAL;
    if (!opts || !topts || !out_training || (!seq && slen > 0)) return fail(ctx, PGPU_EINVAL, "bad arguments");
    if (opts->meta) return fail(ctx, PGPU_ESTATE, "cannot use training sequence in metagenomic mode");
    {   // _TRANSLATION_TABLES, lib.pyx:172
        static const int valid[] = {1, 2, 3, 4, 5, 6, 9, 10, 11, 12, 13, 14, 15, 16, 21, 22, 23, 24, 25, 26, 29, 30, 32, 33};
        bool ok = false;
        for (int v : valid) ok |= v == topts->translation_table;
        if (!ok) return fail(ctx, PGPU_EINVAL, std::to_string(topts->translation_table) + " is not a valid translation table index");
    }
    if (opts->min_gene <= 0) return fail(ctx, PGPU_EINVAL, "`min_gene` must be strictly positive");
    if (opts->min_edge_gene <= 0) return fail(ctx, PGPU_EINVAL, "`min_edge_gene` must be strictly positive");
    if (opts->max_overlap < 0) return fail(ctx, PGPU_EINVAL, "`max_overlap` must be positive");
    if (slen < 20000)   // _MIN_SINGLE_GENOME, lib.pyx:170, 5547-5550
        return fail(ctx, PGPU_EINVAL, "sequence must be at least 20000 characters (" + std::to_string(slen) + " found)");
    if (slen > 0x7ffffff0) return fail(ctx, PGPU_EINVAL, "contig length out of range");
    cudaSetDevice(ctx->device);
    pgpu_result tmp;
    tmp.pinned = ctx->pinned;
    memset(&tmp.stats, 0, sizeof(tmp.stats));
    tmp.summary.resize(1); tmp.gene_off.assign(2, 0); tmp.node_off.assign(2, 0);
    std::vector<RawTraining> T(1);
    TrainRequest req;
    req.tt = topts->translation_table; req.force_nonsd = topts->force_nonsd; req.st_wt = topts->start_weight;
    req.out = &T[0];
    RunPlan plan;
    plan.stage = 1;
    plan.forced_tt = topts->translation_table;
    plan.train = &req;
    const int64_t offs[2] = {0, slen};
    const int64_t launches0 = ctx->launches;
    int rc = run_range(ctx, seq, nullptr, offs, 0, 1, *opts, plan, &tmp, nullptr);
    if (rc) return rc;
    memcpy(out_training, &T[0], sizeof(RawTraining));
    if (stats) { *stats = tmp.stats; stats->kernel_launches = ctx->launches - launches0; }
    return PGPU_OK;
}

int pgpu_compute_skippable(pgpu_ctx *ctx, int n, const int8_t *strand, const uint8_t *type, const int32_t *ndx, int mn,
                           int i, uint8_t *skip) {
    if (!ctx) return PGPU_EINVAL;
    if (n <= 0 || !strand || !type || !ndx || !skip || mn < 0 || i >= n || mn > i) return fail(ctx, PGPU_EINVAL, "bad arguments");
    cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->stream;
    DevPool pool(ctx);
    int8_t *d_s = pool.alloc<int8_t>(n);
    uint8_t *d_t = pool.alloc<uint8_t>(n), *d_k = pool.alloc<uint8_t>(n, true);
    int32_t *d_n = pool.alloc<int32_t>(n);
    if (pool.failed) return PGPU_ENOMEM;
    CK(cudaMemcpyAsync(d_s, strand, n, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_t, type, n, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_n, ndx, n * 4, cudaMemcpyHostToDevice, st));
    launch_skippable(n, d_s, d_t, d_n, mn, i, d_k, st);
    ctx->launches++;
    CK(cudaMemcpyAsync(skip + mn, d_k + mn, i - mn, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return PGPU_OK;
}

// The skip filter with the signature of the reference's plug-in ABI (skippable_t, lib.pxd:120; called per target node by
// BaseConnectionScorer._compute_skippable, lib.pyx:1321-1334): no context argument, so it runs on a process-wide context
// (device PGPU_DEVICE, default 0) created on first use.  A void function cannot report an error: on any failure it
// clears skip[min .. i), which is always a valid answer (nothing is skipped, the connection rules decide).
void pgpu_skippable(const uint8_t *strands, const uint8_t *types, const uint8_t *frames, const int mn, const int i,
                    uint8_t *skip) {
    static std::mutex mu;
    static pgpu_ctx *dctx = nullptr;
    if (!strands || !types || !frames || !skip || mn < 0 || i <= mn) return;
    std::lock_guard<std::mutex> lock(mu);
    const int cnt = i - mn;
    bool ok = false;
    if (!dctx) {
        const char *d = getenv("PGPU_DEVICE");
        if (pgpu_create(d ? atoi(d) : 0, &dctx) != PGPU_OK) dctx = nullptr;
    }
    if (dctx) {
        cudaSetDevice(dctx->device);
        cudaStream_t st = dctx->stream;
        DevPool pool(dctx);
        uint8_t *d_s = pool.alloc<uint8_t>(cnt + 1), *d_t = pool.alloc<uint8_t>(cnt + 1), *d_f = pool.alloc<uint8_t>(cnt + 1);
        uint8_t *d_k = pool.alloc<uint8_t>(cnt);
        if (!pool.failed) {
            ok = cudaMemcpyAsync(d_s, strands + mn, cnt + 1, cudaMemcpyHostToDevice, st) == cudaSuccess &&
                 cudaMemcpyAsync(d_t, types + mn, cnt + 1, cudaMemcpyHostToDevice, st) == cudaSuccess &&
                 cudaMemcpyAsync(d_f, frames + mn, cnt + 1, cudaMemcpyHostToDevice, st) == cudaSuccess;
            if (ok) {
                launch_skippable_plugin(d_s, d_t, d_f, cnt, d_k, st);
                dctx->launches++;
                ok = cudaMemcpyAsync(skip + mn, d_k, cnt, cudaMemcpyDeviceToHost, st) == cudaSuccess &&
                     cudaStreamSynchronize(st) == cudaSuccess;
            }
        }
    }
    if (!ok) memset(skip + mn, 0, cnt);
}

}  // extern "C"
