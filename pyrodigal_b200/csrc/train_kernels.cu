// train_kernels.cu -- device half of GeneFinder.train (lib.pyx:5236-5279): GC frame plot, GC frame bias,
// training-path dicodon statistics and the counting passes of the SD / non-SD start training.  The per-item
// logic lives in train_device.cuh (shared with the CPU emulation used by the tests); this file maps items to
// threads.  All kernels are small integer / byte kernels bound by latency or HBM traffic, no tensor cores.
#include <algorithm>

#include "kernels.cuh"

namespace pgpu {

using namespace train;

// STOP nodes of the (single) extraction through the class-sorted list, so that warps are either busy or exit
__device__ __forceinline__ int stop_item(const DevBatch &B, const TrainView &V, int t) {
    const int n_fe = V.cb[2] - V.cb[1], n_re = V.N.nn - V.cb[3];
    if (t >= n_fe + n_re) return -1;
    return (B.clist + V.node_off)[t < n_fe ? V.cb[1] + t : V.cb[3] + (t - n_fe)];
}

__global__ void __launch_bounds__(256) k_gc_frame(const uint32_t *__restrict__ gcbits, int slen, int ntrip, int8_t *gp) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < ntrip) gc_frame_triplet(gcbits, slen, t, gp);
}

__global__ void __launch_bounds__(128) k_gc_bias(DevBatch B, TrainView V) {
    const int z = stop_item(B, V, blockIdx.x * blockDim.x + threadIdx.x);
    if (z >= 0) gc_bias_orf(z, V.N, V.gp, V.gc_score, V.gc_bias, V.term);
}

// bias[] is a floating point sum over the start nodes in index order (node.c:306-311), one accumulator per GC
// frame: the warp stages 256 addends at a time in shared memory (coalesced), then lanes 0..2 each run through
// them and add the ones of "their" frame in order -- three independent dependent-add chains with the reference's
// rounding, reading from shared memory instead of shuffling every addend through the warp.
__global__ void __launch_bounds__(32) k_bias_sum(TrainView V, double *bias_out) {
    constexpr int kChunk = 256;
    __shared__ double s_t[kChunk];
    __shared__ int8_t s_g[kChunk];
    __shared__ double s_b[3];
    const int lane = threadIdx.x;
    double acc = 0.0;
    for (int base = 0; base < V.N.nn; base += kChunk) {
        const int n = min(kChunk, V.N.nn - base);
        for (int k = lane; k < n; k += 32) {
            const int i = base + k;
            const bool start = !cls_is_stop(V.N.cls[i]);
            s_t[k] = start ? V.term[i] : 0.0;
            s_g[k] = start ? V.gc_bias[i] : (int8_t)-1;
        }
        __syncwarp();
        if (lane < 3)
            for (int k = 0; k < n; k++)
                if (s_g[k] == lane) acc += s_t[k];
        __syncwarp();
    }
    if (lane < 3) s_b[lane] = acc;
    __syncwarp();
    if (lane == 0) {
        const double tot = s_b[0] + s_b[1] + s_b[2];
        bias_out[0] = s_b[0] * (3.0 / tot);
        bias_out[1] = s_b[1] * (3.0 / tot);
        bias_out[2] = s_b[2] * (3.0 / tot);
    }
}

// bias . gc_score of every node, the scoring function of the training DP (_connection.h, final == 0)
__global__ void __launch_bounds__(256) k_gcb(TrainView V, const double *__restrict__ bias, double *__restrict__ gcb) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V.N.nn) return;
    double v = 0.0;
    if (!cls_is_stop(V.N.cls[i])) {
        const double *g = V.gc_score + 3 * (int64_t)i;
        v = bias[0] * g[0] + bias[1] * g[1] + bias[2] * g[2];
    }
    gcb[i] = v;
}

__global__ void k_training_path(DevBatch B, TrainView V, int4 *intervals, int cap, int *n_out) {
    if (blockIdx.x == 0 && threadIdx.x == 0)
        *n_out = training_path(B.chain_ipath[0], V.N, B.traceb, B.ov_mark, B.star_ptr, intervals, cap);
}

// background 6-mers of both strands: block-local histogram in shared memory, flushed once
__global__ void __launch_bounds__(256) k_dicodon_bg(const uint8_t *__restrict__ d, int slen, uint32_t *__restrict__ counts) {
    __shared__ uint32_t h[4096];
    for (int k = threadIdx.x; k < 4096; k += blockDim.x) h[k] = 0;
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < slen - 5; i += gridDim.x * blockDim.x) {
        atomicAdd(&h[mer6(d, slen, i, false)], 1u);
        atomicAdd(&h[mer6(d, slen, i, true)], 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < 4096; k += blockDim.x)
        if (h[k]) atomicAdd(&counts[k], h[k]);
}

// 6-mers of the training genes: one warp per gene, lanes stride over its codons
__global__ void __launch_bounds__(128) k_dicodon_genes(const uint8_t *__restrict__ d, int slen, const int4 *__restrict__ intervals,
                                                        const int *__restrict__ n_intervals, uint32_t *__restrict__ counts,
                                                        unsigned long long *total) {
    const int lane = threadIdx.x & 31;
    const int n = *n_intervals;
    unsigned long long mine = 0;
    for (int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < n; g += (gridDim.x * blockDim.x) >> 5) {
        const int4 iv = intervals[g];
        for (int i = iv.x + 3 * lane; i < iv.y - 5; i += 96) {
            atomicAdd(&counts[mer6(d, slen, i, iv.z < 0)], 1u);
            mine++;
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, off);
    if (lane == 0 && mine) atomicAdd(total, mine);
}

__global__ void __launch_bounds__(256) k_type_background(TrainView V, uint32_t *cnt) {
    __shared__ uint32_t h[3];
    if (threadIdx.x < 3) h[threadIdx.x] = 0;
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < V.N.nn && !cls_is_stop(V.N.cls[i])) atomicAdd(&h[V.N.cls[i] & CLS_TYPE], 1u);
    __syncthreads();
    if (threadIdx.x < 3 && h[threadIdx.x]) atomicAdd(&cnt[C_TBG + threadIdx.x], h[threadIdx.x]);
}

// the small counters are kept per block in shared memory and flushed once
__global__ void __launch_bounds__(128) k_sd_iteration(DevBatch B, TrainView V, SdParams P, uint32_t *cnt) {
    __shared__ uint32_t h[C_TOTAL];
    for (int k = threadIdx.x; k < C_TOTAL; k += blockDim.x) h[k] = 0;
    __syncthreads();
    const int z = stop_item(B, V, blockIdx.x * blockDim.x + threadIdx.x);
    if (z >= 0) sd_orf(z, V.N, V.cscore, V.rbs, V.upc, P, h);
    __syncthreads();
    for (int k = threadIdx.x; k < C_TOTAL; k += blockDim.x)
        if (h[k]) atomicAdd(&cnt[k], h[k]);
}

__global__ void __launch_bounds__(128) k_motif_background(TrainView V, const double *__restrict__ mot_wt, MotParams P,
                                                           uint32_t *bg_cells, uint32_t *cnt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < V.N.nn) motif_background(i, V.N, V.umot, mot_wt, P, V.mot, bg_cells, cnt);
}

__global__ void __launch_bounds__(128) k_motif_orfs(DevBatch B, TrainView V, MotParams P, uint32_t *real_cells, uint32_t *cnt) {
    const int z = stop_item(B, V, blockIdx.x * blockDim.x + threadIdx.x);
    if (z >= 0) motif_orf(z, V.N, V.cscore, V.umot, V.upc, V.mot, P, real_cells, cnt);
}

// ---- launch wrappers -------------------------------------------------------------------------------------
static inline unsigned blocks(int64_t n, int per) { return (unsigned)((n + per - 1) / per); }

void launch_gc_frame(const uint32_t *gcbits, int slen, int8_t *gp, cudaStream_t st) {
    if (slen <= 0) return;
    cudaMemsetAsync(gp, 0xff, slen, st);  // -1: positions without a full triplet
    const int ntrip = slen / 3;
    if (ntrip > 0) k_gc_frame<<<blocks(ntrip, 256), 256, 0, st>>>(gcbits, slen, ntrip, gp);
}
void launch_gc_bias(const DevBatch &B, const TrainView &V, double *bias_out, double *gcb, cudaStream_t st) {
    if (V.N.nn == 0) return;
    k_gc_bias<<<blocks(V.N.nn, 128), 128, 0, st>>>(B, V);
    k_bias_sum<<<1, 32, 0, st>>>(V, bias_out);
    k_gcb<<<blocks(V.N.nn, 256), 256, 0, st>>>(V, bias_out, gcb);
}
void launch_training_path(const DevBatch &B, const TrainView &V, int4 *intervals, int cap, int *n_out, cudaStream_t st) {
    k_training_path<<<1, 32, 0, st>>>(B, V, intervals, cap, n_out);
}
void launch_dicodon(const uint8_t *digits, int slen, const int4 *intervals, const int *n_intervals, int cap,
                    uint32_t *bg_counts, uint32_t *gene_counts, unsigned long long *gene_total, cudaStream_t st) {
    if (slen > 5) k_dicodon_bg<<<std::min(blocks(slen, 256 * 16), 148u * 8u), 256, 0, st>>>(digits, slen, bg_counts);
    if (cap > 0) k_dicodon_genes<<<std::min(blocks((int64_t)cap * 32, 128), 148u * 16u), 128, 0, st>>>(
        digits, slen, intervals, n_intervals, gene_counts, gene_total);
}
void launch_type_background(const TrainView &V, uint32_t *cnt, cudaStream_t st) {
    if (V.N.nn > 0) k_type_background<<<blocks(V.N.nn, 256), 256, 0, st>>>(V, cnt);
}
void launch_sd_iteration(const DevBatch &B, const TrainView &V, const SdParams &P, uint32_t *cnt, cudaStream_t st) {
    if (V.N.nn > 0) k_sd_iteration<<<blocks(V.N.nn, 128), 128, 0, st>>>(B, V, P, cnt);
}
void launch_motif_iteration(const DevBatch &B, const TrainView &V, const double *mot_wt, const MotParams &P,
                            uint32_t *bg_cells, uint32_t *real_cells, uint32_t *cnt, cudaStream_t st) {
    if (V.N.nn == 0) return;
    k_motif_background<<<blocks(V.N.nn, 128), 128, 0, st>>>(V, mot_wt, P, bg_cells, cnt);
    k_motif_orfs<<<blocks(V.N.nn, 128), 128, 0, st>>>(B, V, P, real_cells, cnt);
}

}  // namespace pgpu
