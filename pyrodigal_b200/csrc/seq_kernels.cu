// seq_kernels.cu -- sequence encoding and node extraction (add_nodes) on the device.
//
// Reference semantics: Sequence._build / _mask (src/pyrodigal/lib.pyx:664-713), Nodes._extract
// (lib.pyx:1905-2117) + Nodes._sort (lib.pyx:2489, node.c:1578-1587).
//
// B200 design: the sequence is streamed once (encode: 1 B in, 2 B out per base, 16-byte stores); node
// extraction works on per-(strand, frame) codon bitmaps in scan order, one thread per 32 codons
// (extract_device.cuh): a first pass marks node positions in two bitmaps, a prefix sum over bitmap words
// gives every node its final rank in (ndx, strand) order, and a second pass writes the nodes straight into
// their sorted slots -- no sort pass and no per-node atomics.
#include "extract_device.cuh"
#include "kernels.cuh"

namespace pgpu {

// --------------------------------------------------------------------------------------------------
// encode: ASCII -> digits (A0 G1 C2 T3 N6) + codon codes; per-contig G/C and unknown counts
// --------------------------------------------------------------------------------------------------
constexpr int kTile = 4096;

__device__ __forceinline__ uint8_t encode_base(uint8_t a) {
    switch (a) {
    case 'A': case 'a': return 0;
    case 'G': case 'g': return 1;
    case 'C': case 'c': return 2;
    case 'T': case 't': return 3;
    default: return 6;
    }
}

__global__ void __launch_bounds__(256) k_encode(DevBatch B, const int2 *__restrict__ tiles) {
    __shared__ __align__(16) uint8_t s[kTile + 16];
    const int2 tile = tiles[blockIdx.x];
    const ContigInfo ci = B.contigs[tile.x];
    const int start = tile.y;
    const int n = min(kTile, ci.slen - start);
    const uint8_t *src = B.ascii + ci.aoff + start;
    for (int k = threadIdx.x; k < kTile + 16; k += 256)
        s[k] = (k < n + 2 && start + k < ci.slen) ? encode_base(src[k]) : 0;
    __syncthreads();
    const int k0 = threadIdx.x * 16;
    int gc = 0, unk = 0;
    if (k0 < n) {
        uint32_t d[4] = {0, 0, 0, 0}, c[4] = {0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < 16; k++) {
            int p = k0 + k;
            if (p < n) {
                uint32_t b0 = s[p], b1 = s[p + 1], b2 = s[p + 2];
                gc += (b0 == 1) | (b0 == 2);
                unk += (b0 == 6);
                uint32_t code = (b0 & 3) | ((b1 & 3) << 2) | ((b2 & 3) << 4) | (((b0 | b1 | b2) & 4) << 4);
                d[k >> 2] |= b0 << ((k & 3) * 8);
                c[k >> 2] |= code << ((k & 3) * 8);
            }
        }
        const int64_t o = ci.doff + start + k0;
        *reinterpret_cast<uint4 *>(B.digits + o) = make_uint4(d[0], d[1], d[2], d[3]);
        *reinterpret_cast<uint4 *>(B.cod + o) = make_uint4(c[0], c[1], c[2], c[3]);
    }
    {
        // GC bitmap: 16 bits per thread, two neighbouring threads share a 32-bit word
        uint32_t bits = 0;
        if (k0 < n) {
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const int p = k0 + k;
                if (p < n) { const uint32_t b0 = s[p]; bits |= (uint32_t)(b0 != 0 && b0 != 3) << k; }
            }
        }
        const uint32_t other = __shfl_xor_sync(0xffffffffu, bits, 1);
        if ((threadIdx.x & 1) == 0 && k0 < ((n + 31) & ~31))
            B.gcbits[(ci.doff + start + k0) >> 5] = bits | (other << 16);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        gc += __shfl_down_sync(0xffffffffu, gc, off);
        unk += __shfl_down_sync(0xffffffffu, unk, off);
    }
    if ((threadIdx.x & 31) == 0) {
        if (gc) atomicAdd(&B.gc_count[tile.x], gc);
        if (unk) atomicAdd(&B.unknown[tile.x], unk);
    }
}

// Dicodon (6-mer) index of every position on both strands, so that the coding-score walk needs one 2-byte load per
// codon and no dependency on the previous codon (_sequence.h:207-220: first base in the low bits, N indexes as C).
//   forward  p: codons at p and p+3                       = cod[p] | cod[p+3] << 6
//   reverse  p: reverse codon with its 5' base at p (bases p, p-1, p-2) and the one before it in reverse reading
//               order (5' base at p-3)                     = rc(p) | rc(p-3) << 6
__device__ __forceinline__ int rev_codon_at(const uint8_t *__restrict__ d, const uint8_t *__restrict__ cod, int p) {
    const int c = cod[p - 2];
    if (!(c & 64)) return rev_code(c & 63);
    int r = 0;
    for (int k = 0; k < 3; k++) {
        const int b = d[p - k];
        r |= (b == 6 ? 2 : (b ^ 3)) << (2 * k);
    }
    return r;
}
// frame-planar layout (DevBatch::dic_f / dic_r): the tile covers the elements [1366 j, 1366 (j + 1)) of every plane, i.e. all
// positions of its 4096 bases; consecutive threads write consecutive elements of one plane
__global__ void __launch_bounds__(256) k_dicodon_index(DevBatch B, const int2 *__restrict__ tiles) {
    const int2 tile = tiles[blockIdx.x];
    const ContigInfo ci = B.contigs[tile.x];
    const uint8_t *__restrict__ d = B.digits + ci.doff;
    const uint8_t *__restrict__ cod = B.cod + ci.doff;
    const int P = dic_plane(ci.slen);
    uint16_t *__restrict__ df = B.dic_f + ci.doff;
    uint16_t *__restrict__ dr = B.dic_r + ci.doff;
    constexpr int kPer = (kTile + 2) / 3 + 1;   // 1366 elements per plane and tile
    const int k0 = (tile.y / kTile) * kPer;
    for (int t = threadIdx.x; t < 3 * kPer; t += 256) {
        const int f = t / kPer, k = k0 + t % kPer;
        const int p = 3 * k + f;
        if (p >= ci.slen) {   // slack of the plane: a valid index (k_coding_flat reads whole 16-byte chunks and looks every
            if (k < P) { df[f * P + k] = 0; dr[f * P + (P - 1 - k)] = 0; }   // element up, also the ones it does not add)
            continue;
        }
        df[f * P + k] = (uint16_t)((cod[p] & 63) | ((cod[p + 3] & 63) << 6));     // cod is zero padded past the end
        dr[f * P + (P - 1 - k)] = p >= 5 ? (uint16_t)(rev_codon_at(d, cod, p) | (rev_codon_at(d, cod, p - 3) << 6)) : (uint16_t)0;
    }
    if (tile.y + kTile >= ci.slen)   // last tile of the contig: the slack behind its elements
        for (int t = threadIdx.x; t < 3 * 32; t += 256) {
            const int f = t / 32, k = k0 + kPer + t % 32;
            if (k < P) { df[f * P + k] = 0; dr[f * P + (P - 1 - k)] = 0; }
        }
}

// runs of N: one thread per run start walks to the end of its run (lib.pyx:699-713)
__global__ void k_find_masks(DevBatch B, const int2 *__restrict__ tiles, int min_mask, int4 *out, int cap,
                             int *count) {
    const int2 tile = tiles[blockIdx.x];
    const ContigInfo ci = B.contigs[tile.x];
    const uint8_t *d = B.digits + ci.doff;
    for (int k = threadIdx.x; k < kTile; k += blockDim.x) {
        int p = tile.y + k;
        if (p >= ci.slen) break;
        if (d[p] != 6 || (p > 0 && d[p - 1] == 6)) continue;
        int e = p + 1;
        while (e < ci.slen && d[e] == 6) e++;
        if (e - p >= min_mask || e == ci.slen) {
            int slot = atomicAdd(count, 1);
            if (slot < cap) out[slot] = make_int4(tile.x, p, e, 0);
        }
    }
}

// --------------------------------------------------------------------------------------------------
// node extraction
// --------------------------------------------------------------------------------------------------

// The reference tests a candidate ORF against ONE mask only, the one under a per-frame cursor into the sorted mask
// list (lib.pyx:1959-1966 forward, 2053-2061 reverse).  The ORF end only moves one way during the scan, so the cursor
// is a function of the ORF end and needs no state here:
//   forward: the LAST mask with begin <= last      (then lib.pyx:337-340: begin < last && i < end)
//   reverse: the FIRST mask with end >= left end   (left = slen-last-1; then begin < right && left < end)
// A second N run further inside the ORF is not seen by the reference and must not be seen here.
__device__ __forceinline__ bool masked_fwd(const int32_t *__restrict__ m, int n, int i, int last) {
    int lo = 0, hi = n;   // first mask with begin > last
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (m[2 * mid] > last) hi = mid; else lo = mid + 1;
    }
    return lo > 0 && m[2 * (lo - 1)] < last && i < m[2 * (lo - 1) + 1];
}
__device__ __forceinline__ bool masked_rev(const int32_t *__restrict__ m, int n, int left, int right) {
    int lo = 0, hi = n;   // first mask with end >= left
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (m[2 * mid + 1] >= left) hi = mid; else lo = mid + 1;
    }
    return lo < n && m[2 * lo] < right && left < m[2 * lo + 1];
}

// Warp-cooperative, chunked scan (used for batches with N-run masks; the default path is the bit-parallel
// k_extract_b below): one warp per (extraction, chunk, strand, frame) looks at
// 32 codons of its frame per iteration.  The sequential state of the reference's loop (lib.pyx:1940-2010)
// depends only on the nearest stop codon seen earlier in scan order and on "has a start been emitted since that
// stop", both of which are bit operations on the ballot masks of the block (stops S, qualifying starts Q):
//   lane k: earlier stops Sb = S & lt(k);  nearest = highest bit of Sb  -> last / min_dist / last_real
//   stop lane k: saw = any Q bit strictly between the nearest earlier stop and k (or carried in)
// so 32 codons cost a few dozen instructions instead of 32 dependent iterations.
// Chunks: a frame is cut into runs of kExtractChunkCodons codons (scan order = descending strand coordinate).  The
// state entering a chunk is fully determined by the stretch between the chunk and the nearest stop codon before
// it in scan order, so a warp first looks back for that stop ("pre-roll", on average one or two ballots), replays
// the codons between the stop and its chunk without emitting, and then owns every node whose *triggering* codon
// lies inside the chunk (a start at its own codon; a STOP node at the next stop codon of the frame; the trailing
// STOP node belongs to the last chunk).  One long contig therefore spreads over thousands of warps.
template <bool FILL>
__global__ void __launch_bounds__(128) k_extract_w(DevBatch B, int n_ext, int total_chunks, RunOpts o) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= total_chunks * 6) return;
    const int cidx = w / 6, sf = w % 6, rev = sf / 3, f = sf % 3;
    int e = 0;
    {   // extraction that owns chunk cidx (chunk_off is increasing)
        int hi = n_ext - 1;
        while (e < hi) {
            const int mid = (e + hi + 1) >> 1;
            if (B.exts[mid].chunk_off <= cidx) e = mid; else hi = mid - 1;
        }
    }
    const ExtractInfo X = B.exts[e];
    const int chunk = cidx - X.chunk_off;
    const int slen = X.slen;
    if (slen < 3) return;
    const uint8_t *__restrict__ cod = B.cod + X.doff;
    uint32_t *bits = (rev ? B.bits_rev : B.bits_fwd) + X.woff;
    const uint32_t *__restrict__ bf = B.bits_fwd + X.woff;
    const uint32_t *__restrict__ br = B.bits_rev + X.woff;
    const int32_t *__restrict__ wb = B.wordbase + X.woff;
    const int32_t *__restrict__ masks = B.masks + 2 * (int64_t)X.mask_off;

    auto emit = [&](int pos, int type, int sv, int edge) {
        const int p = rev ? slen - 1 - pos : pos;
        if (!FILL) {
            atomicOr(&bits[p >> 5], 1u << (p & 31));
        } else {
            const int ww = p >> 5, b = p & 31;
            const uint32_t lt = (1u << b) - 1u;
            int slot = wb[ww] + __popc(bf[ww] & lt) + __popc(br[ww] & lt) + (rev ? (int)((bf[ww] >> b) & 1u) : 0);
            int conv = (!o.closed && type != 3 && !edge && (rev ? p >= slen - 3 : p <= 2)) ? CLS_CONV : 0;
            B.ndx[slot] = p;
            B.stop_val[slot] = rev ? slen - 1 - sv : sv;
            B.cls[slot] = (uint8_t)(type | (rev ? CLS_REV : 0) | (edge ? CLS_EDGE : 0) | conv | ((p % 3) << CLS_FRAME_SHIFT));
        }
    };
    // codon code at strand coordinate i (>= 0, <= slen - 3) and whether it holds an unknown base
    auto codon = [&](int i, bool &has_n) {
        int c = rev ? cod[slen - 3 - i] : cod[i];
        has_n = c & 64;
        c &= 63;
        return rev ? rev_code(c) : c;
    };

    int i_top0 = slen - 3;
    i_top0 -= ((i_top0 % 3) - f + 3) % 3;                       // first codon of the frame in scan order
    const bool bottom = chunk == X.n_chunks - 1;
    const int hi_i = i_top0 - 3 * kExtractChunkCodons * chunk;    // first codon this chunk owns
    const int lo_i = bottom ? 0 : hi_i - 3 * (kExtractChunkCodons - 1);

    // state at the top of the frame (warp uniform): lib.pyx:1933-1939
    int last = slen + ((f - slen % 3 + 3) % 3);
    if (!o.closed)
        while (last + 3 > slen) last -= 3;
    bool last_real = false, saw = false;
    int min_dist = o.min_edge_gene;
    int i_start = i_top0;
    if (chunk > 0) {
        // pre-roll: nearest stop codon before the chunk in scan order
        for (int base = hi_i + 3; base <= i_top0; base += 96) {
            const int i = base + 3 * lane;
            bool st = false;
            if (i >= 0 && i <= i_top0) {
                bool has_n;
                const int c = codon(i, has_n);
                st = !has_n && ((X.stopmask >> c) & 1);
            }
            const uint32_t S = __ballot_sync(0xffffffffu, st);
            if (S) {
                last = base + 3 * (__ffs(S) - 1);
                last_real = true;
                min_dist = o.min_gene;
                i_start = last - 3;
                break;
            }
        }
    }
    const uint32_t lt_mask = (1u << lane) - 1u;

    for (int i_top = i_start; i_top >= lo_i; i_top -= 96) {
        const int i = i_top - 3 * lane;
        const bool valid = i >= lo_i;
        const bool own = i <= hi_i;   // nodes triggered above the chunk belong to the previous chunk
        int c = 0;
        bool has_n = true;
        if (valid) c = codon(i, has_n);
        const bool is_stop = valid && !has_n && ((X.stopmask >> c) & 1);
        const bool is_startc = valid && !has_n && ((X.startmask >> c) & 1);
        const uint32_t S = __ballot_sync(0xffffffffu, is_stop);
        const uint32_t Sb = S & lt_mask;
        int my_last = last, my_min = min_dist;
        bool my_real = last_real;
        int h = -1;
        if (Sb) { h = 31 - __clz(Sb); my_last = i_top - 3 * h; my_min = o.min_gene; my_real = true; }
        bool qual = false, q_start = false;
        if (valid && !is_stop && my_last < slen) {
            bool hit = false;
            if (X.n_masks)
                hit = rev ? masked_rev(masks, X.n_masks, slen - my_last - 1, slen - i - 1) : masked_fwd(masks, X.n_masks, i, my_last);
            if (!hit) {
                q_start = is_startc && (my_last - i + 3 >= my_min);
                qual = q_start || (i <= 2 && !o.closed && my_last - i > o.min_edge_gene);
            }
        }
        const uint32_t Q = __ballot_sync(0xffffffffu, qual);
        if (is_stop) {
            const uint32_t seg = lt_mask & ~(h >= 0 ? ((2u << h) - 1u) : 0u);
            const bool saw_here = (Q & seg) != 0 || (h < 0 && saw);
            if (saw_here && own) emit(my_last, 3, i, !my_real);
        } else if (qual && own) {
            if (q_start) { const int b0 = c & 3; emit(i, b0 == 0 ? 0 : (b0 == 1 ? 1 : 2), my_last, 0); }
            else emit(i, 0, my_last, 1);
        }
        // carry the state out of the block
        if (S) {
            const int hl = 31 - __clz(S);
            last = i_top - 3 * hl;
            last_real = true;
            min_dist = o.min_gene;
            saw = (Q & ~((hl == 31) ? 0xffffffffu : ((2u << hl) - 1u))) != 0;
        } else {
            saw = saw || Q != 0;
        }
    }
    if (bottom && saw && lane == 0) emit(last, 3, f - 6, !last_real);
}

// --------------------------------------------------------------------------------------------------
// bit-parallel extraction (extract_device.cuh): k_codon_bits writes the stop / start-codon bitmaps of every
// (extraction, strand, frame) in scan order, k_extract_b<FILL> then handles one word of 32 codons per THREAD.
// Both use the chunk geometry of k_extract_w: one warp per (chunk of kExtractChunkCodons codons, strand, frame).
// --------------------------------------------------------------------------------------------------
constexpr int kChunkWords = kExtractChunkCodons / 32;

__device__ __forceinline__ int chunk_owner(const ExtractInfo *__restrict__ exts, int n_ext, int cidx) {
    int e = 0, hi = n_ext - 1;
    while (e < hi) {
        const int mid = (e + hi + 1) >> 1;
        if (exts[mid].chunk_off <= cidx) e = mid; else hi = mid - 1;
    }
    return e;
}

__global__ void __launch_bounds__(128) k_codon_bits(DevBatch B, int n_ext, int total_chunks) {
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (wid >= total_chunks * 6) return;
    const int cidx = wid / 6, sf = wid % 6, rev = sf / 3, f = sf % 3;
    const ExtractInfo X = B.exts[chunk_owner(B.exts, n_ext, cidx)];
    const int chunk = cidx - X.chunk_off, slen = X.slen;
    if (slen < 3) return;
    int i_top0, n_codons;
    extract_frame_geometry(slen, f, &i_top0, &n_codons);
    const uint8_t *__restrict__ cod = B.cod + X.doff;
    const int64_t base = X.cb_off + (int64_t)sf * X.n_chunks * kChunkWords + (int64_t)chunk * kChunkWords;
    // strand coordinate i = i_top0 - 3u sits at cod[i] (forward) or cod[slen - 3 - i] (reverse): base + u * step
    const uint8_t *__restrict__ cp = rev ? cod + (slen - 3 - i_top0) : cod + i_top0;
    const int step = rev ? 3 : -3;
    const uint64_t stopmask = X.stopmask, startmask = X.startmask;
    const uint8_t *__restrict__ lut = B.codon_lut ? B.codon_lut + 128 * X.lut : nullptr;
    const int lsh = rev ? 2 : 0;
#pragma unroll 1
    for (int r = 0; r < kChunkWords / 32; r++) {
        const int w0 = chunk * kChunkWords + r * 32;
        if (w0 * 32 >= n_codons) break;
        uint32_t myS = 0, myC = 0;
        if (lut) {  // table variant: one dependent byte load instead of the mask arithmetic (uniform branch)
#pragma unroll 1
            for (int k0 = 0; k0 < 32; k0 += 8) {
                int fl[8];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int u = (w0 + k0 + j) * 32 + lane;
                    fl[j] = cp[min(u, n_codons - 1) * step];
                }
#pragma unroll
                for (int j = 0; j < 8; j++) fl[j] = lut[fl[j] & 127] >> lsh;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const bool ok = (w0 + k0 + j) * 32 + lane < n_codons;
                    const uint32_t S = __ballot_sync(0xffffffffu, ok && (fl[j] & 1));
                    const uint32_t C = __ballot_sync(0xffffffffu, ok && (fl[j] & 2));
                    if (lane == k0 + j) { myS = S; myC = C; }
                }
            }
            B.cb_stop[base + r * 32 + lane] = myS;
            B.cb_start[base + r * 32 + lane] = myC;
            continue;
        }
#pragma unroll 1
        for (int k0 = 0; k0 < 32; k0 += 8) {
            // eight independent byte loads first (clamped: slots past the end repeat the last codon and are
            // masked out), then branch-free flags and the ballots
            int cs[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int u = (w0 + k0 + j) * 32 + lane;
                cs[j] = cp[min(u, n_codons - 1) * step];
            }
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int u = (w0 + k0 + j) * 32 + lane;
                int c = cs[j] & 63;
                if (rev) c = rev_code(c);
                const bool ok = u < n_codons && !(cs[j] & 64);
                const uint32_t S = __ballot_sync(0xffffffffu, ok && ((stopmask >> c) & 1ull));
                const uint32_t C = __ballot_sync(0xffffffffu, ok && ((startmask >> c) & 1ull));
                if (lane == k0 + j) { myS = S; myC = C; }
            }
        }
        B.cb_stop[base + r * 32 + lane] = myS;
        B.cb_start[base + r * 32 + lane] = myC;
    }
}

template <bool FILL>
struct BitEmit {
    const DevBatch &B;
    const ExtractInfo &X;
    const uint8_t *cod;
    RunOpts o;
    bool rev;
    __device__ __forceinline__ void put(int pos, int type, int sv, int edge) const {
        const int slen = X.slen;
        const int p = rev ? slen - 1 - pos : pos;
        if (!FILL) {
            atomicOr((rev ? B.bits_rev : B.bits_fwd) + X.woff + (p >> 5), 1u << (p & 31));
        } else {
            const uint32_t *__restrict__ bf = B.bits_fwd + X.woff;
            const uint32_t *__restrict__ br = B.bits_rev + X.woff;
            const int ww = p >> 5, b = p & 31;
            const uint32_t lt = (1u << b) - 1u;
            const int slot = (B.wordbase + X.woff)[ww] + __popc(bf[ww] & lt) + __popc(br[ww] & lt) + (rev ? (int)((bf[ww] >> b) & 1u) : 0);
            const int conv = (!o.closed && type != 3 && !edge && (rev ? p >= slen - 3 : p <= 2)) ? CLS_CONV : 0;
            B.ndx[slot] = p;
            B.stop_val[slot] = rev ? slen - 1 - sv : sv;
            B.cls[slot] = (uint8_t)(type | (rev ? CLS_REV : 0) | (edge ? CLS_EDGE : 0) | conv | ((p % 3) << CLS_FRAME_SHIFT));
        }
    }
    __device__ __forceinline__ void start(int i, int last, int edge) const {
        int type = 0;
        if (FILL && !edge) {
            const int c = rev ? rev_code(cod[X.slen - 3 - i] & 63) : (cod[i] & 63);
            const int b0 = c & 3;
            type = b0 == 0 ? 0 : (b0 == 1 ? 1 : 2);
        }
        put(i, type, last, edge);
    }
    __device__ __forceinline__ void stop(int last, int sv, int edge) const { put(last, 3, sv, edge); }
};

template <bool FILL>
__global__ void __launch_bounds__(128) k_extract_b(DevBatch B, int n_ext, int total_chunks, RunOpts o) {
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (wid >= total_chunks * 6) return;
    const int cidx = wid / 6, sf = wid % 6, rev = sf / 3, f = sf % 3;
    const ExtractInfo X = B.exts[chunk_owner(B.exts, n_ext, cidx)];
    const int chunk = cidx - X.chunk_off, slen = X.slen;
    if (slen < 3) return;
    ExtractFrame F;
    extract_frame_geometry(slen, f, &F.i_top0, &F.n_codons);
    F.n_words = (F.n_codons + 31) >> 5;
    F.f = f; F.closed = o.closed;
    F.d_real = extract_min_codons(o.min_gene); F.d_virt = extract_min_codons(o.min_edge_gene);
    F.min_edge_gene = o.min_edge_gene;
    const int64_t base = X.cb_off + (int64_t)sf * X.n_chunks * kChunkWords;
    F.S = B.cb_stop + base; F.C = B.cb_start + base;
    BitEmit<FILL> em{B, X, B.cod + X.doff, o, rev != 0};
#pragma unroll 1
    for (int r = 0; r < kChunkWords / 32; r++) {
        const int w = chunk * kChunkWords + r * 32 + lane;
        if (w < F.n_words) extract_word(F, w, em);
    }
}

// --------------------------------------------------------------------------------------------------
// exclusive prefix sum of per-word node counts (three-phase: block sums, top scan, apply)
// --------------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanBlock = kScanThreads * kScanItems;

__device__ __forceinline__ int block_exclusive_scan(int v, int *total) {
    __shared__ int warp_sums[kScanThreads / 32];
    __shared__ int block_total;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += y;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int w = lane < kScanThreads / 32 ? warp_sums[lane] : 0;
        int wi = w;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, wi, off);
            if (lane >= off) wi += y;
        }
        if (lane < kScanThreads / 32) warp_sums[lane] = wi - w;
        if (lane == kScanThreads / 32 - 1) block_total = wi;
    }
    __syncthreads();
    int res = incl - v + warp_sums[wid];
    *total = block_total;
    __syncthreads();
    return res;
}

__global__ void __launch_bounds__(kScanThreads) k_count_words(const uint32_t *__restrict__ bf,
                                                              const uint32_t *__restrict__ br, int64_t nwords,
                                                              int *__restrict__ block_sums) {
    const int64_t base = (int64_t)blockIdx.x * kScanBlock + threadIdx.x * kScanItems;
    int s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++)
        if (base + k < nwords) s += __popc(bf[base + k]) + __popc(br[base + k]);
    int total;
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_top(int *__restrict__ block_sums, int nblocks, int *total_out) {
    int carry = 0;
    for (int base = 0; base < nblocks; base += kScanThreads) {
        int idx = base + threadIdx.x;
        int v = idx < nblocks ? block_sums[idx] : 0;
        int total;
        int ex = block_exclusive_scan(v, &total);
        if (idx < nblocks) block_sums[idx] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) *total_out = carry;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_apply(const uint32_t *__restrict__ bf,
                                                             const uint32_t *__restrict__ br, int64_t nwords,
                                                             const int *__restrict__ block_sums,
                                                             int32_t *__restrict__ wordbase) {
    const int64_t base = (int64_t)blockIdx.x * kScanBlock + threadIdx.x * kScanItems;
    int c[kScanItems], s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        c[k] = (base + k < nwords) ? __popc(bf[base + k]) + __popc(br[base + k]) : 0;
        s += c[k];
    }
    int total;
    int ex = block_exclusive_scan(s, &total) + block_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        if (base + k <= nwords) wordbase[base + k] = ex;  // index nwords = sentinel (grand total)
        ex += c[k];
    }
}

// same three phases for a single bit array (GC bitmap)
__global__ void __launch_bounds__(kScanThreads) k_count_words1(const uint32_t *__restrict__ a, int64_t nwords,
                                                               int *__restrict__ block_sums) {
    const int64_t base = (int64_t)blockIdx.x * kScanBlock + threadIdx.x * kScanItems;
    int s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++)
        if (base + k < nwords) s += __popc(a[base + k]);
    int total;
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}
__global__ void __launch_bounds__(kScanThreads) k_scan_apply1(const uint32_t *__restrict__ a, int64_t nwords,
                                                              const int *__restrict__ block_sums,
                                                              int32_t *__restrict__ pre) {
    const int64_t base = (int64_t)blockIdx.x * kScanBlock + threadIdx.x * kScanItems;
    int c[kScanItems], s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        c[k] = (base + k < nwords) ? __popc(a[base + k]) : 0;
        s += c[k];
    }
    int total;
    int ex = block_exclusive_scan(s, &total) + block_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        if (base + k <= nwords) pre[base + k] = ex;
        ex += c[k];
    }
}

// --------------------------------------------------------------------------------------------------
// launch wrappers
// --------------------------------------------------------------------------------------------------
void launch_gc_scan(const DevBatch &B, int64_t nwords, int *block_sums, int *total_out, cudaStream_t st) {
    const int nb = scan_num_blocks(nwords);
    k_count_words1<<<nb, kScanThreads, 0, st>>>(B.gcbits, nwords, block_sums);
    k_scan_top<<<1, kScanThreads, 0, st>>>(block_sums, nb, total_out);
    k_scan_apply1<<<nb, kScanThreads, 0, st>>>(B.gcbits, nwords, block_sums, B.gcpre);
}
void launch_encode(const DevBatch &B, const int2 *tiles, int n_tiles, cudaStream_t st) {
    if (n_tiles > 0) k_encode<<<n_tiles, 256, 0, st>>>(B, tiles);
}
void launch_dicodon_index(const DevBatch &B, const int2 *tiles, int n_tiles, cudaStream_t st) {
    if (n_tiles > 0) k_dicodon_index<<<n_tiles, 256, 0, st>>>(B, tiles);
}
void launch_find_masks(const DevBatch &B, const int2 *tiles, int n_tiles, int min_mask, int4 *out, int cap,
                       int *count, cudaStream_t st) {
    if (n_tiles > 0) k_find_masks<<<n_tiles, 256, 0, st>>>(B, tiles, min_mask, out, cap, count);
}
void launch_extract_mark(const DevBatch &B, int n_ext, int total_chunks, RunOpts o, cudaStream_t st) {
    if (n_ext > 0 && total_chunks > 0)
        k_extract_w<false><<<(unsigned)(((int64_t)total_chunks * 6 * 32 + 127) / 128), 128, 0, st>>>(B, n_ext, total_chunks, o);
}
void launch_codon_bits(const DevBatch &B, int n_ext, int total_chunks, cudaStream_t st) {
    if (n_ext > 0 && total_chunks > 0)
        k_codon_bits<<<(unsigned)(((int64_t)total_chunks * 6 * 32 + 127) / 128), 128, 0, st>>>(B, n_ext, total_chunks);
}
void launch_extract_bits(const DevBatch &B, int n_ext, int total_chunks, RunOpts o, bool fill, cudaStream_t st) {
    if (n_ext == 0 || total_chunks == 0) return;
    const unsigned nb = (unsigned)(((int64_t)total_chunks * 6 * 32 + 127) / 128);
    if (fill) k_extract_b<true><<<nb, 128, 0, st>>>(B, n_ext, total_chunks, o);
    else k_extract_b<false><<<nb, 128, 0, st>>>(B, n_ext, total_chunks, o);
}
void launch_extract_fill(const DevBatch &B, int n_ext, int total_chunks, RunOpts o, cudaStream_t st) {
    if (n_ext > 0 && total_chunks > 0)
        k_extract_w<true><<<(unsigned)(((int64_t)total_chunks * 6 * 32 + 127) / 128), 128, 0, st>>>(B, n_ext, total_chunks, o);
}
int scan_num_blocks(int64_t nwords) { return (int)((nwords + 1 + kScanBlock - 1) / kScanBlock); }
void launch_word_scan(const DevBatch &B, int64_t nwords, int *block_sums, int *total_out, cudaStream_t st) {
    const int nb = scan_num_blocks(nwords);
    k_count_words<<<nb, kScanThreads, 0, st>>>(B.bits_fwd, B.bits_rev, nwords, block_sums);
    k_scan_top<<<1, kScanThreads, 0, st>>>(block_sums, nb, total_out);
    k_scan_apply<<<nb, kScanThreads, 0, st>>>(B.bits_fwd, B.bits_rev, nwords, block_sums, B.wordbase);
}

}  // namespace pgpu
