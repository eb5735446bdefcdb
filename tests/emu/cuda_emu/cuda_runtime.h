// cuda_runtime.h -- TEST INFRASTRUCTURE ONLY: a host stand-in for the CUDA runtime and the SIMT execution model.
//
// tests/emu/build_emu.py rewrites the kernel launches of pyrodigal_b200/csrc/*.cu (`k<<<grid, block, 0, st>>>(args)`)
// into calls of emu::Launch and compiles the UNMODIFIED product sources with g++ against this header into
// tests/emu/libpgpu_emu.so, so that the `-m "not gpu"` tests can run the real kernels (slowly, on small inputs) against
// the oracle in a container without a GPU.  Nothing here is reachable from the product package: pyrodigal_b200/_capi.py
// only ever loads libpyrodigal_b200.so, which needs a CUDA device.
//
// Execution model: the blocks of a launch run one after the other; the threads of a block are fibers (ucontext) that
// run to completion in order and only switch at synchronisation points (__syncthreads, __syncwarp, *_sync warp
// intrinsics), where the live lanes of a warp (or threads of the block) exchange values through a double-buffered
// slot array.  "Device memory" is host memory; streams are ignored (every call is synchronous); events are clocks.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <functional>
#include <type_traits>

#define PGPU_HOST_EMULATION 1
#define __global__
#define __device__
#define __host__
#define __shared__ static
#define __constant__ static
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __noinline__ __attribute__((noinline))

// ---------------------------------------------------------------------------------------------- vector types
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
struct uint3 { unsigned x, y, z; };
struct dim3 { unsigned x = 1, y = 1, z = 1; dim3() {} dim3(unsigned a) : x(a) {} };
static inline int2 make_int2(int x, int y) { return {x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return {x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return {x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return {x, y, z, w}; }

// CUDA's min / max accept mixed integer types
template <class A, class B> constexpr typename std::common_type<A, B>::type min(A a, B b) { return b < a ? b : a; }
template <class A, class B> constexpr typename std::common_type<A, B>::type max(A a, B b) { return a < b ? b : a; }

// ---------------------------------------------------------------------------------------------- SIMT engine
namespace emu {
struct Thread;
extern thread_local Thread *cur;             // the running device thread (fiber)
uint3 thread_idx();
uint3 block_idx();
dim3 block_dim();
dim3 grid_dim();
void sync_block();                           // __syncthreads
// the live lanes named by `mask` publish a value each and wait for one another (lanes that use different masks --
// diverged groups of a warp -- rendezvous independently)
uint64_t warp_exchange(unsigned mask, uint64_t mine, uint64_t out[32], uint32_t *live_mask);
int lane_id();
struct Launch {
    unsigned grid, block;
    Launch(long long g, long long b, size_t = 0, void * = nullptr) : grid((unsigned)g), block((unsigned)b) {}
    void operator<<(const std::function<void()> &body) const;
};
}  // namespace emu

#define threadIdx (emu::thread_idx())
#define blockIdx (emu::block_idx())
#define blockDim (emu::block_dim())
#define gridDim (emu::grid_dim())
static constexpr int warpSize = 32;

static inline void __syncthreads() { emu::sync_block(); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { uint64_t v[32]; uint32_t m; emu::warp_exchange(mask, 0, v, &m); }
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    uint64_t v[32]; uint32_t live;
    emu::warp_exchange(mask, pred ? 1 : 0, v, &live);
    unsigned r = 0;
    for (int l = 0; l < 32; l++) if (((live >> l) & 1u) && v[l]) r |= 1u << l;
    return r;
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __all_sync(unsigned m, int pred) {
    uint64_t v[32]; uint32_t live;
    emu::warp_exchange(m, pred ? 1 : 0, v, &live);
    for (int l = 0; l < 32; l++) if (((live >> l) & 1u) && !v[l]) return 0;
    return 1;
}
template <class T> static inline uint64_t emu_bits(T x) { uint64_t b = 0; static_assert(sizeof(T) <= 8, ""); memcpy(&b, &x, sizeof(T)); return b; }
template <class T> static inline T emu_unbits(uint64_t b) { T x; memcpy(&x, &b, sizeof(T)); return x; }
template <class T> static inline T __shfl_sync(unsigned mask, T var, int src, int width = 32) {
    uint64_t v[32]; uint32_t live;
    emu::warp_exchange(mask, emu_bits(var), v, &live);
    const int lane = emu::lane_id(), base = lane / width * width;
    const int s = base + (src % width + width) % width;
    return ((live >> s) & 1u) ? emu_unbits<T>(v[s]) : var;
}
template <class T> static inline T __shfl_xor_sync(unsigned mask, T var, int lm, int width = 32) {
    uint64_t v[32]; uint32_t live;
    emu::warp_exchange(mask, emu_bits(var), v, &live);
    const int lane = emu::lane_id(), s = lane ^ lm;
    return (s / width == lane / width && ((live >> s) & 1u)) ? emu_unbits<T>(v[s]) : var;
}
template <class T> static inline T __shfl_down_sync(unsigned mask, T var, unsigned d, int width = 32) {
    uint64_t v[32]; uint32_t live;
    emu::warp_exchange(mask, emu_bits(var), v, &live);
    const int lane = emu::lane_id(), s = lane + (int)d;
    return (s / width == lane / width && s < 32 && ((live >> s) & 1u)) ? emu_unbits<T>(v[s]) : var;
}
template <class T> static inline T __shfl_up_sync(unsigned mask, T var, unsigned d, int width = 32) {
    uint64_t v[32]; uint32_t live;
    emu::warp_exchange(mask, emu_bits(var), v, &live);
    const int lane = emu::lane_id(), s = lane - (int)d;
    return (s >= 0 && s / width == lane / width && ((live >> s) & 1u)) ? emu_unbits<T>(v[s]) : var;
}
static inline int __reduce_max_sync(unsigned mask, int x) {
    uint64_t v[32]; uint32_t live;
    emu::warp_exchange(mask, emu_bits(x), v, &live);
    int r = x;
    for (int l = 0; l < 32; l++) if ((live >> l) & 1u) r = std::max(r, emu_unbits<int>(v[l]));
    return r;
}
static inline int __reduce_min_sync(unsigned mask, int x) {
    uint64_t v[32]; uint32_t live;
    emu::warp_exchange(mask, emu_bits(x), v, &live);
    int r = x;
    for (int l = 0; l < 32; l++) if ((live >> l) & 1u) r = std::min(r, emu_unbits<int>(v[l]));
    return r;
}
static inline unsigned __reduce_max_sync(unsigned mask, unsigned x) {
    uint64_t v[32]; uint32_t live;
    emu::warp_exchange(mask, emu_bits(x), v, &live);
    unsigned r = x;
    for (int l = 0; l < 32; l++) if ((live >> l) & 1u) r = std::max(r, emu_unbits<unsigned>(v[l]));
    return r;
}
static inline unsigned __reduce_min_sync(unsigned mask, unsigned x) {
    uint64_t v[32]; uint32_t live;
    emu::warp_exchange(mask, emu_bits(x), v, &live);
    unsigned r = x;
    for (int l = 0; l < 32; l++) if ((live >> l) & 1u) r = std::min(r, emu_unbits<unsigned>(v[l]));
    return r;
}

// ---------------------------------------------------------------------------------------------- scalar intrinsics
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline long long __double_as_longlong(double x) { long long r; memcpy(&r, &x, 8); return r; }
static inline double __longlong_as_double(long long x) { double r; memcpy(&r, &x, 8); return r; }
static inline int __float_as_int(float x) { int r; memcpy(&r, &x, 4); return r; }
static inline float __int_as_float(int x) { float r; memcpy(&r, &x, 4); return r; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline unsigned __brev(unsigned x) {
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i);
    return r;
}
// blocks run one after the other and fibers only switch at sync points: plain read-modify-write is atomic
template <class T, class U> static inline T atomicAdd(T *p, U v) { T o = *p; *p = (T)(o + v); return o; }
template <class T, class U> static inline T atomicOr(T *p, U v) { T o = *p; *p = (T)(o | v); return o; }
template <class T, class U> static inline T atomicMin(T *p, U v) { T o = *p; if ((T)v < o) *p = (T)v; return o; }
template <class T, class U> static inline T atomicMax(T *p, U v) { T o = *p; if ((T)v > o) *p = (T)v; return o; }

// ---------------------------------------------------------------------------------------------- runtime API
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
typedef void *cudaStream_t;
struct emu_event { std::chrono::steady_clock::time_point t; };
typedef emu_event *cudaEvent_t;
typedef void *cudaMemPool_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyHostToHost };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0 };
enum cudaMemPoolAttr { cudaMemPoolAttrReleaseThreshold, cudaMemPoolAttrReservedMemCurrent, cudaMemPoolAttrUsedMemCurrent };

static inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = (void *)1; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (void *)1; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new emu_event(); return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = new emu_event(); return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
    return cudaSuccess;
}
// PGPU_EMU_POISON=1 fills every fresh "device" allocation with a byte pattern: results that still match the oracle do
// not depend on memory the kernels never wrote (device memory from the stream-ordered pool is not zeroed either)
static inline bool emu_poison() { static const bool on = getenv("PGPU_EMU_POISON") && getenv("PGPU_EMU_POISON")[0] == '1'; return on; }
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t n) {
    *p = (T *)malloc(n ? n : 1);
    if (*p && emu_poison()) memset((void *)*p, 0xA5, n ? n : 1);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
template <class T> static inline cudaError_t cudaMallocAsync(T **p, size_t n, cudaStream_t) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaFreeAsync(void *p, cudaStream_t) { free(p); return cudaSuccess; }
template <class T> static inline cudaError_t cudaHostAlloc(T **p, size_t n, unsigned) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { if (n) memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { if (n) memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void *d, int v, size_t n) { if (n) memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = nullptr) { if (n) memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemGetInfo(size_t *f, size_t *t) { *f = size_t(8) << 30; *t = size_t(16) << 30; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetDefaultMemPool(cudaMemPool_t *p, int) { *p = nullptr; return cudaSuccess; }
static inline cudaError_t cudaMemPoolGetAttribute(cudaMemPool_t, cudaMemPoolAttr, void *v) { *(uint64_t *)v = 0; return cudaSuccess; }
static inline cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t, cudaMemPoolAttr, void *) { return cudaSuccess; }
