/*
 * cuda_emu.h -- TEST-ONLY single-process emulation of the small CUDA subset the
 * flake_b200 kernels use, so that kernel *logic* (scans, bit packing, search
 * control flow, barrier placement) can be exercised by `pytest -m "not gpu"` in a
 * container without a GPU.
 *
 * This is NOT a fallback of the product: flake_b200/lib/libflake.so is always
 * built by nvcc for sm_100a and never contains or loads this.  The emulated
 * build (tests/cuda_emu/libflake_emu.so) compiles the very same kernel sources
 * with g++ and -DFLAKE_B200_CUDA_EMU and is only opened by tests/.
 *
 * Model: CTAs run one after another; the threads of a CTA are ucontext fibers
 * scheduled round-robin on the calling OS thread.  __syncthreads and the
 * *_sync warp collectives are rendezvous points; a rendezvous that can never
 * complete (divergent barrier, wrong shuffle mask) is reported as a deadlock
 * and aborts, which makes the emulator stricter than the hardware.
 */
#ifndef FLAKE_B200_CUDA_EMU_H
#define FLAKE_B200_CUDA_EMU_H

#include <ucontext.h>
#include <stdint.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <fenv.h>
#include <functional>
#include <vector>
#include <algorithm>

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

struct __attribute__((aligned(16))) int4 { int x, y, z, w; };
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };
struct __attribute__((aligned(8))) int2 { int x, y; };
struct __attribute__((aligned(8))) uint2 { unsigned x, y; };
struct __attribute__((aligned(16))) ulonglong2 { unsigned long long x, y; };
inline int4 make_int4(int x, int y, int z, int w) { int4 v = {x, y, z, w}; return v; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 v = {x, y, z, w}; return v; }
inline int2 make_int2(int x, int y) { int2 v = {x, y}; return v; }
inline uint2 make_uint2(unsigned x, unsigned y) { uint2 v = {x, y}; return v; }

namespace cuemu {

struct Fiber {
    ucontext_t ctx;
    char *stack = nullptr;
    bool done = false;
    unsigned tid = 0;
};

struct WarpState {
    unsigned arrived = 0;        /* lanes arrived at the current collective */
    unsigned gen = 0;
    uint64_t slot[32];
    unsigned pred_bits = 0;
};

struct CtaState {
    unsigned nthreads = 0, live = 0;
    unsigned bar_arrived = 0, bar_gen = 0;
    std::vector<Fiber> fibers;
    std::vector<WarpState> warps;
    ucontext_t sched;
    Fiber *cur = nullptr;
    bool progress = false;
    unsigned char *dyn_smem = nullptr;
    std::function<void()> body;
};

extern CtaState g_cta;
void yield();
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body);
void cta_barrier();
void warp_barrier(unsigned mask);
[[noreturn]] void die(const char *msg);

inline unsigned lane_id();
inline WarpState &my_warp();

} /* namespace cuemu */

extern dim3 threadIdx, blockIdx, blockDim, gridDim;
static const int warpSize = 32;

inline unsigned cuemu::lane_id() { return threadIdx.x & 31u; }
inline cuemu::WarpState &cuemu::my_warp() { return g_cta.warps[threadIdx.x >> 5]; }

/* ---- qualifiers ----------------------------------------------------- */
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __shared__ static
#define __constant__ static
#define __align__(n) __attribute__((aligned(n)))
#ifndef __restrict__
#define __restrict__ __restrict
#endif

/* ---- synchronisation -------------------------------------------------- */
inline void __syncthreads() { cuemu::cta_barrier(); }
inline void __syncwarp(unsigned mask = 0xffffffffu) { cuemu::warp_barrier(mask); }
inline void __threadfence() {}
inline void __threadfence_block() {}

template <class T> inline uint64_t cuemu_to_bits(T v)
{
    static_assert(sizeof(T) <= 8, "shuffle payload too large");
    uint64_t b = 0; memcpy(&b, &v, sizeof(T)); return b;
}
template <class T> inline T cuemu_from_bits(uint64_t b)
{
    T v; memcpy(&v, &b, sizeof(T)); return v;
}

template <class T> inline T cuemu_exchange(unsigned mask, T v, int src_lane)
{
    cuemu::WarpState &w = cuemu::my_warp();
    w.slot[cuemu::lane_id()] = cuemu_to_bits(v);
    cuemu::warp_barrier(mask);
    T r = v;
    if (src_lane >= 0 && src_lane < 32 && (mask >> src_lane) & 1u)
        r = cuemu_from_bits<T>(w.slot[src_lane]);
    cuemu::warp_barrier(mask);
    return r;
}

template <class T> inline T __shfl_sync(unsigned mask, T v, int src, int width = 32)
{
    int lane = (int)cuemu::lane_id();
    int base = lane & ~(width - 1);
    return cuemu_exchange(mask, v, base + (src & (width - 1)));
}
template <class T> inline T __shfl_down_sync(unsigned mask, T v, unsigned d, int width = 32)
{
    int lane = (int)cuemu::lane_id();
    int src = lane + (int)d;
    if ((src & ~(width - 1)) != (lane & ~(width - 1))) src = lane;
    return cuemu_exchange(mask, v, src);
}
template <class T> inline T __shfl_up_sync(unsigned mask, T v, unsigned d, int width = 32)
{
    int lane = (int)cuemu::lane_id();
    int src = lane - (int)d;
    if (src < (lane & ~(width - 1))) src = lane;
    return cuemu_exchange(mask, v, src);
}
template <class T> inline T __shfl_xor_sync(unsigned mask, T v, int x, int width = 32)
{
    (void)width;
    return cuemu_exchange(mask, v, (int)cuemu::lane_id() ^ x);
}

inline unsigned __ballot_sync(unsigned mask, int pred)
{
    cuemu::WarpState &w = cuemu::my_warp();
    w.slot[cuemu::lane_id()] = pred ? 1u : 0u;
    cuemu::warp_barrier(mask);
    unsigned r = 0;
    for (int l = 0; l < 32; l++) if (((mask >> l) & 1u) && w.slot[l]) r |= 1u << l;
    cuemu::warp_barrier(mask);
    return r;
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline int __all_sync(unsigned mask, int pred) { return (__ballot_sync(mask, pred) & mask) == mask; }
inline unsigned __activemask() { return 0xffffffffu; }

inline unsigned __reduce_add_sync(unsigned mask, unsigned v)
{
    cuemu::WarpState &w = cuemu::my_warp();
    w.slot[cuemu::lane_id()] = v;
    cuemu::warp_barrier(mask);
    unsigned r = 0;
    for (int l = 0; l < 32; l++) if ((mask >> l) & 1u) r += (unsigned)w.slot[l];
    cuemu::warp_barrier(mask);
    return r;
}
inline unsigned __reduce_or_sync(unsigned mask, unsigned v)
{
    cuemu::WarpState &w = cuemu::my_warp();
    w.slot[cuemu::lane_id()] = v;
    cuemu::warp_barrier(mask);
    unsigned r = 0;
    for (int l = 0; l < 32; l++) if ((mask >> l) & 1u) r |= (unsigned)w.slot[l];
    cuemu::warp_barrier(mask);
    return r;
}
inline unsigned __reduce_and_sync(unsigned mask, unsigned v)
{
    cuemu::WarpState &w = cuemu::my_warp();
    w.slot[cuemu::lane_id()] = v;
    cuemu::warp_barrier(mask);
    unsigned r = 0xffffffffu;
    for (int l = 0; l < 32; l++) if ((mask >> l) & 1u) r &= (unsigned)w.slot[l];
    cuemu::warp_barrier(mask);
    return r;
}
inline unsigned __reduce_max_sync(unsigned mask, unsigned v)
{
    cuemu::WarpState &w = cuemu::my_warp();
    w.slot[cuemu::lane_id()] = v;
    cuemu::warp_barrier(mask);
    unsigned r = 0;
    for (int l = 0; l < 32; l++) if ((mask >> l) & 1u) r = std::max(r, (unsigned)w.slot[l]);
    cuemu::warp_barrier(mask);
    return r;
}

inline int __reduce_max_sync(unsigned mask, int v)
{
    cuemu::WarpState &w = cuemu::my_warp();
    w.slot[cuemu::lane_id()] = (unsigned)v;
    cuemu::warp_barrier(mask);
    int r = INT32_MIN;
    for (int l = 0; l < 32; l++) if ((mask >> l) & 1u) r = std::max(r, (int)(unsigned)w.slot[l]);
    cuemu::warp_barrier(mask);
    return r;
}
inline unsigned __reduce_min_sync(unsigned mask, unsigned v)
{
    cuemu::WarpState &w = cuemu::my_warp();
    w.slot[cuemu::lane_id()] = v;
    cuemu::warp_barrier(mask);
    unsigned r = 0xffffffffu;
    for (int l = 0; l < 32; l++) if ((mask >> l) & 1u) r = std::min(r, (unsigned)w.slot[l]);
    cuemu::warp_barrier(mask);
    return r;
}
inline int __reduce_min_sync(unsigned mask, int v)
{
    cuemu::WarpState &w = cuemu::my_warp();
    w.slot[cuemu::lane_id()] = (unsigned)v;
    cuemu::warp_barrier(mask);
    int r = INT32_MAX;
    for (int l = 0; l < 32; l++) if ((mask >> l) & 1u) r = std::min(r, (int)(unsigned)w.slot[l]);
    cuemu::warp_barrier(mask);
    return r;
}

/* ---- atomics (single OS thread: plain read-modify-write) ------------- */
template <class T> inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <class T> inline T atomicOr(T *p, T v)  { T o = *p; *p = o | v; return o; }
template <class T> inline T atomicAnd(T *p, T v) { T o = *p; *p = o & v; return o; }
template <class T> inline T atomicMax(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> inline T atomicMin(T *p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> inline T atomicExch(T *p, T v){ T o = *p; *p = v; return o; }
template <class T> inline T atomicCAS(T *p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }

/* ---- integer intrinsics ---------------------------------------------- */
inline int __clz(int v)  { return v ? __builtin_clz((unsigned)v) : 32; }
inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
inline int __ffs(int v)  { return __builtin_ffs(v); }
inline int __ffsll(long long v) { return __builtin_ffsll(v); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline unsigned __brev(unsigned v)
{
    unsigned r = 0; for (int i = 0; i < 32; i++) if ((v >> i) & 1u) r |= 1u << (31 - i); return r;
}
inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s)
{
    uint64_t t = ((uint64_t)b << 32) | a; unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        unsigned sel = (s >> (4 * i)) & 0xf;
        unsigned byte = (unsigned)(t >> (8 * (sel & 7))) & 0xff;
        if (sel & 8) byte = (byte & 0x80) ? 0xff : 0x00;
        r |= byte << (8 * i);
    }
    return r;
}
inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned sh)
{
    sh &= 31; return sh ? (hi << sh) | (lo >> (32 - sh)) : hi;
}
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh)
{
    sh &= 31; return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
}
inline unsigned __sad(int a, int b, unsigned c)
{
    const long long d = (long long)a - (long long)b;
    return c + (unsigned)(d < 0 ? -d : d);
}
inline int __dp2a_lo(int a, int b, int c)
{
    return c + (int)(int16_t)(a & 0xffff) * (int)(int8_t)(b & 0xff) + (int)(int16_t)((unsigned)a >> 16) * (int)(int8_t)((b >> 8) & 0xff);
}
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
inline int __mulhi(int a, int b) { return (int)(((int64_t)a * b) >> 32); }
template <class T> inline T __ldg(const T *p) { return *p; }
using std::min;
using std::max;
inline unsigned min(unsigned a, int b) { return a < (unsigned)b ? a : (unsigned)b; }
inline unsigned min(int a, unsigned b) { return (unsigned)a < b ? (unsigned)a : b; }
inline unsigned max(unsigned a, int b) { return a > (unsigned)b ? a : (unsigned)b; }
inline unsigned max(int a, unsigned b) { return (unsigned)a > b ? (unsigned)a : b; }

/* ---- FP64 intrinsics (TU is compiled with -ffp-contract=off) ---------- */
inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
/* a + b rounded toward minus infinity */
inline double __dadd_rd(double a, double b)
{
    const int old = fegetround();
    fesetround(FE_DOWNWARD);
    volatile double x = a, y = b;
    volatile double r = x + y;
    fesetround(old);
    return r;
}
inline int __double2loint(double d) { uint64_t u; memcpy(&u, &d, 8); return (int)(uint32_t)u; }
inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
inline double __int2double_rn(int v) { return (double)v; }
inline int __double2int_rz(double x)
{
    if (x != x) return 0;
    if (x >= 2147483647.0) return 2147483647;
    if (x <= -2147483648.0) return (-2147483647 - 1);
    return (int)x;
}

/* ---- runtime API shim --------------------------------------------------- */
typedef int cudaError_t;
typedef void *cudaStream_t;
typedef struct cuemu_event { double t; } *cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDefault = 0, cudaEventBlockingSync = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0 };
/* page-locked or pageable: no difference here; CUEMU_ALL_PINNED=1 makes every host pointer report
 * as page-locked so that the tests can drive both the direct and the staged host paths */
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2 };
struct cudaPointerAttributes { cudaMemoryType type; };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };

inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaGetLastError() { return 0; }
inline cudaError_t cudaPeekAtLastError() { return 0; }
inline cudaError_t cudaSetDevice(int) { return 0; }
inline cudaError_t cudaGetDevice(int *d) { *d = 0; return 0; }
inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return 0; }
inline cudaError_t cudaMalloc(void **p, size_t n) { *p = calloc(n ? n : 1, 1); return *p ? 0 : 2; }
inline cudaError_t cudaFree(void *p) { free(p); return 0; }
inline cudaError_t cudaMallocHost(void **p, size_t n) { *p = calloc(n ? n : 1, 1); return *p ? 0 : 2; }
inline cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) { return cudaMallocHost(p, n); }
inline cudaError_t cudaFreeHost(void *p) { free(p); return 0; }
inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *)
{
    const char *e = getenv("CUEMU_ALL_PINNED");
    a->type = (e && *e == '1') ? cudaMemoryTypeHost : cudaMemoryTypeUnregistered;
    return 0;
}
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return 0; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memmove(d, s, n); return 0; }
inline cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return 0; }
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return 0; }
inline cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = (void *)1; return 0; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (void *)1; return 0; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
inline cudaError_t cudaDeviceSynchronize() { return 0; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new cuemu_event{0}; return 0; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return 0; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = 0) { return 0; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
inline cudaError_t cudaEventQuery(cudaEvent_t) { return 0; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return 0; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return 0; }
template <class F> inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return 0; }
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
inline cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr, int) { *v = 4; return 0; }
struct cudaDeviceProp { int multiProcessorCount; size_t sharedMemPerBlockOptin; char name[64]; int major, minor; };
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int)
{
    memset(p, 0, sizeof *p); p->multiProcessorCount = 148; p->sharedMemPerBlockOptin = 232448;
    strcpy(p->name, "cuda_emu"); p->major = 10; p->minor = 0; return 0;
}
#define cudaMemcpyToSymbol(sym, src, n) (memcpy((void *)&(sym), (src), (n)), 0)

/* launch + dynamic shared memory, see csrc/cuda_compat.h */
#define FB_LAUNCH(kernel, grid, block, smem, stream, ...) \
    cuemu::launch((grid), (block), (smem), [=]() { kernel(__VA_ARGS__); })
#define FB_DYN_SMEM(name) unsigned char *name = cuemu::g_cta.dyn_smem

#endif
