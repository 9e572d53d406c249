/*
 * dev_common.cuh -- device helpers shared by the flake_b200 kernels.
 */
#ifndef FLAKE_B200_DEV_COMMON_CUH
#define FLAKE_B200_DEV_COMMON_CUH

#include "cuda_compat.h"
#include "engine.h"

#define FB_FULL_MASK 0xffffffffu
#define FB_MAX_CH_UNROLL 8

/* ------------------------------------------------------------------ */
/* PCM ingest                                                           */
/* ------------------------------------------------------------------ */
__device__ __forceinline__ int32_t fb_load_pcm(const void *pcm, int fmt, size_t idx)
{
    switch (fmt) {
    case FB_PCM_S16LE:
        return (int32_t)((const int16_t *)pcm)[idx];
    case FB_PCM_S24LE: {
        const uint8_t *b = (const uint8_t *)pcm + idx * 3;
        return (int32_t)((uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)(int32_t)(int8_t)b[2] << 16));
    }
    case FB_PCM_S8:
        return (int32_t)((const int8_t *)pcm)[idx];
    default:
        return ((const int32_t *)pcm)[idx];
    }
}

/* ------------------------------------------------------------------ */
/* warp / block reductions                                              */
/* ------------------------------------------------------------------ */
__device__ __forceinline__ uint64_t fb_warp_sum_u64(uint64_t v)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FB_FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ uint32_t fb_warp_sum_u32(uint32_t v)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FB_FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ uint32_t fb_warp_or_u32(uint32_t v)
{
    for (int o = 16; o > 0; o >>= 1) v |= __shfl_xor_sync(FB_FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ uint32_t fb_warp_max_u32(uint32_t v)
{
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(FB_FULL_MASK, v, o));
    return v;
}

/* All threads of the CTA must call; every thread receives the total.
 * `scratch` holds one uint64 per warp (<= 32). */
__device__ __forceinline__ uint64_t fb_block_sum_u64(uint64_t v, uint64_t *scratch)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nw = (blockDim.x + 31) >> 5;
    v = fb_warp_sum_u64(v);
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    uint64_t t = 0;
    for (int w = 0; w < nw; w++) t += scratch[w];
    __syncthreads();
    return t;
}
__device__ __forceinline__ uint32_t fb_block_or_u32(uint32_t v, uint64_t *scratch)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nw = (blockDim.x + 31) >> 5;
    v = fb_warp_or_u32(v);
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    uint32_t t = 0;
    for (int w = 0; w < nw; w++) t |= (uint32_t)scratch[w];
    __syncthreads();
    return t;
}
__device__ __forceinline__ uint32_t fb_block_max_u32(uint32_t v, uint64_t *scratch)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nw = (blockDim.x + 31) >> 5;
    v = fb_warp_max_u32(v);
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    uint32_t t = 0;
    for (int w = 0; w < nw; w++) t = max(t, (uint32_t)scratch[w]);
    __syncthreads();
    return t;
}

/* exclusive scan of one uint32 per thread over the CTA; *total gets the sum.
 * scratch: one uint32 per warp + 1. */
__device__ __forceinline__ uint32_t fb_block_exscan_u32(uint32_t v, uint32_t *scratch, uint32_t *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nw = (blockDim.x + 31) >> 5;
    uint32_t inc = v;
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(FB_FULL_MASK, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) scratch[warp] = inc;
    __syncthreads();
    uint32_t base = 0, all = 0;
    for (int w = 0; w < nw; w++) {
        uint32_t s = scratch[w];
        if (w < warp) base += s;
        all += s;
    }
    __syncthreads();
    *total = all;
    return base + inc - v;
}

/* Same, with ONE barrier: `scratch` must not be written again before another barrier has
 * been passed by the whole CTA (callers give every scan its own scratch row). */
__device__ __forceinline__ uint32_t fb_block_exscan_u32_once(uint32_t v, uint32_t *scratch, uint32_t *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nw = (blockDim.x + 31) >> 5;
    uint32_t inc = v;
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(FB_FULL_MASK, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) scratch[warp] = inc;
    __syncthreads();
    uint32_t base = 0, all = 0;
    for (int w = 0; w < nw; w++) {
        uint32_t s = scratch[w];
        if (w < warp) base += s;
        all += s;
    }
    *total = all;
    return base + inc - v;
}

/* ------------------------------------------------------------------ */
/* integer helpers with the reference's exact semantics                 */
/* ------------------------------------------------------------------ */
/* floor(log2(v)), 0 for v == 0 -- common.h:53-66 */
__device__ __forceinline__ int fb_ilog2(uint32_t v) { return v ? 31 - __clz((int)v) : 0; }

/* rice.h:48 in uint64, truncated to 32 bits by the callers */
__device__ __forceinline__ uint64_t fb_rice_count64(uint64_t sum, int n, int k)
{
    return (uint64_t)((int64_t)n * (int64_t)(k + 1)) + ((sum - (uint64_t)(n >> 1)) >> k);
}

/* rice.c:30-45 literally: first strict minimum of the uint32-truncated cost over k = 0..30.
 * Out of line: only sums >= 2^31 (where the truncation matters) come here. */
__device__ __noinline__ int fb_rice_k_scan(uint64_t sum, int n)
{
    int best = 0;
    uint32_t best_bits = 0xffffffffu;
#pragma unroll 1
    for (int k = 0; k <= 30; k++) {
        uint32_t b = (uint32_t)fb_rice_count64(sum, n, k);
        if (b < best_bits) { best_bits = b; best = k; }
    }
    return best;
}

/* rice.c:30-45.  The cost is convex in k while nothing wraps, which gives the closed form:
 * smallest k with ((sum - n/2) >> k) <= 2n; otherwise scan. */
__device__ __forceinline__ int fb_rice_k(uint64_t sum, int n)
{
    if (sum < 0x80000000ull) {
        int64_t s = (int64_t)sum - (int64_t)(n >> 1);
        uint32_t t = 2u * (uint32_t)n;
        if (s <= (int64_t)t) return 0;
        uint32_t su = (uint32_t)s;                     /* 0 < s < 2^31 */
        int k = (32 - __clz((int)su)) - (32 - __clz((int)t));   /* t >= 1 here since s > t >= 0 ... */
        if (k < 0) k = 0;
        if ((su >> k) > t) k++;
        return k > 30 ? 30 : k;
    }
    return fb_rice_k_scan(sum, n);
}

/* the same two for sums known to be below 2^31 (and n <= 65535): all 32-bit.  `sum - n/2` is
 * negative only where k == 0, where wrapping modulo 2^32 equals the reference's 64-bit wrap
 * truncated to uint32. */
__device__ __forceinline__ int fb_rice_k_t(uint32_t sum, int n)
{
    const int s = (int)sum - (n >> 1);
    const int t = 2 * n;
    if (s <= t) return 0;
    int k = __clz(t) - __clz(s);                        /* 0 <= t < s < 2^31 */
    if (k < 0) k = 0;
    if ((s >> k) > t) k++;
    return k > 30 ? 30 : k;
}

__device__ __forceinline__ uint32_t fb_rice_count_t(uint32_t sum, int n, int k)
{
    return (uint32_t)(n * (k + 1)) + ((sum - (uint32_t)(n >> 1)) >> k);
}

__device__ __forceinline__ int fb_rice_k_t(unsigned long long sum, int n) { return fb_rice_k((uint64_t)sum, n); }

__device__ __forceinline__ uint32_t fb_rice_count_t(unsigned long long sum, int n, int k)
{
    return (uint32_t)fb_rice_count64((uint64_t)sum, n, k);
}

/* rice.c:148-155 */
__device__ __forceinline__ int fb_limit_porder(int p, int n, int order)
{
    int lim = __ffs(n) - 1;                 /* ctz(n) == log2i(n ^ (n-1)) */
    if (lim < p) p = lim;
    if (order > 0) {
        int l2 = fb_ilog2((uint32_t)(n / order));
        if (l2 < p) p = l2;
    }
    return p;
}

__device__ __forceinline__ uint32_t fb_zigzag(int32_t v)
{
    return ((uint32_t)v << 1) ^ (uint32_t)(v >> 31);
}

/* encode.c:522-527 */
__device__ __forceinline__ int fb_verbatim_size(const FbConfig &cfg, int n)
{
    if (cfg.channels == 2) return 16 + ((n * (cfg.bps + cfg.bps + 1) + 7) >> 3);
    return 16 + ((n * cfg.channels * cfg.bps + 7) >> 3);
}

#endif
