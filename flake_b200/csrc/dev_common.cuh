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

__device__ __forceinline__ int fb_pcm_bytes(int fmt)
{
    return fmt == FB_PCM_S16LE ? 2 : fmt == FB_PCM_S24LE ? 3 : fmt == FB_PCM_S8 ? 1 : 4;
}

/* decorrelated value of channel c (0 or 1) of a stereo sample, encode.c:648-694; `mode` is the
 * frame's ch_mode (1 left/right, 8 left/side, 9 right/side, 10 mid/side) */
__device__ __forceinline__ int32_t fb_stereo_value(int mode, int c, int32_t l, int32_t r)
{
    const int32_t side = (int32_t)((uint32_t)l - (uint32_t)r);
    if (c == 0) return mode == 10 ? (int32_t)((uint32_t)l + (uint32_t)r) >> 1 : (mode == 9 ? side : l);
    return mode == 1 || mode == 9 ? r : side;
}

/*
 * The same transform as one linear form, value = (alpha * l + beta * r) >> gamma with the wasted
 * bits folded into the shift ((x >> a) >> b == x >> (a + b) for arithmetic shifts):
 *   left (1, 0, w)   right (0, 1, w)   side (1, -1, w)   mid (1, 1, 1 + w)
 * packed as alpha | beta << 8 | gamma << 16 (alpha, beta as signed bytes): for packed 16-bit
 * stereo the word (r << 16 | l) goes through ONE two-way dot product (dp2a) and one shift,
 * whatever the mode.  32-bit wrap-around like the reference's int arithmetic.
 */
__device__ __forceinline__ int fb_stereo_coef(int mode, int c, int wasted)
{
    int a = 1, b = 0, g = wasted;
    if (c == 0) {
        if (mode == 10) { b = 1; g = wasted + 1; }
        else if (mode == 9) b = -1;
    } else {
        if (mode == 1 || mode == 9) { a = 0; b = 1; }
        else b = -1;
    }
    return (a & 0xff) | ((b & 0xff) << 8) | (g << 16);
}
__device__ __forceinline__ int32_t fb_stereo_apply(int coef, int32_t l, int32_t r)
{
    const int a = (int)(int8_t)(coef & 0xff), b = (int)(int8_t)((coef >> 8) & 0xff);
    return (int32_t)((uint32_t)a * (uint32_t)l + (uint32_t)b * (uint32_t)r) >> (coef >> 16);
}
/* w = r << 16 | (l & 0xffff): a packed 16-bit stereo sample pair as it lies in memory */
__device__ __forceinline__ int32_t fb_stereo_apply16(int coef, uint32_t w)
{
    return __dp2a_lo((int)w, coef, 0) >> (coef >> 16);
}

/* Channel layouts whose subframes are read from deinterleaved int32 planes written by k_prep
 * (more than two channels); mono and stereo subframes are read from the packed PCM. */
__host__ __device__ __forceinline__ bool fb_uses_planes(int channels) { return channels > 2; }

/*
 * Sample p of subframe channel c as the analysis and the packer see it: deinterleaved
 * (encode.c:541-553), decorrelated when the frame is stereo (encode.c:648-694), wasted bits
 * shifted out (encode.c:558-593) -- straight from the packed PCM (no int32 plane exists in
 * global memory).  `ebase` = interleaved element index of the frame's sample 0.
 */
__device__ __forceinline__ int32_t fb_pcm_sample(const void *pcm, int fmt, size_t ebase, int C, int c,
                                                 int mode, int wasted, int p)
{
    int32_t v;
    if (C == 2) {
        const size_t e = ebase + 2 * (size_t)p;
        v = fb_stereo_value(mode, c, fb_load_pcm(pcm, fmt, e), fb_load_pcm(pcm, fmt, e + 1));
    } else {
        v = fb_load_pcm(pcm, fmt, ebase + (size_t)p * (size_t)C + (size_t)c);
    }
    return v >> wasted;
}

/* ------------------------------------------------------------------ */
/* frame tile: the packed PCM of a frame streamed through shared memory  */
/* ------------------------------------------------------------------ */
/*
 * The consumers of a frame (k_prep: statistics; k_search: the subframe's plane) read its packed
 * PCM through a two-stage shared-memory ring filled by TMA bulk copies (cp.async.bulk, one
 * instruction per chunk issued by thread 0, completion counted on an mbarrier): chunk k+2 is
 * in flight while chunk k is consumed.  A chunk holds a whole number of 16-sample groups
 * (16 * channels * bytes-per-sample divides it) so that no sample straddles two chunks.
 * Frames whose first byte is not 16-byte aligned (odd block sizes) are copied by the threads
 * instead; the consumer code is the same.
 */
#define FB_TILE_CHUNK_BYTES 8192
#define FB_TILE_STAGES 2
#define FB_TILE_BYTES (FB_TILE_CHUNK_BYTES * FB_TILE_STAGES)

struct FbTile {
    const uint8_t *src;         /* first byte of the frame in global memory */
    uint8_t *buf;               /* shared, FB_TILE_BYTES, 16-byte aligned */
    fb_mbar_t *bar;             /* shared, FB_TILE_STAGES */
    uint32_t total;             /* bytes of the frame */
    uint32_t chunk;             /* bytes per chunk */
    uint32_t nchunks;
    uint32_t chunk_samples;     /* inter-channel samples per chunk (multiple of 16) */
    unsigned long long avail;   /* bytes of the PCM buffer from src on: a copy may be rounded up to 16 within it */
    bool tma;
};

__device__ __forceinline__ void fb_tile_issue(const FbTile &t, uint32_t k)
{
    const uint32_t off = k * t.chunk;
    const uint32_t bytes = min(t.chunk, t.total - off);
    uint32_t cp = (bytes + 15u) & ~15u;
    if ((unsigned long long)off + cp > t.avail) cp = bytes & ~15u;      /* the tail is copied by threads */
    fb_mbar_expect_tx(&t.bar[k % FB_TILE_STAGES], cp);
    if (cp) fb_bulk_g2s(t.buf + (k % FB_TILE_STAGES) * FB_TILE_CHUNK_BYTES, t.src + off, cp, &t.bar[k % FB_TILE_STAGES]);
}

/* every thread of the CTA calls; returns after the barriers are initialised and the first chunks requested */
__device__ __forceinline__ void fb_tile_begin(FbTile &t, const void *pcm, int fmt, size_t ebase, int n, int C,
                                              unsigned long long pcm_bytes, uint8_t *buf, fb_mbar_t *bar)
{
    const uint32_t bps = (uint32_t)fb_pcm_bytes(fmt);
    const uint32_t unit = 16u * (uint32_t)C * bps;
    const unsigned long long byte_off = (unsigned long long)ebase * bps;
    t.src = (const uint8_t *)pcm + byte_off;
    t.buf = buf; t.bar = bar;
    t.total = (uint32_t)n * (uint32_t)C * bps;
    t.chunk = (FB_TILE_CHUNK_BYTES / unit) * unit;
    t.chunk_samples = t.chunk / ((uint32_t)C * bps);
    t.nchunks = (t.total + t.chunk - 1u) / t.chunk;
    t.avail = pcm_bytes > byte_off ? pcm_bytes - byte_off : 0ull;
    t.tma = (((size_t)t.src) & 15u) == 0;
    if (t.tma) {
        if (threadIdx.x == 0) {
            for (int s = 0; s < FB_TILE_STAGES; s++) fb_mbar_init(&bar[s], 1);
            fb_mbar_init_fence();
            for (uint32_t k = 0; k < FB_TILE_STAGES && k < t.nchunks; k++) fb_tile_issue(t, k);
        }
        __syncthreads();
    }
}

/* wait for chunk k; returns its first byte in shared memory, *bytes = valid bytes */
__device__ __forceinline__ const uint8_t *fb_tile_acquire(const FbTile &t, uint32_t k, uint32_t *bytes)
{
    const uint32_t off = k * t.chunk;
    const uint32_t nb = min(t.chunk, t.total - off);
    uint8_t *dst = t.buf + (k % FB_TILE_STAGES) * FB_TILE_CHUNK_BYTES;
    *bytes = nb;
    if (t.tma) {
        fb_mbar_wait(&t.bar[k % FB_TILE_STAGES], (k / FB_TILE_STAGES) & 1u);
        uint32_t cp = (nb + 15u) & ~15u;
        if ((unsigned long long)off + cp > t.avail) {                /* CTA-uniform: the last bytes of the buffer */
            cp = nb & ~15u;
            for (uint32_t i = cp + threadIdx.x; i < nb; i += blockDim.x) dst[i] = t.src[off + i];
            __syncthreads();
        }
    } else {
        for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) dst[i] = t.src[off + i];
        __syncthreads();
    }
    return dst;
}

/* Every thread has finished reading chunk k: its stage is refilled with chunk k + FB_TILE_STAGES.
 * The barrier in front of the refill is skipped when nothing will be refilled (the frame fits
 * the ring: 16-bit stereo blocks of 4096) unless the caller passes state from chunk to chunk
 * through shared memory (`always_sync`). */
__device__ __forceinline__ void fb_tile_release(const FbTile &t, uint32_t k, bool always_sync = true)
{
    if (always_sync || k + FB_TILE_STAGES < t.nchunks) __syncthreads();          /* CTA-uniform */
    if (t.tma && threadIdx.x == 0 && k + FB_TILE_STAGES < t.nchunks) fb_tile_issue(t, k + FB_TILE_STAGES);
}

/* element e (interleaved index inside the chunk) of a staged chunk */
__device__ __forceinline__ int32_t fb_tile_elem(const uint8_t *tile, int fmt, uint32_t e)
{
    return fb_load_pcm(tile, fmt, e);
}

/* four consecutive stereo samples (l, r pairs) starting at pair index q (multiple of 4) of a chunk */
__device__ __forceinline__ void fb_tile_stereo4(const uint8_t *tile, int fmt, uint32_t q, int32_t *l, int32_t *r)
{
    if (fmt == FB_PCM_S16LE) {
        const uint4 v = *reinterpret_cast<const uint4 *>(tile + 4u * q);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; k++) { l[k] = (int32_t)(w[k] << 16) >> 16; r[k] = (int32_t)w[k] >> 16; }
    } else if (fmt == FB_PCM_S24LE) {
        /* 24 bytes, 8-byte aligned: three 64-bit loads; a sample is three bytes moved to the top of a
         * word and shifted back down with sign */
        const uint2 *src = reinterpret_cast<const uint2 *>(tile + 6u * q);
        const uint2 a = src[0], b = src[1], c = src[2];
        const uint32_t w[6] = {a.x, a.y, b.x, b.y, c.x, c.y};
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const uint32_t w0 = w[3 * h], w1 = w[3 * h + 1], w2 = w[3 * h + 2];
            l[2 * h]     = (int32_t)__byte_perm(w0, 0u, 0x2100) >> 8;
            r[2 * h]     = (int32_t)__byte_perm(w0, w1, 0x5430) >> 8;
            l[2 * h + 1] = (int32_t)__byte_perm(w1, w2, 0x4320) >> 8;
            r[2 * h + 1] = (int32_t)w2 >> 8;
        }
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++) { l[k] = fb_load_pcm(tile, fmt, 2u * (q + k)); r[k] = fb_load_pcm(tile, fmt, 2u * (q + k) + 1u); }
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
