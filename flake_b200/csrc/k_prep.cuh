/*
 * k_prep.cuh -- frame table, VBS split decision and the per-frame "prepare"
 * kernel: PCM ingest + deinterleave (encode.c:541-553), stereo decorrelation
 * estimate and transform (encode.c:598-694), wasted bits (encode.c:558-593),
 * CONSTANT detection (optimize.c:143-151).
 *
 * Data layout written here: for a frame starting at sample `start` with n
 * samples and C channels, channel c's plane is the int32 run
 *     smp[start*C + c*n .. start*C + (c+1)*n)
 * so the planes of a chunk tile the scratch buffer exactly like the
 * interleaved input does.
 */
#ifndef FLAKE_B200_K_PREP_CUH
#define FLAKE_B200_K_PREP_CUH

#include "dev_common.cuh"

#ifndef FB_PREP_THREADS
#define FB_PREP_THREADS 128      /* with 12 CTAs per SM: 0.42 ms per C2 stream; 256 x 6: 0.48; 512 x 3: 0.70 */
#endif
#ifndef FB_PREP_MINBLOCKS
#define FB_PREP_MINBLOCKS 12    /* <= 42 registers */
#endif

/* staging-slot geometry: frame f of a chunk gets a slot that is large enough
 * for its VERBATIM encoding (16-byte aligned, 96 bytes of header slack). */
__host__ __device__ __forceinline__ uint32_t fb_slot_offset(uint32_t frame_index, uint32_t start,
                                                            int channels, int bps)
{
    uint64_t bits = (uint64_t)start * (uint64_t)(channels * bps + 1);
    return (uint32_t)(16u * (frame_index * 6u + (uint32_t)((bits + 127u) >> 7)));
}

/* fixed block size: frame f = block f (encode.c:979-1005 without VBS) */
__global__ void k_frames_fixed(FbConfig cfg, uint32_t nsamples, uint32_t first_number,
                               FbFrame *frames, uint32_t *nframes)
{
    const uint32_t B = (uint32_t)cfg.block_size;
    const uint32_t nf = (nsamples + B - 1) / B;
    for (uint32_t f = blockIdx.x * blockDim.x + threadIdx.x; f < nf; f += gridDim.x * blockDim.x) {
        FbFrame fr;
        fr.start = f * B;
        fr.n = min(B, nsamples - fr.start);
        fr.number = first_number + (cfg.allow_vbs ? fr.start : f);
        fr.slot = fb_slot_offset(f, fr.start, cfg.channels, cfg.bps);
        frames[f] = fr;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *nframes = nf;
}

/*
 * VBS split decision, one CTA per block -- vbs.c:36-83.
 * Writes sizes[block][8] (zero padded) and counts[block].
 * Blocks that do not qualify (n % 8, n < 128: encode.c:997-998) get one frame.
 */
__global__ void __launch_bounds__(FB_PREP_THREADS)
k_vbs_split(FbConfig cfg, const void *pcm, int fmt, uint32_t nsamples,
            uint32_t *sizes, uint32_t *counts)
{
    __shared__ uint64_t red[32];
    __shared__ long long energy[8];
    const uint32_t B = (uint32_t)cfg.block_size;
    const uint32_t blk = blockIdx.x;
    const uint32_t start = blk * B;
    if (start >= nsamples) return;
    const uint32_t n = min(B, nsamples - start);
    const int C = cfg.channels;
    uint32_t *my_sizes = sizes + (size_t)blk * 8;

    if ((n % 8u) != 0 || n < 128u) {
        if (threadIdx.x == 0) {
            for (int i = 0; i < 8; i++) my_sizes[i] = 0;
            my_sizes[0] = n; counts[blk] = 1;
        }
        return;
    }
    const uint32_t sec = n / 8;
    for (int s = 0; s < 8; s++) {
        /* sum over channels and j = 2..sec-1 of |x[j] - 2x[j-1] + x[j-2]|, int32 wrap + abs(int) */
        const size_t base = ((size_t)start + (size_t)s * sec) * (size_t)C;
        const uint32_t items = (sec - 2) * (uint32_t)C;
        uint64_t acc = 0;
        for (uint32_t it = threadIdx.x; it < items; it += blockDim.x) {
            const size_t idx = base + 2u * (uint32_t)C + it;      /* element (j, ch) with j >= 2 */
            uint32_t x0 = (uint32_t)fb_load_pcm(pcm, fmt, idx);
            uint32_t x1 = (uint32_t)fb_load_pcm(pcm, fmt, idx - (size_t)C);
            uint32_t x2 = (uint32_t)fb_load_pcm(pcm, fmt, idx - 2 * (size_t)C);
            int32_t v = (int32_t)(x0 - 2u * x1 + x2);
            int32_t a = v < 0 ? (int32_t)(0u - (uint32_t)v) : v;
            acc += (uint64_t)(int64_t)a;
        }
        acc = fb_block_sum_u64(acc, red);
        if (threadIdx.x == 0) energy[s] = (long long)acc / C + 1;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t out[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        int nf = 0;
        for (int s = 0; s < 8; s++) {
            bool begin = (s == 0);
            if (s > 0) {
                /* abs() on a truncated int and a 32-bit multiply */
                int32_t t = (int32_t)(uint32_t)(uint64_t)(energy[s - 1] - energy[s]);
                int32_t a = t < 0 ? (int32_t)(0u - (uint32_t)t) : t;
                int32_t prod = (int32_t)((uint32_t)a * 200u);
                begin = ((long long)prod / energy[s - 1]) > 50;
            }
            if (begin) nf++;
            out[nf - 1] += sec;
        }
        if (nf <= 1) { out[0] = n; nf = 1; for (int i = 1; i < 8; i++) out[i] = 0; }
        for (int i = 0; i < 8; i++) my_sizes[i] = out[i];
        counts[blk] = (uint32_t)nf;
    }
}

/* frame table from the per-block split: single CTA, serial over warps' chunks */
__global__ void __launch_bounds__(1024)
k_frames_vbs(FbConfig cfg, uint32_t nsamples, uint32_t first_number,
             const uint32_t *sizes, const uint32_t *counts, FbFrame *frames, uint32_t *nframes)
{
    __shared__ uint32_t scan_scratch[33];
    __shared__ uint32_t carry;
    const uint32_t B = (uint32_t)cfg.block_size;
    const uint32_t nblocks = (nsamples + B - 1) / B;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nblocks; base += blockDim.x) {
        const uint32_t blk = base + threadIdx.x;
        const uint32_t cnt = blk < nblocks ? counts[blk] : 0u;
        uint32_t total;
        const uint32_t ex = fb_block_exscan_u32(cnt, scan_scratch, &total);
        const uint32_t first = carry + ex;
        if (blk < nblocks) {
            uint32_t start = blk * B;
            for (uint32_t j = 0; j < cnt; j++) {
                FbFrame fr;
                fr.start = start;
                fr.n = sizes[(size_t)blk * 8 + j];
                fr.number = first_number + (cfg.allow_vbs ? start : first + j);
                fr.slot = fb_slot_offset(first + j, start, cfg.channels, cfg.bps);
                frames[first + j] = fr;
                start += fr.n;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *nframes = carry;
}

/* ------------------------------------------------------------------ */
/* prepare: one CTA per frame                                           */
/* ------------------------------------------------------------------ */
#define FB_PREP_RUN 8

/* Samples i-2 .. i+7 of a stereo frame into lw[0..9] / rw[0..9] (index 2 = sample i);
 * positions outside [0, n) read as 0.  `ibase` = interleaved element index of the frame's
 * sample 0.  Whole in-range runs of packed s16 (two 16-byte loads) and int32 (four) are
 * loaded vectorised when the address allows. */
__device__ __forceinline__ void fb_load_stereo_run(const void *pcm, int fmt, size_t ibase, int i, int n,
                                                   int32_t *lw, int32_t *rw)
{
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const int idx = i - 2 + k;
        const bool ok = idx >= 0 && idx < n;
        lw[k] = ok ? fb_load_pcm(pcm, fmt, ibase + 2 * (size_t)idx) : 0;
        rw[k] = ok ? fb_load_pcm(pcm, fmt, ibase + 2 * (size_t)idx + 1) : 0;
    }
    const size_t e0 = ibase + 2 * (size_t)i;
    if (i + FB_PREP_RUN <= n && fmt == FB_PCM_S16LE && ((((size_t)pcm) + e0 * 2) & 15u) == 0) {
        const uint4 *src = reinterpret_cast<const uint4 *>((const uint8_t *)pcm + e0 * 2);
        const uint4 a = src[0], b = src[1];
        const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int k = 0; k < 8; k++) { lw[2 + k] = (int32_t)(w[k] << 16) >> 16; rw[2 + k] = (int32_t)w[k] >> 16; }
    } else if (i + FB_PREP_RUN <= n && fmt == FB_PCM_S32 && ((((size_t)pcm) + e0 * 4) & 15u) == 0) {
        const int4 *src = reinterpret_cast<const int4 *>((const uint8_t *)pcm + e0 * 4);
#pragma unroll
        for (int g = 0; g < 4; g++) {
            const int4 v = src[g];
            lw[2 + 2 * g] = v.x; rw[2 + 2 * g] = v.y; lw[3 + 2 * g] = v.z; rw[3 + 2 * g] = v.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < FB_PREP_RUN; k++) {
            const bool ok = i + k < n;
            lw[2 + k] = ok ? fb_load_pcm(pcm, fmt, e0 + 2 * (size_t)k) : 0;
            rw[2 + k] = ok ? fb_load_pcm(pcm, fmt, e0 + 2 * (size_t)k + 1) : 0;
        }
    }
}

/* decorrelated pair (a, b) of a stereo sample, encode.c:648-694; MODE is the frame's ch_mode */
template <int MODE>
__device__ __forceinline__ void fb_decorrelate(int32_t l, int32_t r, int32_t &a, int32_t &b)
{
    if (MODE == 10)     { a = (int32_t)((uint32_t)l + (uint32_t)r) >> 1; b = (int32_t)((uint32_t)l - (uint32_t)r); }
    else if (MODE == 8) { a = l; b = (int32_t)((uint32_t)l - (uint32_t)r); }
    else if (MODE == 9) { a = (int32_t)((uint32_t)l - (uint32_t)r); b = r; }
    else                { a = l; b = r; }
}

/* plane statistics of one channel: OR of all samples (wasted bits), minimum and maximum
 * (|sample| bound for k_search's 32-bit test; min == max is the CONSTANT test) */
struct FbPlaneStat { uint32_t orv; int32_t mn, mx; };

/* deinterleave + decorrelate a stereo frame in 8-sample runs, gather the statistics */
template <int MODE>
__device__ __forceinline__ void fb_prep_stereo(const void *pcm, int fmt, size_t ibase, int n, int32_t *plane,
                                               FbPlaneStat &sa, FbPlaneStat &sb)
{
    const int tid = threadIdx.x, T = blockDim.x;
    const bool planes_aligned = ((((size_t)plane) | ((size_t)n * 4u)) & 15u) == 0;
    for (int i = tid * FB_PREP_RUN; i < n; i += T * FB_PREP_RUN) {
        int32_t lw[FB_PREP_RUN + 2], rw[FB_PREP_RUN + 2], av[FB_PREP_RUN], bv[FB_PREP_RUN];
        fb_load_stereo_run(pcm, fmt, ibase, i, n, lw, rw);
#pragma unroll
        for (int k = 0; k < FB_PREP_RUN; k++) fb_decorrelate<MODE>(lw[k + 2], rw[k + 2], av[k], bv[k]);
        if (i + FB_PREP_RUN <= n) {
#pragma unroll
            for (int k = 0; k < FB_PREP_RUN; k++) {
                sa.orv |= (uint32_t)av[k]; sa.mn = min(sa.mn, av[k]); sa.mx = max(sa.mx, av[k]);
                sb.orv |= (uint32_t)bv[k]; sb.mn = min(sb.mn, bv[k]); sb.mx = max(sb.mx, bv[k]);
            }
            if (planes_aligned) {
                int4 *pa = reinterpret_cast<int4 *>(plane + i), *pb = reinterpret_cast<int4 *>(plane + n + i);
                pa[0] = make_int4(av[0], av[1], av[2], av[3]); pa[1] = make_int4(av[4], av[5], av[6], av[7]);
                pb[0] = make_int4(bv[0], bv[1], bv[2], bv[3]); pb[1] = make_int4(bv[4], bv[5], bv[6], bv[7]);
            } else {
#pragma unroll
                for (int k = 0; k < FB_PREP_RUN; k++) { plane[i + k] = av[k]; plane[n + i + k] = bv[k]; }
            }
        } else {
#pragma unroll
            for (int k = 0; k < FB_PREP_RUN; k++) {
                if (i + k < n) {
                    sa.orv |= (uint32_t)av[k]; sa.mn = min(sa.mn, av[k]); sa.mx = max(sa.mx, av[k]);
                    sb.orv |= (uint32_t)bv[k]; sb.mn = min(sb.mn, bv[k]); sb.mx = max(sb.mx, bv[k]);
                    plane[i + k] = av[k]; plane[n + i + k] = bv[k];
                }
            }
        }
    }
}

__global__ void __launch_bounds__(FB_PREP_THREADS, FB_PREP_MINBLOCKS)
k_prep(FbConfig cfg, const void *pcm, int fmt, const FbFrame *frames, const uint32_t *nframes,
       int32_t *smp, FbSub *subs, uint8_t *ch_modes)
{
    __shared__ uint64_t s_sum[FB_PREP_THREADS / 32][4];
    __shared__ uint32_t s_or[FB_PREP_THREADS / 32][FB_MAX_CH_UNROLL];
    __shared__ int32_t s_mn[FB_PREP_THREADS / 32][FB_MAX_CH_UNROLL], s_mx[FB_PREP_THREADS / 32][FB_MAX_CH_UNROLL];
    __shared__ int s_mode;
    __shared__ int s_wasted[FB_MAX_CH_UNROLL];
    const uint32_t f = blockIdx.x;
    if (f >= *nframes) return;
    const FbFrame fr = frames[f];
    const int n = (int)fr.n, C = cfg.channels;
    const size_t ibase = (size_t)fr.start * (size_t)C;      /* interleaved index of sample 0 */
    int32_t *plane = smp + ibase;
    const int tid = threadIdx.x, T = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nw = (T + 31) >> 5;

    /* ---- stereo mode estimate, encode.c:598-643 ---------------------- */
    int mode = 0;                                            /* NOT_STEREO */
    if (C == 2) {
        mode = 1;                                            /* LEFT_RIGHT */
        if (n > 32 && cfg.stereo_method == 1) {
            uint64_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;
            for (int i = tid * FB_PREP_RUN; i < n; i += T * FB_PREP_RUN) {
                int32_t lw[FB_PREP_RUN + 2], rw[FB_PREP_RUN + 2];
                fb_load_stereo_run(pcm, fmt, ibase, i, n, lw, rw);
                /* 32-bit sums inside a run: 8 terms below 2^28 each when the packed format bounds
                 * the samples to 24 bits; int32 input may hold anything and is summed in 64 bits */
                uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
                const bool whole = i >= 2 && i + FB_PREP_RUN <= n;
#pragma unroll
                for (int k = 0; k < FB_PREP_RUN; k++) {
                    if (!whole && (i + k < 2 || i + k >= n)) continue;
                    const int32_t lt = (int32_t)((uint32_t)lw[k + 2] - 2u * (uint32_t)lw[k + 1] + (uint32_t)lw[k]);
                    const int32_t rt = (int32_t)((uint32_t)rw[k + 2] - 2u * (uint32_t)rw[k + 1] + (uint32_t)rw[k]);
                    const int32_t m = (int32_t)((uint32_t)lt + (uint32_t)rt) >> 1;
                    const int32_t d = (int32_t)((uint32_t)lt - (uint32_t)rt);
                    if (fmt != FB_PCM_S32) {
                        a0 += (uint32_t)(lt < 0 ? (int32_t)(0u - (uint32_t)lt) : lt);
                        a1 += (uint32_t)(rt < 0 ? (int32_t)(0u - (uint32_t)rt) : rt);
                        a2 += (uint32_t)(m < 0 ? (int32_t)(0u - (uint32_t)m) : m);
                        a3 += (uint32_t)(d < 0 ? (int32_t)(0u - (uint32_t)d) : d);
                    } else {
                        s0 += (uint64_t)(int64_t)(lt < 0 ? (int32_t)(0u - (uint32_t)lt) : lt);
                        s1 += (uint64_t)(int64_t)(rt < 0 ? (int32_t)(0u - (uint32_t)rt) : rt);
                        s2 += (uint64_t)(int64_t)(m < 0 ? (int32_t)(0u - (uint32_t)m) : m);
                        s3 += (uint64_t)(int64_t)(d < 0 ? (int32_t)(0u - (uint32_t)d) : d);
                    }
                }
                s0 += a0; s1 += a1; s2 += a2; s3 += a3;
            }
            s0 = fb_warp_sum_u64(s0); s1 = fb_warp_sum_u64(s1); s2 = fb_warp_sum_u64(s2); s3 = fb_warp_sum_u64(s3);
            if (lane == 0) { s_sum[warp][0] = s0; s_sum[warp][1] = s1; s_sum[warp][2] = s2; s_sum[warp][3] = s3; }
            __syncthreads();
            if (tid == 0) {
                uint64_t s[4] = {0, 0, 0, 0}, score[4];
                for (int w = 0; w < nw; w++)
                    for (int q = 0; q < 4; q++) s[q] += s_sum[w][q];
                for (int i = 0; i < 4; i++) {
                    /* k from the uint32-truncated search, cost kept in uint64 (encode.c:617-620) */
                    const uint64_t two = 2 * s[i];
                    const int best = fb_rice_k(two, n);
                    s[i] = fb_rice_count64(two, n, best);
                }
                score[0] = s[0] + s[1]; score[1] = s[0] + s[3];
                score[2] = s[1] + s[3]; score[3] = s[2] + s[3];
                int best = 0;
                for (int i = 1; i < 4; i++) if (score[i] < score[best]) best = i;
                s_mode = best == 0 ? 1 : (best == 1 ? 8 : (best == 2 ? 9 : 10));
            }
            __syncthreads();
            mode = s_mode;
        }
    }

    /* ---- deinterleave + decorrelate, gather plane statistics ----------- */
    FbPlaneStat st[FB_MAX_CH_UNROLL];
#pragma unroll
    for (int c = 0; c < FB_MAX_CH_UNROLL; c++) { st[c].orv = 0; st[c].mn = 0x7fffffff; st[c].mx = (int32_t)0x80000000; }

    if (C == 2) {
        if (mode == 10)     fb_prep_stereo<10>(pcm, fmt, ibase, n, plane, st[0], st[1]);
        else if (mode == 8) fb_prep_stereo<8>(pcm, fmt, ibase, n, plane, st[0], st[1]);
        else if (mode == 9) fb_prep_stereo<9>(pcm, fmt, ibase, n, plane, st[0], st[1]);
        else                fb_prep_stereo<1>(pcm, fmt, ibase, n, plane, st[0], st[1]);
    } else {
        for (int i = tid; i < n; i += T) {
#pragma unroll
            for (int c = 0; c < FB_MAX_CH_UNROLL; c++) {
                if (c < C) {
                    const int32_t a = fb_load_pcm(pcm, fmt, ibase + (size_t)i * C + c);
                    plane[(size_t)c * n + i] = a;
                    st[c].orv |= (uint32_t)a; st[c].mn = min(st[c].mn, a); st[c].mx = max(st[c].mx, a);
                }
            }
        }
    }

    /* ---- wasted bits (encode.c:558-593) + subframe records -------------- */
#pragma unroll
    for (int c = 0; c < FB_MAX_CH_UNROLL; c++) {
        if (c < C) {
            const uint32_t o = __reduce_or_sync(FB_FULL_MASK, st[c].orv);
            const int32_t lo = __reduce_min_sync(FB_FULL_MASK, st[c].mn);
            const int32_t hi = __reduce_max_sync(FB_FULL_MASK, st[c].mx);
            if (lane == 0) { s_or[warp][c] = o; s_mn[warp][c] = lo; s_mx[warp][c] = hi; }
        }
    }
    __syncthreads();
    if (tid < C) {
        const int c = tid;
        uint32_t o = 0;
        int32_t lo = 0x7fffffff, hi = (int32_t)0x80000000;
        for (int w = 0; w < nw; w++) { o |= s_or[w][c]; lo = min(lo, s_mn[w][c]); hi = max(hi, s_mx[w][c]); }
        int wasted = 0;
        if (o) {
            wasted = __ffs((int)o) - 1;
            if (wasted >= cfg.bps - 1) wasted = 0;
        }
        s_wasted[c] = wasted;
        /* bound on |sample|: ~v for negative v (two's complement magnitude - 1) */
        const uint32_t m = max((uint32_t)(hi < 0 ? ~hi : hi), (uint32_t)(lo < 0 ? ~lo : lo));
        FbSub *sb = &subs[(size_t)f * C + c];
        int obits = cfg.bps;
        if ((mode == 10 || mode == 8) && c == 1) obits++;
        if (mode == 9 && c == 0) obits++;
        sb->obits = obits - wasted;
        sb->wasted = wasted;
        sb->is_const = lo == hi ? 1 : 0;
        sb->first = lo >> wasted;                    /* read only when is_const: every sample equals lo */
        sb->maxabs = (m >> wasted) + 1u;
        sb->type = -1; sb->order = 0; sb->shift = 0; sb->method = 0; sb->porder = 0;
        sb->est_order = 0; sb->est_bits = 0;
    }
    if (tid == 0) ch_modes[f] = (uint8_t)mode;
    __syncthreads();
#pragma unroll
    for (int c = 0; c < FB_MAX_CH_UNROLL; c++) {
        if (c < C) {
            const int wasted = s_wasted[c];
            if (wasted) {
                int32_t *pl = plane + (size_t)c * n;
                if (C == 2) {                                       /* own elements only: same runs as above */
                    for (int i = tid * FB_PREP_RUN; i < n; i += T * FB_PREP_RUN)
                        for (int k = 0; k < FB_PREP_RUN && i + k < n; k++) pl[i + k] >>= wasted;
                } else {
                    for (int i = tid; i < n; i += T) pl[i] >>= wasted;
                }
            }
        }
    }
}

#endif
