/*
 * k_prep.cuh -- frame table, VBS split decision and the per-frame "prepare"
 * kernel: PCM ingest + deinterleave (encode.c:541-553), stereo decorrelation
 * estimate and transform (encode.c:598-694), wasted bits (encode.c:558-593),
 * CONSTANT detection (optimize.c:143-151).
 *
 * k_prep reads the packed PCM through a TMA-filled shared-memory tile and writes
 * decisions: the stereo mode per frame and, per subframe, wasted bits, sample width,
 * the CONSTANT flag and the magnitude bound.  For mono and stereo input that is all:
 * the consumers deinterleave and decorrelate the packed PCM themselves and no int32
 * plane is kept in global memory.  With more than two channels a consumer of ONE
 * channel would have to pull the whole interleaved frame through its shared memory
 * (8 channels: eight times the bytes it needs), so there k_prep also deinterleaves
 * into int32 planes, once, for k_lpc / k_search / k_pack to read (fb_uses_planes).
 */
#ifndef FLAKE_B200_K_PREP_CUH
#define FLAKE_B200_K_PREP_CUH

#include "dev_common.cuh"

#ifndef FB_PREP_THREADS
#define FB_PREP_THREADS 128      /* with 12 CTAs per SM: 0.42 ms per C2 stream; 256 x 6: 0.48; 512 x 3: 0.70 */
#endif
#ifndef FB_PREP_MINBLOCKS
#define FB_PREP_MINBLOCKS 8     /* <= 64 registers; 8 x 16 KB of tile staging per SM */
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
/* Also clears what k_pack counts in: its ticket (nframes[2]), the look-back status of every frame and
 * the chunk summary. */
__global__ void k_frames_fixed(FbConfig cfg, uint32_t nsamples, uint32_t first_number,
                               FbFrame *frames, uint32_t *nframes, unsigned long long *status, FbSummary *summary)
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
        status[f] = 0ull;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        nframes[0] = nf; nframes[1] = 0; nframes[2] = 0; nframes[3] = 0;
        summary->nframes = 0; summary->max_frame_bytes = 0; summary->total_bytes = 0;
        summary->verbatim_frames = 0; summary->min_frame_inv = 0;
    }
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
    __shared__ uint64_t red[8 * 8];                              /* [warp][section], FB_PREP_THREADS <= 256 */
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
    /* Sum over channels and j = 2..sec-1 of |x[j] - 2 x[j-1] + x[j-2]| (int32 wrap + abs(int), vbs.c:52-60).
     * A thread owns a channel and a range of consecutive j, so that x[j-1] and x[j-2] stay in registers
     * (one load per element instead of three; a 24-bit sample is three byte loads); the threads of a
     * channel group read neighbouring channels of the same sample.  Unsigned 64-bit adds in any order. */
    const int G = (int)blockDim.x / C;                           /* ranges per section (blockDim >= 8 channels) */
    const int c = (int)threadIdx.x % C, g = (int)threadIdx.x / C;
    const uint32_t len = (sec - 2u + (uint32_t)G - 1u) / (uint32_t)G;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = ((int)blockDim.x + 31) >> 5;
    for (int s = 0; s < 8; s++) {
        uint64_t acc = 0;
        const uint32_t j0 = 2u + (uint32_t)g * len, j1 = min(sec, j0 + len);
        if (g < G && j0 < j1) {
            const size_t e0 = ((size_t)start + (size_t)s * sec) * (size_t)C + (size_t)c;   /* element (0, c) of the section */
            uint32_t x2 = (uint32_t)fb_load_pcm(pcm, fmt, e0 + (size_t)(j0 - 2u) * (size_t)C);
            uint32_t x1 = (uint32_t)fb_load_pcm(pcm, fmt, e0 + (size_t)(j0 - 1u) * (size_t)C);
#pragma unroll 4
            for (uint32_t j = j0; j < j1; j++) {
                const uint32_t x0 = (uint32_t)fb_load_pcm(pcm, fmt, e0 + (size_t)j * (size_t)C);
                const int32_t v = (int32_t)(x0 - 2u * x1 + x2);
                const int32_t a = v < 0 ? (int32_t)(0u - (uint32_t)v) : v;
                acc += (uint64_t)(int64_t)a;
                x2 = x1; x1 = x0;
            }
        }
        acc = fb_warp_sum_u64(acc);
        if (lane == 0) red[warp * 8 + s] = acc;
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        uint64_t acc = 0;
        for (int w = 0; w < nw; w++) acc += red[w * 8 + (int)threadIdx.x];
        energy[threadIdx.x] = (long long)acc / C + 1;
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
/* prepare: one CTA per frame, READ-ONLY over the packed PCM             */
/* ------------------------------------------------------------------ */
/*
 * One pass over the frame's packed PCM, staged through shared memory by TMA bulk copies
 * (FbTile, dev_common.cuh).  Nothing is written back but the decisions: no int32 plane exists
 * in global memory -- k_lpc, k_search and k_pack deinterleave and decorrelate the packed PCM
 * themselves from the frame's stereo decision (encode.c:541-553, 648-694).
 *
 *   stereo: the four estimate sums of calc_decorr_scores (encode.c:598-643) and, because the
 *           decision is not known yet, the statistics of all four candidate channels
 *           (left, right, mid, side); the chosen pair's statistics become the subframe records;
 *   other:  per-channel statistics.
 * Statistics per plane: OR of all samples (`ctz(OR)` = wasted bits, encode.c:558-593),
 * minimum and maximum (min == max is the CONSTANT test, optimize.c:143-151; max(|min|,|max|)
 * bounds k_search's 32-bit arithmetic test).
 */
#define FB_PREP_RUN 8

struct FbPlaneStat { uint32_t orv; int32_t mn, mx; };

__device__ __forceinline__ void fb_stat_add(FbPlaneStat &s, int32_t v)
{
    s.orv |= (uint32_t)v; s.mn = min(s.mn, v); s.mx = max(s.mx, v);
}

__global__ void __launch_bounds__(FB_PREP_THREADS, FB_PREP_MINBLOCKS)
k_prep(FbConfig cfg, const void *pcm, int fmt, unsigned long long pcm_bytes, const FbFrame *frames,
       const uint32_t *nframes, int32_t *planes, FbSub *subs, uint8_t *ch_modes)
{
    __shared__ __align__(16) uint8_t s_tile[FB_TILE_BYTES];
    __shared__ fb_mbar_t s_bar[FB_TILE_STAGES];
    __shared__ int32_t s_carry[2][4];                 /* last two (l, r) pairs of a chunk, by chunk parity */
    __shared__ uint64_t s_sum[FB_PREP_THREADS / 32][4];
    __shared__ uint32_t s_or[FB_PREP_THREADS / 32][FB_MAX_CH_UNROLL];
    __shared__ int32_t s_mn[FB_PREP_THREADS / 32][FB_MAX_CH_UNROLL], s_mx[FB_PREP_THREADS / 32][FB_MAX_CH_UNROLL];
    __shared__ int s_wasted[FB_MAX_CH_UNROLL];
    const uint32_t f = blockIdx.x;
    if (f >= *nframes) return;
    const FbFrame fr = frames[f];
    const int n = (int)fr.n, C = cfg.channels;
    const size_t ibase = (size_t)fr.start * (size_t)C;      /* interleaved index of sample 0 */
    const int tid = threadIdx.x, T = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nw = (T + 31) >> 5;

    FbTile tile;
    fb_tile_begin(tile, pcm, fmt, ibase, n, C, pcm_bytes, s_tile, s_bar);

    /* planes 0..3 of a stereo frame: left, right, mid, side; else one per channel */
    FbPlaneStat st[FB_MAX_CH_UNROLL];
#pragma unroll
    for (int c = 0; c < FB_MAX_CH_UNROLL; c++) { st[c].orv = 0; st[c].mn = 0x7fffffff; st[c].mx = (int32_t)0x80000000; }
    uint64_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    const bool estimate = C == 2 && n > 32 && cfg.stereo_method == 1;

    for (uint32_t k = 0; k < tile.nchunks; k++) {
        uint32_t nb;
        const uint8_t *t = fb_tile_acquire(tile, k, &nb);
        const int base = (int)(k * tile.chunk_samples);               /* first sample of the chunk */
        const int cs = min((int)tile.chunk_samples, n - base);        /* samples in the chunk */
        if (C == 2) {
            for (int q = tid * FB_PREP_RUN; q < cs; q += T * FB_PREP_RUN) {
                /* lw[2 + j] = sample base + q + j; lw[0], lw[1] = the two before it */
                int32_t lw[FB_PREP_RUN + 2], rw[FB_PREP_RUN + 2];
                if (q + FB_PREP_RUN <= cs) {
                    fb_tile_stereo4(t, fmt, (uint32_t)q, lw + 2, rw + 2);
                    fb_tile_stereo4(t, fmt, (uint32_t)q + 4u, lw + 6, rw + 6);
                } else {
#pragma unroll
                    for (int j = 0; j < FB_PREP_RUN; j++) {
                        const bool ok = q + j < cs;
                        lw[2 + j] = ok ? fb_tile_elem(t, fmt, 2u * (uint32_t)(q + j)) : 0;
                        rw[2 + j] = ok ? fb_tile_elem(t, fmt, 2u * (uint32_t)(q + j) + 1u) : 0;
                    }
                }
                if (q >= 2) {
                    lw[0] = fb_tile_elem(t, fmt, 2u * (uint32_t)(q - 2)); rw[0] = fb_tile_elem(t, fmt, 2u * (uint32_t)(q - 2) + 1u);
                    lw[1] = fb_tile_elem(t, fmt, 2u * (uint32_t)(q - 1)); rw[1] = fb_tile_elem(t, fmt, 2u * (uint32_t)(q - 1) + 1u);
                } else {                                              /* q == 0: the previous chunk's tail */
                    lw[0] = k ? s_carry[(k - 1) & 1][0] : 0; rw[0] = k ? s_carry[(k - 1) & 1][1] : 0;
                    lw[1] = k ? s_carry[(k - 1) & 1][2] : 0; rw[1] = k ? s_carry[(k - 1) & 1][3] : 0;
                }
                if (q + FB_PREP_RUN == cs) {                          /* the last run of a chunk leaves the carry (a chunk
                                                                       * that is followed by another one is whole runs) */
                    s_carry[k & 1][0] = lw[FB_PREP_RUN]; s_carry[k & 1][1] = rw[FB_PREP_RUN];
                    s_carry[k & 1][2] = lw[FB_PREP_RUN + 1]; s_carry[k & 1][3] = rw[FB_PREP_RUN + 1];
                }
                /* 32-bit sums inside a run: 8 terms below 2^28 each when the packed format bounds
                 * the samples to 24 bits; int32 input may hold anything and is summed in 64 bits */
                uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
                for (int j = 0; j < FB_PREP_RUN; j++) {
                    if (q + j >= cs) continue;
                    const int32_t l = lw[j + 2], r = rw[j + 2];
                    fb_stat_add(st[0], l);
                    fb_stat_add(st[1], r);
                    fb_stat_add(st[2], (int32_t)((uint32_t)l + (uint32_t)r) >> 1);
                    fb_stat_add(st[3], (int32_t)((uint32_t)l - (uint32_t)r));
                    if (!estimate || base + q + j < 2) continue;
                    const int32_t lt = (int32_t)((uint32_t)l - 2u * (uint32_t)lw[j + 1] + (uint32_t)lw[j]);
                    const int32_t rt = (int32_t)((uint32_t)r - 2u * (uint32_t)rw[j + 1] + (uint32_t)rw[j]);
                    const int32_t m = (int32_t)((uint32_t)lt + (uint32_t)rt) >> 1;
                    const int32_t d = (int32_t)((uint32_t)lt - (uint32_t)rt);
                    if (fmt != FB_PCM_S32) {
                        /* |x| + acc in one instruction (VABSDIFF); lt, rt stay below 2^26 for packed
                         * input, so |lt - rt| needs no wrap-around */
                        a0 = __sad(lt, 0, a0);
                        a1 = __sad(rt, 0, a1);
                        a2 = __sad(m, 0, a2);
                        a3 = __sad(lt, rt, a3);
                    } else {
                        s0 += (uint64_t)(int64_t)(lt < 0 ? (int32_t)(0u - (uint32_t)lt) : lt);
                        s1 += (uint64_t)(int64_t)(rt < 0 ? (int32_t)(0u - (uint32_t)rt) : rt);
                        s2 += (uint64_t)(int64_t)(m < 0 ? (int32_t)(0u - (uint32_t)m) : m);
                        s3 += (uint64_t)(int64_t)(d < 0 ? (int32_t)(0u - (uint32_t)d) : d);
                    }
                }
                s0 += a0; s1 += a1; s2 += a2; s3 += a3;
            }
        } else {
            /* deinterleave (encode.c:541-553): channel c of the frame is the plane
             * [ibase + c n, + n); consecutive threads write consecutive words of a plane */
            const bool to_planes = fb_uses_planes(C);
            for (int q = tid; q < cs; q += T) {
#pragma unroll
                for (int c = 0; c < FB_MAX_CH_UNROLL; c++)
                    if (c < C) {
                        const int32_t v = fb_tile_elem(t, fmt, (uint32_t)q * (uint32_t)C + (uint32_t)c);
                        fb_stat_add(st[c], v);
                        if (to_planes) planes[ibase + (size_t)c * (size_t)n + (size_t)(base + q)] = v;
                    }
            }
        }
        fb_tile_release(tile, k);
    }

    /* ---- reductions ------------------------------------------------------------------ */
    if (estimate) {
        s0 = fb_warp_sum_u64(s0); s1 = fb_warp_sum_u64(s1); s2 = fb_warp_sum_u64(s2); s3 = fb_warp_sum_u64(s3);
        if (lane == 0) { s_sum[warp][0] = s0; s_sum[warp][1] = s1; s_sum[warp][2] = s2; s_sum[warp][3] = s3; }
    }
    const int nplanes = C == 2 ? 4 : C;
#pragma unroll
    for (int c = 0; c < FB_MAX_CH_UNROLL; c++) {
        if (c < nplanes) {
            const uint32_t o = __reduce_or_sync(FB_FULL_MASK, st[c].orv);
            const int32_t lo = __reduce_min_sync(FB_FULL_MASK, st[c].mn);
            const int32_t hi = __reduce_max_sync(FB_FULL_MASK, st[c].mx);
            if (lane == 0) { s_or[warp][c] = o; s_mn[warp][c] = lo; s_mx[warp][c] = hi; }
        }
    }
    __syncthreads();

    /* ---- stereo mode, encode.c:598-643 (every thread of warp 0 computes the same) ----- */
    int mode = 0;                                            /* NOT_STEREO */
    if (C == 2) {
        mode = 1;                                            /* LEFT_RIGHT */
        if (estimate) {
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
            mode = best == 0 ? 1 : (best == 1 ? 8 : (best == 2 ? 9 : 10));
        }
    }

    /* ---- wasted bits (encode.c:558-593) + subframe records ---------------------------- */
    if (tid < C) {
        const int c = tid;
        /* the plane this channel carries after decorrelation */
        int pl = c;
        if (C == 2) pl = c == 0 ? (mode == 10 ? 2 : (mode == 9 ? 3 : 0)) : ((mode == 1 || mode == 9) ? 1 : 3);
        uint32_t o = 0;
        int32_t lo = 0x7fffffff, hi = (int32_t)0x80000000;
        for (int w = 0; w < nw; w++) { o |= s_or[w][pl]; lo = min(lo, s_mn[w][pl]); hi = max(hi, s_mx[w][pl]); }
        int wasted = 0;
        if (o) {
            wasted = __ffs((int)o) - 1;
            if (wasted >= cfg.bps - 1) wasted = 0;
        }
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
        s_wasted[c] = wasted;
    }
    if (tid == 0) ch_modes[f] = (uint8_t)mode;
    /* planes hold the samples the subframe codes: wasted bits shifted out (encode.c:586-590).
     * Rare, so a second walk over the few planes concerned (still in L2) instead of a shift in
     * every consumer. */
    if (fb_uses_planes(C)) {
        __syncthreads();
        for (int c = 0; c < C; c++) {
            const int w = s_wasted[c];
            if (w == 0) continue;                           /* CTA-uniform */
            int32_t *pl = planes + ibase + (size_t)c * (size_t)n;
            for (int i = tid; i < n; i += T) pl[i] >>= w;
        }
    }
}

#endif
