/*
 * k_search.cuh -- subframe type / predictor order / Rice parameter search, one
 * CTA per subframe (optimize.c:124-276, rice.c:30-187).
 *
 * A candidate predictor is costed exactly as the reference does: residual
 * (optimize.c:34-122), zig-zag sums of the finest partition level
 * (rice.c:76-95), the pairwise-summed pyramid (rice.c:96-102), one Rice
 * parameter per partition (rice.c:30-74) and the smallest total over the
 * allowed partition orders with ties going to the HIGHER order (rice.c:128-135).
 * Only the sums leave registers; the residual itself is stored once, for the
 * chosen predictor.
 */
#ifndef FLAKE_B200_K_SEARCH_CUH
#define FLAKE_B200_K_SEARCH_CUH

#include "dev_common.cuh"
#include "k_lpc.cuh"

#ifndef FB_SEARCH_THREADS
#define FB_SEARCH_THREADS 64     /* 16 samples per thread and tile; measured best of 32/64/128/256 on B200 */
#endif

struct FbSearchShared {
    unsigned long long sums[512];   /* level L lives at [(1<<L)-1, (1<<(L+1))-1) */
    uint8_t  kbuf[512];
    uint32_t lvl_bits[9];
    uint32_t lvl_rice2[9];
    int32_t  coef[FB_MAX_ORDER];
    int32_t  shift;
    uint32_t result;
    int32_t  best_porder;
    int32_t  best_method;
};

/* optimize.c:34-68, one sample */
__device__ __forceinline__ int32_t fb_fixed_residual(const int32_t *x, int i, int order)
{
    const long long a = x[i];
    switch (order) {
    case 0:  return (int32_t)a;
    case 1:  return (int32_t)(a - x[i - 1]);
    case 2:  return (int32_t)(a - 2LL * x[i - 1] + x[i - 2]);
    case 3:  return (int32_t)(a - 3LL * x[i - 1] + 3LL * x[i - 2] - x[i - 3]);
    default: return (int32_t)(a - 4LL * x[i - 1] + 6LL * x[i - 2] - 4LL * x[i - 3] + x[i - 4]);
    }
}

/* optimize.c:70-122, one sample */
__device__ __forceinline__ int32_t fb_lpc_residual(const int32_t *x, int i, int order,
                                                   const int32_t *coef, int shift)
{
    long long pred = 0;
    for (int j = 0; j < order; j++)
        pred += (long long)coef[j] * (long long)x[i - 1 - j];
    return (int32_t)((long long)x[i] - (pred >> shift));
}

/* ------------------------------------------------------------------ */
/* candidate costing                                                    */
/* ------------------------------------------------------------------ */
/* partition-order limits of a candidate (rice.c:148-171).  The finest-level sums and
 * the per-level accumulators are zero on entry: the kernel zeroes them once and
 * fb_eval_finish re-zeroes what it consumed. */
__device__ __forceinline__ void fb_eval_begin(FbSearchShared &S, int n, int order, int pmin_cfg,
                                              int pmax_cfg, int &pmin, int &pmax)
{
    pmin = fb_limit_porder(pmin_cfg, n, order);
    pmax = fb_limit_porder(pmax_cfg, n, order);
}

/* Partition sums of every allowed level straight from the finest level (the
 * pairwise pyramid of rice.c:96-102 is exact integer addition, so the order of
 * summation is free), per-partition Rice parameter (rice.c:47-74), best partition
 * order with ties to the higher one (rice.c:128-135), total (rice.c:157-171). */
__device__ __forceinline__ uint32_t fb_eval_finish(FbSearchShared &S, int n, int is_lpc, int order,
                                                   int obits, int pmin, int pmax, FbSub *store)
{
    const int tid = threadIdx.x, T = blockDim.x;
    const int nparts = 1 << pmax;
    const unsigned long long *fine = &S.sums[nparts - 1];
    __syncthreads();                                   /* finest sums complete */
    const int e_first = (1 << pmin) - 1, e_end = (1 << (pmax + 1)) - 1;
    for (int e = e_first + tid; e < e_end; e += T) {
        const int L = fb_ilog2((uint32_t)(e + 1));
        const int j = e - ((1 << L) - 1);
        const int span = 1 << (pmax - L);
        unsigned long long sum = 0;
        for (int q = 0; q < span; q++) sum += fine[j * span + q];
        const int cnt = (n >> L) - (j == 0 ? order : 0);
        const int k = fb_rice_k(sum, cnt);
        S.kbuf[e] = (uint8_t)k;
        atomicAdd(&S.lvl_bits[L], (uint32_t)fb_rice_count64(sum, cnt, k));
        if (k > 14) atomicOr(&S.lvl_rice2[L], 1u);
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t best = 0xffffffffu;
        int bl = pmin;
        for (int L = pmin; L <= pmax; L++) {
            const uint32_t b = S.lvl_bits[L] + 4u * (1u << L);
            if (b <= best) { best = b; bl = L; }
        }
        const uint32_t method = S.lvl_rice2[bl];
        uint32_t total = (uint32_t)(order * obits + 2);
        if (is_lpc) total += 4u + 5u + (uint32_t)order * 15u;
        total += best;
        total += method + 4u;
        S.result = total;
        S.best_porder = bl;
        S.best_method = (int32_t)method;
        if (store) { store->porder = bl; store->method = (int32_t)method; store->est_bits = total; }
    }
    __syncthreads();
    if (store) {
        const int np = 1 << S.best_porder;
        for (int j = tid; j < np; j += T) store->params[j] = S.kbuf[np - 1 + j];
    }
    const uint32_t result = S.result;
    /* leave the accumulators zeroed for the next candidate (its first write to them
     * comes after the barrier that follows the coefficient load) */
    for (int e = tid; e < nparts; e += T) S.sums[nparts - 1 + e] = 0;
    if (tid < 9) { S.lvl_bits[tid] = 0; S.lvl_rice2[tid] = 0; }
    __syncthreads();
    return result;
}

/*
 * Generic candidate evaluation: any block size, samples through a generic
 * pointer (global memory for blocks that do not fit shared memory).
 * Every thread of the CTA calls it and gets the total
 * (calc_rice_params_fixed / _lpc return value, rice.c:157-187).
 * res_out != NULL: also store the residual (warm-up = samples).
 * store   != NULL: also record method / porder / params for the packer.
 */
__device__ __noinline__ uint32_t fb_evaluate(FbSearchShared &S, const int32_t *x, int n, int is_lpc, int order,
                                int obits, int pmin_cfg, int pmax_cfg,
                                int32_t *res_out, FbSub *store)
{
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31;
    int pmin, pmax;
    fb_eval_begin(S, n, order, pmin_cfg, pmax_cfg, pmin, pmax);
    const int nparts = 1 << pmax, psize = n >> pmax;

    const int shift = S.shift;
    for (int base = tid - lane; base < n; base += T) {
        const int i = base + lane;
        unsigned long long u = 0;
        if (i < n) {
            int32_t r;
            if (i < order) r = x[i];
            else r = is_lpc ? fb_lpc_residual(x, i, order, S.coef, shift)
                            : fb_fixed_residual(x, i, order);
            if (res_out) res_out[i] = r;
            if (i >= order) u = fb_zigzag(r);
        }
        /* partition index, monotone across the warp; idle lanes carry u = 0 */
        int ic = i < order ? order : i;
        if (ic > n - 1) ic = n - 1;
        const int p = ic / psize;
        /* segmented inclusive sum over runs of equal p */
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long tu = __shfl_up_sync(FB_FULL_MASK, u, o);
            const int tp = __shfl_up_sync(FB_FULL_MASK, p, o);
            if (lane >= o && tp == p) u += tu;
        }
        const int pn = __shfl_down_sync(FB_FULL_MASK, p, 1);
        if ((lane == 31 || pn != p) && u)
            atomicAdd(&S.sums[nparts - 1 + p], u);
    }
    return fb_eval_finish(S, n, is_lpc, order, obits, pmin, pmax, store);
}

/* ------------------------------------------------------------------ */
/* fast path: block staged in shared memory in a skewed layout          */
/* ------------------------------------------------------------------ */
#define FB_RUN 16                 /* consecutive samples per thread and tile */
#define FB_HIST 32                /* zero samples in front of sample 0 */

/* logical index (sample i lives at logical i + FB_HIST) -> word offset; every
 * 16-sample run is followed by 4 pad words so that 128-bit loads of threads
 * whose runs are 16 samples apart hit distinct bank groups */
__host__ __device__ __forceinline__ int fb_skew(int logical) { return logical + ((logical >> 4) << 2); }
__host__ __device__ __forceinline__ int fb_skew_words(int n) { return (fb_skew(n + FB_HIST + FB_RUN) + 8 + 1) & ~1; }
/* staged plane + one 64-bit zig-zag sum per 16-sample run, in 32-bit words */
__host__ __device__ __forceinline__ int fb_search_smem_words(int n) { return fb_skew_words(n) + 2 * (((n + FB_RUN - 1) / FB_RUN) + 2); }

/*
 * Residual of the samples [i0, i0+16) with a register sliding window.
 * P = predictor order rounded up to a multiple of 4 (coefficients beyond the
 * order are zero).  WIDE: 64-bit prediction and sums (always exact);
 * !WIDE: 32-bit, used only when the caller proved nothing can overflow.
 */
template <int P, bool WIDE>
__device__ __forceinline__ void fb_run_residual(FbSearchShared &S, const int32_t *xs, int n, int order,
                                                int psize, int nparts, int tile_base,
                                                int32_t *res_out, unsigned long long *runsum)
{
    const int tid = threadIdx.x;
    const int i0 = tile_base + tid * FB_RUN;
    if (i0 >= n) return;
    int32_t c[P];
#pragma unroll
    for (int j = 0; j < P; j++) c[j] = S.coef[j];
    const int shift = S.shift;

    int32_t w[P + FB_RUN];
#pragma unroll
    for (int g = 0; g < (P + FB_RUN) / 4; g++) {
        const int4 v = *reinterpret_cast<const int4 *>(xs + fb_skew(i0 + FB_HIST - P + 4 * g));
        w[4 * g] = v.x; w[4 * g + 1] = v.y; w[4 * g + 2] = v.z; w[4 * g + 3] = v.w;
    }

    /* residuals of the run, branch free */
    int32_t r[FB_RUN];
#pragma unroll
    for (int k = 0; k < FB_RUN; k++) {
        int32_t rv;
        if (WIDE) {
            long long pred = 0;
#pragma unroll
            for (int j = 0; j < P; j++) pred += (long long)c[j] * (long long)w[P + k - 1 - j];
            rv = (int32_t)((long long)w[P + k] - (pred >> shift));
        } else {
            int32_t pred = 0;
#pragma unroll
            for (int j = 0; j < P; j++) pred += c[j] * w[P + k - 1 - j];
            rv = w[P + k] - (pred >> shift);
        }
        r[k] = (i0 + k < order) ? w[P + k] : rv;      /* warm-up samples pass through */
    }
    if (res_out) {
        int32_t *dst = res_out + i0;
        if (i0 + FB_RUN <= n && (((size_t)dst) & 15u) == 0) {
#pragma unroll
            for (int g = 0; g < FB_RUN / 4; g++)
                reinterpret_cast<int4 *>(dst)[g] = make_int4(r[4 * g], r[4 * g + 1], r[4 * g + 2], r[4 * g + 3]);
        } else {
#pragma unroll
            for (int k = 0; k < FB_RUN; k++) if (i0 + k < n) dst[k] = r[k];
        }
    }

    /* zig-zag sums (rice.c:76-95).  Partitions that are whole multiples of the run length
     * (the rule for every power-of-two block size): one sum per run, folded into the
     * partition sums after the tiles -- no atomics on the hot loop. */
    if (runsum) {
        unsigned long long acc;
        if (i0 >= order && i0 + FB_RUN <= n) {
            if (WIDE) {
                acc = 0;
#pragma unroll
                for (int k = 0; k < FB_RUN; k++) acc += fb_zigzag(r[k]);
            } else {
                uint32_t a32 = 0;
#pragma unroll
                for (int k = 0; k < FB_RUN; k++) a32 += fb_zigzag(r[k]);
                acc = a32;
            }
        } else {
            acc = 0;
#pragma unroll
            for (int k = 0; k < FB_RUN; k++)
                if (i0 + k >= order && i0 + k < n) acc += fb_zigzag(r[k]);
        }
        runsum[i0 >> 4] = acc;
        return;
    }
    /* general partition sizes: walk the run, flushing at partition boundaries */
    {
        const int istart = i0 < order ? order : i0;
        int pcur = istart / psize;
        int nb = (pcur + 1) * psize;
        unsigned long long acc = 0;
#pragma unroll 1
        for (int k = 0; k < FB_RUN; k++) {
            const int i = i0 + k;
            if (i < order || i >= n) continue;
            if (i >= nb) {
                if (acc) atomicAdd(&S.sums[nparts - 1 + pcur], acc);
                acc = 0; pcur++; nb += psize;
            }
            /* r[] lives in registers: select without dynamic indexing */
            int32_t rv = 0;
#pragma unroll
            for (int q = 0; q < FB_RUN; q++) rv = (q == k) ? r[q] : rv;
            acc += fb_zigzag(rv);
        }
        if (acc) atomicAdd(&S.sums[nparts - 1 + pcur], acc);
    }
}

template <int P, bool WIDE>
__device__ __noinline__ void fb_tiles(FbSearchShared &S, const int32_t *xs, int n, int order, int psize,
                                      int nparts, int32_t *res_out, unsigned long long *runsum)
{
    for (int tile = 0; tile < n; tile += (int)blockDim.x * FB_RUN)
        fb_run_residual<P, WIDE>(S, xs, n, order, psize, nparts, tile, res_out, runsum);
}

/*
 * Fast evaluation.  S.coef holds the coefficients zero-padded to 32 and S.shift
 * the shift (fixed predictors: binomial coefficients, shift 0).  `maxabs`
 * bounds |sample| and decides whether 32-bit arithmetic is provably exact:
 *   |pred| <= sum|c| * maxabs < 2^31, and
 *   |residual| <= maxabs + (sum|c|*maxabs >> shift) + 1 < 2^26 so that a run's
 *   zig-zag sum fits 32 bits.
 */
template <int MAXP>
__device__ uint32_t fb_evaluate_fast(FbSearchShared &S, const int32_t *xs, int n, int is_lpc, int order,
                                     int obits, int pmin_cfg, int pmax_cfg, uint32_t maxabs,
                                     int32_t *res_out, FbSub *store)
{
    int pmin, pmax;
    fb_eval_begin(S, n, order, pmin_cfg, pmax_cfg, pmin, pmax);
    const int nparts = 1 << pmax, psize = n >> pmax;
    unsigned long long sumabs = 0;
    for (int j = 0; j < order; j++) { const int32_t v = S.coef[j]; sumabs += (unsigned long long)(v < 0 ? -(long long)v : v); }
    const unsigned long long pm = sumabs * (unsigned long long)maxabs;
    const bool narrow = pm < 0x80000000ull &&
                        ((unsigned long long)maxabs + (pm >> S.shift) + 1ull) < (1ull << 26);
    const int P = (order + 3) & ~3;
    /* per-run sums are usable when every partition is a whole number of runs */
    unsigned long long *runsum = (psize % FB_RUN) == 0
        ? reinterpret_cast<unsigned long long *>(const_cast<int32_t *>(xs) + fb_skew_words(n)) : nullptr;
#define FB_CASE(PP)                                                                           \
    case PP:                                                                                  \
        if (narrow) fb_tiles<PP, false>(S, xs, n, order, psize, nparts, res_out, runsum);             \
        else        fb_tiles<PP, true>(S, xs, n, order, psize, nparts, res_out, runsum);              \
        break;
    switch (P) {
        case 0:
        FB_CASE(4) FB_CASE(8) FB_CASE(12)
        default:
            if (MAXP > 12) {
                switch (P) {
                    FB_CASE(16) FB_CASE(20) FB_CASE(24) FB_CASE(28)
                    default:
                        if (narrow) fb_tiles<32, false>(S, xs, n, order, psize, nparts, res_out, runsum);
                        else        fb_tiles<32, true>(S, xs, n, order, psize, nparts, res_out, runsum);
                        break;
                }
            }
            break;
    }
#undef FB_CASE
    if (runsum) {
        __syncthreads();
        const int per = psize / FB_RUN;
        for (int p = threadIdx.x; p < nparts; p += blockDim.x) {
            unsigned long long sum = 0;
            for (int q = 0; q < per; q++) sum += runsum[p * per + q];
            S.sums[nparts - 1 + p] = sum;
        }
    }
    return fb_eval_finish(S, n, is_lpc, order, obits, pmin, pmax, store);
}

/* one candidate through whichever path the block size allows */
#define FB_EVAL(is_lpc_, order_, res_, store_)                                                      \
    (fast ? fb_evaluate_fast<MAXP>(S, xs, n, (is_lpc_), (order_), obits, pmin, pmax, maxabs, (res_), (store_)) \
          : fb_evaluate(S, xg, n, (is_lpc_), (order_), obits, pmin, pmax, (res_), (store_)))

template <int MAXP>
__global__ void __launch_bounds__(FB_SEARCH_THREADS)
k_search(FbConfig cfg, const FbFrame *frames, const uint32_t *nframes, const int32_t *smp,
         int32_t *res, FbSub *subs, const int32_t *coefs, const int32_t *shifts, int smem_ints)
{
    FB_DYN_SMEM(dyn);
    __shared__ FbSearchShared S;

    const int C = cfg.channels;
    const uint32_t sf = blockIdx.x;
    const uint32_t f = sf / (uint32_t)C;
    const int c = (int)(sf % (uint32_t)C);
    if (f >= *nframes) return;
    const FbFrame fr = frames[f];
    const int n = (int)fr.n;
    FbSub *sb = &subs[sf];
    const int tid = threadIdx.x, T = blockDim.x;
    const size_t off = (size_t)fr.start * C + (size_t)c * n;
    const int32_t *xg = smp + off;
    int32_t *rg = res + off;
    const int obits = sb->obits;
    const uint32_t maxabs = sb->maxabs;

    /* CONSTANT, optimize.c:143-151 */
    if (sb->is_const) {
        if (tid == 0) { sb->type = 0; sb->order = 0; sb->est_bits = (uint32_t)obits; }
        return;
    }
    /* VERBATIM, optimize.c:154-158 */
    if (n < 5 || cfg.prediction_type == 0) {
        if (tid == 0) { sb->type = 1; sb->order = 0; sb->est_bits = (uint32_t)(obits * n); }
        return;
    }

    /* stage the plane: skewed layout, FB_HIST zero samples in front (fast path) */
    const bool fast = fb_search_smem_words(n) <= smem_ints;
    int32_t *xs = (int32_t *)dyn;
    if (fast) {
        for (int L = tid; L < FB_HIST; L += T) xs[fb_skew(L)] = 0;
        const int n4 = (((size_t)xg) & 15u) == 0 ? (n & ~3) : 0;
        /* 16-byte groups never straddle a skew pad: asynchronous copies straight into place */
        for (int i = 4 * tid; i < n4; i += 4 * T) fb_cp_async16(xs + fb_skew(i + FB_HIST), xg + i);
#pragma unroll 8
        for (int i = n4 + tid; i < n; i += T) xs[fb_skew(i + FB_HIST)] = xg[i];
        for (int i = n + tid; i < n + FB_RUN; i += T) xs[fb_skew(i + FB_HIST)] = 0;
        fb_cp_async_wait_all();
    }
    for (int e = tid; e < 512; e += T) S.sums[e] = 0;
    if (tid < 9) { S.lvl_bits[tid] = 0; S.lvl_rice2[tid] = 0; }
    if (tid < FB_MAX_ORDER) S.coef[tid] = 0;
    if (tid == 0) S.shift = 0;
    __syncthreads();

    const int pmin = cfg.min_porder, pmax = cfg.max_porder;
    int min_order = cfg.min_order, max_order = cfg.max_order;

    /* FIXED, optimize.c:168-190: the fixed predictors are LPC with binomial
     * coefficients and shift 0 (optimize.c:44-66) */
    if (cfg.prediction_type == 1 || n <= max_order) {
        if (max_order > 4) max_order = 4;
        int opt = min_order;
        uint32_t best = 0xffffffffu;
#define FB_SET_FIXED(order_)                                                                  \
        do {                                                                                  \
            __syncthreads();                                                                  \
            if (tid < 4) {                                                                    \
                const int o_ = (order_);                                                      \
                const int32_t bc[5][4] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {2, -1, 0, 0}, {3, -3, 1, 0}, {4, -6, 4, -1}}; \
                S.coef[tid] = bc[o_ > 4 ? 4 : o_][tid];                                       \
            }                                                                                 \
            __syncthreads();                                                                  \
        } while (0)
        for (int i = min_order; i <= max_order; i++) {
            FB_SET_FIXED(i);
            const uint32_t b = FB_EVAL(0, i, nullptr, nullptr);
            if (b < best) { best = b; opt = i; }
        }
        if (opt > 4) opt = 4;   /* min_order > 4 with a tiny last block: undefined in the reference */
        if (tid == 0) { sb->type = 8; sb->order = opt; }
        FB_SET_FIXED(opt);
        FB_EVAL(0, opt, rg, sb);
#undef FB_SET_FIXED
        return;
    }

    /* LPC, optimize.c:193-275 */
    const int32_t *co = coefs + (size_t)sf * FB_MAX_ORDER * FB_MAX_ORDER;
    const int32_t *so = shifts + (size_t)sf * FB_MAX_ORDER;
    const int om = cfg.order_method;
    int opt_order;                       /* 0-based index while searching */

#define FB_EVAL_INDEX(idx, out_bits)                                             \
    do {                                                                         \
        __syncthreads();                                                         \
        if (tid < FB_MAX_ORDER) S.coef[tid] = tid <= (idx) ? co[(idx) * FB_MAX_ORDER + tid] : 0; \
        if (tid == 0) S.shift = so[(idx)];                                       \
        __syncthreads();                                                         \
        (out_bits) = FB_EVAL(1, (idx) + 1, nullptr, nullptr);                    \
    } while (0)

    if (om == 0) {
        opt_order = max_order - 1;
    } else if (om == 1) {
        opt_order = sb->est_order - 1;
    } else if (om >= 2 && om <= 4) {
        const int levels = 1 << (om - 1);
        uint32_t best = 0xffffffffu;
        opt_order = max_order - 1;
        for (int i = levels - 1; i >= 0; i--) {
            int order = min_order + (((max_order - min_order + 1) * (i + 1)) / levels) - 2;
            if (order < 0) order = 0;
            uint32_t b;
            FB_EVAL_INDEX(order, b);
            if (b < best) { best = b; opt_order = order; }
        }
    } else if (om == 5) {
        uint32_t best = 0xffffffffu;
        opt_order = 0;
        for (int i = 0; i < max_order; i++) {
            uint32_t b;
            FB_EVAL_INDEX(i, b);
            if (b < best) { best = b; opt_order = i; }
        }
    } else {
        /* log search, optimize.c:241-261 */
        uint32_t best = 0xffffffffu, done = 0;
        opt_order = min_order - 1 + (max_order - min_order) / 3;
        for (int step = 16; step > 0; step >>= 1) {
            const int last = opt_order;
            for (int i = last - step; i <= last + step; i += step) {
                if (i < min_order - 1 || i >= max_order || ((done >> i) & 1u)) continue;
                uint32_t b;
                FB_EVAL_INDEX(i, b);
                done |= 1u << i;
                if (b < best) { best = b; opt_order = i; }
            }
        }
    }

    /* final pass for the chosen order, optimize.c:266-275 */
    {
        const int idx = opt_order;
        __syncthreads();
        if (tid < FB_MAX_ORDER) {
            const int32_t v = tid <= idx ? co[idx * FB_MAX_ORDER + tid] : 0;
            S.coef[tid] = v;
            sb->coefs[tid] = v;
        }
        if (tid == 0) {
            S.shift = so[idx];
            sb->type = 32; sb->order = idx + 1; sb->shift = so[idx];
        }
        __syncthreads();
        FB_EVAL(1, idx + 1, rg, sb);
    }
#undef FB_EVAL_INDEX
}
#undef FB_EVAL

#endif
