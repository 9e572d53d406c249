/*
 * k_search.cuh -- subframe type / predictor order / Rice parameter search, one
 * CTA per subframe (optimize.c:124-276, rice.c:30-187).
 *
 * A candidate predictor is costed exactly as the reference does: residual
 * (optimize.c:34-122), zig-zag sums of the finest partition level
 * (rice.c:76-95), the pairwise-summed pyramid (rice.c:96-102), one Rice
 * parameter per partition (rice.c:30-74) and the smallest total over the
 * allowed partition orders with ties going to the HIGHER order (rice.c:128-135).
 *
 * Shape of one candidate evaluation:
 *   tiles     every thread owns 16-sample runs; the run and its history come from the
 *             staged plane with 128-bit shared loads, the residuals stay in registers,
 *             one zig-zag sum per run goes to shared memory;
 *   barrier
 *   finish    warp 0 alone folds the run sums into the partition pyramid with shuffles
 *             (a lane holds 1..8 partitions of the finest level, then pairs merge),
 *             picks parameter and partition order, posts the total;
 *   barrier
 * All candidate coefficient rows are staged once per subframe, so nothing else
 * separates two candidates.  The Rice parameters of the best candidate so far are
 * kept beside the search, which makes the last pass (optimize.c:266-275) a pure
 * residual store.
 */
#ifndef FLAKE_B200_K_SEARCH_CUH
#define FLAKE_B200_K_SEARCH_CUH

#include "dev_common.cuh"
#include "k_lpc.cuh"

#ifndef FB_SEARCH_THREADS
#define FB_SEARCH_THREADS 64     /* 16 samples per thread and tile; measured best of 32/64/128/256 on B200 */
#endif

template <int MAXP>
struct FbSearchShared {
    unsigned long long sums[256];   /* finest-level partition sums when runs do not tile the partitions */
    uint8_t  kbuf[512];             /* candidate: parameter of partition j at level L at [(1<<L)-1+j] */
    uint8_t  kbest[256];            /* best candidate so far: parameters at its partition order */
    int32_t  coef[MAXP][MAXP];      /* candidate rows (row = order-1), zero padded */
    int32_t  shift[MAXP];
    uint32_t sumabs[MAXP];          /* sum |coef| per row */
    uint32_t result;
    int32_t  porder, method;        /* of the candidate just finished */
    int32_t  best_porder, best_method;
    uint32_t best_bits;
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
/* finish: partition pyramid, Rice parameters, best partition order     */
/* ------------------------------------------------------------------ */
/*
 * Warp 0 only.  Finest-level sums come from `runsum` (one per 16-sample run, `per` runs per
 * partition) or, when runsum == NULL, from S.sums.  The pairwise pyramid of rice.c:96-102 is
 * exact integer addition, so the order of summation is free.  Levels are visited from the
 * finest down and a level replaces the best only when strictly smaller, which is the
 * reference's ascending scan with `<=` (ties to the higher order, rice.c:128-135).
 * Posts S.result (rice.c:157-187 total), S.porder, S.method and the parameters in S.kbuf.
 */
template <int MAXP>
__device__ __forceinline__ void fb_finish_warp0(FbSearchShared<MAXP> &S, const unsigned long long *runsum,
                                                int per, int n, int is_lpc, int order, int obits,
                                                int pmin, int pmax)
{
    const int lane = threadIdx.x & 31;
    unsigned long long s[8];
#pragma unroll
    for (int v = 0; v < 8; v++) s[v] = 0;
    if (pmax >= 5) {
        const int V = 1 << (pmax - 5);
#pragma unroll
        for (int v = 0; v < 8; v++) {
            if (v < V) {
                const int p = lane * V + v;
                if (runsum) {
                    unsigned long long a = 0;
                    for (int q = 0; q < per; q++) a += runsum[p * per + q];
                    s[v] = a;
                } else {
                    s[v] = S.sums[p];
                }
            }
        }
    } else {
        const int g = 32 >> pmax, j = lane / g, sub = lane % g;
        unsigned long long a = 0;
        if (runsum) {
            for (int q = sub; q < per; q += g) a += runsum[j * per + q];
            for (int o = 1; o < g; o <<= 1) a += __shfl_xor_sync(FB_FULL_MASK, a, o);
        } else {
            a = S.sums[j];
        }
        s[0] = a;
    }

    uint32_t best = 0xffffffffu;
    int bl = pmin, bmethod = 0;
    for (int L = pmax; L >= pmin; L--) {
        uint32_t bits = 0;
        int flag = 0;
        if (L >= 5) {
            const int Vc = 1 << (L - 5);
#pragma unroll
            for (int v = 0; v < 8; v++) {
                if (v < Vc) {
                    const int j = lane * Vc + v;
                    const int cnt = (n >> L) - (j == 0 ? order : 0);
                    const int k = fb_rice_k(s[v], cnt);
                    S.kbuf[(1 << L) - 1 + j] = (uint8_t)k;
                    bits += (uint32_t)fb_rice_count64(s[v], cnt, k);
                    flag |= (k > 14);
                }
            }
        } else {
            const int g = 32 >> L, j = lane / g;
            const int cnt = (n >> L) - (j == 0 ? order : 0);
            const int k = fb_rice_k(s[0], cnt);
            if ((lane % g) == 0) {
                S.kbuf[(1 << L) - 1 + j] = (uint8_t)k;
                bits = (uint32_t)fb_rice_count64(s[0], cnt, k);
                flag = (k > 14);
            }
        }
        const uint32_t lvl = __reduce_add_sync(FB_FULL_MASK, bits);
        const int r2 = __any_sync(FB_FULL_MASK, flag) ? 1 : 0;
        const uint32_t b = lvl + 4u * (1u << L);
        if (b < best) { best = b; bl = L; bmethod = r2; }
        if (L > pmin) {
            if (L > 5) {
                const int Vh = 1 << (L - 6);
#pragma unroll
                for (int v = 0; v < 4; v++)
                    if (v < Vh) s[v] = s[2 * v] + s[2 * v + 1];
            } else {
                s[0] += __shfl_xor_sync(FB_FULL_MASK, s[0], 1 << (5 - L));
            }
        }
    }
    if (lane == 0) {
        uint32_t total = (uint32_t)(order * obits + 2);
        if (is_lpc) total += 4u + 5u + (uint32_t)order * 15u;
        total += best;
        total += (uint32_t)bmethod + 4u;
        S.result = total;
        S.porder = bl;
        S.method = bmethod;
    }
}

/* the candidate just finished is the best so far: warp 0 keeps its parameters */
template <int MAXP>
__device__ __forceinline__ void fb_keep_best(FbSearchShared<MAXP> &S, uint32_t bits)
{
    if (threadIdx.x < 32) {
        __syncwarp();
        const int bl = S.porder, np = 1 << bl;
        for (int j = threadIdx.x; j < np; j += 32) S.kbest[j] = S.kbuf[np - 1 + j];
        if (threadIdx.x == 0) { S.best_porder = bl; S.best_method = S.method; S.best_bits = bits; }
        __syncwarp();
    }
}

/* every thread: write the kept decision to the subframe record */
template <int MAXP>
__device__ __forceinline__ void fb_store_best(FbSearchShared<MAXP> &S, FbSub *sb)
{
    __syncthreads();
    const int np = 1 << S.best_porder;
    for (int j = threadIdx.x; j < np; j += blockDim.x) sb->params[j] = S.kbest[j];
    if (threadIdx.x == 0) { sb->porder = S.best_porder; sb->method = S.best_method; sb->est_bits = S.best_bits; }
}

/*
 * Generic candidate evaluation: any block size, samples through a generic
 * pointer (global memory for blocks that do not fit shared memory).
 * Every thread of the CTA calls it and gets the total
 * (calc_rice_params_fixed / _lpc return value, rice.c:157-187).
 * want_sums: cost the candidate; res_out != NULL: store the residual (warm-up = samples).
 */
template <int MAXP>
__device__ __noinline__ uint32_t fb_evaluate(FbSearchShared<MAXP> &S, const int32_t *x, int n, int is_lpc, int order,
                                             int row, int obits, int pmin_cfg, int pmax_cfg,
                                             int32_t *res_out, bool want_sums)
{
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31;
    const int pmin = fb_limit_porder(pmin_cfg, n, order);
    const int pmax = fb_limit_porder(pmax_cfg, n, order);
    const int nparts = 1 << pmax, psize = n >> pmax;
    const int32_t *coef = S.coef[row];
    const int shift = S.shift[row];
    if (want_sums) {
        for (int e = tid; e < nparts; e += T) S.sums[e] = 0;
        __syncthreads();
    }
    for (int base = tid - lane; base < n; base += T) {
        const int i = base + lane;
        unsigned long long u = 0;
        if (i < n) {
            int32_t r;
            if (i < order) r = x[i];
            else r = is_lpc ? fb_lpc_residual(x, i, order, coef, shift)
                            : fb_fixed_residual(x, i, order);
            if (res_out) res_out[i] = r;
            if (i >= order) u = fb_zigzag(r);
        }
        if (!want_sums) continue;
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
            atomicAdd(&S.sums[p], u);
    }
    if (!want_sums) return 0;
    __syncthreads();
    if (tid < 32) fb_finish_warp0<MAXP>(S, nullptr, 0, n, is_lpc, order, obits, pmin, pmax);
    __syncthreads();
    return S.result;
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
/* word offset of logical (16 m + d) relative to that of logical 16 m; d may be negative */
__host__ __device__ constexpr int fb_skew_delta(int d) { return d + 4 * (d >= 0 ? d / 16 : -((-d + 15) / 16)); }

/*
 * Residual of the samples [i0, i0+16) with a register sliding window.
 * P = predictor order rounded up to a multiple of 4 (coefficients beyond the
 * order are zero).  WIDE: 64-bit prediction and sums (always exact);
 * !WIDE: 32-bit, used only when the caller proved nothing can overflow.
 */
template <int MAXP, int P, bool WIDE>
__device__ __forceinline__ void fb_run_residual(FbSearchShared<MAXP> &S, const int32_t *xs, int n, int order,
                                                int row, int psize, int tile_base,
                                                int32_t *res_out, unsigned long long *runsum, bool want_sums)
{
    const int tid = threadIdx.x;
    const int i0 = tile_base + tid * FB_RUN;
    if (i0 >= n) return;
    int32_t c[P];
#pragma unroll
    for (int g = 0; g < P / 4; g++) {
        const int4 v = *reinterpret_cast<const int4 *>(&S.coef[row][4 * g]);
        c[4 * g] = v.x; c[4 * g + 1] = v.y; c[4 * g + 2] = v.z; c[4 * g + 3] = v.w;
    }
    const int shift = S.shift[row];

    int32_t w[P + FB_RUN];
    const int32_t *xr = xs + fb_skew(i0 + FB_HIST);           /* i0 + FB_HIST is a multiple of 16 */
#pragma unroll
    for (int g = 0; g < (P + FB_RUN) / 4; g++) {
        const int4 v = *reinterpret_cast<const int4 *>(xr + fb_skew_delta(4 * g - P));
        w[4 * g] = v.x; w[4 * g + 1] = v.y; w[4 * g + 2] = v.z; w[4 * g + 3] = v.w;
    }

    /* residuals of the run, branch free */
    int32_t r[FB_RUN];
#pragma unroll
    for (int k = 0; k < FB_RUN; k++) {
        if (WIDE) {
            long long pred = 0;
#pragma unroll
            for (int j = 0; j < P; j++) pred += (long long)c[j] * (long long)w[P + k - 1 - j];
            r[k] = (int32_t)((long long)w[P + k] - (pred >> shift));
        } else {
            int32_t pred = 0;
#pragma unroll
            for (int j = 0; j < P; j++) pred += c[j] * w[P + k - 1 - j];
            r[k] = w[P + k] - (pred >> shift);
        }
    }
    if (i0 < order) {                                         /* warm-up samples pass through */
#pragma unroll
        for (int k = 0; k < FB_RUN; k++) if (i0 + k < order) r[k] = w[P + k];
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
    if (!want_sums) return;

    /* zig-zag sums (rice.c:76-95).  Partitions that are whole multiples of the run length
     * (the rule for every power-of-two block size): one sum per run, folded into the
     * partition sums by the finishing warp -- no atomics on the hot loop. */
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
                if (acc) atomicAdd(&S.sums[pcur], acc);
                acc = 0; pcur++; nb += psize;
            }
            /* r[] lives in registers: select without dynamic indexing */
            int32_t rv = 0;
#pragma unroll
            for (int q = 0; q < FB_RUN; q++) rv = (q == k) ? r[q] : rv;
            acc += fb_zigzag(rv);
        }
        if (acc) atomicAdd(&S.sums[pcur], acc);
    }
}

template <int MAXP, int P, bool WIDE>
__device__ __noinline__ void fb_tiles(FbSearchShared<MAXP> &S, const int32_t *xs, int n, int order, int row, int psize,
                                      int32_t *res_out, unsigned long long *runsum, bool want_sums)
{
    for (int tile = 0; tile < n; tile += (int)blockDim.x * FB_RUN)
        fb_run_residual<MAXP, P, WIDE>(S, xs, n, order, row, psize, tile, res_out, runsum, want_sums);
}

/*
 * Fast evaluation of the candidate in row `row` of S.coef (fixed predictors: binomial
 * coefficients, shift 0).  `maxabs` bounds |sample| and decides whether 32-bit
 * arithmetic is provably exact:
 *   |pred| <= sum|c| * maxabs < 2^31, and
 *   |residual| <= maxabs + (sum|c|*maxabs >> shift) + 1 < 2^26 so that a run's
 *   zig-zag sum fits 32 bits.
 */
template <int MAXP>
__device__ __noinline__ uint32_t fb_evaluate_fast(FbSearchShared<MAXP> &S, const int32_t *xs, int n, int is_lpc, int order,
                                                  int row, int obits, int pmin_cfg, int pmax_cfg, uint32_t maxabs,
                                                  int32_t *res_out, bool want_sums)
{
    const int pmin = fb_limit_porder(pmin_cfg, n, order);
    const int pmax = fb_limit_porder(pmax_cfg, n, order);
    const int nparts = 1 << pmax, psize = n >> pmax;
    const unsigned long long pm = (unsigned long long)S.sumabs[row] * (unsigned long long)maxabs;
    const bool narrow = pm < 0x80000000ull &&
                        ((unsigned long long)maxabs + (pm >> S.shift[row]) + 1ull) < (1ull << 26);
    const int P = (order + 3) & ~3;
    /* per-run sums are usable when every partition is a whole number of runs */
    unsigned long long *runsum = (psize % FB_RUN) == 0
        ? reinterpret_cast<unsigned long long *>(const_cast<int32_t *>(xs) + fb_skew_words(n)) : nullptr;
    if (want_sums && !runsum) {
        for (int e = threadIdx.x; e < nparts; e += blockDim.x) S.sums[e] = 0;
        __syncthreads();
    }
#define FB_CASE(PP)                                                                                         \
    case PP:                                                                                                \
        if (narrow) fb_tiles<MAXP, PP, false>(S, xs, n, order, row, psize, res_out, runsum, want_sums);     \
        else        fb_tiles<MAXP, PP, true>(S, xs, n, order, row, psize, res_out, runsum, want_sums);      \
        break;
    switch (P) {
        case 0:
        FB_CASE(4) FB_CASE(8) FB_CASE(12)
        default:
            if constexpr (MAXP > 12) {
                switch (P) {
                    FB_CASE(16) FB_CASE(20) FB_CASE(24) FB_CASE(28)
                    default:
                        if (narrow) fb_tiles<MAXP, 32, false>(S, xs, n, order, row, psize, res_out, runsum, want_sums);
                        else        fb_tiles<MAXP, 32, true>(S, xs, n, order, row, psize, res_out, runsum, want_sums);
                        break;
                }
            }
            break;
    }
#undef FB_CASE
    if (!want_sums) return 0;
    __syncthreads();
    if (threadIdx.x < 32) fb_finish_warp0<MAXP>(S, runsum, psize / FB_RUN, n, is_lpc, order, obits, pmin, pmax);
    __syncthreads();
    return S.result;
}

/* one candidate through whichever path the block size allows */
#define FB_EVAL(is_lpc_, order_, row_, res_, sums_)                                                           \
    (fast ? fb_evaluate_fast<MAXP>(S, xs, n, (is_lpc_), (order_), (row_), obits, pmin, pmax, maxabs, (res_), (sums_)) \
          : fb_evaluate<MAXP>(S, xg, n, (is_lpc_), (order_), (row_), obits, pmin, pmax, (res_), (sums_)))

template <int MAXP>
__global__ void __launch_bounds__(FB_SEARCH_THREADS)
k_search(FbConfig cfg, const FbFrame *frames, const uint32_t *nframes, const int32_t *smp,
         int32_t *res, FbSub *subs, const int32_t *coefs, const int32_t *shifts, int smem_ints)
{
    FB_DYN_SMEM(dyn);
    __shared__ __align__(16) FbSearchShared<MAXP> S;

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
    }

    const int pmin = cfg.min_porder, pmax = cfg.max_porder;
    int min_order = cfg.min_order, max_order = cfg.max_order;
    const bool fixed = (cfg.prediction_type == 1 || n <= max_order);

    /* candidate rows: binomial coefficients (optimize.c:44-66 is LPC with shift 0) or the
     * quantised rows of k_lpc; sum |c| per row for the 32-bit exactness test */
    if (fixed) {
        if (tid < 5) {
            const int32_t bc[5][4] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {2, -1, 0, 0}, {3, -3, 1, 0}, {4, -6, 4, -1}};
            const uint32_t sa[5] = {0, 1, 3, 7, 15};
#pragma unroll
            for (int j = 0; j < 4; j++) S.coef[tid][j] = bc[tid][j];
            S.shift[tid] = 0;
            S.sumabs[tid] = sa[tid];
        }
    } else {
        const int32_t *co = coefs + (size_t)sf * FB_MAX_ORDER * FB_MAX_ORDER;
        const int32_t *so = shifts + (size_t)sf * FB_MAX_ORDER;
        for (int e = tid; e < MAXP * MAXP; e += T) {
            const int rowi = e / MAXP, j = e % MAXP;
            S.coef[rowi][j] = (rowi < max_order && j <= rowi) ? co[rowi * FB_MAX_ORDER + j] : 0;
        }
        for (int rowi = tid; rowi < MAXP; rowi += T) {
            uint32_t sa = 0;
            if (rowi < max_order)
                for (int j = 0; j <= rowi; j++) {
                    const int32_t v = co[rowi * FB_MAX_ORDER + j];
                    sa += (uint32_t)(v < 0 ? -v : v);                   /* <= 32 * 16383 */
                }
            S.sumabs[rowi] = sa;
            S.shift[rowi] = rowi < max_order ? so[rowi] : 0;
        }
    }
    if (fast) fb_cp_async_wait_all();
    __syncthreads();

    /* FIXED, optimize.c:168-190: row = order (row 0 is the all-zero order-0 predictor) */
    if (fixed) {
        if (max_order > 4) max_order = 4;
        int opt = min_order;
        uint32_t best = 0xffffffffu;
        for (int i = min_order; i <= max_order; i++) {
            const uint32_t b = FB_EVAL(0, i, i, nullptr, true);
            if (b < best) { best = b; opt = i; fb_keep_best<MAXP>(S, b); }
        }
        if (opt > 4 || best == 0xffffffffu) {   /* min_order > 4 with a tiny last block: undefined in the reference */
            if (opt > 4) opt = 4;
            const uint32_t b = FB_EVAL(0, opt, opt, nullptr, true);
            fb_keep_best<MAXP>(S, b);
        }
        if (tid == 0) { sb->type = 8; sb->order = opt; }
        FB_EVAL(0, opt, opt, rg, false);
        fb_store_best<MAXP>(S, sb);
        return;
    }

    /* LPC, optimize.c:193-275 */
    const int om = cfg.order_method;
    int opt_order;                       /* 0-based index while searching */
    bool have_best = false;

#define FB_EVAL_INDEX(idx, out_bits) (out_bits) = FB_EVAL(1, (idx) + 1, (idx), nullptr, true)

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
            if (b < best) { best = b; opt_order = order; have_best = true; fb_keep_best<MAXP>(S, b); }
        }
    } else if (om == 5) {
        uint32_t best = 0xffffffffu;
        opt_order = 0;
        for (int i = 0; i < max_order; i++) {
            uint32_t b;
            FB_EVAL_INDEX(i, b);
            if (b < best) { best = b; opt_order = i; have_best = true; fb_keep_best<MAXP>(S, b); }
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
                if (b < best) { best = b; opt_order = i; have_best = true; fb_keep_best<MAXP>(S, b); }
            }
        }
    }
#undef FB_EVAL_INDEX

    /* final pass for the chosen order, optimize.c:266-275: the costing of a searched order is
     * already known, only its residual is missing */
    {
        const int idx = opt_order;
        if (tid < FB_MAX_ORDER) sb->coefs[tid] = (tid <= idx && tid < MAXP) ? S.coef[idx][tid] : 0;
        if (tid == 0) { sb->type = 32; sb->order = idx + 1; sb->shift = S.shift[idx]; }
        if (have_best) {
            FB_EVAL(1, idx + 1, idx, rg, false);
        } else {
            const uint32_t b = FB_EVAL(1, idx + 1, idx, rg, true);
            fb_keep_best<MAXP>(S, b);
        }
        fb_store_best<MAXP>(S, sb);
    }
}
#undef FB_EVAL

#endif
