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

#define FB_SEARCH_THREADS 256

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

/*
 * Cost one candidate; every thread of the CTA calls it and gets the total
 * (calc_rice_params_fixed / _lpc return value, rice.c:157-187).
 * res_out != NULL: also store the residual (warm-up = samples).
 * store   != NULL: also record method / porder / params for the packer.
 */
__device__ uint32_t fb_evaluate(FbSearchShared &S, const int32_t *x, int n, int is_lpc, int order,
                                int obits, int pmin_cfg, int pmax_cfg,
                                int32_t *res_out, FbSub *store)
{
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31;
    const int pmin = fb_limit_porder(pmin_cfg, n, order);
    const int pmax = fb_limit_porder(pmax_cfg, n, order);
    const int nparts = 1 << pmax, psize = n >> pmax;

    for (int e = tid; e < nparts; e += T) S.sums[nparts - 1 + e] = 0;
    if (tid < 9) { S.lvl_bits[tid] = 0; S.lvl_rice2[tid] = 0; }
    __syncthreads();

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
    __syncthreads();

    for (int L = pmax - 1; L >= pmin; L--) {
        const int cnt = 1 << L;
        for (int j = tid; j < cnt; j += T)
            S.sums[cnt - 1 + j] = S.sums[2 * cnt - 1 + 2 * j] + S.sums[2 * cnt + 2 * j];
        __syncthreads();
    }

    const int e_first = (1 << pmin) - 1, e_end = (1 << (pmax + 1)) - 1;
    for (int e = e_first + tid; e < e_end; e += T) {
        const int L = fb_ilog2((uint32_t)(e + 1));
        const int j = e - ((1 << L) - 1);
        const int cnt = (n >> L) - (j == 0 ? order : 0);
        const unsigned long long sum = S.sums[e];
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
    return S.result;
}

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

    const int32_t *x = xg;
    if (n <= smem_ints) {
        int32_t *xs = (int32_t *)dyn;
        for (int i = tid; i < n; i += T) xs[i] = xg[i];
        x = xs;
    }
    if (tid == 0) S.shift = 0;
    __syncthreads();

    const int pmin = cfg.min_porder, pmax = cfg.max_porder;
    int min_order = cfg.min_order, max_order = cfg.max_order;

    /* FIXED, optimize.c:168-190 */
    if (cfg.prediction_type == 1 || n <= max_order) {
        if (max_order > 4) max_order = 4;
        int opt = min_order;
        uint32_t best = 0xffffffffu;
        for (int i = min_order; i <= max_order; i++) {
            const uint32_t b = fb_evaluate(S, x, n, 0, i, obits, pmin, pmax, nullptr, nullptr);
            if (b < best) { best = b; opt = i; }
        }
        if (opt > 4) opt = 4;   /* min_order > 4 with a tiny last block: undefined in the reference */
        if (tid == 0) { sb->type = 8; sb->order = opt; }
        fb_evaluate(S, x, n, 0, opt, obits, pmin, pmax, rg, sb);
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
        (out_bits) = fb_evaluate(S, x, n, 1, (idx) + 1, obits, pmin, pmax, nullptr, nullptr); \
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
        fb_evaluate(S, x, n, 1, idx + 1, obits, pmin, pmax, rg, sb);
    }
#undef FB_EVAL_INDEX
}

#endif
