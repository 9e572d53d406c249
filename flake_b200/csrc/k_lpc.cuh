/*
 * k_lpc.cuh -- FP64 LPC analysis, one CTA per subframe (lpc.c):
 *   window (lpc.c:28-40) -> autocorrelation (lpc.c:46-71) ->
 *   Levinson-Durbin (lpc.c:77-117) or Schur order estimate (lpc.c:125-162) ->
 *   15-bit quantisation with error feedback (lpc.c:167-219).
 *
 * Bit-exactness rules: every FP64 operation is an
 * individually rounded IEEE op (__dmul_rn/__dadd_rn/..., never contracted to
 * FMA) and every sum is accumulated in the reference's order: one thread owns
 * one (lag, accumulator) chain and walks it sequentially; parallelism comes
 * from lags x accumulators x subframes, not from splitting a sum.
 */
#ifndef FLAKE_B200_K_LPC_CUH
#define FLAKE_B200_K_LPC_CUH

#include "dev_common.cuh"

#define FB_LPC_THREADS 96      /* launch bound; actual block = 32 * ceil(2*(lag+1)/32) */
#define FB_MAX_ORDER 32

/* x86-64 cvttsd2si semantics of the reference's `int q = double` */
__device__ __forceinline__ int32_t fb_trunc_to_int(double x)
{
    if (!(x > -2147483649.0 && x < 2147483648.0)) return (int32_t)0x80000000u;
    return __double2int_rz(x);
}

/* lpc.c:77-117; rows 0..max_order-1 of lpc[][] */
__device__ void fb_levinson(const double *autoc, int max_order, const double *refl,
                            double (*lpc)[FB_MAX_ORDER])
{
    double a[FB_MAX_ORDER];
    for (int i = 0; i < FB_MAX_ORDER; i++) a[i] = 0.0;
    double err = autoc ? autoc[0] : 1.0;
    for (int i = 0; i < max_order; i++) {
        double r;
        if (refl) {
            r = refl[i];
        } else {
            r = -autoc[i + 1];
            for (int j = 0; j < i; j++)
                r = __dsub_rn(r, __dmul_rn(a[j], autoc[i - j]));
            r = __ddiv_rn(r, err);
            err = __dmul_rn(err, __dsub_rn(1.0, __dmul_rn(r, r)));
        }
        a[i] = r;
        const int half = i >> 1;
        int j;
        for (j = 0; j < half; j++) {
            const double t = a[j];
            a[j] = __dadd_rn(a[j], __dmul_rn(r, a[i - 1 - j]));
            a[i - 1 - j] = __dadd_rn(a[i - 1 - j], __dmul_rn(r, t));
        }
        if (i & 1)
            a[j] = __dadd_rn(a[j], __dmul_rn(a[j], r));
        for (j = 0; j <= i; j++)
            lpc[i][j] = -a[j];
    }
}

/* lpc.c:125-162 */
__device__ int fb_schur_estimate(const double *autoc, int max_order, double (*lpc)[FB_MAX_ORDER])
{
    double g0[FB_MAX_ORDER], g1[FB_MAX_ORDER], refl[FB_MAX_ORDER];
    for (int i = 0; i < max_order; i++) g0[i] = g1[i] = autoc[i + 1];
    double e = autoc[0];
    refl[0] = __ddiv_rn(-g1[0], e);
    e = __dadd_rn(e, __dmul_rn(g1[0], refl[0]));
    for (int i = 1; i < max_order; i++) {
        for (int j = 0; j < max_order - i; j++) {
            const double g1n = g1[j + 1];
            g1[j] = __dadd_rn(g1n, __dmul_rn(refl[i - 1], g0[j]));
            g0[j] = __dadd_rn(__dmul_rn(g1n, refl[i - 1]), g0[j]);
        }
        refl[i] = __ddiv_rn(-g1[0], e);
        e = __dadd_rn(e, __dmul_rn(g1[0], refl[i]));
    }
    int est = 1;
    for (int i = max_order - 1; i >= 0; i--)
        if (fabs(refl[i]) > 0.10) { est = i + 1; break; }
    fb_levinson(nullptr, est, refl, lpc);
    return est;
}

/* lpc.c:167-219 with precision 15 (encode.c:443) */
__device__ void fb_quantize(double *in, int order, int32_t *out, int32_t *shift)
{
    const int32_t qmax = (1 << 14) - 1;
    double cmax = 0.0;
    for (int i = 0; i < order; i++) {
        const double d = fabs(in[i]);
        if (d > cmax) cmax = d;
    }
    if (__dmul_rn(cmax, 32768.0) < 1.0) {
        *shift = 0;
        for (int i = 0; i < order; i++) out[i] = 0;
        return;
    }
    int sh = 15;
    while (__dmul_rn(cmax, (double)(1 << sh)) > (double)qmax && sh > 0) sh--;
    if (sh == 0 && cmax > (double)qmax) {
        const double scale = __ddiv_rn((double)qmax, cmax);
        for (int i = 0; i < order; i++) in[i] = __dmul_rn(in[i], scale);
    }
    double err = 0.0;
    const double mul = (double)(1 << sh);
    for (int i = 0; i < order; i++) {
        err = __dadd_rn(err, __dmul_rn(in[i], mul));
        int32_t q = fb_trunc_to_int(__dadd_rn(err, 0.5));
        if (q <= -qmax) q = -qmax + 1;
        if (q > qmax) q = qmax;
        err = __dsub_rn(err, (double)q);
        out[i] = q;
    }
    *shift = sh;
}

#define FB_LPC_CHUNK 512          /* window samples produced per round */
#define FB_LPC_HIST  32           /* samples of the previous round kept in front (>= max lag) */
#define FB_LPC_BUF   (FB_LPC_HIST + FB_LPC_CHUNK)
#define FB_LPC_RING  (2 * FB_LPC_BUF)   /* two buffers, alternating */

/* data1[p] of lpc.c:46-56: windowed sample p, 0 at p == n.  The window value depends
 * only on min(p, n-1-p) (lpc.c:33-39), so it is recomputed per position instead of being
 * kept for the mirrored sample.  Odd n: the centre sample is uninitialised in the
 * reference; defined as 0.0 here (parity-exempt). */
__device__ __forceinline__ double fb_windowed(const int32_t *__restrict__ x, int p, int n, int half, double cc)
{
    if (p >= n || ((n & 1) && p == half)) return 0.0;
    const int i = p < half ? p : n - 1 - p;
    const double d = __dsub_rn(cc, (double)i);
    const double win = __dsub_rn(1.0, __dmul_rn(d, d));
    return __dmul_rn((double)x[p], win);
}

/*
 * coefs_out: [subframe][32][32] int32, shift_out: [subframe][32].
 * Block = 32 * ceil(2*(lag+1)/32) threads: thread ch owns chain (lag ch>>1, accumulator ch&1)
 * of lpc.c:57-68 -- one warp per subframe up to order 15.
 *
 * The chains of all lags advance in lockstep over the sample index, so only a sliding
 * window of the windowed signal is live: it is produced 512 samples at a time into a
 * 1024-entry shared ring (8 KB) straight from the int32 plane, whatever the block size.
 * Dynamic shared memory: ring[1024] doubles, then lpc[lag][32] doubles.
 */
__global__ void __launch_bounds__(FB_LPC_THREADS)
k_lpc(FbConfig cfg, const FbFrame *frames, const uint32_t *nframes, const int32_t *smp,
      FbSub *subs, int32_t *coefs_out, int32_t *shift_out)
{
    FB_DYN_SMEM(dyn);
    __shared__ double s_autoc[2 * (FB_MAX_ORDER + 1)];
    __shared__ int s_est;
    double *ring = reinterpret_cast<double *>(dyn);
    double (*s_lpc)[FB_MAX_ORDER] = reinterpret_cast<double (*)[FB_MAX_ORDER]>(ring + FB_LPC_RING);

    const int C = cfg.channels;
    const uint32_t sf = blockIdx.x;
    const uint32_t f = sf / (uint32_t)C;
    const int c = (int)(sf % (uint32_t)C);
    if (f >= *nframes) return;
    const FbFrame fr = frames[f];
    const int n = (int)fr.n;
    FbSub *sb = &subs[sf];
    const int lag = cfg.max_order;
    /* same gate as optimize.c:143-193: only the LPC branch needs coefficients */
    if (sb->is_const || n < 5 || cfg.prediction_type != 2 || n <= lag) return;

    const int32_t *__restrict__ x = smp + (size_t)fr.start * C + (size_t)c * n;
    const int tid = threadIdx.x, T = blockDim.x;
    const double cc = __dsub_rn(__ddiv_rn(2.0, __dsub_rn((double)n, 1.0)), 1.0);
    const int half = n >> 1;

    /* chain state */
    const bool active = tid < 2 * (lag + 1);
    const int ci = tid >> 1, ca = tid & 1;
    double s = 1.0;                               /* lpc.c:58-59: both accumulators start at 1.0 */
    int j = lag + 1 + ca;                         /* next tail term of this chain */
    const int last = n - 1;

    int which = 0;
    for (int base = 0; base <= n; base += FB_LPC_CHUNK, which ^= 1) {
        double *buf = ring + which * FB_LPC_BUF;              /* buf[FB_LPC_HIST + (p - base)] = data1[p] */
        const double *prev = ring + (which ^ 1) * FB_LPC_BUF;
        const int end = min(base + FB_LPC_CHUNK, n + 1);      /* positions [base, end) */
        /* produce: plane loads first (independent), then the FP64 window arithmetic */
        for (int p0 = base + tid; p0 < end; p0 += 8 * T) {
            int32_t xv[8];
#pragma unroll
            for (int q = 0; q < 8; q++) { const int p = p0 + q * T; xv[q] = p < n ? x[p] : 0; }
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int p = p0 + q * T;
                if (p < end) {
                    double v = 0.0;
                    if (p < n && !((n & 1) && p == half)) {
                        const int i = p < half ? p : n - 1 - p;
                        const double d = __dsub_rn(cc, (double)i);
                        v = __dmul_rn((double)xv[q], __dsub_rn(1.0, __dmul_rn(d, d)));
                    }
                    buf[FB_LPC_HIST + (p - base)] = v;
                }
            }
        }
        if (tid < FB_LPC_HIST)                                  /* carry the last 32 samples over */
            buf[tid] = base ? prev[FB_LPC_CHUNK + tid] : 0.0;
        __syncthreads();
        if (active) {
            const double *w = buf + FB_LPC_HIST - base;          /* w[p] = data1[p], p >= base - 32 */
            if (base == 0 && ca == 0)                            /* head terms, lpc.c:60-61 */
                for (int q = 0; q <= lag - ci; q++)
                    s = __dadd_rn(s, __dmul_rn(w[q + ci], w[q]));
            const int lim = min(end - 1, last);
            /* tail terms in order; the products do not depend on the running sum, so the
             * loads and multiplies of the next 8 terms are issued while the strictly
             * ordered add chain of the current 8 drains */
            if (j + 14 <= lim) {
                const double *pu = w + j, *pv = w + j - ci;
                double u0[8], v0[8];
#pragma unroll
                for (int q = 0; q < 8; q++) { u0[q] = pu[2 * q]; v0[q] = pv[2 * q]; }
                while (j + 30 <= lim) {
                    pu += 16; pv += 16;
                    double u1[8], v1[8];
#pragma unroll
                    for (int q = 0; q < 8; q++) { u1[q] = pu[2 * q]; v1[q] = pv[2 * q]; }
                    double pr[8];
#pragma unroll
                    for (int q = 0; q < 8; q++) pr[q] = __dmul_rn(u0[q], v0[q]);
#pragma unroll
                    for (int q = 0; q < 8; q++) s = __dadd_rn(s, pr[q]);
#pragma unroll
                    for (int q = 0; q < 8; q++) { u0[q] = u1[q]; v0[q] = v1[q]; }
                    j += 16;
                }
#pragma unroll
                for (int q = 0; q < 8; q++) s = __dadd_rn(s, __dmul_rn(u0[q], v0[q]));
                j += 16;
            }
            for (; j <= lim; j += 2)
                s = __dadd_rn(s, __dmul_rn(w[j], w[j - ci]));
        }
        __syncthreads();
    }
    if (active) s_autoc[tid] = s;
    __syncthreads();
    /* fold the two accumulators (autoc[i] = temp + temp2) in place */
    if (tid == 0)
        for (int i = 0; i <= lag; i++)
            s_autoc[i] = __dadd_rn(s_autoc[2 * i], s_autoc[2 * i + 1]);
    __syncthreads();

    const int om = cfg.order_method;
    if (tid == 0) {
        int est = lag;
        if (om == 1) est = fb_schur_estimate(s_autoc, lag, s_lpc);
        else fb_levinson(s_autoc, lag, nullptr, s_lpc);
        s_est = est;
        sb->est_order = est;
    }
    __syncthreads();

    int32_t *co = coefs_out + (size_t)sf * FB_MAX_ORDER * FB_MAX_ORDER;
    int32_t *so = shift_out + (size_t)sf * FB_MAX_ORDER;
    if (om == 0 || om == 1) {
        if (tid == 0) {
            const int i = s_est - 1;
            fb_quantize(s_lpc[i], i + 1, co + i * FB_MAX_ORDER, so + i);
        }
    } else {
        for (int i = tid; i < lag; i += T)
            fb_quantize(s_lpc[i], i + 1, co + i * FB_MAX_ORDER, so + i);
    }
}

#endif
