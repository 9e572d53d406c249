/*
 * k_lpc.cuh -- FP64 LPC analysis (lpc.c):
 *   window (lpc.c:28-40) -> autocorrelation (lpc.c:46-71) ->
 *   Levinson-Durbin (lpc.c:77-117) or Schur order estimate (lpc.c:125-162) ->
 *   15-bit quantisation with error feedback (lpc.c:167-219).
 *
 * Bit-exactness rules: every FP64 operation is an
 * individually rounded IEEE op (__dmul_rn/__dadd_rn/..., never contracted to
 * FMA) and every sum is accumulated in the reference's order.  The reference
 * keeps TWO accumulators per lag (temp over the odd tail terms, temp2 over the
 * even ones, lpc.c:57-68), so a subframe offers exactly two independent
 * strictly-ordered walks over the samples, each carrying lag+1 sums.
 *
 * Work decomposition: a LANE PAIR per subframe.  Lane `ca` of the pair owns
 * accumulator `ca` of every lag: lag+1 independent add chains per thread (the
 * instruction-level parallelism that hides the FP64 latency), fed from a
 * register ring of the last lag+1 windowed samples.  Per step a lane windows
 * ONE sample (its own position) and receives its partner's through one
 * shuffle -- the windowed signal never touches memory.  The int32 planes are
 * staged through shared memory a tile at a time by the whole warp (coalesced
 * row loads, conflict-free row stride), so a warp = 16 subframes in lockstep.
 * Shared-memory traffic per sample-step is one 32-bit load per lane, against
 * the 2 x 64-bit loads per multiply-add of a thread-per-chain layout.
 */
#ifndef FLAKE_B200_K_LPC_CUH
#define FLAKE_B200_K_LPC_CUH

#include "dev_common.cuh"

#define FB_MAX_ORDER 32
#ifndef FB_LPC_WARPS
#define FB_LPC_WARPS 1                                  /* warps per CTA: 17 resident warps per SM at 118 registers (2: 16) */
#endif
#define FB_LPC_THREADS (32 * FB_LPC_WARPS)
#define FB_LPC_SUBS_PER_CTA (16 * FB_LPC_WARPS)         /* a warp owns 16 subframes */
#define FB_LPC_ROW 34                                    /* staged row stride in words, == 2 (mod 32) */

/* x86-64 cvttsd2si semantics of the reference's `int q = double` */
__device__ __forceinline__ int32_t fb_trunc_to_int(double x)
{
    if (!(x > -2147483649.0 && x < 2147483648.0)) return (int32_t)0x80000000u;
    return __double2int_rz(x);
}

/* geometry of the register ring for lags 0..ML (ML even) */
template <int ML> struct FbLpcGeom {
    static constexpr int K = ML + 2;                     /* ring entries: d[q-ML-1 .. q] */
    static constexpr int U = K / 2;                      /* steps until the ring indices repeat */
    static constexpr int R = (2 * U >= 32) ? 1 : 32 / (2 * U);
    static constexpr int TL = 2 * U * R;                 /* sample positions per staged tile (<= 34) */
    static constexpr int LD = (TL + 31) / 32;            /* 32-lane loads per staged row */
};

/* w_i of lpc.c:33-39 for i = min(p, n-1-p) */
__device__ __forceinline__ double fb_welch(int i, double cc)
{
    const double d = __dsub_rn(cc, (double)i);
    return __dsub_rn(1.0, __dmul_rn(d, d));
}

/* data1[p] of lpc.c:46-56 for a sample value xv at position p.  The caller passes xv == 0
 * for p >= n and for the centre of an odd block (uninitialised in the reference, defined as
 * zero here, parity-exempt), which makes the product +-0.0: adding it never changes a sum
 * that started at 1.0.  (A per-block-size table of the window was measured: the loads cost
 * more than the three FP64 operations they save, 0.92 vs 0.85 ms per C2 stream.) */
__device__ __forceinline__ double fb_windowed(int32_t xv, int p, int n, double cc)
{
    return __dmul_rn((double)xv, fb_welch(min(p, n - 1 - p), cc));
}

/* lpc.c:167-219 with precision 15 (encode.c:443).  `in` is modified like the reference's
 * row (the rescale branch).  `order` may differ per lane. */
template <int ML>
__device__ __forceinline__ void fb_quantize(double (&in)[ML], int order, int32_t (&out)[ML], int32_t &shift)
{
    const int32_t qmax = (1 << 14) - 1;
    double cmax = 0.0;
#pragma unroll
    for (int i = 0; i < ML; i++)
        if (i < order) { const double d = fabs(in[i]); if (d > cmax) cmax = d; }
    if (__dmul_rn(cmax, 32768.0) < 1.0) {
        shift = 0;
#pragma unroll
        for (int i = 0; i < ML; i++) out[i] = 0;
        return;
    }
    int sh = 15;
    while (__dmul_rn(cmax, (double)(1 << sh)) > (double)qmax && sh > 0) sh--;
    if (sh == 0 && cmax > (double)qmax) {
        const double scale = __ddiv_rn((double)qmax, cmax);
#pragma unroll
        for (int i = 0; i < ML; i++) if (i < order) in[i] = __dmul_rn(in[i], scale);
    }
    double err = 0.0;
    const double mul = (double)(1 << sh);
#pragma unroll
    for (int i = 0; i < ML; i++) {
        out[i] = 0;
        if (i < order) {
            err = __dadd_rn(err, __dmul_rn(in[i], mul));
            int32_t q = fb_trunc_to_int(__dadd_rn(err, 0.5));
            if (q <= -qmax) q = -qmax + 1;
            if (q > qmax) q = qmax;
            err = __dsub_rn(err, (double)q);
            out[i] = q;
        }
    }
    shift = sh;
}

/* one Levinson step (lpc.c:92-111): order index i, reflection/prediction coefficient r */
template <int ML>
__device__ __forceinline__ void fb_levinson_update(double (&a)[ML], int i, double r)
{
    const int half = i >> 1;
    a[i] = r;
#pragma unroll
    for (int j = 0; j < ML / 2; j++) {
        if (j < half) {
            const double t = a[j];
            a[j] = __dadd_rn(a[j], __dmul_rn(r, a[i - 1 - j]));
            a[i - 1 - j] = __dadd_rn(a[i - 1 - j], __dmul_rn(r, t));
        }
    }
    if (i & 1) a[half] = __dadd_rn(a[half], __dmul_rn(a[half], r));
}

/* A row's shift (0 .. 15) travels with the sum of its |coefficients| (< 2^19, bits 8 and up): k_search
 * needs both for its 32-bit exactness tests and would otherwise add up every row of every subframe behind
 * a barrier of its own. */
template <int ML>
__device__ __forceinline__ void fb_store_row(int32_t *co, int32_t *so, int rowi, const int32_t (&q)[ML], int32_t sh)
{
    uint32_t sa = 0;
#pragma unroll
    for (int j = 0; j < ML; j++)
        if (j <= rowi) { co[rowi * FB_MAX_ORDER + j] = q[j]; sa += (uint32_t)(q[j] < 0 ? -q[j] : q[j]); }
    so[rowi] = (int32_t)((uint32_t)sh | (sa << 8));
}

/*
 * From the autocorrelation to the quantised coefficient rows: Levinson-Durbin (lpc.c:77-117) or the
 * Schur order estimate (lpc.c:125-156), the 15-bit quantiser (lpc.c:167-219).  Every calling lane
 * computes the same recursion; `writer` stores the rows.  only_row >= 0 (latency kernel): this
 * lane quantises and stores row only_row alone, so that the rows of one subframe are quantised
 * side by side instead of one after the other.
 */
template <int ML>
__device__ __forceinline__ void fb_lpc_rows(const FbConfig &cfg, const double (&autoc)[ML + 1], int lag, uint32_t sf,
                                            bool writer, FbSub *sb, int32_t *coefs_out, int32_t *shift_out,
                                            int only_row = -1)
{
    int32_t *co = coefs_out + (size_t)sf * FB_MAX_ORDER * FB_MAX_ORDER;
    int32_t *so = shift_out + (size_t)sf * FB_MAX_ORDER;
    const int om = cfg.order_method;
    double a[ML];
#pragma unroll
    for (int i = 0; i < ML; i++) a[i] = 0.0;

    if (om == 1) {
        /* Schur recursion and order estimate, lpc.c:125-154 */
        double g0[ML], g1[ML], refl[ML];
#pragma unroll
        for (int i = 0; i < ML; i++) { g0[i] = g1[i] = autoc[i + 1]; refl[i] = 0.0; }
        double e = autoc[0];
        refl[0] = __ddiv_rn(-g1[0], e);
        e = __dadd_rn(e, __dmul_rn(g1[0], refl[0]));
#pragma unroll
        for (int i = 1; i < ML; i++) {
            if (i < lag) {
#pragma unroll
                for (int j = 0; j < ML - i; j++) {
                    if (j < lag - i) {
                        const double g1n = g1[j + 1];
                        g1[j] = __dadd_rn(g1n, __dmul_rn(refl[i - 1], g0[j]));
                        g0[j] = __dadd_rn(__dmul_rn(g1n, refl[i - 1]), g0[j]);
                    }
                }
                refl[i] = __ddiv_rn(-g1[0], e);
                e = __dadd_rn(e, __dmul_rn(g1[0], refl[i]));
            }
        }
        int est = 1;
#pragma unroll
        for (int i = 0; i < ML; i++)
            if (i < lag && fabs(refl[i]) > 0.10) est = i + 1;           /* highest such i */
        /* Levinson from the reflection coefficients up to the estimate, lpc.c:156 */
#pragma unroll
        for (int i = 0; i < ML; i++)
            if (i < est) fb_levinson_update<ML>(a, i, refl[i]);
        double rowv[ML];
        int32_t q[ML], sh;
#pragma unroll
        for (int j = 0; j < ML; j++) rowv[j] = -a[j];
        fb_quantize<ML>(rowv, est, q, sh);
        if (writer) {
            fb_store_row<ML>(co, so, est - 1, q, sh);
            sb->est_order = est;
        }
        return;
    }

    /* Levinson-Durbin, lpc.c:77-117; rows are quantised as they appear */
    double err = autoc[0];
#pragma unroll
    for (int i = 0; i < ML; i++) {
        if (i < lag) {
            double r = -autoc[i + 1];
#pragma unroll
            for (int j = 0; j < i; j++)
                r = __dsub_rn(r, __dmul_rn(a[j], autoc[i - j]));
            r = __ddiv_rn(r, err);
            err = __dmul_rn(err, __dsub_rn(1.0, __dmul_rn(r, r)));
            fb_levinson_update<ML>(a, i, r);
            if ((om >= 2 || i == lag - 1) && (only_row < 0 || only_row == i)) {
                double rowv[ML];
                int32_t q[ML], sh;
#pragma unroll
                for (int j = 0; j < ML; j++) rowv[j] = j <= i ? -a[j] : 0.0;
                fb_quantize<ML>(rowv, i + 1, q, sh);
                if (writer) fb_store_row<ML>(co, so, i, q, sh);
            }
        }
    }
    if (writer && only_row < 0) sb->est_order = lag;
}

/*
 * coefs_out: [subframe][32][32] int32 (row = order-1, entries 0..order-1 written),
 * shift_out: [subframe][32].  Grid: ceil(subframes / FB_LPC_SUBS_PER_CTA) CTAs.
 */
/* how the tile loader reads the packed PCM */
#define FB_LPC_LK_GENERIC    0   /* any layout: one (branchy) fb_pcm_sample per row and position */
#define FB_LPC_LK_S16_STEREO 1   /* 16-bit stereo: ONE 32-bit load per sample pair serves both rows of a frame */
#define FB_LPC_LK_S24_STEREO 2   /* 24-bit stereo: three 16-bit loads per sample pair */
#define FB_LPC_LK_PLANES     3   /* more than two channels: k_prep's deinterleaved int32 planes (fb_uses_planes) */

template <int ML, int LK>
__global__ void __launch_bounds__(FB_LPC_THREADS)
k_lpc(FbConfig cfg, const FbFrame *frames, const uint32_t *nframes, const void *pcm, int fmt,
      const int32_t *planes, const uint8_t *ch_modes, FbSub *subs, int32_t *coefs_out, int32_t *shift_out)
{
    typedef FbLpcGeom<ML> G;
    __shared__ int32_t s_rows[FB_LPC_WARPS][2][16 * FB_LPC_ROW];
    __shared__ unsigned long long s_ebase[FB_LPC_WARPS][16];   /* interleaved element index of the frame's sample 0 */
    __shared__ int s_n[FB_LPC_WARPS][16];
    __shared__ int s_hole[FB_LPC_WARPS][16];          /* centre of an odd block, else -1 */
    __shared__ int s_cmw[FB_LPC_WARPS][16];           /* channel | ch_mode << 8 | wasted bits << 16 */
    __shared__ int s_nf[FB_LPC_WARPS][16];            /* samples of the row's frame (s_n is 0 for a row that needs no analysis) */
    __shared__ int s_coef[FB_LPC_WARPS][16];          /* stereo rows: the transform as a linear form (fb_stereo_coef) */

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ca = lane & 1, row = lane >> 1;
    const int C = cfg.channels, lag = cfg.max_order;
    const uint32_t nsubs = *nframes * (uint32_t)C;
    const uint32_t sf = blockIdx.x * FB_LPC_SUBS_PER_CTA + (uint32_t)(warp * 16 + row);

    /* same gate as optimize.c:143-193: only the LPC branch needs coefficients */
    /* The subframe's samples are read straight from the packed PCM: deinterleaved, decorrelated by
     * the frame's stereo decision and shifted by the wasted bits on the fly (fb_pcm_sample). */
    int n = 0, cmw = 0, nfr = 0;
    size_t ebase = 0;
    FbSub *sb = nullptr;
    if (sf < nsubs) {
        const uint32_t f = sf / (uint32_t)C;
        const int c = (int)(sf % (uint32_t)C);
        const FbFrame fr = frames[f];
        sb = &subs[sf];
        nfr = (int)fr.n;
        ebase = (size_t)fr.start * C;
        if (LK == FB_LPC_LK_PLANES) ebase += (size_t)c * (size_t)nfr;     /* first word of the subframe's plane */
        cmw = c | ((int)ch_modes[f] << 8) | (sb->wasted << 16);
        if (!(sb->is_const || fr.n < 5u || cfg.prediction_type != 2 || (int)fr.n <= lag)) n = nfr;
    }
    const int nmax = (int)__reduce_max_sync(FB_FULL_MASK, (unsigned)n);
    if (nmax == 0) return;                               /* warp-uniform */
    if (ca == 0) {
        s_ebase[warp][row] = ebase; s_n[warp][row] = n; s_nf[warp][row] = nfr;
        s_hole[warp][row] = (nfr & 1) ? (nfr >> 1) : -1; s_cmw[warp][row] = cmw;
        s_coef[warp][row] = fb_stereo_coef((cmw >> 8) & 0xff, cmw & 0xff, cmw >> 16);
    }
    __syncwarp();
    /* sample p of this lane pair's subframe as the analysis sees it (zero outside the block and at
     * the centre of an odd block) */
    auto x_at = [&](int p) -> int32_t {
        if (!(p >= 0 && p < n && !((n & 1) && p == (n >> 1)))) return 0;
        if (LK == FB_LPC_LK_PLANES) return planes[ebase + (size_t)p];
        return fb_pcm_sample(pcm, fmt, ebase, C, cmw & 0xff, (cmw >> 8) & 0xff, cmw >> 16, p);
    };

    const double cc = __dsub_rn(__ddiv_rn(2.0, __dsub_rn((double)n, 1.0)), 1.0);

    /* ---- head terms (lpc.c:60-61): both lanes compute them, lane 1 starts over at 1.0 ---- */
    double acc[ML + 1];
    {
        double d0[ML + 1];
#pragma unroll
        for (int p = 0; p <= ML; p++) d0[p] = fb_windowed(x_at(p), p, n, cc);
#pragma unroll
        for (int i = 0; i <= ML; i++) {
            double s = 1.0;
#pragma unroll
            for (int j = 0; j <= ML - i; j++)
                if (j <= lag - i) s = __dadd_rn(s, __dmul_rn(d0[j + i], d0[j]));
            acc[i] = ca ? 1.0 : s;
        }
    }

    /* ---- ring of the windowed samples in front of the first tail term ------------------
     * Lane ca walks q = lag+1+ca, +2, ...; before a step the ring holds d[q-2-k] at logical
     * k.  Lane 0 uses its partner's value one step late (d[q-1] was the partner's previous
     * position), lane 1 uses it at once. */
    double D[G::K];
#pragma unroll
    for (int k = 0; k < G::K; k++) {
        const int p = lag - 1 - k + ca;
        D[k] = fb_windowed(x_at(p), p, n, cc);
    }
    double other_prev = fb_windowed(x_at(lag), lag, n, cc);

    const int P0 = lag + 1;
    const int ntiles = (nmax - P0 + G::TL - 1) / G::TL;
    int32_t pre[16 * G::LD];
    int32_t *rows0 = s_rows[warp][0], *rows1 = s_rows[warp][1];

    /* The tile loader.  Stereo layouts: rows 2k and 2k+1 of the warp are the two channels of one
     * frame (a warp's first subframe is even), so one load of the (left, right) pair at position p
     * feeds both rows -- half the loads of reading two planes; everything inside is branch free so
     * that the 16 x LD loads of a tile are in flight together while the previous tile is analysed.
     * Other layouts take fb_pcm_sample per row. */
#define FB_LPC_LOAD_TILE(tb)                                                              \
    do {                                                                                  \
        if (LK == FB_LPC_LK_PLANES) {                                                     \
            _Pragma("unroll")                                                             \
            for (int r = 0; r < 16; r++) {                                                \
                const int32_t *rp = planes + (size_t)s_ebase[warp][r];                    \
                const int rn = s_n[warp][r], rh = s_hole[warp][r];                        \
                _Pragma("unroll")                                                         \
                for (int l = 0; l < G::LD; l++) {                                         \
                    const int p = (tb) + lane + 32 * l;                                   \
                    pre[r * G::LD + l] = (lane + 32 * l < G::TL && p < rn && p != rh) ? rp[p] : 0; \
                }                                                                         \
            }                                                                             \
        } else if (LK != FB_LPC_LK_GENERIC) {                                             \
            _Pragma("unroll")                                                             \
            for (int rp = 0; rp < 8; rp++) {                                              \
                const size_t pe = (size_t)s_ebase[warp][2 * rp] >> 1;                     \
                const int nf_ = s_nf[warp][2 * rp], rh = s_hole[warp][2 * rp];            \
                const int n0_ = s_n[warp][2 * rp], n1_ = s_n[warp][2 * rp + 1];           \
                const int c0_ = s_coef[warp][2 * rp], c1_ = s_coef[warp][2 * rp + 1];     \
                _Pragma("unroll")                                                         \
                for (int l = 0; l < G::LD; l++) {                                         \
                    const int p = (tb) + lane + 32 * l;                                   \
                    const bool ok = lane + 32 * l < G::TL && p < nf_ && p != rh;          \
                    int32_t a_, b_;                                                       \
                    if (LK == FB_LPC_LK_S16_STEREO) {                                     \
                        const uint32_t w_ = ok ? reinterpret_cast<const uint32_t *>(pcm)[pe + (size_t)p] : 0u; \
                        a_ = fb_stereo_apply16(c0_, w_); b_ = fb_stereo_apply16(c1_, w_); \
                    } else {                                                              \
                        const uint16_t *h_ = reinterpret_cast<const uint16_t *>(pcm) + 3 * (pe + (size_t)p); \
                        const uint32_t h0 = ok ? h_[0] : 0u, h1 = ok ? h_[1] : 0u, h2 = ok ? h_[2] : 0u; \
                        const int32_t lv = (int32_t)((h0 | (h1 << 16)) << 8) >> 8;        \
                        const int32_t rv = (int32_t)(((h1 >> 8) | (h2 << 8)) << 8) >> 8;  \
                        a_ = fb_stereo_apply(c0_, lv, rv); b_ = fb_stereo_apply(c1_, lv, rv); \
                    }                                                                     \
                    pre[(2 * rp) * G::LD + l] = p < n0_ ? a_ : 0;                         \
                    pre[(2 * rp + 1) * G::LD + l] = p < n1_ ? b_ : 0;                     \
                }                                                                         \
            }                                                                             \
        } else {                                                                          \
            _Pragma("unroll")                                                             \
            for (int r = 0; r < 16; r++) {                                                \
                const size_t re = (size_t)s_ebase[warp][r];                               \
                const int rn = s_n[warp][r], rh = s_hole[warp][r], rc = s_cmw[warp][r];   \
                _Pragma("unroll")                                                         \
                for (int l = 0; l < G::LD; l++) {                                         \
                    const int p = (tb) + lane + 32 * l;                                   \
                    pre[r * G::LD + l] = (lane + 32 * l < G::TL && p < rn && p != rh)     \
                        ? fb_pcm_sample(pcm, fmt, re, C, rc & 0xff, (rc >> 8) & 0xff, rc >> 16, p) : 0; \
                }                                                                         \
            }                                                                             \
        }                                                                                 \
    } while (0)
#define FB_LPC_STORE_TILE(dst)                                                            \
    do {                                                                                  \
        _Pragma("unroll")                                                                 \
        for (int r = 0; r < 16; r++) {                                                    \
            _Pragma("unroll")                                                             \
            for (int l = 0; l < G::LD; l++)                                               \
                if (lane + 32 * l < G::TL) (dst)[r * FB_LPC_ROW + lane + 32 * l] = pre[r * G::LD + l]; \
        }                                                                                 \
    } while (0)

    if (ntiles > 0) {
        FB_LPC_LOAD_TILE(P0);
        FB_LPC_STORE_TILE(rows0);
    }
    __syncwarp();
    for (int t = 0; t < ntiles; t++) {
        const int tb = P0 + t * G::TL;
        const int32_t *cur = (t & 1) ? rows1 : rows0;
        int32_t *nxt = (t & 1) ? rows0 : rows1;
        if (t + 1 < ntiles) FB_LPC_LOAD_TILE(tb + G::TL);          /* in flight during the tile */
        const int32_t *mine = cur + row * FB_LPC_ROW + ca;
#pragma unroll
        for (int b = 0; b < G::R; b++) {
#pragma unroll
            for (int u = 0; u < G::U; u++) {
                const int off = b * 2 * G::U + 2 * u;
                const double own = fb_windowed(mine[off], tb + off + ca, n, cc);
                const double other = __shfl_xor_sync(FB_FULL_MASK, own, 1);
                const int pos0 = G::K - 2 - 2 * u;                  /* physical slot of logical 0 */
                D[pos0] = own;
                D[pos0 + 1] = ca ? other : other_prev;
                other_prev = other;
#pragma unroll
                for (int i = 0; i <= ML; i++)
                    acc[i] = __dadd_rn(acc[i], __dmul_rn(own, D[(pos0 + i) % G::K]));
            }
        }
        if (t + 1 < ntiles) FB_LPC_STORE_TILE(nxt);
        __syncwarp();
    }
#undef FB_LPC_LOAD_TILE
#undef FB_LPC_STORE_TILE

    /* autoc[i] = temp + temp2 (lpc.c:67) */
    double autoc[ML + 1];
#pragma unroll
    for (int i = 0; i <= ML; i++)
        autoc[i] = __dadd_rn(acc[i], __shfl_xor_sync(FB_FULL_MASK, acc[i], 1));

    if (n == 0) return;
    fb_lpc_rows<ML>(cfg, autoc, lag, sf, ca == 0, sb, coefs_out, shift_out);
}

/*
 * Latency form of the same analysis, for passes of a few blocks (flake_encode_frame: one block
 * per synchronous call).  k_lpc is built for throughput -- a lane pair walks a whole subframe,
 * 152 us for one 4096-sample block when there are no other warps to hide behind.  Here a WARP
 * owns a subframe and a LANE owns one accumulator chain of one lag (lpc.c:57-68: `temp` over
 * the head and the odd tail terms, `temp2` over the even ones; lags beyond 15 take a second and
 * third chain per lane, NS), so the 2 (lag + 1) strictly ordered sums advance side by side and the
 * time is one chain's: n / 2 dependent additions.  The windowed signal is computed once, by all
 * lanes, into shared memory; each step reads data1[j] (broadcast) and data1[j - i].  Two 64-bit
 * shared loads per multiply-add make this form LSU-bound in bulk (3.3 ms per C2 stream in round
 * 1), which is why it is only launched for small passes.  Same operations in the same order as
 * k_lpc: byte-identical rows.  Dynamic shared memory: (n + 2) doubles.
 */
#define FB_LPC_LAT_MAX_SUBFRAMES 2048        /* engine.cu: passes with no more subframes take this kernel (crossover ~3500) */

template <int ML>
__global__ void __launch_bounds__(32)
k_lpc_lat(FbConfig cfg, const FbFrame *frames, const uint32_t *nframes, const void *pcm, int fmt,
          const int32_t *planes, const uint8_t *ch_modes, FbSub *subs, int32_t *coefs_out, int32_t *shift_out)
{
    constexpr int NS = (2 * (ML + 1) + 31) / 32;            /* chains per lane: 2 (ML + 1) accumulators over 32 lanes */
    FB_DYN_SMEM(dyn);
    double *d = reinterpret_cast<double *>(dyn);
    const int lane = threadIdx.x;
    const int C = cfg.channels, lag = cfg.max_order;
    const uint32_t sf = blockIdx.x;
    if (sf >= *nframes * (uint32_t)C) return;
    const uint32_t f = sf / (uint32_t)C;
    const int c = (int)(sf % (uint32_t)C);
    const FbFrame fr = frames[f];
    FbSub *sb = &subs[sf];
    /* same gate as optimize.c:143-193: only the LPC branch needs coefficients */
    if (sb->is_const || fr.n < 5u || cfg.prediction_type != 2 || (int)fr.n <= lag) return;
    const int n = (int)fr.n;
    const int mode = ch_modes[f], wasted = sb->wasted;
    const size_t ebase = (size_t)fr.start * C;

    /* data1[] of lpc.c:46-56; the centre of an odd block is zero (see fb_windowed) */
    const double cc = __dsub_rn(__ddiv_rn(2.0, __dsub_rn((double)n, 1.0)), 1.0);
    /* eight positions per lane at a time: their loads are in flight together (one at a time, the
     * load latency of 128 dependent round trips was most of the kernel) */
    for (int p0 = lane; p0 < n; p0 += 32 * 8) {
        int32_t xv[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int p = p0 + 32 * u;
            xv[u] = 0;
            if (p < n && !((n & 1) && p == (n >> 1)))
                xv[u] = fb_uses_planes(C) ? planes[ebase + (size_t)c * (size_t)n + (size_t)p]
                                          : fb_pcm_sample(pcm, fmt, ebase, C, c, mode, wasted, p);
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int p = p0 + 32 * u;
            if (p < n) d[p] = fb_windowed(xv[u], p, n, cc);
        }
    }
    if (lane < 2) d[n + lane] = 0.0;                      /* data1[len] = 0 */
    __syncwarp();

    /* chain q = 2 i + ca of lane (q % 32), slot (q / 32) */
    double acc[NS];
    int lagi[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) {
        const int q = lane + 32 * s;
        const int i = q >> 1, ca = q & 1;
        lagi[s] = i <= lag ? i : -1;
        double t = 1.0;
        if (i <= lag && ca == 0)
            for (int j = 0; j <= lag - i; j++) t = __dadd_rn(t, __dmul_rn(d[j + i], d[j]));
        acc[s] = t;
    }
    const int ca = lane & 1;                               /* the same for every slot: 32 is even */
    /* temp: j = lag+1, lag+3, ...; temp2: the positions after them.  Eight steps at a time: the
     * loads and products of a batch are independent, only the additions form the chain */
    int j = lag + 1 + ca;
    const int jend = n - 1 + ca;                           /* last position, inclusive */
    /* dl[s][j] = data1[j - lag of the chain]; a lane without a chain walks lag 0 and its sums are never
     * read: no select per product, the loads are base + constant */
    const double *dl[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) dl[s] = d - (lagi[s] >= 0 ? lagi[s] : 0);
    for (; j + 14 <= jend; j += 16) {
        double pr[NS][8];
        const double *dj = d + j;
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const double x = dj[2 * u];
#pragma unroll
            for (int s = 0; s < NS; s++) pr[s][u] = __dmul_rn(x, dl[s][j + 2 * u]);
        }
#pragma unroll
        for (int s = 0; s < NS; s++) {
#pragma unroll
            for (int u = 0; u < 8; u++) acc[s] = __dadd_rn(acc[s], pr[s][u]);
        }
    }
    for (; j <= jend; j += 2) {
        const double x = d[j];
#pragma unroll
        for (int s = 0; s < NS; s++) acc[s] = __dadd_rn(acc[s], __dmul_rn(x, dl[s][j]));
    }

    /* autoc[i] = temp + temp2 (lpc.c:67), gathered to every lane */
    double autoc[ML + 1];
#pragma unroll
    for (int s = 0; s < NS; s++) {
        const double pair = __dadd_rn(__shfl_sync(FB_FULL_MASK, acc[s], lane & ~1), __shfl_sync(FB_FULL_MASK, acc[s], lane | 1));
#pragma unroll
        for (int i = 16 * s; i < 16 * s + 16; i++)
            if (i <= ML) autoc[i] = __shfl_sync(FB_FULL_MASK, pair, 2 * (i - 16 * s));
    }
    /* row `lane` is quantised by lane `lane`; the order estimate needs one row only */
    const int om = cfg.order_method;
    fb_lpc_rows<ML>(cfg, autoc, lag, sf, om == 1 ? lane == 0 : lane < lag, sb, coefs_out, shift_out,
                    om == 1 ? -1 : (om >= 2 ? lane : (lane == 0 ? lag - 1 : 99)));
    if (om != 1 && lane == 0) sb->est_order = lag;
}

#endif
