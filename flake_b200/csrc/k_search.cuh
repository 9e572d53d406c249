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
 * Candidates are evaluated in GROUPS of up to three (order-32 kernel: four; one finishing warp each) whose members do not depend on
 * each other's result (the order searches of optimize.c:205-261 are replayed on
 * the stored costs afterwards, so the decisions are the reference's):
 *   tiles     for each member, every thread owns 16-sample runs; the run and its
 *             history come from the staged plane with 128-bit shared loads, the
 *             residuals stay in registers, one zig-zag sum per run goes to shared
 *             memory (one buffer per member);
 *   barrier
 *   finish    one WARP per member folds the run sums into the partition pyramid with
 *             shuffles (a lane holds 1..8 partitions of the finest level, then pairs
 *             merge), picks parameter and partition order, posts the total;
 *   barrier
 * All candidate coefficient rows are staged once per subframe.  The Rice parameters
 * of the best candidate so far are kept beside the search, which makes the last pass
 * (optimize.c:266-275) a pure residual store.
 */
#ifndef FLAKE_B200_K_SEARCH_CUH
#define FLAKE_B200_K_SEARCH_CUH

#include "dev_common.cuh"
#include "k_lpc.cuh"
#include <type_traits>

#ifndef FB_SEARCH_THREADS
#define FB_SEARCH_THREADS 128    /* 16 samples per thread and tile; measured best of 64/96/128/256 on B200 */
#endif

/* Candidates evaluated between two decisions, one finishing warp each (at most 4: FbOrders).
 * The order <= 12 kernel takes three: its usual search (log, C2) never has more than three
 * independent candidates and the fourth run-sum array only costs shared memory (3.174 vs 3.187 ms);
 * the order-32 kernel takes four (exhaustive search of C3: 6.02 -> 5.56 ms). */
#define FB_GROUP_MAX 4
#define FB_PLAN_SMEM_NODES 12    /* nodes of the log-search plan staged in shared memory (all of them up to order 12) */
#define FB_GROUP_OF(MAXP) ((MAXP) > 12 ? 4 : 3)

/* dev builds (-DFB_SEARCH_PROF): cycles per phase as seen by thread 0, summed over all CTAs */
#ifdef FB_SEARCH_PROF
__device__ unsigned long long g_sprof[16];
#define FB_PROF_DECL long long prof_t = clock64()
#define FB_PROF(i) do { if (threadIdx.x == 0) { const long long now_ = clock64(); atomicAdd(&g_sprof[(i)], (unsigned long long)(now_ - prof_t)); prof_t = now_; } } while (0)
#define FB_PROF_ARG , long long &prof_t
#define FB_PROF_PASS , prof_t
#else
#define FB_PROF_DECL do { } while (0)
#define FB_PROF(i) do { } while (0)
#define FB_PROF_ARG
#define FB_PROF_PASS
#endif

#define FB_F64_MAGIC 6755399441055744.0          /* 1.5 * 2^52 */
/* Which kernel takes the FP64 bodies.  Measured (ms per pass, search stage): the order-32 kernel
 * is bound by the multiplies (C3, exhaustive search over 32 orders: 10.74 with IMAD.WIDE, 8.25
 * with DFMA); the order <= 12 kernel is bound by barriers and latency, not by the multiply pipe,
 * and the int -> double conversions of the window only add to its critical path (C2 3.21 -> 4.09,
 * C4 4.00 -> 4.25), so it keeps IMAD.WIDE. */
#ifndef FB_F64_WIDE
#define FB_F64_WIDE(MAXP) ((MAXP) > 12)
#endif

template <int MAXP>
struct FbSearchShared {
    unsigned long long sums[256];   /* finest-level partition sums when runs do not tile the partitions */
    uint8_t  kbuf[FB_GROUP_OF(MAXP)][512];   /* group member: parameter of partition j at level L at [(1<<L)-1+j] */
    uint8_t  kbest[256];            /* best candidate so far: parameters at its partition order */
    int32_t  coef[MAXP][MAXP];      /* candidate rows (row = order-1), zero padded */
    double   coefd[FB_F64_WIDE(MAXP) ? MAXP : 1][FB_F64_WIDE(MAXP) ? MAXP : 1];   /* the same rows as coef * 2^-shift
                                     * (exact), for the FP64 bodies (fb_floor_lo32); only where they are used */
    int32_t  shift[MAXP];
    uint32_t sumabs[MAXP];          /* sum |coef| per row */
    uint8_t  narrow_of[MAXP];       /* row can be costed in 32-bit arithmetic */
    uint8_t  pmin_of[MAXP + 1], pmax_of[MAXP + 1];   /* partition-order limits per predictor order (rice.c:148-171) */
    uint8_t  sum32_of[MAXP];        /* every partition sum of the row's residual is below 2^31 (32-bit pyramid) */
    uint32_t result[FB_GROUP_MAX];      /* of the group members just finished */
    int32_t  porder[FB_GROUP_MAX], method[FB_GROUP_MAX];
    int32_t  best_porder, best_method;
    uint32_t best_bits;
    FbPlanNode plan[FB_PLAN_SMEM_NODES];
};

/*
 * The 64-bit predictor on the FP64 pipe.  Where 32-bit arithmetic cannot be proven exact (24-bit
 * audio: |prediction| reaches 2^37) the reference's int64 sum costs an IMAD.WIDE per tap, a quarter
 * rate instruction (32 lanes/clk/SM).  B200's DFMA runs at 64 lanes/clk/SM and is exact here:
 * |coef| < 2^14, |sample| < 2^32, <= 32 taps, so every product and partial sum is an integer
 * below 2^51 (times the power of two 2^-shift folded into the coefficients, which only moves the
 * exponent).  floor(pred / 2^shift) -- the arithmetic right shift of optimize.c:84-118 -- is one
 * add of 1.5 * 2^52 rounded toward minus infinity: the low mantissa word of the sum is the floor
 * modulo 2^32, which is all the residual (an int32) keeps.
 */
__device__ __forceinline__ int32_t fb_floor_lo32(double y)
{
    return __double2loint(__dadd_rd(y, FB_F64_MAGIC));
}

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
 * One whole warp, for group member `slot`.  F[] holds the finest-level sums: one per
 * 16-sample run (`per` runs per partition of level pmax), or the partition sums themselves
 * (per == 1, S.sums).  The pairwise pyramid of rice.c:96-102 is exact integer addition, so a
 * partition of level L is simply the sum of its per << (pmax - L) entries of F.  Levels of 32
 * partitions or more are walked lane-strided; the six coarser levels are costed side by side,
 * one partition per lane group, merged by shuffles.  Levels are compared from the finest down
 * and replace the best only when strictly smaller, which is the reference's ascending scan
 * with `<=` (ties to the higher order, rice.c:128-135).
 * Posts S.result[slot] (rice.c:157-187 total), S.porder[slot], S.method[slot] and the
 * parameters of every level in S.kbuf[slot].
 */
template <int MAXP, typename SumT>
__device__ __forceinline__ void fb_finish_body(FbSearchShared<MAXP> &S, int slot, const unsigned long long *F,
                                               int per, int n, int is_lpc, int order, int obits,
                                               int pmin, int pmax)
{
    const int lane = threadIdx.x & 31;
    uint8_t *kbuf = S.kbuf[slot];
    uint32_t best = 0xffffffffu;
    int bl = pmin, bmethod = 0;
    SumT t5 = 0;

    /* levels with 32 partitions or more */
#pragma unroll 1
    for (int L = pmax; L >= 5; L--) {
        const int span = per << (pmax - L);
        uint32_t bits = 0;
        int flag = 0;
#pragma unroll 1
        for (int j = lane; j < (1 << L); j += 32) {
            const unsigned long long *src = F + j * span;
            SumT sum = 0;
#pragma unroll 4
            for (int q = 0; q < span; q++) sum += (SumT)src[q];
            const int cnt = (n >> L) - (j == 0 ? order : 0);
            const int k = fb_rice_k_t(sum, cnt);
            kbuf[(1 << L) - 1 + j] = (uint8_t)k;
            bits += fb_rice_count_t(sum, cnt, k);
            flag |= (k > 14);
            t5 = sum;                                     /* L == 5: the lane's own partition */
        }
        const uint32_t tot = __reduce_add_sync(FB_FULL_MASK, bits);
        const int r2 = __any_sync(FB_FULL_MASK, flag) ? 1 : 0;
        const uint32_t b = tot + 4u * (1u << L);
        if (L >= pmin && b < best) { best = b; bl = L; bmethod = r2; }
    }
    if (pmax < 5) {
        /* the lanes of a group share partition lane / g of level pmax */
        const int g = 32 >> pmax, j = lane / g, sub = lane % g;
        SumT a = 0;
        for (int q = sub; q < per; q += g) a += (SumT)F[j * per + q];
        for (int o = 1; o < g; o <<= 1) a += __shfl_xor_sync(FB_FULL_MASK, a, o);
        t5 = a;
    }

    /* levels 4..0 (and level pmax < 5): t[L] = sum of the partition this lane belongs to.
     * With pmax < 5 the lanes start out holding level pmax, so the merges above it are skipped. */
    if (pmin < 5) {
        SumT t[5];
        {
            const SumT o = __shfl_xor_sync(FB_FULL_MASK, t5, 1);
            t[4] = t5 + (5 <= pmax ? o : (SumT)0);
        }
#pragma unroll
        for (int L = 4; L >= 1; L--) {
            const SumT o = __shfl_xor_sync(FB_FULL_MASK, t[L], 1 << (5 - L));
            t[L - 1] = t[L] + (L <= pmax ? o : (SumT)0);
        }
        uint32_t bits[5];
        int ks[5];
#pragma unroll
        for (int L = 0; L < 5; L++) {
            const int g = 32 >> L, j = lane / g;
            const int cnt = (n >> L) - (j == 0 ? order : 0);
            const int k = fb_rice_k_t(t[L], cnt);
            const bool mine = (lane % g) == 0 && L <= pmax && L >= pmin;
            ks[L] = k;
            bits[L] = mine ? fb_rice_count_t(t[L], cnt, k) : 0u;
        }
#pragma unroll
        for (int L = 4; L >= 0; L--) {
            const int g = 32 >> L, j = lane / g;
            const bool mine = (lane % g) == 0 && L <= pmax && L >= pmin;
            if (mine) kbuf[(1 << L) - 1 + j] = (uint8_t)ks[L];
            const uint32_t tot = __reduce_add_sync(FB_FULL_MASK, bits[L]);
            const int r2 = __any_sync(FB_FULL_MASK, mine && ks[L] > 14) ? 1 : 0;
            const uint32_t b = tot + 4u * (1u << L);
            if (L <= pmax && L >= pmin && b < best) { best = b; bl = L; bmethod = r2; }
        }
    }

    if (lane == 0) {
        uint32_t total = (uint32_t)(order * obits + 2);
        if (is_lpc) total += 4u + 5u + (uint32_t)order * 15u;
        total += best;
        total += (uint32_t)bmethod + 4u;
        S.result[slot] = total;
        S.porder[slot] = bl;
        S.method[slot] = bmethod;
    }
}


/* the always-exact 64-bit finish, out of line: audio whose zig-zag residuals sum to 2^31 or more
 * over one block (mean |residual| of 2^18 at 4096 samples) is the rare case */
template <int MAXP>
__device__ __noinline__ void fb_finish_wide(FbSearchShared<MAXP> &S, int slot, const unsigned long long *F,
                                            int per, int n, int is_lpc, int order, int obits, int pmin, int pmax)
{
    fb_finish_body<MAXP, unsigned long long>(S, slot, F, per, n, is_lpc, order, obits, pmin, pmax);
}

/*
 * One whole warp.  When the block's zig-zag total is provably below 2^31 (S.sum32_of, from the
 * magnitude bound of the plane and the row's coefficients), every partition sum of every
 * level is below 2^31 and the whole finish is exact in 32-bit arithmetic (the costs are uint32
 * in the reference, rice.c:30-45, and `sum - n/2` can only wrap where k = 0, where the wrap is
 * the same modulo 2^32); otherwise the 64-bit body.
 */
template <int MAXP>
__device__ __noinline__ void fb_finish_warp(FbSearchShared<MAXP> &S, int slot, const unsigned long long *F,
                                            int per, int n, int is_lpc, int order, int obits, int pmin, int pmax)
{
    /* S.sum32_of: 2 * (bound on |residual|) * n < 2^31 was proven when the row was staged (a scan
     * of the run sums here instead: 3.24 vs 3.16 ms) */
    if (!S.sum32_of[is_lpc ? order - 1 : order]) {
        fb_finish_wide<MAXP>(S, slot, F, per, n, is_lpc, order, obits, pmin, pmax);
        return;
    }
    fb_finish_body<MAXP, uint32_t>(S, slot, F, per, n, is_lpc, order, obits, pmin, pmax);
}

/*
 * The finish for the common shape: 32-bit run sums (F32, one word per run, written by the 32-bit
 * run bodies), n a power of two from 512 to FB_FAST_MAX_N, every partition sum below 2^31
 * (S.sum32_of).  Levels 5 and up: a partition is 1, 2, 4 or 8 consecutive run sums -- one or two
 * vector loads -- partitions lane-strided; level 5 leaves every lane the sum of its own
 * partition.  Levels 4 .. 0: five xor shuffles give every lane the sum of the partition it
 * belongs to, and EVERY lane costs one partition (rice.c:30-74): lane l takes level 5 - c,
 * c = 1 + ctz(l), partition l >> c -- 16 + 8 + 4 + 2 + 1 lanes, one Rice computation for five
 * levels; the level totals are xor steps that keep the lowest set bit of the lane index in
 * place.  A parameter above 14 (RICE2) is counted in bits 20 and up of the same word.  The best
 * level is the smallest total, the finer one on a tie (rice.c:128-135): one minimum over
 * (bits, 31 - level, method).  Same posts as fb_finish_body.
 */
#define FB_FAST_MAX_N 8192

__device__ __forceinline__ uint32_t fb_level_key(uint32_t x, int L)
{
    return (((x & 0xfffffu) + (4u << L)) << 6) | ((uint32_t)(31 - L) << 1) | ((x >> 20) ? 1u : 0u);
}

template <int MAXP>
__device__ __noinline__ void fb_finish_fast(FbSearchShared<MAXP> &S, int slot, const uint32_t *F32, int n, int is_lpc,
                                            int order, int obits, int pmin, int pmax)
{
    const int lane = threadIdx.x & 31;
    uint8_t *kbuf = S.kbuf[slot];
    const int ltop = 27 - __clz(n);                               /* the level whose partitions are single runs */
    uint32_t key = 0xffffffffu, a3 = 0;
    const int lfirst = pmax < 5 ? 5 : (pmax < ltop ? pmax : ltop);
#pragma unroll 1
    for (int L = lfirst; L >= 5; L--) {
        const int e = ltop - L;
        uint32_t x = 0;
#pragma unroll 1
        for (int j = lane; j < (1 << L); j += 32) {
            const uint32_t *src = F32 + (j << e);
            uint32_t sum;
            if (e >= 3) {                                         /* 8 run sums, 16 in blocks of 8192 */
                sum = 0;
#pragma unroll 1
                for (int q4 = 0; q4 < (1 << e); q4 += 8) {
                    const uint4 p = *reinterpret_cast<const uint4 *>(src + q4), q = *reinterpret_cast<const uint4 *>(src + q4 + 4);
                    sum += ((p.x + p.y) + (p.z + p.w)) + ((q.x + q.y) + (q.z + q.w));
                }
            } else if (e == 2) {
                const uint4 p = *reinterpret_cast<const uint4 *>(src);
                sum = (p.x + p.y) + (p.z + p.w);
            } else if (e == 1) {
                const uint2 p = *reinterpret_cast<const uint2 *>(src);
                sum = p.x + p.y;
            } else {
                sum = src[0];
            }
            const int cnt = (16 << e) - (j == 0 ? order : 0);     /* runs of 16 samples (FB_RUN) */
            const int k = fb_rice_k_t(sum, cnt);
            kbuf[(1 << L) - 1 + j] = (uint8_t)k;
            x += fb_rice_count_t(sum, cnt, k) + (k > 14 ? (1u << 20) : 0u);
            a3 = sum;
        }
        x = __reduce_add_sync(FB_FULL_MASK, x);
        if (L <= pmax && L >= pmin) key = min(key, fb_level_key(x, L));
    }
    /* across the lanes: levels 4 .. 0 */
    if (pmin < 5) {
        const uint32_t c1 = a3 + __shfl_xor_sync(FB_FULL_MASK, a3, 1);
        const uint32_t c2 = c1 + __shfl_xor_sync(FB_FULL_MASK, c1, 2);
        const uint32_t c3 = c2 + __shfl_xor_sync(FB_FULL_MASK, c2, 4);
        const uint32_t c4 = c3 + __shfl_xor_sync(FB_FULL_MASK, c3, 8);
        const uint32_t c5 = c4 + __shfl_xor_sync(FB_FULL_MASK, c4, 16);
        const int c = lane ? __ffs(lane) : 6;
        const int L = 5 - c;
        const uint32_t sum = c == 1 ? c1 : (c == 2 ? c2 : (c == 3 ? c3 : (c == 4 ? c4 : c5)));
        const int j = lane >> c;
        const int cnt = (n >> (L < 0 ? 0 : L)) - (j == 0 ? order : 0);
        const int k = fb_rice_k_t(sum, cnt);
        const bool use = lane != 0 && L <= pmax && L >= pmin;
        uint32_t x = use ? fb_rice_count_t(sum, cnt, k) + (k > 14 ? (1u << 20) : 0u) : 0u;
        if (use) kbuf[(1 << L) - 1 + j] = (uint8_t)k;
#pragma unroll
        for (int o = 1; o < 5; o++) {
            const uint32_t t = __shfl_xor_sync(FB_FULL_MASK, x, 1 << o);
            if (o >= c) x += t;
        }
        if (use) key = min(key, fb_level_key(x, L));
    }
    key = __reduce_min_sync(FB_FULL_MASK, key);
    if (lane == 0) {
        const bool none = key == 0xffffffffu;                     /* no level allowed: fb_finish_body's untouched best */
        const uint32_t bmethod = none ? 0u : key & 1u;
        uint32_t total = (uint32_t)(order * obits + 2);
        if (is_lpc) total += 4u + 5u + (uint32_t)order * 15u;
        total += (none ? 0xffffffffu : key >> 6) + bmethod + 4u;
        S.result[slot] = total;
        S.porder[slot] = none ? pmin : 31 - (int)((key >> 1) & 31u);
        S.method[slot] = (int)bmethod;
    }
}

/*
 * The same finish over 64-bit run sums, exact for any input (24-bit audio, whose residual bound
 * cannot prove 32-bit partition sums; full-scale noise, where they really do not fit): the
 * structure of fb_finish_fast -- vector loads, a partition per lane, one Rice computation for the
 * five coarse levels -- with the always-exact arithmetic of fb_finish_body<unsigned long long>
 * (costs truncated to uint32 as the reference computes them, rice.c:30-45; levels compared from
 * the finest down, replaced only when strictly smaller).  n a power of two from 512 to FB_FAST_MAX_N.
 */
template <int MAXP>
__device__ __noinline__ void fb_finish_fast64(FbSearchShared<MAXP> &S, int slot, const unsigned long long *F, int n,
                                              int is_lpc, int order, int obits, int pmin, int pmax)
{
    const int lane = threadIdx.x & 31;
    uint8_t *kbuf = S.kbuf[slot];
    const int ltop = 27 - __clz(n);                               /* the level whose partitions are single runs */
    uint32_t best = 0xffffffffu;
    int bl = pmin, bmethod = 0;
    unsigned long long a3 = 0;
    const int lfirst = pmax < 5 ? 5 : (pmax < ltop ? pmax : ltop);
#pragma unroll 1
    for (int L = lfirst; L >= 5; L--) {
        const int e = ltop - L;
        uint32_t bits = 0;
        int flag = 0;
#pragma unroll 1
        for (int j = lane; j < (1 << L); j += 32) {
            const unsigned long long *src = F + (j << e);
            unsigned long long sum = 0;
            if (e == 0) sum = src[0];
            else {
#pragma unroll 1
                for (int q = 0; q < (1 << e); q += 2) {
                    const ulonglong2 p = *reinterpret_cast<const ulonglong2 *>(src + q);
                    sum += p.x + p.y;
                }
            }
            const int cnt = (16 << e) - (j == 0 ? order : 0);     /* runs of 16 samples (FB_RUN) */
            const int k = fb_rice_k_t(sum, cnt);
            kbuf[(1 << L) - 1 + j] = (uint8_t)k;
            bits += fb_rice_count_t(sum, cnt, k);
            flag |= (k > 14);
            a3 = sum;                                             /* L == 5 comes last: the lane's own partition */
        }
        const uint32_t b = __reduce_add_sync(FB_FULL_MASK, bits) + 4u * (1u << L);
        const int r2 = __any_sync(FB_FULL_MASK, flag) ? 1 : 0;
        if (L <= pmax && L >= pmin && b < best) { best = b; bl = L; bmethod = r2; }
    }
    /* across the lanes: levels 4 .. 0, lane l costs partition l >> c of level 5 - c, c = 1 + ctz(l) */
    if (pmin < 5) {
        const unsigned long long c1 = a3 + __shfl_xor_sync(FB_FULL_MASK, a3, 1);
        const unsigned long long c2 = c1 + __shfl_xor_sync(FB_FULL_MASK, c1, 2);
        const unsigned long long c3 = c2 + __shfl_xor_sync(FB_FULL_MASK, c2, 4);
        const unsigned long long c4 = c3 + __shfl_xor_sync(FB_FULL_MASK, c3, 8);
        const unsigned long long c5 = c4 + __shfl_xor_sync(FB_FULL_MASK, c4, 16);
        const int c = lane ? __ffs(lane) : 6;
        const int Lc = 5 - c;
        const unsigned long long sum = c == 1 ? c1 : (c == 2 ? c2 : (c == 3 ? c3 : (c == 4 ? c4 : c5)));
        const int j = lane >> c;
        const int cnt = (n >> (Lc < 0 ? 0 : Lc)) - (j == 0 ? order : 0);
        const int k = fb_rice_k_t(sum, cnt);
        const bool use = lane != 0 && Lc <= pmax && Lc >= pmin;
        uint32_t x = use ? fb_rice_count_t(sum, cnt, k) : 0u;
        if (use) kbuf[(1 << Lc) - 1 + j] = (uint8_t)k;
        const uint32_t over = __ballot_sync(FB_FULL_MASK, use && k > 14);
#pragma unroll
        for (int o = 1; o < 5; o++) {
            const uint32_t t = __shfl_xor_sync(FB_FULL_MASK, x, 1 << o);
            if (o >= c) x += t;
        }
        /* lane 2^(c-1) holds the total of level 5 - c; the lanes of that role are those = 2^(c-1) mod 2^c */
#pragma unroll
        for (int cc = 1; cc <= 5; cc++) {
            const int L = 5 - cc;
            const uint32_t b = __shfl_sync(FB_FULL_MASK, x, 1 << (cc - 1)) + 4u * (1u << L);
            const uint32_t role = cc == 1 ? 0xaaaaaaaau : (cc == 2 ? 0x44444444u : (cc == 3 ? 0x10101010u : (cc == 4 ? 0x01000100u : 0x00010000u)));
            const int r2 = (over & role) ? 1 : 0;
            if (L <= pmax && L >= pmin && b < best) { best = b; bl = L; bmethod = r2; }
        }
    }
    if (lane == 0) {
        uint32_t total = (uint32_t)(order * obits + 2);
        if (is_lpc) total += 4u + 5u + (uint32_t)order * 15u;
        total += best;
        total += (uint32_t)bmethod + 4u;
        S.result[slot] = total;
        S.porder[slot] = bl;
        S.method[slot] = bmethod;
    }
}

/* group member `slot` is the best candidate so far: warp 0 keeps its parameters */
template <int MAXP>
__device__ __forceinline__ void fb_keep_best(FbSearchShared<MAXP> &S, int slot, uint32_t bits)
{
    if (threadIdx.x < 32) {
        __syncwarp();
        const int bl = S.porder[slot], np = 1 << bl;
        for (int j = threadIdx.x; j < np; j += 32) S.kbest[j] = S.kbuf[slot][np - 1 + j];
        if (threadIdx.x == 0) { S.best_porder = bl; S.best_method = S.method[slot]; S.best_bits = bits; }
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
 * Every thread of the CTA calls it; the total (calc_rice_params_fixed / _lpc
 * return value, rice.c:157-187) lands in S.result[slot].
 * want_sums: cost the candidate; res_out != NULL: store the residual (warm-up = samples).
 */
template <int MAXP>
__device__ __noinline__ void fb_evaluate(FbSearchShared<MAXP> &S, int slot, const int32_t *x, int n, int is_lpc,
                                         int order, int row, int obits, int pmin_cfg, int pmax_cfg,
                                         int32_t *res_out, bool want_sums)
{
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31;
    const int pmin = fb_limit_porder(pmin_cfg, n, order);
    const int pmax = fb_limit_porder(pmax_cfg, n, order);
    const int nparts = 1 << pmax, psize = n >> pmax;
    const int32_t *coef = S.coef[row];
    const int shift = S.shift[row];
    if (want_sums) {
        __syncthreads();                                  /* the previous member's finish is done with S.sums */
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
    if (!want_sums) return;
    __syncthreads();
    if (tid < 32) fb_finish_warp<MAXP>(S, slot, S.sums, 1, n, is_lpc, order, obits, pmin, pmax);
}

/* ------------------------------------------------------------------ */
/* fast path: block staged in shared memory in a skewed layout          */
/* ------------------------------------------------------------------ */
#ifndef FB_RUN
#define FB_RUN 16                 /* consecutive samples per thread and tile: 16 or 8 */
#endif
#define FB_HIST 32                /* zero samples in front of sample 0 */

/* logical index (sample i lives at logical i + FB_HIST) -> word offset.  The 128-bit loads of
 * eight neighbouring threads must hit the eight distinct 16-byte bank groups:
 *   runs of 16: 4 pad words after every 16 samples (granule 5t + const);
 *   runs of 8:  4 pad words after every 32 samples (granule 2t + (2t >> 3) + const). */
#if FB_RUN == 16
__host__ __device__ __forceinline__ int fb_skew(int logical) { return logical + ((logical >> 4) << 2); }
#else
__host__ __device__ __forceinline__ int fb_skew(int logical) { return logical + ((logical >> 5) << 2); }
#endif
__host__ __device__ __forceinline__ int fb_skew_words(int n) { return (fb_skew(n + FB_HIST + 16) + 8 + 3) & ~3; }
/* 64-bit zig-zag sums, one per 16-sample run, per group member: words */
__host__ __device__ __forceinline__ int fb_runsum_words(int n) { return 2 * (((n + FB_RUN - 1) / FB_RUN) + 2); }
/* staged plane + (the run sums of a whole group | the TMA tile ring the plane is unpacked from: the
 * ring is dead once the plane is staged, the run sums are only written after that), in 32-bit words */
__host__ __device__ __forceinline__ int fb_search_smem_words(int n, int group)
{
    const int rs = group * fb_runsum_words(n), tile = FB_TILE_BYTES / 4;
    return fb_skew_words(n) + (rs > tile ? rs : tile);
}
/* word offset of logical (16 m + d) relative to that of logical 16 m; d may be negative */
__host__ __device__ constexpr int fb_skew_delta(int d) { return d + 4 * (d >= 0 ? d / 16 : -((-d + 15) / 16)); }

/* window of a run: groups 0..G-1 of four samples, group g at logical offset 4g - P from the run */
template <int P, int G> struct FbWindow {
    /* xr: shared address of the run's first sample (runs of 16) or of the staged plane (runs
     * of 8, L0 = logical index of the run's first sample) */
    static __device__ __forceinline__ void load(fb_sptr xr, int L0, int32_t *w)
    {
        FbWindow<P, G - 1>::load(xr, L0, w);
#if FB_RUN == 16
        const int4 v = fb_lds128<4 * fb_skew_delta(4 * (G - 1) - P)>(xr);
#else
        const int4 v = fb_lds128_at(xr, 4 * fb_skew(L0 + 4 * (G - 1) - P));
#endif
        w[4 * (G - 1)] = v.x; w[4 * (G - 1) + 1] = v.y; w[4 * (G - 1) + 2] = v.z; w[4 * (G - 1) + 3] = v.w;
    }
};
template <int P> struct FbWindow<P, 0> {
    static __device__ __forceinline__ void load(fb_sptr, int, int32_t *) {}
};

#define FB_SUMS  1                /* cost the candidate: zig-zag sums */
#define FB_STORE 2                /* write the residual to global memory */

/* Cold paths of a run, out of line so that the hot body stays small (the kernel is bound by
 * instruction fetch before anything else). */

/* a run that crosses the end of the block: per sample, 64-bit, straight from the staged plane */
template <int MAXP>
__device__ __noinline__ void fb_run_tail(FbSearchShared<MAXP> &S, const int32_t *xs, int n, int order, int row,
                                         int psize, int i0, int32_t *res_out, unsigned long long *runsum, int mode)
{
    const int32_t *coef = S.coef[row];
    const int shift = S.shift[row];
    unsigned long long acc = 0;
    int pcur = -1;
    for (int i = i0; i < n && i < i0 + FB_RUN; i++) {
        const int32_t xi = xs[fb_skew(i + FB_HIST)];
        int32_t r = xi;
        if (i >= order) {
            long long pred = 0;
            for (int j = 0; j < order; j++)
                pred += (long long)coef[j] * (long long)xs[fb_skew(i - 1 - j + FB_HIST)];
            r = (int32_t)((long long)xi - (pred >> shift));
        }
        if (mode & FB_STORE) res_out[i] = r;
        if ((mode & FB_SUMS) && i >= order) {
            if (!runsum) {
                const int pi = i / psize;
                if (pi != pcur) {
                    if (acc) atomicAdd(&S.sums[pcur], acc);
                    acc = 0; pcur = pi;
                }
            }
            acc += fb_zigzag(r);
        }
    }
    if (mode & FB_SUMS) {
        if (runsum) runsum[i0 / FB_RUN] = acc;
        else if (acc) atomicAdd(&S.sums[pcur], acc);
    }
}

/* full run whose partitions are not whole runs: walk it, flushing at partition boundaries */
struct FbRunRegs { int4 g[FB_RUN / 4]; };

template <int MAXP>
__device__ __noinline__ void fb_run_scatter(FbSearchShared<MAXP> &S, int i0, int order, int psize, FbRunRegs rr)
{
    int32_t r[FB_RUN];
    for (int g = 0; g < FB_RUN / 4; g++) { r[4 * g] = rr.g[g].x; r[4 * g + 1] = rr.g[g].y; r[4 * g + 2] = rr.g[g].z; r[4 * g + 3] = rr.g[g].w; }
    unsigned long long acc = 0;
    int pcur = -1;
    for (int k = 0; k < FB_RUN; k++) {
        const int i = i0 + k;
        if (i < order) continue;
        const int pi = i / psize;
        if (pi != pcur) {
            if (acc) atomicAdd(&S.sums[pcur], acc);
            acc = 0; pcur = pi;
        }
        acc += fb_zigzag(r[k]);
    }
    if (acc) atomicAdd(&S.sums[pcur], acc);
}

/* full run to a destination that is not 16-byte aligned */
__device__ __noinline__ void fb_run_store_scalar(int32_t *dst, FbRunRegs rr)
{
    for (int g = 0; g < FB_RUN / 4; g++) { dst[4 * g] = rr.g[g].x; dst[4 * g + 1] = rr.g[g].y; dst[4 * g + 2] = rr.g[g].z; dst[4 * g + 3] = rr.g[g].w; }
}

/*
 * Residual of the full run [i0, i0+16) with a register sliding window.
 * P = predictor order rounded up to a multiple of 4 (coefficients beyond the
 * order are zero).  WIDE: 64-bit prediction and sums (always exact);
 * !WIDE: 32-bit, used only when the caller proved nothing can overflow.
 */
template <int MAXP, int P, bool WIDE, int MODE>
__device__ __forceinline__ void fb_run_residual(FbSearchShared<MAXP> &S, const int32_t *xs, int order,
                                                int row, int psize, int i0,
                                                int32_t *res_out, unsigned long long *runsum)
{
    int32_t c[P];
#pragma unroll
    for (int g = 0; g < P / 4; g++) {
        const int4 v = *reinterpret_cast<const int4 *>(&S.coef[row][4 * g]);
        c[4 * g] = v.x; c[4 * g + 1] = v.y; c[4 * g + 2] = v.z; c[4 * g + 3] = v.w;
    }
    const int shift = S.shift[row];

    int32_t w[P + FB_RUN];
#if FB_RUN == 16
    const fb_sptr xr = fb_to_sptr(xs + fb_skew(i0 + FB_HIST));  /* i0 + FB_HIST is a multiple of 16 */
#else
    const fb_sptr xr = fb_to_sptr(xs);
#endif
    FbWindow<P, (P + FB_RUN) / 4>::load(xr, i0 + FB_HIST, w);

    /* residuals of the run, branch free */
    int32_t r[FB_RUN];
    if constexpr (WIDE && FB_F64_WIDE(MAXP)) {
        /* the 64-bit prediction on the FP64 pipe, see fb_floor_lo32 */
        const double *cd = S.coefd[row];
        double wd[P + FB_RUN - 1], pd[FB_RUN];
#pragma unroll
        for (int q = 0; q < P + FB_RUN - 1; q++) wd[q] = (double)w[q];
#pragma unroll
        for (int k = 0; k < FB_RUN; k++) pd[k] = 0.0;
#pragma unroll
        for (int j = 0; j < P; j++) {
            const double cj = cd[j];
#pragma unroll
            for (int k = 0; k < FB_RUN; k++) pd[k] = __fma_rn(cj, wd[P + k - 1 - j], pd[k]);
        }
#pragma unroll
        for (int k = 0; k < FB_RUN; k++) r[k] = (int32_t)((uint32_t)w[P + k] - (uint32_t)fb_floor_lo32(pd[k]));
    } else {
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
    }
    /* zig-zag sum of the run (rice.c:76-95) */
    unsigned long long acc = 0;
    if (MODE & FB_SUMS) {
        if (WIDE) {
#pragma unroll
            for (int k = 0; k < FB_RUN; k++) acc += fb_zigzag(r[k]);
        } else {
            uint32_t a32 = 0;
#pragma unroll
            for (int k = 0; k < FB_RUN; k++) a32 += fb_zigzag(r[k]);
            acc = a32;
        }
    }
    if (i0 < order) {                                         /* warm-up samples pass through, uncounted */
        FB_COLD_BLOCK();
        /* order <= P: beyond the first run only orders above 16 have warm-up samples */
#pragma unroll
        for (int k = 0; k < (P < FB_RUN ? P : FB_RUN); k++) {
            if (i0 + k < order) {
                if (MODE & FB_SUMS) acc -= fb_zigzag(r[k]);
                r[k] = w[P + k];
            }
        }
    }
    if (MODE & FB_STORE) {
        int32_t *dst = res_out + i0;
        if ((((size_t)dst) & 15u) == 0) {
#pragma unroll
            for (int g = 0; g < FB_RUN / 4; g++)
                reinterpret_cast<int4 *>(dst)[g] = make_int4(r[4 * g], r[4 * g + 1], r[4 * g + 2], r[4 * g + 3]);
        } else {
            FbRunRegs rr;
#pragma unroll
            for (int g = 0; g < FB_RUN / 4; g++) rr.g[g] = make_int4(r[4 * g], r[4 * g + 1], r[4 * g + 2], r[4 * g + 3]);
            fb_run_store_scalar(dst, rr);
        }
    }
    if (MODE & FB_SUMS) {
        /* partitions that are whole multiples of the run length (the rule for every
         * power-of-two block size): one sum per run, folded into the partition sums by the
         * finishing warp -- no atomics on the hot loop */
        if (runsum) runsum[i0 / FB_RUN] = acc;
        else {
            FbRunRegs rr;
#pragma unroll
            for (int g = 0; g < FB_RUN / 4; g++) rr.g[g] = make_int4(r[4 * g], r[4 * g + 1], r[4 * g + 2], r[4 * g + 3]);
            fb_run_scatter<MAXP>(S, i0, order, psize, rr);
        }
    }
}

template <int MAXP, int P, bool WIDE, int MODE>
__device__ __noinline__ void fb_tiles(FbSearchShared<MAXP> &S, const int32_t *xs, int n, int order, int row, int psize,
                                      int32_t *res_out, unsigned long long *runsum)
{
    for (int i0 = (int)threadIdx.x * FB_RUN; i0 < n; i0 += (int)blockDim.x * FB_RUN) {
        if (i0 + FB_RUN <= n) fb_run_residual<MAXP, P, WIDE, MODE>(S, xs, order, row, psize, i0, res_out, runsum);
        else fb_run_tail<MAXP>(S, xs, n, order, row, psize, i0, res_out, runsum, MODE);
    }
}

/*
 * Residual pass of the candidate in row `row` of S.coef over the staged block (fixed
 * predictors: binomial coefficients, shift 0); no barrier inside.  `maxabs` bounds |sample|
 * and decides whether 32-bit arithmetic is provably exact:
 *   |pred| <= sum|c| * maxabs < 2^31, and
 *   |residual| <= maxabs + (sum|c|*maxabs >> shift) + 1 < 2^26 so that a run's
 *   zig-zag sum fits 32 bits.
 * MODE: FB_SUMS, FB_STORE or both.  runsum == NULL with FB_SUMS: partition sums by atomics
 * into S.sums (zeroed by the caller).
 */
template <int MAXP, int MODE>
__device__ __noinline__ void fb_residual_pass(FbSearchShared<MAXP> &S, const int32_t *xs, int n, int order, int row,
                                              int psize, uint32_t maxabs, int32_t *res_out,
                                              unsigned long long *runsum)
{
    const unsigned long long pm = (unsigned long long)S.sumabs[row] * (unsigned long long)maxabs;
    const bool narrow = pm < 0x80000000ull &&
                        ((unsigned long long)maxabs + (pm >> S.shift[row]) + 1ull) < (1ull << 26);
    const int P = (order + 3) & ~3;
#define FB_CASE(PP)                                                                                         \
    case PP:                                                                                                \
        if (narrow) fb_tiles<MAXP, PP, false, MODE>(S, xs, n, order, row, psize, res_out, runsum);          \
        else        fb_tiles<MAXP, PP, true, MODE>(S, xs, n, order, row, psize, res_out, runsum);           \
        break;
    switch (P) {
        case 0:
        FB_CASE(4) FB_CASE(8) FB_CASE(12)
        default:
            if constexpr (MAXP > 12) {
                switch (P) {
                    FB_CASE(16) FB_CASE(20) FB_CASE(24) FB_CASE(28)
                    default:
                        if (narrow) fb_tiles<MAXP, 32, false, MODE>(S, xs, n, order, row, psize, res_out, runsum);
                        else        fb_tiles<MAXP, 32, true, MODE>(S, xs, n, order, row, psize, res_out, runsum);
                        break;
                }
            }
            break;
    }
#undef FB_CASE
}

/* the orders of a group's members, 8 bits each (orders <= 32): travels in a register, where an
 * int array would live in local memory */
typedef uint32_t FbOrders;
__device__ __forceinline__ int fb_order_of(FbOrders o, int m) { return (int)((o >> (8 * m)) & 0xffu); }
__device__ __forceinline__ FbOrders fb_order_put(FbOrders o, int m, int order) { return o | ((FbOrders)order << (8 * m)); }

/*
 * Costing pass of a whole group over the staged block: every run's window is loaded ONCE and
 * each member's predictor applied to it (one code path for all members: P covers the highest
 * order of the group, lower orders have zero coefficients above theirs).  Sums go to the
 * member's run-sum buffer.  The kernel is bound by instruction fetch and per-run latency, not by
 * the multiplies: the zero taps are free, the shared window and the single body are not.
 */
template <int MAXP, int P, bool WIDE, bool RS32>
__device__ __noinline__ void fb_tiles_group(FbSearchShared<MAXP> &S, const int32_t *xs, int n, int is_lpc, int count,
                                            FbOrders ord, unsigned long long *runsum0, int rstride)
{
    for (int i0 = (int)threadIdx.x * FB_RUN; i0 < n; i0 += (int)blockDim.x * FB_RUN) {
        if (i0 + FB_RUN > n) {                                    /* block tail: per member, per sample */
            for (int m = 0; m < count; m++)
                fb_run_tail<MAXP>(S, xs, n, fb_order_of(ord, m), is_lpc ? fb_order_of(ord, m) - 1 : fb_order_of(ord, m), n, i0, nullptr, runsum0 + m * rstride, FB_SUMS);
            continue;
        }
        int32_t w[P + FB_RUN];
#if FB_RUN == 16
        const fb_sptr xr = fb_to_sptr(xs + fb_skew(i0 + FB_HIST));
#else
        const fb_sptr xr = fb_to_sptr(xs);
#endif
        FbWindow<P, (P + FB_RUN) / 4>::load(xr, i0 + FB_HIST, w);
        if constexpr (WIDE && FB_F64_WIDE(MAXP)) {
            /* FP64 bodies: the window is converted once for all members of the group */
            double wd[P + FB_RUN - 1];
#pragma unroll
            for (int q = 0; q < P + FB_RUN - 1; q++) wd[q] = (double)w[q];
#pragma unroll 1
            for (int m = 0; m < count; m++) {
                const int order = fb_order_of(ord, m), row = is_lpc ? order - 1 : order;
                const double *cd = S.coefd[row];
                double pd[FB_RUN];
#pragma unroll
                for (int k = 0; k < FB_RUN; k++) pd[k] = 0.0;
#pragma unroll
                for (int j = 0; j < P; j++) {
                    const double cj = cd[j];
#pragma unroll
                    for (int k = 0; k < FB_RUN; k++) pd[k] = __fma_rn(cj, wd[P + k - 1 - j], pd[k]);
                }
                const int wlim = order - i0;                      /* samples of this run below the order */
                unsigned long long acc = 0;
#pragma unroll
                for (int k = 0; k < FB_RUN; k++) {
                    const int32_t rk = (int32_t)((uint32_t)w[P + k] - (uint32_t)fb_floor_lo32(pd[k]));
                    if (!(k < P && k < wlim)) acc += fb_zigzag(rk);  /* warm-up samples are not counted */
                }
                runsum0[m * rstride + i0 / FB_RUN] = acc;
            }
        } else {
#pragma unroll 1
            for (int m = 0; m < count; m++) {
                const int order = fb_order_of(ord, m), row = is_lpc ? order - 1 : order;
                int32_t c[P];
#pragma unroll
                for (int g = 0; g < P / 4; g++) {
                    const int4 v = *reinterpret_cast<const int4 *>(&S.coef[row][4 * g]);
                    c[4 * g] = v.x; c[4 * g + 1] = v.y; c[4 * g + 2] = v.z; c[4 * g + 3] = v.w;
                }
                const int shift = S.shift[row];
                const int wlim = order - i0;                      /* samples of this run below the order */
                unsigned long long acc = 0;
                uint32_t a32 = 0;
#pragma unroll
                for (int k = 0; k < FB_RUN; k++) {
                    if (WIDE) {
                        long long pred = 0;
#pragma unroll
                        for (int j = 0; j < P; j++) pred += (long long)c[j] * (long long)w[P + k - 1 - j];
                        const int32_t rk = (int32_t)((long long)w[P + k] - (pred >> shift));
                        acc += fb_zigzag(rk);
                        if (k < P && k < wlim) acc -= fb_zigzag(rk);       /* warm-up samples are not counted */
                    } else {
                        int32_t pred = 0;
#pragma unroll
                        for (int j = 0; j < P; j++) pred += c[j] * w[P + k - 1 - j];
                        /* zigzag(r) = (|4r + 1| - 1) / 2, |r| < 2^26: the sum of sixteen |4r + 1| fits 32 bits;
                         * |x| + acc is one instruction (VABSDIFF); a warm-up sample (order <= P) counts as r = 0.
                         * (Measured and dropped: the warm-up runs costed apart by sixteen threads per run so
                         * that this body does not know the case, 3.58 vs 3.16 ms.) */
                        int32_t rk = w[P + k] - (pred >> shift);
                        if (k < P && k < wlim) rk = 0;
                        a32 = __sad(4 * rk + 1, 0, a32);
                    }
                }
                /* RS32: one 32-bit word per run, the layout fb_finish_fast reads with 128-bit loads */
                if constexpr (RS32 && !WIDE) reinterpret_cast<uint32_t *>(runsum0)[m * 2 * rstride + i0 / FB_RUN] = (a32 - FB_RUN) >> 1;
                else runsum0[m * rstride + i0 / FB_RUN] = WIDE ? acc : (unsigned long long)((a32 - FB_RUN) >> 1);
            }
        }
    }
}

/* returns true when the run sums were left as 32-bit words (fb_finish_fast finishes) */
template <int MAXP>
__device__ __forceinline__ bool fb_residual_group(FbSearchShared<MAXP> &S, const int32_t *xs, int n, int is_lpc, int count,
                                                  FbOrders ord, bool fastable, unsigned long long *runsum0, int rstride)
{
    bool narrow = true, sum32 = true;
    int omax = 0;
    for (int m = 0; m < count; m++) {
        const int order = fb_order_of(ord, m);
        narrow = narrow && S.narrow_of[is_lpc ? order - 1 : order];
        sum32 = sum32 && S.sum32_of[is_lpc ? order - 1 : order];
        omax = max(omax, order);
    }
    /* one body per kernel for orders up to 12; the order-32 kernel keeps a body per 4 taps */
    const int P = (MAXP <= 12) ? 12 : ((omax + 3) & ~3);
#define FB_CASE(PP)                                                                                         \
    case PP:                                                                                                \
        if (narrow) fb_tiles_group<MAXP, PP, false, false>(S, xs, n, is_lpc, count, ord, runsum0, rstride); \
        else        fb_tiles_group<MAXP, PP, true, false>(S, xs, n, is_lpc, count, ord, runsum0, rstride);  \
        break;
    if constexpr (MAXP <= 12) {
        if (narrow && sum32 && fastable) {
            fb_tiles_group<MAXP, 12, false, true>(S, xs, n, is_lpc, count, ord, runsum0, rstride);
            return true;
        }
        switch (P) { default: FB_CASE(12) }
    } else {
        switch (P) {
            case 0:
            FB_CASE(4) FB_CASE(8) FB_CASE(12) FB_CASE(16) FB_CASE(20) FB_CASE(24) FB_CASE(28)
            default: FB_CASE(32)
        }
    }
#undef FB_CASE
    return false;
}

/* what a group evaluation needs to know about the subframe */
struct FbSearchCtx {
    const int32_t *xs;          /* staged plane (fast) */
    const int32_t *xg;          /* plane in global memory */
    int n, obits, pmin, pmax, is_lpc;
    uint32_t maxabs;
    bool fast, tileable;        /* tileable: every partition size is a whole number of runs */
    bool fastable;              /* fast, tileable, n a power of two from 512 to FB_FAST_MAX_N: fb_finish_fast applies */
};

/*
 * Cost `count` (<= FB_GROUP_OF(MAXP)) candidates of orders ord[]; totals land in S.result[0..count).
 * Every thread of the CTA calls it.  res_out != NULL (count == 1 only): the same pass also
 * stores the residual -- the orders that are not searched (optimize.c:196-204) need one pass.
 */
template <int MAXP>
__device__ __noinline__ void fb_eval_group(FbSearchShared<MAXP> &S, const FbSearchCtx X, int count, FbOrders ord,
                                           int32_t *res_out FB_PROF_ARG)
{
    const int tid = threadIdx.x;
    if (count <= 0) return;
    if (X.fast && X.tileable) {
        unsigned long long *runsum0 = reinterpret_cast<unsigned long long *>(const_cast<int32_t *>(X.xs) + fb_skew_words(X.n));
        const int rstride = fb_runsum_words(X.n) / 2;
        bool rs32 = false;
        if (res_out) {
            const int order = fb_order_of(ord, 0), row = X.is_lpc ? order - 1 : order;
            const int pmax = S.pmax_of[order];
            fb_residual_pass<MAXP, FB_SUMS | FB_STORE>(S, X.xs, X.n, order, row, X.n >> pmax, X.maxabs, res_out, runsum0);
        } else {
            rs32 = fb_residual_group<MAXP>(S, X.xs, X.n, X.is_lpc, count, ord, X.fastable, runsum0, rstride);
        }
        FB_PROF(1);
        __syncthreads();
        FB_PROF(2);
        for (int s = tid >> 5; s < count; s += (int)(blockDim.x >> 5)) {
            const int order = fb_order_of(ord, s);
            const int pmin = S.pmin_of[order], pmax = S.pmax_of[order];
            if (rs32) fb_finish_fast<MAXP>(S, s, reinterpret_cast<const uint32_t *>(runsum0) + s * 2 * rstride, X.n, X.is_lpc, order, X.obits, pmin, pmax);
            else if (X.fastable && !res_out) fb_finish_fast64<MAXP>(S, s, runsum0 + s * rstride, X.n, X.is_lpc, order, X.obits, pmin, pmax);
            else fb_finish_warp<MAXP>(S, s, runsum0 + s * rstride, (X.n >> pmax) / FB_RUN, X.n, X.is_lpc, order, X.obits, pmin, pmax);
        }
        FB_PROF(3);
        __syncthreads();
        FB_PROF(4);
        return;
    }
    for (int s = 0; s < count; s++) {
        const int order = fb_order_of(ord, s), row = X.is_lpc ? order - 1 : order;
        if (X.fast) {
            const int pmin = fb_limit_porder(X.pmin, X.n, order);
            const int pmax = fb_limit_porder(X.pmax, X.n, order);
            __syncthreads();
            for (int e = tid; e < (1 << pmax); e += (int)blockDim.x) S.sums[e] = 0;
            __syncthreads();
            if (res_out) fb_residual_pass<MAXP, FB_SUMS | FB_STORE>(S, X.xs, X.n, order, row, X.n >> pmax, X.maxabs, res_out, nullptr);
            else fb_residual_pass<MAXP, FB_SUMS>(S, X.xs, X.n, order, row, X.n >> pmax, X.maxabs, nullptr, nullptr);
            __syncthreads();
            if (tid < 32) fb_finish_warp<MAXP>(S, s, S.sums, 1, X.n, X.is_lpc, order, X.obits, pmin, pmax);
        } else {
            fb_evaluate<MAXP>(S, s, X.xg, X.n, X.is_lpc, order, row, X.obits, X.pmin, X.pmax, res_out, true);
        }
    }
    __syncthreads();
}

/* residual of the chosen predictor to global memory (optimize.c:266-275), no costing */
template <int MAXP>
__device__ __forceinline__ void fb_store_residual(FbSearchShared<MAXP> &S, const FbSearchCtx X, int order, int32_t *rg)
{
    const int row = X.is_lpc ? order - 1 : order;
    if (X.fast) fb_residual_pass<MAXP, FB_STORE>(S, X.xs, X.n, order, row, X.n, X.maxabs, rg, nullptr);
    else        fb_evaluate<MAXP>(S, 0, X.xg, X.n, X.is_lpc, order, row, X.obits, X.pmin, X.pmax, rg, false);
}

#ifndef FB_SEARCH_MINBLOCKS_WIDE
#define FB_SEARCH_MINBLOCKS_WIDE 3   /* the order-32 bodies hold a 48-sample window and 32 coefficients */
#endif
#ifndef FB_SEARCH_MINBLOCKS
#define FB_SEARCH_MINBLOCKS 4    /* <= 128 registers, four CTAs per SM: 3.69 ms per C2 stream; 5 -> 3.79, 6 -> 3.97 (spills) */
#endif

#define FB_SEARCH_ANY 0          /* everything decided at run time */
#define FB_SEARCH_CD  1          /* stereo, packed 16-bit, LPC, log order search */

template <int MAXP, int SPEC>
__global__ void __launch_bounds__(FB_SEARCH_THREADS, (MAXP > 12 ? FB_SEARCH_MINBLOCKS_WIDE : FB_SEARCH_MINBLOCKS))
k_search(FbConfig cfg, const FbFrame *frames, const uint32_t *nframes, const void *pcm, int fmt,
         unsigned long long pcm_bytes, const uint8_t *ch_modes, int32_t *plane_scratch,
         int32_t *res, FbSub *subs, const int32_t *coefs, const int32_t *shifts, const FbPlanNode *plan,
         int smem_ints)
{
    FB_DYN_SMEM(dyn);
    __shared__ __align__(16) FbSearchShared<MAXP> S;
    __shared__ fb_mbar_t s_bar[FB_TILE_STAGES];

    /* SPEC: what the instantiation knows at compile time.  The kernel is bound by instruction
     * fetch before anything else, and the generic one carries every PCM layout, channel count,
     * predictor type and order search; FB_SEARCH_CD is the CD-audio shape of the presets 8, 9
     * and 11 (stereo, packed s16le, LPC, log search): 13.5 k -> 12.3 k SASS instructions,
     * 3.22 -> 3.04 ms per C2 stream. */
    if (SPEC == FB_SEARCH_CD) { fmt = FB_PCM_S16LE; cfg.channels = 2; cfg.order_method = 6; cfg.prediction_type = 2; }
    const int C = SPEC == FB_SEARCH_CD ? 2 : cfg.channels;
    const uint32_t sf = blockIdx.x;
    const uint32_t f = sf / (uint32_t)C;
    const int c = (int)(sf % (uint32_t)C);
    /* the grid never exceeds the frame table (engine.cu sizes both for the same upper bound), so the
     * records are requested together with the frame count instead of one round trip later */
    FbSub *sb = &subs[sf];
    const uint32_t nf = *nframes;
    const FbFrame fr = frames[f];
    const int sb_obits = sb->obits, sb_const = sb->is_const, sb_wasted = sb->wasted;
    const uint32_t sb_maxabs = sb->maxabs;
    const int mode = ch_modes[f];
    if (f >= nf) return;
    const int n = (int)fr.n;
    const int tid = threadIdx.x, T = blockDim.x;
    FB_PROF_DECL;
    const size_t ebase = (size_t)fr.start * C;               /* interleaved element index of the frame's sample 0 */
    const size_t off = ebase + (size_t)c * n;
    int32_t *rg = res + off;
    const int obits = sb_obits;

    /* CONSTANT, optimize.c:143-151 */
    if (sb_const) {
        if (tid == 0) { sb->type = 0; sb->order = 0; sb->est_bits = (uint32_t)obits; }
        return;
    }
    /* VERBATIM, optimize.c:154-158 */
    if (n < 5 || cfg.prediction_type == 0) {
        if (tid == 0) { sb->type = 1; sb->order = 0; sb->est_bits = (uint32_t)(obits * n); }
        return;
    }

    /*
     * Stage the subframe's plane.  Mono / stereo: the frame's packed PCM comes in through the TMA
     * tile ring (cp.async.bulk + mbarrier, two chunks in flight) and is deinterleaved, decorrelated
     * by the frame's stereo decision and shifted by the wasted bits on its way into the skewed
     * int32 layout (FB_HIST zero samples in front) that the run bodies read with 128-bit loads
     * (encode.c:541-553, 648-694, 558-593 fused into the load).  More channels: k_prep's int32
     * plane of this channel, asynchronous 16-byte copies straight into place.
     */
    const bool fast = fb_search_smem_words(n, FB_GROUP_OF(MAXP)) <= smem_ints;
    const bool from_planes = fb_uses_planes(C);
    int32_t *xs = (int32_t *)dyn;
    const int32_t *xg = from_planes ? plane_scratch + off : nullptr;
    FbTile tile;
    if (fast && !from_planes)                                 /* the frame's PCM is requested first ... */
        fb_tile_begin(tile, pcm, fmt, ebase, n, C, pcm_bytes, (uint8_t *)(xs + fb_skew_words(n)), s_bar);
    if (fast && from_planes) {
        for (int L = tid; L < FB_HIST; L += T) xs[fb_skew(L)] = 0;
        const int n4 = (((size_t)xg) & 15u) == 0 ? (n & ~3) : 0;
        /* 16-byte groups never straddle a skew pad */
        for (int i = 4 * tid; i < n4; i += 4 * T) fb_cp_async16(xs + fb_skew(i + FB_HIST), xg + i);
#pragma unroll 8
        for (int i = n4 + tid; i < n; i += T) xs[fb_skew(i + FB_HIST)] = xg[i];
        for (int i = n + tid; i < n + 16; i += T) xs[fb_skew(i + FB_HIST)] = 0;
    }
    int min_order = cfg.min_order, max_order = cfg.max_order;
    const bool fixed = (cfg.prediction_type == 1 || n <= max_order);

    FbSearchCtx X;
    X.xs = xs; X.xg = xg; X.n = n; X.obits = obits;
    X.pmin = cfg.min_porder; X.pmax = cfg.max_porder; X.is_lpc = fixed ? 0 : 1;
    X.maxabs = sb_maxabs; X.fast = fast;
    {
        /* the finest partition any candidate can use: larger ones are multiples of it */
        const int pfin = fb_limit_porder(cfg.max_porder, n, 0);
        X.tileable = ((n >> pfin) % FB_RUN) == 0;
        X.fastable = FB_RUN == 16 && fast && X.tileable && (n & (n - 1)) == 0 && n >= 512 && n <= FB_FAST_MAX_N;
    }

    /* candidate rows: binomial coefficients (optimize.c:44-66 is LPC with shift 0) or the
     * quantised rows of k_lpc; sum |c| per row for the 32-bit exactness test */
    if (fixed) {
        if (tid < 5) {
            const int32_t bc[5][4] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {2, -1, 0, 0}, {3, -3, 1, 0}, {4, -6, 4, -1}};
            const uint32_t sa[5] = {0, 1, 3, 7, 15};
            for (int j = 0; j < MAXP; j++) {                 /* bodies may cover more taps */
                S.coef[tid][j] = j < 4 ? bc[tid][j] : 0;
                if (FB_F64_WIDE(MAXP)) S.coefd[tid][j] = j < 4 ? (double)bc[tid][j] : 0.0;
            }
            S.shift[tid] = 0;
            S.sumabs[tid] = sa[tid];
            const unsigned long long rb = (unsigned long long)sb_maxabs + (unsigned long long)sa[tid] * sb_maxabs + 1ull;
            S.narrow_of[tid] = (unsigned long long)sa[tid] * sb_maxabs < 0x80000000ull && rb < (1ull << 26);
            S.sum32_of[tid] = 2ull * rb * (unsigned long long)n < 0x80000000ull;
        }
    } else {
        const int32_t *co = coefs + (size_t)sf * FB_MAX_ORDER * FB_MAX_ORDER;
        const int32_t *so = shifts + (size_t)sf * FB_MAX_ORDER;
        for (int e = tid; e < MAXP * MAXP; e += T) {
            const int rowi = e / MAXP, j = e % MAXP;
            S.coef[rowi][j] = (rowi < max_order && j <= rowi) ? co[rowi * FB_MAX_ORDER + j] : 0;
        }
        /* shift | sum |c| << 8 per row (k_lpc, fb_store_row); the 32-bit exactness tests right here, in front of
         * the staging barrier */
        for (int rowi = tid; rowi < MAXP; rowi += T) {
            const uint32_t v = rowi < max_order ? (uint32_t)so[rowi] : 0u;
            const int sh = (int)(v & 0xffu);
            const uint32_t sa = v >> 8;                                 /* <= 32 * 16383 */
            S.shift[rowi] = sh;
            S.sumabs[rowi] = sa;
            const unsigned long long pm = (unsigned long long)sa * (unsigned long long)sb_maxabs;
            const unsigned long long rb = (unsigned long long)sb_maxabs + (pm >> sh) + 1ull;   /* bound on |residual| */
            S.narrow_of[rowi] = pm < 0x80000000ull && rb < (1ull << 26);
            S.sum32_of[rowi] = 2ull * rb * (unsigned long long)n < 0x80000000ull;   /* zig-zag <= 2 rb: every partition sum < 2^31 */
        }
    }
    if (cfg.order_method == 6 && tid < (int)(FB_PLAN_SMEM_NODES * sizeof(FbPlanNode) / 4))   /* engine.cu pads the plan */
        reinterpret_cast<uint32_t *>(S.plan)[tid] = reinterpret_cast<const uint32_t *>(plan)[tid];
    for (int o = tid; o <= MAXP; o += T) {
        S.pmin_of[o] = (uint8_t)fb_limit_porder(cfg.min_porder, n, o);
        S.pmax_of[o] = (uint8_t)fb_limit_porder(cfg.max_porder, n, o);
    }
    /* ... and unpacked after the candidate rows have been requested, so that both latencies overlap */
    if (fast && from_planes) {
        fb_cp_async_wait_all();
    } else if (fast) {
        for (int L = tid; L < FB_HIST; L += T) xs[fb_skew(L)] = 0;
        for (int i = ((n + 3) & ~3) + tid; i < n + 16; i += T) xs[fb_skew(i + FB_HIST)] = 0;
        for (uint32_t k = 0; k < tile.nchunks; k++) {
            uint32_t nb;
            const uint8_t *t = fb_tile_acquire(tile, k, &nb);
            const int base = (int)(k * tile.chunk_samples);
            const int cs = min((int)tile.chunk_samples, n - base);
            const int cs4 = cs & ~3;                                  /* whole groups of four samples */
            const int coef = C == 2 ? fb_stereo_coef(mode, c, sb_wasted) : 0;
            int32_t *dst = xs + fb_skew(base + FB_HIST);               /* base is a multiple of 16 */
            if (C == 2 && fmt == FB_PCM_S16LE) {
                /* a 128-bit load brings four (left, right) pairs; one dp2a + one shift per sample */
                for (int q = 4 * tid; q < cs4; q += 4 * T) {
                    const uint4 w = *reinterpret_cast<const uint4 *>(t + 4 * q);
                    *reinterpret_cast<int4 *>(dst + fb_skew(q)) =
                        make_int4(fb_stereo_apply16(coef, w.x), fb_stereo_apply16(coef, w.y),
                                  fb_stereo_apply16(coef, w.z), fb_stereo_apply16(coef, w.w));
                }
            } else if (C == 2) {
                for (int q = 4 * tid; q < cs4; q += 4 * T) {
                    int32_t l[4], r[4];
                    fb_tile_stereo4(t, fmt, (uint32_t)q, l, r);
                    *reinterpret_cast<int4 *>(dst + fb_skew(q)) =
                        make_int4(fb_stereo_apply(coef, l[0], r[0]), fb_stereo_apply(coef, l[1], r[1]),
                                  fb_stereo_apply(coef, l[2], r[2]), fb_stereo_apply(coef, l[3], r[3]));
                }
            } else {
                for (int q = 4 * tid; q < cs4; q += 4 * T) {
                    int32_t v[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) v[j] = fb_tile_elem(t, fmt, (uint32_t)(q + j) * (uint32_t)C + (uint32_t)c) >> sb_wasted;
                    *reinterpret_cast<int4 *>(dst + fb_skew(q)) = make_int4(v[0], v[1], v[2], v[3]);
                }
            }
            /* the last one to three samples of the block; the rest of their group is zero */
            if (tid < 4 && cs4 < cs) {
                const int q = cs4 + tid;
                int32_t v = 0;
                if (q < cs) {
                    if (C == 2) v = fb_stereo_apply(coef, fb_tile_elem(t, fmt, 2u * (uint32_t)q), fb_tile_elem(t, fmt, 2u * (uint32_t)q + 1u));
                    else v = fb_tile_elem(t, fmt, (uint32_t)q * (uint32_t)C + (uint32_t)c) >> sb_wasted;
                }
                dst[fb_skew(q)] = v;
            }
            fb_tile_release(tile, k, false);     /* chunks land in disjoint parts of the plane */
        }
    } else if (!from_planes) {
        /* blocks too large for shared memory (correctness path): the plane is materialised in a
         * global scratch by this CTA and read back through generic pointers */
        int32_t *pl = plane_scratch + off;
        for (int i = tid; i < n; i += T) pl[i] = fb_pcm_sample(pcm, fmt, ebase, C, c, mode, sb_wasted, i);
        xg = pl; X.xg = pl;
        __syncthreads();
    }

    __syncthreads();
    if (!fixed && FB_F64_WIDE(MAXP)) {
        /* the rows for the FP64 bodies: coefficient * 2^-shift, exact */
        for (int e = tid; e < MAXP * MAXP; e += T) {
            const int rowi = e / MAXP, j = e % MAXP;
            S.coefd[rowi][j] = (double)S.coef[rowi][j] * fb_exp2_neg(S.shift[rowi]);
        }
        __syncthreads();
    }
    FB_PROF(0);

    FbOrders ord;

    /* FIXED, optimize.c:168-190: row = order (row 0 is the all-zero order-0 predictor) */
    if (fixed) {
        if (max_order > 4) max_order = 4;
        int opt = min_order;
        uint32_t best = 0xffffffffu;
        for (int base = min_order; base <= max_order; base += FB_GROUP_OF(MAXP)) {
            int cnt = 0;
            ord = 0;
            for (int i = base; i <= max_order && cnt < FB_GROUP_OF(MAXP); i++) ord = fb_order_put(ord, cnt++, i);
            fb_eval_group<MAXP>(S, X, cnt, ord, nullptr FB_PROF_PASS);
            int bs = -1;
            for (int s = 0; s < cnt; s++) {
                const uint32_t b = S.result[s];
                if (b < best) { best = b; opt = fb_order_of(ord, s); bs = s; }
            }
            if (bs >= 0) fb_keep_best<MAXP>(S, bs, best);       /* one copy per group: only the last winner matters */
        }
        if (best == 0xffffffffu) {      /* min_order > 4 with a tiny last block: undefined in the reference */
            opt = opt > 4 ? 4 : opt;
            ord = (FbOrders)opt;
            fb_eval_group<MAXP>(S, X, 1, ord, rg FB_PROF_PASS);
            fb_keep_best<MAXP>(S, 0, S.result[0]);
        } else {
            fb_store_residual<MAXP>(S, X, opt, rg);
        }
        if (tid == 0) { sb->type = 8; sb->order = opt; }
        fb_store_best<MAXP>(S, sb);
        return;
    }

    /* LPC, optimize.c:193-275.  Orders are 0-based indices (order - 1) while searching. */
    const int om = cfg.order_method;
    int opt_order;
    uint32_t best = 0xffffffffu;

    if (om == 0) {
        opt_order = max_order - 1;
    } else if (om == 1) {
        opt_order = sb->est_order - 1;
    } else if (om >= 2 && om <= 4) {
        /* optimize.c:205-222: `levels` fixed orders, highest first */
        const int levels = 1 << (om - 1);
        opt_order = max_order - 1;
        for (int base = levels - 1; base >= 0; base -= FB_GROUP_OF(MAXP)) {
            int cnt = 0;
            ord = 0;
            for (int i = base; i >= 0 && cnt < FB_GROUP_OF(MAXP); i--) {
                int order = min_order + (((max_order - min_order + 1) * (i + 1)) / levels) - 2;
                if (order < 0) order = 0;
                ord = fb_order_put(ord, cnt, order + 1); cnt++;
            }
            fb_eval_group<MAXP>(S, X, cnt, ord, nullptr FB_PROF_PASS);
            int bs = -1;
            for (int s = 0; s < cnt; s++) {
                const uint32_t b = S.result[s];
                if (b < best) { best = b; opt_order = fb_order_of(ord, s) - 1; bs = s; }
            }
            if (bs >= 0) fb_keep_best<MAXP>(S, bs, best);
        }
    } else if (om == 5) {
        /* optimize.c:223-240: every order */
        opt_order = 0;
        for (int base = 0; base < max_order; base += FB_GROUP_OF(MAXP)) {
            int cnt = 0;
            ord = 0;
            for (int i = base; i < max_order && cnt < FB_GROUP_OF(MAXP); i++) ord = fb_order_put(ord, cnt++, i + 1);
            fb_eval_group<MAXP>(S, X, cnt, ord, nullptr FB_PROF_PASS);
            int bs = -1;
            for (int s = 0; s < cnt; s++) {
                const uint32_t b = S.result[s];
                if (b < best) { best = b; opt_order = base + s; bs = s; }
            }
            if (bs >= 0) fb_keep_best<MAXP>(S, bs, best);
        }
    } else {
        /* log search, optimize.c:241-261.  A step's candidates are last-step, last, last+step; which
         * orders are costed next depends only on the decisions so far, so the decision tree comes
         * tabulated from the host (FbPlanNode, engine.h): cost the node's group, take the first
         * strict minimum over its members IN THEIR ORDER, follow the child of the winner.  That is
         * the reference's replay: the planner merges steps into a node only while the new
         * candidates of every step are the same whichever order is the best by then, and lists
         * them step by step, ascending inside a step -- the sequence in which optimize.c:249-258
         * compares them with `<` (the walk over steps and neighbours that used to stand here, with
         * its done-mask and its search for each candidate's slot, was 8 % of a CTA's life). */
        uint32_t node = 0;
        opt_order = (int)plan[0].start_order;
        while (node != FB_PLAN_END) {
            const FbPlanNode nd = node < FB_PLAN_SMEM_NODES ? S.plan[node] : plan[node];
            const int cnt = (int)nd.cnt;
            ord = nd.ord;
            FB_PROF(5);
            fb_eval_group<MAXP>(S, X, cnt, ord, nullptr FB_PROF_PASS);
            int bs = -1;
#pragma unroll
            for (int s = 0; s < FB_GROUP_OF(MAXP); s++) {
                if (s < cnt) {
                    const uint32_t b = S.result[s];
                    if (b < best) { best = b; opt_order = fb_order_of(ord, s) - 1; bs = s; }
                }
            }
            if (bs >= 0) fb_keep_best<MAXP>(S, bs, best);
            node = nd.child[bs + 1];
            FB_PROF(6);
        }
    }

    /* final pass for the chosen order, optimize.c:266-275: the costing of a searched order is
     * already known, only its residual is missing */
    {
        const int idx = opt_order;
        if (tid < FB_MAX_ORDER) sb->coefs[tid] = (tid <= idx && tid < MAXP) ? S.coef[idx][tid] : 0;
        if (tid == 0) { sb->type = 32; sb->order = idx + 1; sb->shift = S.shift[idx]; }
        if (best == 0xffffffffu) {                       /* no search ran (or nothing beat 2^32-1) */
            ord = (FbOrders)(idx + 1);
            fb_eval_group<MAXP>(S, X, 1, ord, rg FB_PROF_PASS);
            fb_keep_best<MAXP>(S, 0, S.result[0]);
        } else {
            fb_store_residual<MAXP>(S, X, idx + 1, rg);
        }
        FB_PROF(7);
        fb_store_best<MAXP>(S, sb);
        FB_PROF(8);
#ifdef FB_SEARCH_PROF
        if (tid == 0) atomicAdd(&g_sprof[15], 1ull);
#endif
    }
}

#endif
