/*
 * flake_oracle.c -- CPU restatement of Flake's FLAC encoding hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see flake_oracle.h).  Written from the reference's
 * documented behaviour as an independent program: stages
 * are separated the way the GPU pipeline separates them (prepare -> LPC
 * analysis -> order/Rice search -> pack) so that each stage's output can be
 * compared with the CUDA kernels'.
 *
 * Compile with -ffp-contract=off: the reference is built -std=c99 for baseline
 * x86-64 (scalar SSE2 doubles, no FMA), and the FP64 stage must round the same
 * way.
 */
#include "flake_oracle.h"

#include <stdlib.h>
#include <string.h>
#include <math.h>

/* ------------------------------------------------------------------ */
/* small helpers                                                      */
/* ------------------------------------------------------------------ */

/* floor(log2(v)), 0 for v==0 -- common.h:53-66 */
static int ilog2_u32(uint32_t v)
{
    int r = 0;
    while (v >>= 1) r++;
    return r;
}

static int count_trailing_zeros(uint32_t v)   /* v != 0 */
{
    int c = 0;
    while (!(v & 1)) { v >>= 1; c++; }
    return c;
}

/* ------------------------------------------------------------------ */
/* MSB-first bit sink (bitio.h:33-141 semantics, own implementation)   */
/* ------------------------------------------------------------------ */
typedef struct {
    uint8_t *buf;
    size_t cap;
    size_t pos;      /* whole bytes emitted */
    uint64_t acc;    /* pending bits, right aligned */
    int nacc;        /* number of pending bits (< 8 after every put) */
    int overflow;
} BitSink;

static void bs_init(BitSink *b, uint8_t *buf, size_t cap)
{
    b->buf = buf; b->cap = cap; b->pos = 0; b->acc = 0; b->nacc = 0;
    b->overflow = 0;
}

static void bs_put(BitSink *b, int nbits, uint32_t val)
{
    if (nbits <= 0) return;
    if (nbits < 32) val &= (1u << nbits) - 1u;
    b->acc = (b->acc << nbits) | val;
    b->nacc += nbits;
    while (b->nacc >= 8) {
        uint8_t byte = (uint8_t)(b->acc >> (b->nacc - 8));
        if (b->pos < b->cap) b->buf[b->pos] = byte; else b->overflow = 1;
        b->pos++;
        b->nacc -= 8;
    }
    b->acc &= (1ull << b->nacc) - 1ull;
}

static void bs_zeros(BitSink *b, uint32_t count)
{
    while (count >= 24) { bs_put(b, 24, 0); count -= 24; }
    if (count) bs_put(b, (int)count, 0);
}

static void bs_align(BitSink *b)
{
    if (b->nacc) bs_put(b, 8 - b->nacc, 0);
}

/* ------------------------------------------------------------------ */
/* CRC-8 / CRC-16 -- crc.c:24-92 (MSB first, init 0, no xor-out)       */
/* ------------------------------------------------------------------ */
uint8_t orc_crc8(const uint8_t *d, size_t n)
{
    uint8_t c = 0;
    for (size_t i = 0; i < n; i++) {
        c ^= d[i];
        for (int b = 0; b < 8; b++)
            c = (c & 0x80) ? (uint8_t)((c << 1) ^ 0x07) : (uint8_t)(c << 1);
    }
    return c;
}

uint16_t orc_crc16(const uint8_t *d, size_t n)
{
    uint16_t c = 0;
    for (size_t i = 0; i < n; i++) {
        c ^= (uint16_t)(d[i] << 8);
        for (int b = 0; b < 8; b++)
            c = (c & 0x8000) ? (uint16_t)((c << 1) ^ 0x8005) : (uint16_t)(c << 1);
    }
    return c;
}

/* ------------------------------------------------------------------ */
/* MD5 (RFC 1321), own implementation; md5.c                           */
/* ------------------------------------------------------------------ */
typedef struct { uint32_t h[4]; uint64_t nbytes; uint8_t tail[64]; } OrcMd5;

static const uint32_t md5_k[64] = {
    0xd76aa478,0xe8c7b756,0x242070db,0xc1bdceee,0xf57c0faf,0x4787c62a,0xa8304613,0xfd469501,
    0x698098d8,0x8b44f7af,0xffff5bb1,0x895cd7be,0x6b901122,0xfd987193,0xa679438e,0x49b40821,
    0xf61e2562,0xc040b340,0x265e5a51,0xe9b6c7aa,0xd62f105d,0x02441453,0xd8a1e681,0xe7d3fbc8,
    0x21e1cde6,0xc33707d6,0xf4d50d87,0x455a14ed,0xa9e3e905,0xfcefa3f8,0x676f02d9,0x8d2a4c8a,
    0xfffa3942,0x8771f681,0x6d9d6122,0xfde5380c,0xa4beea44,0x4bdecfa9,0xf6bb4b60,0xbebfbc70,
    0x289b7ec6,0xeaa127fa,0xd4ef3085,0x04881d05,0xd9d4d039,0xe6db99e5,0x1fa27cf8,0xc4ac5665,
    0xf4292244,0x432aff97,0xab9423a7,0xfc93a039,0x655b59c3,0x8f0ccc92,0xffeff47d,0x85845dd1,
    0x6fa87e4f,0xfe2ce6e0,0xa3014314,0x4e0811a1,0xf7537e82,0xbd3af235,0x2ad7d2bb,0xeb86d391 };
static const uint8_t md5_s[64] = {
    7,12,17,22,7,12,17,22,7,12,17,22,7,12,17,22, 5,9,14,20,5,9,14,20,5,9,14,20,5,9,14,20,
    4,11,16,23,4,11,16,23,4,11,16,23,4,11,16,23, 6,10,15,21,6,10,15,21,6,10,15,21,6,10,15,21 };

static void md5_block(uint32_t h[4], const uint8_t *p)
{
    uint32_t w[16], a = h[0], b = h[1], c = h[2], d = h[3];
    for (int i = 0; i < 16; i++)
        w[i] = (uint32_t)p[4*i] | ((uint32_t)p[4*i+1] << 8) |
               ((uint32_t)p[4*i+2] << 16) | ((uint32_t)p[4*i+3] << 24);
    for (int i = 0; i < 64; i++) {
        uint32_t f; int g;
        if (i < 16)      { f = (b & c) | (~b & d); g = i; }
        else if (i < 32) { f = (d & b) | (~d & c); g = (5*i + 1) & 15; }
        else if (i < 48) { f = b ^ c ^ d;          g = (3*i + 5) & 15; }
        else             { f = c ^ (b | ~d);       g = (7*i) & 15; }
        uint32_t t = a + f + md5_k[i] + w[g];
        a = d; d = c; c = b;
        b = b + ((t << md5_s[i]) | (t >> (32 - md5_s[i])));
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d;
}

static void md5_begin(OrcMd5 *m)
{
    m->h[0] = 0x67452301; m->h[1] = 0xefcdab89; m->h[2] = 0x98badcfe; m->h[3] = 0x10325476;
    m->nbytes = 0;
}

static void md5_feed(OrcMd5 *m, const uint8_t *p, size_t n)
{
    size_t used = (size_t)(m->nbytes & 63);
    m->nbytes += n;
    if (used) {
        size_t room = 64 - used;
        if (n < room) { memcpy(m->tail + used, p, n); return; }
        memcpy(m->tail + used, p, room);
        md5_block(m->h, m->tail);
        p += room; n -= room;
    }
    while (n >= 64) { md5_block(m->h, p); p += 64; n -= 64; }
    memcpy(m->tail, p, n);
}

static void md5_end(OrcMd5 *m, uint8_t out[16])
{
    uint64_t bits = m->nbytes << 3;
    size_t used = (size_t)(m->nbytes & 63);
    m->tail[used++] = 0x80;
    if (used > 56) { memset(m->tail + used, 0, 64 - used); md5_block(m->h, m->tail); used = 0; }
    memset(m->tail + used, 0, 56 - used);
    for (int i = 0; i < 8; i++) m->tail[56 + i] = (uint8_t)(bits >> (8*i));
    md5_block(m->h, m->tail);
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) out[4*i + j] = (uint8_t)(m->h[i] >> (8*j));
}

/* md5.c:281-320: little-endian, ceil(bps/8) bytes per sample, interleaved */
void orc_md5_pcm(const int32_t *pcm, int channels, int bps, uint64_t nsamples,
                 uint8_t digest[16])
{
    OrcMd5 m; md5_begin(&m);
    int bytes = (bps + 7) >> 3;
    uint8_t chunk[4096 * 4];
    uint64_t total = nsamples * (uint64_t)channels, i = 0;
    while (i < total) {
        size_t k = 0;
        while (i < total && k + 4 <= sizeof chunk) {
            uint32_t x = (uint32_t)pcm[i++];
            for (int b = 0; b < bytes; b++) { chunk[k++] = (uint8_t)x; x >>= 8; }
        }
        md5_feed(&m, chunk, k);
    }
    md5_end(&m, digest);
}

/* flake_encode_init calls write_headers() BEFORE md5_init() on a calloc'ed
 * context (encode.c:391, 458-469), so the provisional STREAMINFO carries the
 * digest of an all-zero MD5 state finalised with zero length. */
void orc_md5_zero_ctx(uint8_t digest[16])
{
    OrcMd5 m; memset(&m, 0, sizeof m);
    md5_end(&m, digest);
}

/* ------------------------------------------------------------------ */
/* parameters                                                          */
/* ------------------------------------------------------------------ */
int orc_set_defaults(OrcParams *p, int level)
{
    /* rows of the table in SURVEY.md 5.1 == encode.c:171-263 */
    static const struct { int bs, ptype, omin, omax, ometh, pomax, stereo, vbs; } T[13] = {
        /*0*/ {1152, 1, 2,  2, 1, 3, 0, 0},
        /*1*/ {1152, 1, 2,  4, 1, 3, 1, 0},
        /*2*/ {1152, 1, 0,  4, 1, 3, 1, 0},
        /*3*/ {4096, 2, 1,  6, 1, 4, 0, 0},
        /*4*/ {4096, 2, 1,  8, 1, 4, 1, 0},
        /*5*/ {4096, 2, 1,  8, 1, 5, 1, 0},
        /*6*/ {4096, 2, 1,  8, 1, 6, 1, 0},
        /*7*/ {4096, 2, 1,  8, 3, 6, 1, 0},
        /*8*/ {4096, 2, 1, 12, 6, 6, 1, 0},
        /*9*/ {4096, 2, 1, 12, 6, 8, 1, 1},
        /*10*/{4096, 2, 1, 12, 5, 8, 1, 1},
        /*11*/{8192, 2, 1, 32, 6, 8, 1, 1},
        /*12*/{8192, 2, 1, 32, 5, 8, 1, 1},
    };
    if (!p || level < 0 || level > 12) return -1;
    p->block_size = T[level].bs;
    p->prediction_type = T[level].ptype;
    p->min_order = T[level].omin;
    p->max_order = T[level].omax;
    p->order_method = T[level].ometh;
    p->min_porder = 0;
    p->max_porder = T[level].pomax;
    p->stereo_method = T[level].stereo;
    p->variable_block_size = T[level].vbs;
    p->allow_vbs = T[level].vbs;
    p->padding_size = 8192;
    return 0;
}

int orc_validate(const OrcParams *p)
{
    int subset = 0;
    if (!p) return -1;
    if (p->channels < 1 || p->channels > 8) return -1;
    if (p->sample_rate < 1 || p->sample_rate > 655350) return -1;
    if (p->bps < 4 || p->bps > 32) return -1;
    if (p->bps < 8 || p->bps > 24 || (p->bps % 4)) subset = 1;
    if (p->order_method < 0 || p->order_method > 6) return -1;
    if (p->stereo_method < 0 || p->stereo_method > 1) return -1;
    if (p->block_size < 16 || p->block_size > 65535) return -1;
    if (p->sample_rate <= 48000 && p->block_size > 4608) subset = 1;
    if (p->prediction_type < 0 || p->prediction_type > 2) return -1;
    if (p->min_order > p->max_order) return -1;
    if (p->prediction_type == 1) {
        if (p->min_order < 0 || p->min_order > 4) return -1;
        if (p->max_order < 0 || p->max_order > 4) return -1;
    } else {
        if (p->min_order < 1 || p->min_order > 32) return -1;
        if (p->max_order < 1 || p->max_order > 32) return -1;
        if (p->sample_rate <= 48000 && p->max_order > 12) subset = 1;
    }
    if (p->min_porder > p->max_porder) return -1;
    if (p->min_porder < 0 || p->min_porder > 8) return -1;
    if (p->max_porder < 0 || p->max_porder > 8) return -1;
    if (p->padding_size < 0 || p->padding_size >= (1 << 24)) return -1;
    if (p->variable_block_size < 0 || p->variable_block_size > 1) return -1;
    if (p->variable_block_size > 0 && !p->allow_vbs) return -1;
    if (p->block_size < 128 && p->allow_vbs) return -1;
    return subset;
}

/* ------------------------------------------------------------------ */
/* stage 2: FP64 LPC analysis -- lpc.c                                  */
/* ------------------------------------------------------------------ */

/* lpc.c:28-71.  The odd-length centre sample is uninitialised heap memory in
 * the reference; here it is 0.0 (parity-exempt, SURVEY.md 8c hazard 1). */
void orc_autocorr(const int32_t *smp, int n, int lag, double *autoc)
{
    double *w = (double *)calloc((size_t)n + 16, sizeof(double));
    double c = (2.0 / (n - 1.0)) - 1.0;
    for (int i = 0; i < (n >> 1); i++) {
        double d = c - i;
        double win = 1.0 - (d * d);
        w[i] = smp[i] * win;
        w[n - 1 - i] = smp[n - 1 - i] * win;
    }
    w[n] = 0.0;
    for (int i = 0; i <= lag; i++) {
        double s0 = 1.0, s1 = 1.0;
        for (int j = 0; j <= lag - i; j++)
            s0 += w[j + i] * w[j];
        for (int j = lag + 1; j <= n - 1; j += 2) {
            s0 += w[j] * w[j - i];
            s1 += w[j + 1] * w[j + 1 - i];
        }
        autoc[i] = s0 + s1;
    }
    free(w);
}

/* Levinson-Durbin, all orders -- lpc.c:77-117.  refl != NULL: consume given
 * reflection coefficients instead (the EST path). */
static void levinson(const double *autoc, int max_order, const double *refl,
                     double lpc[ORC_MAX_ORDER][ORC_MAX_ORDER])
{
    double a[ORC_MAX_ORDER];
    double err = autoc ? autoc[0] : 1.0;
    memset(a, 0, sizeof a);
    for (int i = 0; i < max_order; i++) {
        double r;
        if (refl) {
            r = refl[i];
        } else {
            r = -autoc[i + 1];
            for (int j = 0; j < i; j++)
                r -= a[j] * autoc[i - j];
            r /= err;
            err *= 1.0 - (r * r);
        }
        a[i] = r;
        int half = i >> 1, j;
        for (j = 0; j < half; j++) {
            double t = a[j];
            a[j] += r * a[i - 1 - j];
            a[i - 1 - j] += r * t;
        }
        if (i & 1)
            a[j] += a[j] * r;
        for (j = 0; j <= i; j++)
            lpc[i][j] = -a[j];
    }
}

/* Schur recursion + order estimate -- lpc.c:125-162 */
static int schur_estimate(const double *autoc, int max_order,
                          double lpc[ORC_MAX_ORDER][ORC_MAX_ORDER])
{
    double g0[ORC_MAX_ORDER], g1[ORC_MAX_ORDER], refl[ORC_MAX_ORDER], e;
    for (int i = 0; i < max_order; i++) g0[i] = g1[i] = autoc[i + 1];
    e = autoc[0];
    refl[0] = -g1[0] / e;
    e += g1[0] * refl[0];
    for (int i = 1; i < max_order; i++) {
        for (int j = 0; j < max_order - i; j++) {
            g1[j] = g1[j + 1] + refl[i - 1] * g0[j];
            g0[j] = g1[j + 1] * refl[i - 1] + g0[j];
        }
        refl[i] = -g1[0] / e;
        e += g1[0] * refl[i];
    }
    int est = 1;
    for (int i = max_order - 1; i >= 0; i--)
        if (fabs(refl[i]) > 0.10) { est = i + 1; break; }
    levinson(NULL, est, refl, lpc);
    return est;
}

/* x86-64 cvttsd2si: out-of-range and NaN give INT_MIN */
static int32_t trunc_to_int(double x)
{
    if (!(x > -2147483649.0 && x < 2147483648.0)) return INT32_MIN;
    return (int32_t)x;
}

/* lpc.c:167-219, precision fixed at 15 (encode.c:443) */
static void quantize(double *in, int order, int32_t *out, int *shift)
{
    const int32_t qmax = (1 << 14) - 1;
    double cmax = 0.0;
    for (int i = 0; i < order; i++) {
        double d = fabs(in[i]);
        if (d > cmax) cmax = d;
    }
    if (cmax * 32768.0 < 1.0) {
        *shift = 0;
        memset(out, 0, sizeof(int32_t) * (size_t)order);
        return;
    }
    int sh = 15;
    while (cmax * (double)(1 << sh) > qmax && sh > 0) sh--;
    if (sh == 0 && cmax > qmax) {
        double scale = ((double)qmax) / cmax;
        for (int i = 0; i < order; i++) in[i] *= scale;
    }
    double err = 0;
    for (int i = 0; i < order; i++) {
        err += in[i] * (double)(1 << sh);
        int32_t q = trunc_to_int(err + 0.5);
        if (q <= -qmax) q = -qmax + 1;
        if (q > qmax) q = qmax;
        err -= q;
        out[i] = q;
    }
    *shift = sh;
}

/* lpc.c:224-257.  coefs: [32][32] row-major, shift[32]. */
int orc_lpc_calc(const int32_t *smp, int n, int max_order, int omethod,
                 int32_t *coefs, int *shift)
{
    double autoc[ORC_MAX_ORDER + 1];
    double lpc[ORC_MAX_ORDER][ORC_MAX_ORDER];
    int est = max_order;
    memset(lpc, 0, sizeof lpc);
    orc_autocorr(smp, n, max_order, autoc);
    if (omethod == 1) est = schur_estimate(autoc, max_order, lpc);
    else levinson(autoc, max_order, NULL, lpc);
    if (omethod == 0 || omethod == 1) {
        int i = est - 1;
        quantize(lpc[i], i + 1, coefs + i * ORC_MAX_ORDER, &shift[i]);
    } else {
        for (int i = 0; i < max_order; i++)
            quantize(lpc[i], i + 1, coefs + i * ORC_MAX_ORDER, &shift[i]);
    }
    return est;
}

/* ------------------------------------------------------------------ */
/* stage 3: residual + Rice search -- optimize.c, rice.c                */
/* ------------------------------------------------------------------ */

/* rice.h:48 evaluated in uint64 then truncated */
static uint32_t rice_bits_u32(uint64_t sum, int n, int k)
{
    uint64_t v = (uint64_t)((int64_t)n * (k + 1)) + ((sum - (uint64_t)(n >> 1)) >> k);
    return (uint32_t)v;
}

/* rice.c:30-45: first strict minimum over k = 0..30 */
int orc_rice_k(uint64_t sum, int n)
{
    int best = 0;
    uint32_t best_bits = UINT32_MAX;
    for (int k = 0; k <= 30; k++) {
        uint32_t b = rice_bits_u32(sum, n, k);
        if (b < best_bits) { best_bits = b; best = k; }
    }
    return best;
}

/* rice.c:148-155 */
static int limit_porder(int p, int n, int order)
{
    int lim = count_trailing_zeros((uint32_t)n);
    if (lim < p) p = lim;
    if (order > 0) {
        int l2 = ilog2_u32((uint32_t)(n / order));
        if (l2 < p) p = l2;
    }
    return p;
}

/* rice.c:47-139 + 157-187 */
uint32_t orc_rice_cost(const int32_t *res, int n, int pred_order, int obits,
                       int pmin, int pmax, int is_lpc,
                       int *method_out, int *porder_out, int *params_out)
{
    static uint64_t sums[9][ORC_MAX_PARTS];     /* not re-entrant; tests are serial */
    uint64_t (*S)[ORC_MAX_PARTS] = sums;
    pmin = limit_porder(pmin, n, pred_order);
    pmax = limit_porder(pmax, n, pred_order);

    /* finest level (rice.c:76-95) */
    int parts = 1 << pmax, psize = n >> pmax;
    for (int p = 0; p < parts; p++) {
        int lo = p ? p * psize : pred_order, hi = (p + 1) * psize;
        uint64_t s = 0;
        for (int i = lo; i < hi; i++) {
            int32_t v = res[i];
            uint32_t u = ((uint32_t)v << 1) ^ (uint32_t)(v >> 31);
            s += u;
        }
        S[pmax][p] = s;
    }
    for (int lv = pmax - 1; lv >= pmin; lv--)
        for (int p = 0; p < (1 << lv); p++)
            S[lv][p] = S[lv + 1][2 * p] + S[lv + 1][2 * p + 1];

    uint32_t best_bits = UINT32_MAX;
    int best_po = pmin, best_method = 0;
    int tmp[ORC_MAX_PARTS];
    for (int po = pmin; po <= pmax; po++) {
        int np = 1 << po, method = 0;
        uint32_t bits = 0;
        for (int p = 0; p < np; p++) {
            int cnt = (n >> po) - (p ? 0 : pred_order);
            int k = orc_rice_k(S[po][p], cnt);
            tmp[p] = k;
            if (k > 14) method = 1;
            bits += rice_bits_u32(S[po][p], cnt, k);
        }
        bits += 4u * (uint32_t)np;
        if (bits <= best_bits) {             /* ties -> higher porder (rice.c:131) */
            best_bits = bits; best_po = po; best_method = method;
            if (params_out) memcpy(params_out, tmp, sizeof(int) * (size_t)np);
        }
    }
    if (method_out) *method_out = best_method;
    if (porder_out) *porder_out = best_po;

    uint32_t total = (uint32_t)(pred_order * obits + 2);
    if (is_lpc) total += 4 + 5 + (uint32_t)pred_order * 15u;
    total += best_bits;
    total += (uint32_t)best_method + 4;
    return total;
}

/* optimize.c:34-68 */
void orc_residual_fixed(int32_t *res, const int32_t *smp, int n, int order)
{
    for (int i = 0; i < order && i < n; i++) res[i] = smp[i];
    for (int i = order; i < n; i++) {
        int64_t a = smp[i], b = order > 0 ? smp[i-1] : 0, c = order > 1 ? smp[i-2] : 0;
        int64_t d = order > 2 ? smp[i-3] : 0, e = order > 3 ? smp[i-4] : 0, r;
        switch (order) {
        case 0:  r = a; break;
        case 1:  r = a - b; break;
        case 2:  r = a - 2*b + c; break;
        case 3:  r = a - 3*b + 3*c - d; break;
        default: r = a - 4*b + 6*c - 4*d + e; break;
        }
        res[i] = (int32_t)r;
    }
}

/* optimize.c:70-122 */
void orc_residual_lpc(int32_t *res, const int32_t *smp, int n, int order,
                      const int32_t *coefs, int shift)
{
    for (int i = 0; i < order && i < n; i++) res[i] = smp[i];
    for (int i = order; i < n; i++) {
        int64_t pred = 0;
        for (int j = order - 1; j >= 0; j--)       /* same term order as the switch */
            pred += (int64_t)coefs[j] * (int64_t)smp[i - 1 - j];
        res[i] = (int32_t)((int64_t)smp[i] - (pred >> shift));
    }
}

typedef struct {
    const OrcParams *p;
    int n;
    int32_t *smp;      /* decorrelated, wasted-shifted samples */
    int32_t *res;
    OrcSubframe *sf;
} SubJob;

/* optimize.c:124-276 */
static void choose_subframe(SubJob *J)
{
    const OrcParams *P = J->p;
    OrcSubframe *sf = J->sf;
    const int n = J->n;
    int32_t *smp = J->smp, *res = J->res;
    int i;

    for (i = 1; i < n; i++) if (smp[i] != smp[0]) break;
    if (i == n) {                                   /* CONSTANT */
        sf->type = 0; res[0] = smp[0]; sf->est_bits = (uint32_t)sf->obits;
        return;
    }
    if (n < 5 || P->prediction_type == 0) {         /* VERBATIM */
        sf->type = 1; memcpy(res, smp, sizeof(int32_t) * (size_t)n);
        sf->est_bits = (uint32_t)(sf->obits * n);
        return;
    }

    int min_order = P->min_order, max_order = P->max_order;
    int pmin = P->min_porder, pmax = P->max_porder;

    if (P->prediction_type == 1 || n <= max_order) { /* FIXED */
        uint32_t bits[5];
        if (max_order > 4) max_order = 4;
        int opt = min_order;
        /* NB: when min_order > 4 (LPC preset, tiny last block) the reference
         * indexes bits[] out of range; every preset has min_order <= 4. */
        bits[opt > 4 ? 4 : opt] = UINT32_MAX;
        for (i = min_order; i <= max_order; i++) {
            orc_residual_fixed(res, smp, n, i);
            bits[i] = orc_rice_cost(res, n, i, sf->obits, pmin, pmax, 0,
                                    &sf->method, &sf->porder, sf->params);
            if (bits[i] < bits[opt]) opt = i;
        }
        sf->order = opt; sf->type = 8;
        if (opt != max_order) {
            orc_residual_fixed(res, smp, n, opt);
            sf->est_bits = orc_rice_cost(res, n, opt, sf->obits, pmin, pmax, 0,
                                         &sf->method, &sf->porder, sf->params);
        } else {
            sf->est_bits = bits[opt];
        }
        return;
    }

    /* LPC */
    static int32_t coefs[ORC_MAX_ORDER][ORC_MAX_ORDER];
    int shift[ORC_MAX_ORDER];
    int om = P->order_method, opt_order;
    int est = orc_lpc_calc(smp, n, max_order, om, &coefs[0][0], shift);

#define EVAL(idx) ( orc_residual_lpc(res, smp, n, (idx)+1, coefs[idx], shift[idx]), \
                    orc_rice_cost(res, n, (idx)+1, sf->obits, pmin, pmax, 1, \
                                  &sf->method, &sf->porder, sf->params) )
    if (om == 0) {
        opt_order = max_order;
    } else if (om == 1) {
        opt_order = est;
    } else if (om >= 2 && om <= 4) {
        int levels = 1 << (om - 1);
        uint32_t bits[8];
        int opt_index = levels - 1;
        opt_order = max_order - 1;
        bits[opt_index] = UINT32_MAX;
        for (i = opt_index; i >= 0; i--) {
            int order = min_order + (((max_order - min_order + 1) * (i + 1)) / levels) - 2;
            if (order < 0) order = 0;
            bits[i] = EVAL(order);
            if (bits[i] < bits[opt_index]) { opt_index = i; opt_order = order; }
        }
        opt_order++;
    } else if (om == 5) {
        uint32_t bits[ORC_MAX_ORDER];
        opt_order = 0;
        bits[0] = UINT32_MAX;
        for (i = 0; i < max_order; i++) {
            bits[i] = EVAL(i);
            if (bits[i] < bits[opt_order]) opt_order = i;
        }
        opt_order++;
    } else {
        uint32_t bits[ORC_MAX_ORDER];
        opt_order = min_order - 1 + (max_order - min_order) / 3;
        memset(bits, 0xff, sizeof bits);
        for (int step = 16; step > 0; step >>= 1) {
            int last = opt_order;
            for (i = last - step; i <= last + step; i += step) {
                if (i < min_order - 1 || i >= max_order || bits[i] < UINT32_MAX) continue;
                bits[i] = EVAL(i);
                if (bits[i] < bits[opt_order]) opt_order = i;
            }
        }
        opt_order++;
    }
#undef EVAL
    sf->order = opt_order; sf->type = 32;
    sf->shift = shift[opt_order - 1];
    for (i = 0; i < opt_order; i++) sf->coefs[i] = coefs[opt_order - 1][i];
    orc_residual_lpc(res, smp, n, opt_order, sf->coefs, sf->shift);
    sf->est_bits = orc_rice_cost(res, n, opt_order, sf->obits, pmin, pmax, 1,
                                 &sf->method, &sf->porder, sf->params);
}

/* ------------------------------------------------------------------ */
/* stage 1: prepare -- encode.c:541-694                                 */
/* ------------------------------------------------------------------ */
static uint64_t rice_bits_u64(uint64_t sum, int n, int k)
{
    return (uint64_t)((int64_t)n * (k + 1)) + ((sum - (uint64_t)(n >> 1)) >> k);
}

/* encode.c:598-643 */
static int stereo_mode(const int32_t *l, const int32_t *r, int n)
{
    uint64_t s[4] = {0, 0, 0, 0}, score[4];
    for (int i = 2; i < n; i++) {
        int32_t lt = l[i] - 2 * l[i-1] + l[i-2];
        int32_t rt = r[i] - 2 * r[i-1] + r[i-2];
        s[2] += (uint64_t)(int64_t)abs((lt + rt) >> 1);
        s[3] += (uint64_t)(int64_t)abs(lt - rt);
        s[0] += (uint64_t)(int64_t)abs(lt);
        s[1] += (uint64_t)(int64_t)abs(rt);
    }
    for (int i = 0; i < 4; i++) {
        int k = orc_rice_k(2 * s[i], n);
        s[i] = rice_bits_u64(2 * s[i], n, k);
    }
    score[0] = s[0] + s[1]; score[1] = s[0] + s[3];
    score[2] = s[1] + s[3]; score[3] = s[2] + s[3];
    int best = 0;
    for (int i = 1; i < 4; i++) if (score[i] < score[best]) best = i;
    static const int modes[4] = {1, 8, 9, 10};
    return modes[best];
}

/* encode.c:558-593 */
static int strip_wasted_bits(int32_t *smp, int n, int bps)
{
    int wasted = bps - 1;
    for (int i = 0; i < n && wasted; i++) {
        uint32_t s = (uint32_t)smp[i];
        if (s) {
            int b = count_trailing_zeros(s);
            if (b < wasted) wasted = b;
        }
    }
    if (wasted == bps - 1) return 0;
    if (wasted) for (int i = 0; i < n; i++) smp[i] >>= wasted;
    return wasted;
}

/* ------------------------------------------------------------------ */
/* stage 4: pack -- encode.c:700-917                                    */
/* ------------------------------------------------------------------ */
static const int rate_table[16]  = {0,0,0,0,8000,16000,22050,24000,32000,44100,48000,96000,0,0,0,0};
static const int depth_table[8]  = {0,8,12,0,16,20,24,0};
static const int bsize_table[15] = {0,192,576,1152,2304,4608,0,0,256,512,1024,2048,4096,8192,16384};

static void rate_codes(int rate, int *c0, int *c1)
{
    *c0 = 0; *c1 = 0;
    for (int i = 4; i < 12; i++) if (rate == rate_table[i]) { *c0 = i; return; }
    if (rate % 1000 == 0 && rate <= 255000) { *c0 = 12; *c1 = rate / 1000; }
    else if (rate % 10 == 0 && rate <= 655350) { *c0 = 14; *c1 = rate / 10; }
    else if (rate < 65535) { *c0 = 13; *c1 = rate; }
}

static int depth_code(int bps)
{
    for (int i = 1; i < 8; i++) if (bps == depth_table[i]) return i;
    return 0;
}

/* encode.c:700-716 */
static void put_utf8(BitSink *b, uint32_t v)
{
    if (v < 0x80) { bs_put(b, 8, v); return; }
    int bytes = (ilog2_u32(v) + 4) / 5;
    int shift = (bytes - 1) * 6;
    bs_put(b, 8, (256 - (256 >> bytes)) | (v >> shift));
    while (shift >= 6) { shift -= 6; bs_put(b, 8, 0x80 | ((v >> shift) & 0x3f)); }
}

static void put_signed(BitSink *b, int bits, int32_t v)
{
    /* bitio.h:110-115: mask with (1ULL<<bits)-1, bits may be 32 for side ch of 32-bit? no: <=31 asserted */
    if (bits >= 32) bs_put(b, 32, (uint32_t)v);
    else bs_put(b, bits, (uint32_t)v & (uint32_t)((1ull << bits) - 1));
}

static void put_rice(BitSink *b, int k, int32_t v)
{
    uint32_t u = ((uint32_t)v << 1) ^ (uint32_t)(v >> 31);
    bs_zeros(b, u >> k);
    bs_put(b, 1, 1);
    if (k) bs_put(b, k, u & ((1u << k) - 1u));
}

static int pack_frame(const OrcParams *P, int n, int ch_mode, uint32_t number,
                      OrcSubframe *sf, int32_t **res, uint8_t *out, int cap)
{
    BitSink b; bs_init(&b, out, (size_t)cap);
    int sr0, sr1; rate_codes(P->sample_rate, &sr0, &sr1);
    int bs0 = -1, bs1 = -1;
    for (int i = 0; i < 15; i++) if (n == bsize_table[i]) { bs0 = i; break; }
    if (bs0 < 0) { bs0 = (n <= 256) ? 6 : 7; bs1 = n - 1; }

    bs_put(&b, 15, 0x7ffc);
    bs_put(&b, 1, (uint32_t)P->allow_vbs);
    bs_put(&b, 4, (uint32_t)bs0);
    bs_put(&b, 4, (uint32_t)sr0);
    bs_put(&b, 4, (uint32_t)(ch_mode == 0 ? P->channels - 1 : ch_mode));
    bs_put(&b, 3, (uint32_t)depth_code(P->bps));
    bs_put(&b, 1, 0);
    put_utf8(&b, number);
    if (bs1 >= 0) bs_put(&b, bs1 < 256 ? 8 : 16, (uint32_t)bs1);
    if (sr1 > 0)  bs_put(&b, sr1 < 256 ? 8 : 16, (uint32_t)sr1);
    bs_put(&b, 8, orc_crc8(out, b.pos <= (size_t)cap ? b.pos : (size_t)cap));

    for (int c = 0; c < P->channels; c++) {
        OrcSubframe *s = &sf[c];
        int code = s->type;
        if (s->type == 8)  code = 8 | s->order;
        if (s->type == 32) code = 32 | (s->order - 1);
        bs_put(&b, 1, 0);
        bs_put(&b, 6, (uint32_t)code);
        if (s->wasted) { bs_put(&b, 1, 1); bs_zeros(&b, (uint32_t)(s->wasted - 1)); bs_put(&b, 1, 1); }
        else bs_put(&b, 1, 0);

        const int32_t *r = res[c];
        if (s->type == 0) {
            put_signed(&b, s->obits, r[0]);
        } else if (s->type == 1) {
            for (int i = 0; i < n; i++) put_signed(&b, s->obits, r[i]);
        } else {
            for (int i = 0; i < s->order; i++) put_signed(&b, s->obits, r[i]);
            if (s->type == 32) {
                bs_put(&b, 4, 14);
                put_signed(&b, 5, s->shift);
                for (int i = 0; i < s->order; i++) put_signed(&b, 15, s->coefs[i]);
            }
            bs_put(&b, 2, (uint32_t)s->method);
            bs_put(&b, 4, (uint32_t)s->porder);
            int psize = n >> s->porder, j = s->order;
            for (int p = 0; p < (1 << s->porder); p++) {
                int k = s->params[p];
                bs_put(&b, 4 + s->method, (uint32_t)k);
                int end = (p + 1) * psize;
                for (; j < end && j < n; j++) put_rice(&b, k, r[j]);
            }
        }
    }
    bs_align(&b);
    if (b.overflow) return -2;
    uint16_t crc = orc_crc16(out, b.pos);
    bs_put(&b, 16, crc);
    if (b.overflow) return -2;
    return (int)b.pos;
}

static int verbatim_bound(const OrcParams *P, int n)
{
    if (P->channels == 2) return 16 + ((n * (P->bps + P->bps + 1) + 7) >> 3);
    return 16 + ((n * P->channels * P->bps + 7) >> 3);
}

uint32_t orc_initial_max_frame_size(const OrcParams *p)
{
    return (uint32_t)verbatim_bound(p, p->block_size);
}

/* encode.c:919-977 */
int orc_encode_frame(const OrcParams *P, const int32_t *pcm, int n,
                     uint32_t number, uint8_t *out, int out_cap, OrcFrameInfo *info)
{
    if (!P || !pcm || !out || n < 1 || n > 65535) return -1;
    const int C = P->channels;
    OrcFrameInfo local;
    if (!info) info = &local;
    memset(info, 0, sizeof *info);
    info->blocksize = n;

    int32_t *smp[ORC_MAX_CH], *res[ORC_MAX_CH];
    for (int c = 0; c < C; c++) {
        smp[c] = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
        res[c] = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
        for (int i = 0; i < n; i++) smp[c][i] = pcm[(size_t)i * C + c];
        info->sub[c].obits = P->bps;
    }

    /* stereo decorrelation, encode.c:648-694 */
    int mode;
    if (C != 2) mode = 0;
    else if (n <= 32 || P->stereo_method == 0) mode = 1;
    else mode = stereo_mode(smp[0], smp[1], n);
    if (mode == 10) {
        for (int i = 0; i < n; i++) {
            int32_t l = smp[0][i], r = smp[1][i];
            smp[0][i] = (l + r) >> 1; smp[1][i] = l - r;
        }
        info->sub[1].obits++;
    } else if (mode == 8) {
        for (int i = 0; i < n; i++) smp[1][i] = smp[0][i] - smp[1][i];
        info->sub[1].obits++;
    } else if (mode == 9) {
        for (int i = 0; i < n; i++) smp[0][i] = smp[0][i] - smp[1][i];
        info->sub[0].obits++;
    }
    info->ch_mode = mode;

    for (int c = 0; c < C; c++) {
        int w = strip_wasted_bits(smp[c], n, P->bps);
        info->sub[c].wasted = w;
        info->sub[c].obits -= w;
    }
    for (int c = 0; c < C; c++) {
        SubJob J = { P, n, smp[c], res[c], &info->sub[c] };
        choose_subframe(&J);
    }

    int vsize = verbatim_bound(P, n);
    int cap = vsize * 3 / 2 + 64;
    if (cap > out_cap) cap = out_cap;
    int nb = pack_frame(P, n, mode, number, info->sub, res, out, cap);
    if (nb < 0 || nb > vsize) {
        /* encode.c:949-964 + optimize.c:278-289 */
        for (int c = 0; c < C; c++) {
            info->sub[c].type = 1;
            memcpy(res[c], smp[c], sizeof(int32_t) * (size_t)n);
        }
        info->verbatim_fallback = 1;
        nb = pack_frame(P, n, mode, number, info->sub, res, out, out_cap);
        if (nb < 0) nb = -1;
    }
    info->nbytes = nb;
    for (int c = 0; c < C; c++) { free(smp[c]); free(res[c]); }
    return nb;
}

/* vbs.c:36-83 with the 32-bit abs()/multiply quirk */
int orc_vbs_split(const int32_t *pcm, int channels, int block_size, int sizes[8])
{
    int n = block_size / 8;
    int64_t e[8];
    for (int s = 0; s < 8; s++) {
        const int32_t *base = pcm + (size_t)s * n * channels;
        int64_t acc = 0;
        for (int c = 0; c < channels; c++)
            for (int j = 2; j < n; j++) {
                int32_t v = (int32_t)((uint32_t)base[(size_t)j * channels + c]
                          - 2u * (uint32_t)base[(size_t)(j-1) * channels + c]
                          + (uint32_t)base[(size_t)(j-2) * channels + c]);
                acc += (v < 0) ? -(int64_t)v : (int64_t)v;   /* abs(int) then widened */
            }
        e[s] = acc / channels + 1;
    }
    int nfr = 0;
    memset(sizes, 0, 8 * sizeof(int));
    for (int s = 0; s < 8; s++) {
        int start = (s == 0);
        if (s > 0) {
            int32_t t = (int32_t)(uint32_t)(uint64_t)(e[s-1] - e[s]);
            int32_t a = (t < 0) ? (int32_t)(0u - (uint32_t)t) : t;
            int32_t prod = (int32_t)((uint32_t)a * 200u);
            start = ((int64_t)prod / e[s-1]) > 50;
        }
        if (start) nfr++;
        sizes[nfr - 1] += n;
    }
    return nfr;
}

/* flake_encode_frame loop, encode.c:979-1008 + vbs.c:85-119 */
int64_t orc_encode_stream(const OrcParams *P, const int32_t *pcm, uint64_t nsamples,
                          uint8_t *out, size_t out_cap, uint32_t *frame_len,
                          uint32_t *frame_bs, uint32_t frame_cap, uint32_t *nframes,
                          uint32_t *max_frame_size)
{
    if (orc_validate(P) < 0) return -1;
    const int C = P->channels, B = P->block_size;
    uint32_t counter = 0, nf = 0, maxfs = orc_initial_max_frame_size(P);
    size_t pos = 0;
    for (uint64_t start = 0; start < nsamples; start += (uint64_t)B) {
        int n = (int)((nsamples - start < (uint64_t)B) ? nsamples - start : (uint64_t)B);
        const int32_t *blk = pcm + start * (uint64_t)C;
        int sizes[8], parts = 1;
        sizes[0] = n;
        if (P->variable_block_size > 0 && (n % 8) == 0 && n >= 128) {
            parts = orc_vbs_split(blk, C, n, sizes);
            if (parts <= 1) { parts = 1; sizes[0] = n; }
        }
        int off = 0;
        for (int f = 0; f < parts; f++) {
            if (nf >= frame_cap) return -1;
            size_t room = out_cap - pos;
            int cap = room > (size_t)INT32_MAX ? INT32_MAX : (int)room;
            int nb = orc_encode_frame(P, blk + (size_t)off * C, sizes[f], counter,
                                      out + pos, cap, NULL);
            if (nb < 0) return -1;
            if ((uint32_t)nb > maxfs) maxfs = (uint32_t)nb;
            if (frame_len) frame_len[nf] = (uint32_t)nb;
            if (frame_bs)  frame_bs[nf] = (uint32_t)sizes[f];
            nf++;
            pos += (size_t)nb;
            counter += P->allow_vbs ? (uint32_t)sizes[f] : 1u;
            off += sizes[f];
        }
    }
    if (nframes) *nframes = nf;
    if (max_frame_size) *max_frame_size = maxfs;
    return (int64_t)pos;
}

/* ------------------------------------------------------------------ */
/* stream header / STREAMINFO -- encode.c:52-156, metadata.c:32-84     */
/* ------------------------------------------------------------------ */
static void streaminfo_body(const OrcParams *p, uint32_t max_frame, const uint8_t md5[16],
                            uint8_t out[34])
{
    BitSink b; bs_init(&b, out, 34);
    uint32_t min_bs = (p->variable_block_size || p->allow_vbs) ? 16u : (uint32_t)p->block_size;
    bs_put(&b, 16, min_bs);
    bs_put(&b, 16, (uint32_t)p->block_size);
    bs_put(&b, 24, 0);
    bs_put(&b, 24, max_frame);
    bs_put(&b, 20, (uint32_t)p->sample_rate);
    bs_put(&b, 3, (uint32_t)(p->channels - 1));
    bs_put(&b, 5, (uint32_t)(p->bps - 1));
    bs_put(&b, 4, 0);
    bs_put(&b, 32, p->total_samples);
    memcpy(out + 18, md5, 16);
}

void orc_streaminfo(const OrcParams *p, uint32_t max_frame_size, const uint8_t md5[16],
                    uint8_t out[34])
{
    streaminfo_body(p, max_frame_size, md5, out);
}

int orc_write_header(const OrcParams *p, uint8_t *h, int cap)
{
    static const char vendor[] = "Flake SVN";
    int vlen = (int)strlen(vendor), vc = 4 + vlen + 4;
    int need = 4 + 38 + 4 + vc + (p->padding_size > 0 ? 4 + p->padding_size : 0);
    if (cap < need) return -1;
    memset(h, 0, (size_t)need);
    int pos = 0;
    memcpy(h, "fLaC", 4); pos = 4;
    h[pos] = 0x00; h[pos+1] = 0; h[pos+2] = 0; h[pos+3] = 34;
    uint8_t md5[16]; orc_md5_zero_ctx(md5);
    streaminfo_body(p, orc_initial_max_frame_size(p), md5, h + pos + 4);
    pos += 38;
    int last = (p->padding_size == 0);
    h[pos] = (uint8_t)((last << 7) | 4);
    h[pos+1] = (uint8_t)(vc >> 16); h[pos+2] = (uint8_t)(vc >> 8); h[pos+3] = (uint8_t)vc;
    pos += 4;
    h[pos] = (uint8_t)vlen; pos += 4;
    memcpy(h + pos, vendor, (size_t)vlen); pos += vlen;
    pos += 4;                                  /* zero entries */
    if (p->padding_size > 0) {
        h[pos] = 0x80 | 1;
        h[pos+1] = (uint8_t)(p->padding_size >> 16);
        h[pos+2] = (uint8_t)(p->padding_size >> 8);
        h[pos+3] = (uint8_t)p->padding_size;
        pos += 4 + p->padding_size;
    }
    return pos;
}
