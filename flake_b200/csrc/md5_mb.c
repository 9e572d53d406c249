/*
 * md5_mb.c -- multi-buffer MD5: the digests of up to 16 INDEPENDENT streams advanced
 * together, one stream per 32-bit SIMD lane (AVX2: two interleaved sets of 8 lanes).
 *
 * Why: MD5 (libflake/md5.c, fed from flake_encode_frame, encode.c:1006) is a serial chain
 * per stream -- about 0.78 GB/s on one host core -- while one B200 encodes 100+ GB/s of
 * PCM.  One stream cannot go faster, but a corpus of streams (100 x 1 h) has
 * one chain per file, and chains of different files are independent: putting 16 of them in
 * the lanes of one core's vector unit makes that core hash 16 files at once.  Used by
 * flake_b200_encode_corpus only; a single stream keeps the scalar code of md5.c.
 *
 * The message words of lane l are bytes of stream l, so each 64-byte block step first
 * transposes an (8 lanes x 8 words) tile twice; the 64 rounds then run on vectors holding
 * word j of all lanes.  Two sets of 8 lanes are interleaved instruction by instruction: a
 * single set is bound by the latency of the round's dependency chain, two fill the ports.
 */
#include "md5_mb.h"

#include <string.h>

#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define FB_HAVE_AVX2_BUILD 1
#else
#define FB_HAVE_AVX2_BUILD 0
#endif

static const uint32_t K[64] = {
    0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501,
    0x698098d8, 0x8b44f7af, 0xffff5bb1, 0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821,
    0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8,
    0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a,
    0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70,
    0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665,
    0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d, 0x85845dd1,
    0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};
static const uint8_t WI[64] = {
    0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 1, 6, 11, 0, 5, 10, 15, 4, 9, 14, 3, 8, 13, 2, 7, 12,
    5, 8, 11, 14, 1, 4, 7, 10, 13, 0, 3, 6, 9, 12, 15, 2, 0, 7, 14, 5, 12, 3, 10, 1, 8, 15, 6, 13, 4, 11, 2, 9};
static const uint8_t SH[64] = {
    7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 5, 9, 14, 20, 5, 9, 14, 20, 5, 9, 14, 20, 5, 9, 14, 20,
    4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21};

#if FB_HAVE_AVX2_BUILD

/* words 0..7 of eight lanes (rows r[0..7]) -> w[j] = word j of every lane */
__attribute__((target("avx2"))) static inline void transpose8(const __m256i r[8], __m256i w[8])
{
    const __m256i t0 = _mm256_unpacklo_epi32(r[0], r[1]), t1 = _mm256_unpackhi_epi32(r[0], r[1]);
    const __m256i t2 = _mm256_unpacklo_epi32(r[2], r[3]), t3 = _mm256_unpackhi_epi32(r[2], r[3]);
    const __m256i t4 = _mm256_unpacklo_epi32(r[4], r[5]), t5 = _mm256_unpackhi_epi32(r[4], r[5]);
    const __m256i t6 = _mm256_unpacklo_epi32(r[6], r[7]), t7 = _mm256_unpackhi_epi32(r[6], r[7]);
    const __m256i u0 = _mm256_unpacklo_epi64(t0, t2), u1 = _mm256_unpackhi_epi64(t0, t2);
    const __m256i u2 = _mm256_unpacklo_epi64(t1, t3), u3 = _mm256_unpackhi_epi64(t1, t3);
    const __m256i u4 = _mm256_unpacklo_epi64(t4, t6), u5 = _mm256_unpackhi_epi64(t4, t6);
    const __m256i u6 = _mm256_unpacklo_epi64(t5, t7), u7 = _mm256_unpackhi_epi64(t5, t7);
    w[0] = _mm256_permute2x128_si256(u0, u4, 0x20); w[4] = _mm256_permute2x128_si256(u0, u4, 0x31);
    w[1] = _mm256_permute2x128_si256(u1, u5, 0x20); w[5] = _mm256_permute2x128_si256(u1, u5, 0x31);
    w[2] = _mm256_permute2x128_si256(u2, u6, 0x20); w[6] = _mm256_permute2x128_si256(u2, u6, 0x31);
    w[3] = _mm256_permute2x128_si256(u3, u7, 0x20); w[7] = _mm256_permute2x128_si256(u3, u7, 0x31);
}

#define ROLV(x, s) _mm256_or_si256(_mm256_slli_epi32((x), (s)), _mm256_srli_epi32((x), 32 - (s)))

/* one round step on both sets; FN(b, c, d) is the round's boolean function */
#define STEP(FN, a, b, c, d, i)                                                                   \
    do {                                                                                          \
        const __m256i k_ = _mm256_set1_epi32((int)K[i]);                                          \
        __m256i x0 = _mm256_add_epi32(_mm256_add_epi32(a##0, k_), _mm256_add_epi32(w0[WI[i]], FN(b##0, c##0, d##0))); \
        __m256i x1 = _mm256_add_epi32(_mm256_add_epi32(a##1, k_), _mm256_add_epi32(w1[WI[i]], FN(b##1, c##1, d##1))); \
        a##0 = _mm256_add_epi32(ROLV(x0, SH[i]), b##0);                                           \
        a##1 = _mm256_add_epi32(ROLV(x1, SH[i]), b##1);                                           \
    } while (0)
#define FF(b, c, d) _mm256_xor_si256((d), _mm256_and_si256((b), _mm256_xor_si256((c), (d))))
#define GG(b, c, d) _mm256_xor_si256((c), _mm256_and_si256((d), _mm256_xor_si256((b), (c))))
#define HH(b, c, d) _mm256_xor_si256((b), _mm256_xor_si256((c), (d)))
#define II(b, c, d) _mm256_xor_si256((c), _mm256_or_si256((b), _mm256_xor_si256((d), ones)))
#define ROUND4(FN, i)                                                                             \
    STEP(FN, a, b, c, d, (i)); STEP(FN, d, a, b, c, (i) + 1); STEP(FN, c, d, a, b, (i) + 2); STEP(FN, b, c, d, a, (i) + 3)

/* state[j][l] = word j of lane l's digest; data[l] = lane l's bytes (nblocks * 64 each) */
__attribute__((target("avx2"))) static void md5_mb16_avx2(uint32_t state[4][32], const uint8_t *const data[32], size_t nblocks)
{
    __m256i a0 = _mm256_loadu_si256((const __m256i *)&state[0][0]), a1 = _mm256_loadu_si256((const __m256i *)&state[0][8]);   /* lanes 0..15 of 32 */
    __m256i b0 = _mm256_loadu_si256((const __m256i *)&state[1][0]), b1 = _mm256_loadu_si256((const __m256i *)&state[1][8]);
    __m256i c0 = _mm256_loadu_si256((const __m256i *)&state[2][0]), c1 = _mm256_loadu_si256((const __m256i *)&state[2][8]);
    __m256i d0 = _mm256_loadu_si256((const __m256i *)&state[3][0]), d1 = _mm256_loadu_si256((const __m256i *)&state[3][8]);
    const __m256i ones = _mm256_set1_epi32(-1);
    for (size_t blk = 0; blk < nblocks; blk++) {
        __m256i w0[16], w1[16], r[8];
        const size_t off = blk * 64;
        for (int h = 0; h < 2; h++) {
            for (int l = 0; l < 8; l++) r[l] = _mm256_loadu_si256((const __m256i *)(data[l] + off + 32 * h));
            transpose8(r, w0 + 8 * h);
            for (int l = 0; l < 8; l++) r[l] = _mm256_loadu_si256((const __m256i *)(data[8 + l] + off + 32 * h));
            transpose8(r, w1 + 8 * h);
        }
        if ((blk & 7) == 0)
            for (int l = 0; l < 16; l++) _mm_prefetch((const char *)(data[l] + off + 1024), _MM_HINT_T0);
        const __m256i sa0 = a0, sb0 = b0, sc0 = c0, sd0 = d0, sa1 = a1, sb1 = b1, sc1 = c1, sd1 = d1;
        ROUND4(FF, 0);  ROUND4(FF, 4);  ROUND4(FF, 8);  ROUND4(FF, 12);
        ROUND4(GG, 16); ROUND4(GG, 20); ROUND4(GG, 24); ROUND4(GG, 28);
        ROUND4(HH, 32); ROUND4(HH, 36); ROUND4(HH, 40); ROUND4(HH, 44);
        ROUND4(II, 48); ROUND4(II, 52); ROUND4(II, 56); ROUND4(II, 60);
        a0 = _mm256_add_epi32(a0, sa0); b0 = _mm256_add_epi32(b0, sb0); c0 = _mm256_add_epi32(c0, sc0); d0 = _mm256_add_epi32(d0, sd0);
        a1 = _mm256_add_epi32(a1, sa1); b1 = _mm256_add_epi32(b1, sb1); c1 = _mm256_add_epi32(c1, sc1); d1 = _mm256_add_epi32(d1, sd1);
    }
    _mm256_storeu_si256((__m256i *)&state[0][0], a0); _mm256_storeu_si256((__m256i *)&state[0][8], a1);
    _mm256_storeu_si256((__m256i *)&state[1][0], b0); _mm256_storeu_si256((__m256i *)&state[1][8], b1);
    _mm256_storeu_si256((__m256i *)&state[2][0], c0); _mm256_storeu_si256((__m256i *)&state[2][8], c1);
    _mm256_storeu_si256((__m256i *)&state[3][0], d0); _mm256_storeu_si256((__m256i *)&state[3][8], d1);
}
/* ---- AVX-512: 16 lanes per vector, rotate and three-input logic are single instructions --- */
/* sixteen words of sixteen lanes (rows r[0..15]) -> w[j] = word j of every lane */
__attribute__((target("avx512f"))) static inline void transpose16(const __m512i r[16], __m512i w[16])
{
    __m512i u[4][4];
    for (int g = 0; g < 4; g++) {
        const __m512i t0 = _mm512_unpacklo_epi32(r[4 * g], r[4 * g + 1]), t1 = _mm512_unpackhi_epi32(r[4 * g], r[4 * g + 1]);
        const __m512i t2 = _mm512_unpacklo_epi32(r[4 * g + 2], r[4 * g + 3]), t3 = _mm512_unpackhi_epi32(r[4 * g + 2], r[4 * g + 3]);
        u[g][0] = _mm512_unpacklo_epi64(t0, t2); u[g][1] = _mm512_unpackhi_epi64(t0, t2);
        u[g][2] = _mm512_unpacklo_epi64(t1, t3); u[g][3] = _mm512_unpackhi_epi64(t1, t3);
    }
    for (int q = 0; q < 4; q++) {
        const __m512i lo01 = _mm512_shuffle_i32x4(u[0][q], u[1][q], 0x88), hi01 = _mm512_shuffle_i32x4(u[0][q], u[1][q], 0xdd);
        const __m512i lo23 = _mm512_shuffle_i32x4(u[2][q], u[3][q], 0x88), hi23 = _mm512_shuffle_i32x4(u[2][q], u[3][q], 0xdd);
        w[q]      = _mm512_shuffle_i32x4(lo01, lo23, 0x88);
        w[q + 8]  = _mm512_shuffle_i32x4(lo01, lo23, 0xdd);
        w[q + 4]  = _mm512_shuffle_i32x4(hi01, hi23, 0x88);
        w[q + 12] = _mm512_shuffle_i32x4(hi01, hi23, 0xdd);
    }
}

/* truth tables with operands (b, c, d): F = b ? c : d, G = d ? b : c, H = b ^ c ^ d, I = c ^ (b | ~d) */
#define T_F 0xca
#define T_G 0xe4
#define T_H 0x96
#define T_I 0x39
#define STEP512(S, TT, a, b, c, d, i)                                                             \
    do {                                                                                          \
        __m512i x = _mm512_add_epi32(_mm512_add_epi32(a##S, _mm512_set1_epi32((int)K[i])),        \
                                     _mm512_add_epi32(w##S[WI[i]], _mm512_ternarylogic_epi32(b##S, c##S, d##S, TT))); \
        a##S = _mm512_add_epi32(_mm512_rol_epi32(x, SH[i]), b##S);                                \
    } while (0)
#define STEP512X2(TT, a, b, c, d, i) do { STEP512(0, TT, a, b, c, d, i); STEP512(1, TT, a, b, c, d, i); } while (0)
#define ROUND4_512(ST, TT, i)                                                                     \
    ST(TT, a, b, c, d, (i)); ST(TT, d, a, b, c, (i) + 1); ST(TT, c, d, a, b, (i) + 2); ST(TT, b, c, d, a, (i) + 3)
#define ALL_ROUNDS_512(ST)                                                                        \
    ROUND4_512(ST, T_F, 0);  ROUND4_512(ST, T_F, 4);  ROUND4_512(ST, T_F, 8);  ROUND4_512(ST, T_F, 12);   \
    ROUND4_512(ST, T_G, 16); ROUND4_512(ST, T_G, 20); ROUND4_512(ST, T_G, 24); ROUND4_512(ST, T_G, 28);   \
    ROUND4_512(ST, T_H, 32); ROUND4_512(ST, T_H, 36); ROUND4_512(ST, T_H, 40); ROUND4_512(ST, T_H, 44);   \
    ROUND4_512(ST, T_I, 48); ROUND4_512(ST, T_I, 52); ROUND4_512(ST, T_I, 56); ROUND4_512(ST, T_I, 60)

/* two interleaved sets of 16 lanes: state[j][l], l < 32 */
__attribute__((target("avx512f"))) static void md5_mb32_avx512(uint32_t state[4][32], const uint8_t *const data[32], size_t nblocks)
{
    __m512i a0 = _mm512_loadu_si512(&state[0][0]), a1 = _mm512_loadu_si512(&state[0][16]);
    __m512i b0 = _mm512_loadu_si512(&state[1][0]), b1 = _mm512_loadu_si512(&state[1][16]);
    __m512i c0 = _mm512_loadu_si512(&state[2][0]), c1 = _mm512_loadu_si512(&state[2][16]);
    __m512i d0 = _mm512_loadu_si512(&state[3][0]), d1 = _mm512_loadu_si512(&state[3][16]);
    for (size_t blk = 0; blk < nblocks; blk++) {
        __m512i w0[16], w1[16], r[16];
        const size_t off = blk * 64;
        for (int l = 0; l < 16; l++) r[l] = _mm512_loadu_si512(data[l] + off);
        transpose16(r, w0);
        for (int l = 0; l < 16; l++) r[l] = _mm512_loadu_si512(data[16 + l] + off);
        transpose16(r, w1);
        if ((blk & 3) == 0)
            for (int l = 0; l < 32; l++) _mm_prefetch((const char *)(data[l] + off + 768), _MM_HINT_T0);
        const __m512i sa0 = a0, sb0 = b0, sc0 = c0, sd0 = d0, sa1 = a1, sb1 = b1, sc1 = c1, sd1 = d1;
#define ST2(TT, a, b, c, d, i) STEP512X2(TT, a, b, c, d, i)
        ALL_ROUNDS_512(ST2);
#undef ST2
        a0 = _mm512_add_epi32(a0, sa0); b0 = _mm512_add_epi32(b0, sb0); c0 = _mm512_add_epi32(c0, sc0); d0 = _mm512_add_epi32(d0, sd0);
        a1 = _mm512_add_epi32(a1, sa1); b1 = _mm512_add_epi32(b1, sb1); c1 = _mm512_add_epi32(c1, sc1); d1 = _mm512_add_epi32(d1, sd1);
    }
    _mm512_storeu_si512(&state[0][0], a0); _mm512_storeu_si512(&state[0][16], a1);
    _mm512_storeu_si512(&state[1][0], b0); _mm512_storeu_si512(&state[1][16], b1);
    _mm512_storeu_si512(&state[2][0], c0); _mm512_storeu_si512(&state[2][16], c1);
    _mm512_storeu_si512(&state[3][0], d0); _mm512_storeu_si512(&state[3][16], d1);
}

/* one set of 16 lanes (fewer streams per core: more cores busy on a small corpus) */
__attribute__((target("avx512f"))) static void md5_mb16_avx512(uint32_t state[4][32], const uint8_t *const data[32], size_t nblocks)
{
    __m512i a0 = _mm512_loadu_si512(&state[0][0]), b0 = _mm512_loadu_si512(&state[1][0]);
    __m512i c0 = _mm512_loadu_si512(&state[2][0]), d0 = _mm512_loadu_si512(&state[3][0]);
    for (size_t blk = 0; blk < nblocks; blk++) {
        __m512i w0[16], r[16];
        const size_t off = blk * 64;
        for (int l = 0; l < 16; l++) r[l] = _mm512_loadu_si512(data[l] + off);
        transpose16(r, w0);
        if ((blk & 3) == 0)
            for (int l = 0; l < 16; l++) _mm_prefetch((const char *)(data[l] + off + 768), _MM_HINT_T0);
        const __m512i sa0 = a0, sb0 = b0, sc0 = c0, sd0 = d0;
#define ST1(TT, a, b, c, d, i) STEP512(0, TT, a, b, c, d, i)
        ALL_ROUNDS_512(ST1);
#undef ST1
        a0 = _mm512_add_epi32(a0, sa0); b0 = _mm512_add_epi32(b0, sb0); c0 = _mm512_add_epi32(c0, sc0); d0 = _mm512_add_epi32(d0, sd0);
    }
    _mm512_storeu_si512(&state[0][0], a0); _mm512_storeu_si512(&state[1][0], b0);
    _mm512_storeu_si512(&state[2][0], c0); _mm512_storeu_si512(&state[3][0], d0);
}
#endif /* FB_HAVE_AVX2_BUILD */

static int g_mb_lanes = -1;

/* most streams one call advances together on this CPU: 32 (AVX-512), 16 (AVX2) or 1 */
int fb_md5_mb_lanes(void)
{
    if (g_mb_lanes < 0) {
        int lanes = 1;
#if FB_HAVE_AVX2_BUILD
        __builtin_cpu_init();
        if (__builtin_cpu_supports("avx2")) lanes = 16;
        if (__builtin_cpu_supports("avx512f")) lanes = 32;
#endif
        g_mb_lanes = lanes;
    }
    return g_mb_lanes;
}

void fb_md5_mb_update(FbMd5 *const ctx[], const uint8_t *const data[], const size_t len[], int n)
{
    size_t done[FB_MD5_MB_MAX];
    if (n > FB_MD5_MB_MAX) n = FB_MD5_MB_MAX;
    for (int i = 0; i < n; i++) done[i] = 0;
#if FB_HAVE_AVX2_BUILD
    const int lanes = fb_md5_mb_lanes();
    if (n >= 2 && lanes > 1) {
        /* bring every stream to a block boundary, then run the common whole blocks together */
        size_t common = (size_t)-1;
        for (int i = 0; i < n; i++) {
            const size_t used = (size_t)(ctx[i]->nbytes & 63u);
            if (used) {
                const size_t fill = 64 - used < len[i] ? 64 - used : len[i];
                fb_md5_update(ctx[i], data[i], fill);
                done[i] = fill;
            }
            const size_t blocks = (ctx[i]->nbytes & 63u) ? 0 : (len[i] - done[i]) >> 6;
            if (blocks < common) common = blocks;
        }
        if (common > 0 && common != (size_t)-1) {
            /* streams are taken in groups of the widest kernel that is not mostly idle lanes */
            int i0 = 0;
            while (i0 < n) {
                const int left = n - i0;
                const int width = (lanes == 32 && left > 16) ? 32 : 16;
                const int take = left < width ? left : width;
                uint32_t st[4][32];
                const uint8_t *p[32];
                for (int l = 0; l < 32; l++) {
                    const int i = i0 + (l < take ? l : 0);           /* idle lanes shadow the group's first stream */
                    for (int j = 0; j < 4; j++) st[j][l] = ctx[i]->h[j];
                    p[l] = data[i] + done[i];
                }
                if (width == 32) md5_mb32_avx512(st, p, common);
                else if (lanes == 32) md5_mb16_avx512(st, p, common);
                else md5_mb16_avx2(st, p, common);
                for (int l = 0; l < take; l++)
                    for (int j = 0; j < 4; j++) ctx[i0 + l]->h[j] = st[j][l];
                i0 += take;
            }
            for (int i = 0; i < n; i++) { ctx[i]->nbytes += common << 6; done[i] += common << 6; }
        }
    }
#endif
    for (int i = 0; i < n; i++)
        if (done[i] < len[i]) fb_md5_update(ctx[i], data[i] + done[i], len[i] - done[i]);
}
