/* md5.c -- RFC 1321, see md5.h */
#include "md5.h"
#include <string.h>

#define ROL(x, s) (((x) << (s)) | ((x) >> (32 - (s))))
#define F1(b, c, d) ((d) ^ ((b) & ((c) ^ (d))))
#define F2(b, c, d) ((c) ^ ((d) & ((b) ^ (c))))
#define F3(b, c, d) ((b) ^ ((c) ^ (d)))
#define F4(b, c, d) ((c) ^ ((b) | ~(d)))
#define RND(f, a, b, c, d, w, k, s) do { (a) += f((b), (c), (d)) + (w) + (k); (a) = ROL((a), (s)) + (b); } while (0)
/* round 2: G(b,c,d) = (b & d) | (c & ~d) with disjoint terms, so the part that does not
 * depend on b (the value produced by the previous step) is added early and only one AND
 * and one ADD sit between b and the rotate */
#define RND2(a, b, c, d, w, k, s) do { (a) += ((c) & ~(d)) + (w) + (k); (a) += ((b) & (d)); (a) = ROL((a), (s)) + (b); } while (0)

static void md5_blocks(uint32_t h[4], const uint8_t *p, size_t nblocks)
{
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3];
    while (nblocks--) {
        uint32_t w[16];
        memcpy(w, p, 64);                       /* little-endian host (x86-64) */
        const uint32_t a0 = a, b0 = b, c0 = c, d0 = d;
        RND(F1, a, b, c, d, w[0],  0xd76aa478,  7); RND(F1, d, a, b, c, w[1],  0xe8c7b756, 12);
        RND(F1, c, d, a, b, w[2],  0x242070db, 17); RND(F1, b, c, d, a, w[3],  0xc1bdceee, 22);
        RND(F1, a, b, c, d, w[4],  0xf57c0faf,  7); RND(F1, d, a, b, c, w[5],  0x4787c62a, 12);
        RND(F1, c, d, a, b, w[6],  0xa8304613, 17); RND(F1, b, c, d, a, w[7],  0xfd469501, 22);
        RND(F1, a, b, c, d, w[8],  0x698098d8,  7); RND(F1, d, a, b, c, w[9],  0x8b44f7af, 12);
        RND(F1, c, d, a, b, w[10], 0xffff5bb1, 17); RND(F1, b, c, d, a, w[11], 0x895cd7be, 22);
        RND(F1, a, b, c, d, w[12], 0x6b901122,  7); RND(F1, d, a, b, c, w[13], 0xfd987193, 12);
        RND(F1, c, d, a, b, w[14], 0xa679438e, 17); RND(F1, b, c, d, a, w[15], 0x49b40821, 22);

        RND2(a, b, c, d, w[1],  0xf61e2562,  5); RND2(d, a, b, c, w[6],  0xc040b340,  9);
        RND2(c, d, a, b, w[11], 0x265e5a51, 14); RND2(b, c, d, a, w[0],  0xe9b6c7aa, 20);
        RND2(a, b, c, d, w[5],  0xd62f105d,  5); RND2(d, a, b, c, w[10], 0x02441453,  9);
        RND2(c, d, a, b, w[15], 0xd8a1e681, 14); RND2(b, c, d, a, w[4],  0xe7d3fbc8, 20);
        RND2(a, b, c, d, w[9],  0x21e1cde6,  5); RND2(d, a, b, c, w[14], 0xc33707d6,  9);
        RND2(c, d, a, b, w[3],  0xf4d50d87, 14); RND2(b, c, d, a, w[8],  0x455a14ed, 20);
        RND2(a, b, c, d, w[13], 0xa9e3e905,  5); RND2(d, a, b, c, w[2],  0xfcefa3f8,  9);
        RND2(c, d, a, b, w[7],  0x676f02d9, 14); RND2(b, c, d, a, w[12], 0x8d2a4c8a, 20);

        RND(F3, a, b, c, d, w[5],  0xfffa3942,  4); RND(F3, d, a, b, c, w[8],  0x8771f681, 11);
        RND(F3, c, d, a, b, w[11], 0x6d9d6122, 16); RND(F3, b, c, d, a, w[14], 0xfde5380c, 23);
        RND(F3, a, b, c, d, w[1],  0xa4beea44,  4); RND(F3, d, a, b, c, w[4],  0x4bdecfa9, 11);
        RND(F3, c, d, a, b, w[7],  0xf6bb4b60, 16); RND(F3, b, c, d, a, w[10], 0xbebfbc70, 23);
        RND(F3, a, b, c, d, w[13], 0x289b7ec6,  4); RND(F3, d, a, b, c, w[0],  0xeaa127fa, 11);
        RND(F3, c, d, a, b, w[3],  0xd4ef3085, 16); RND(F3, b, c, d, a, w[6],  0x04881d05, 23);
        RND(F3, a, b, c, d, w[9],  0xd9d4d039,  4); RND(F3, d, a, b, c, w[12], 0xe6db99e5, 11);
        RND(F3, c, d, a, b, w[15], 0x1fa27cf8, 16); RND(F3, b, c, d, a, w[2],  0xc4ac5665, 23);

        RND(F4, a, b, c, d, w[0],  0xf4292244,  6); RND(F4, d, a, b, c, w[7],  0x432aff97, 10);
        RND(F4, c, d, a, b, w[14], 0xab9423a7, 15); RND(F4, b, c, d, a, w[5],  0xfc93a039, 21);
        RND(F4, a, b, c, d, w[12], 0x655b59c3,  6); RND(F4, d, a, b, c, w[3],  0x8f0ccc92, 10);
        RND(F4, c, d, a, b, w[10], 0xffeff47d, 15); RND(F4, b, c, d, a, w[1],  0x85845dd1, 21);
        RND(F4, a, b, c, d, w[8],  0x6fa87e4f,  6); RND(F4, d, a, b, c, w[15], 0xfe2ce6e0, 10);
        RND(F4, c, d, a, b, w[6],  0xa3014314, 15); RND(F4, b, c, d, a, w[13], 0x4e0811a1, 21);
        RND(F4, a, b, c, d, w[4],  0xf7537e82,  6); RND(F4, d, a, b, c, w[11], 0xbd3af235, 10);
        RND(F4, c, d, a, b, w[2],  0x2ad7d2bb, 15); RND(F4, b, c, d, a, w[9],  0xeb86d391, 21);
        a += a0; b += b0; c += c0; d += d0;
        p += 64;
    }
    h[0] = a; h[1] = b; h[2] = c; h[3] = d;
}

void fb_md5_init(FbMd5 *m)
{
    m->h[0] = 0x67452301u; m->h[1] = 0xefcdab89u; m->h[2] = 0x98badcfeu; m->h[3] = 0x10325476u;
    m->nbytes = 0;
}

void fb_md5_zero(FbMd5 *m) { memset(m, 0, sizeof *m); }

void fb_md5_update(FbMd5 *m, const void *data, size_t n)
{
    const uint8_t *p = (const uint8_t *)data;
    size_t used = (size_t)(m->nbytes & 63u);
    m->nbytes += n;
    if (used) {
        size_t room = 64 - used;
        if (n < room) { memcpy(m->tail + used, p, n); return; }
        memcpy(m->tail + used, p, room);
        md5_blocks(m->h, m->tail, 1);
        p += room; n -= room;
    }
    if (n >= 64) { md5_blocks(m->h, p, n >> 6); p += n & ~(size_t)63; n &= 63; }
    memcpy(m->tail, p, n);
}

void fb_md5_final(const FbMd5 *m0, uint8_t out[16])
{
    FbMd5 m = *m0;
    uint64_t bits = m.nbytes << 3;
    size_t used = (size_t)(m.nbytes & 63u);
    m.tail[used++] = 0x80;
    if (used > 56) { memset(m.tail + used, 0, 64 - used); md5_blocks(m.h, m.tail, 1); used = 0; }
    memset(m.tail + used, 0, 56 - used);
    for (int i = 0; i < 8; i++) m.tail[56 + i] = (uint8_t)(bits >> (8 * i));
    md5_blocks(m.h, m.tail, 1);
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) out[4 * i + j] = (uint8_t)(m.h[i] >> (8 * j));
}

void fb_md5_update_s32(FbMd5 *m, const int32_t *s, size_t count, int bps)
{
    const int bytes = (bps + 7) >> 3;
    if (bytes == 4) { fb_md5_update(m, s, count * 4); return; }
    uint8_t buf[3 * 4096];
    while (count) {
        size_t take = count < 4096 ? count : 4096, k = 0;
        if (bytes == 2) {
            for (size_t i = 0; i < take; i++) { uint32_t x = (uint32_t)s[i]; buf[k++] = (uint8_t)x; buf[k++] = (uint8_t)(x >> 8); }
        } else if (bytes == 3) {
            for (size_t i = 0; i < take; i++) { uint32_t x = (uint32_t)s[i]; buf[k++] = (uint8_t)x; buf[k++] = (uint8_t)(x >> 8); buf[k++] = (uint8_t)(x >> 16); }
        } else {
            for (size_t i = 0; i < take; i++) buf[k++] = (uint8_t)s[i];
        }
        fb_md5_update(m, buf, k);
        s += take; count -= take;
    }
}
