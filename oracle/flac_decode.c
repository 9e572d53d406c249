/*
 * flac_decode.c -- test-only FLAC stream decoder.
 *
 * TEST INFRASTRUCTURE ONLY (see flake_oracle.h).  The reference ships no
 * decoder and its acceptance check is the external `flac -t`
 * (util/flake-test.sh:10,30), which is not installed here; this file replaces
 * it for the "decodes bit-exactly" gate: it parses every frame, checks CRC-8
 * and CRC-16, reconstructs the PCM and (with a stream header) verifies the
 * STREAMINFO MD5.  Follows the published FLAC format (frame header, subframe
 * types CONSTANT / VERBATIM / FIXED / LPC, partitioned Rice / Rice2 residual
 * with escape codes, stereo decorrelation modes, wasted bits).
 *
 * error codes: 1 bad marker/metadata, 2 lost sync, 3 reserved field, 4 CRC-8,
 * 5 CRC-16, 6 truncated, 7 output overflow, 8 bad subframe, 9 header/streaminfo
 * mismatch.
 */
#include "flake_oracle.h"

#include <stdlib.h>
#include <string.h>

typedef struct {
    const uint8_t *d;
    size_t len;       /* bytes */
    size_t pos;       /* bit position */
    int eof;
} BitSrc;

static uint32_t br_bit(BitSrc *b)
{
    if ((b->pos >> 3) >= b->len) { b->eof = 1; return 0; }
    uint32_t v = (b->d[b->pos >> 3] >> (7 - (b->pos & 7))) & 1u;
    b->pos++;
    return v;
}

static uint32_t br_bits(BitSrc *b, int n)      /* n <= 32 */
{
    uint64_t v = 0;
    while (n > 0) {
        if ((b->pos >> 3) >= b->len) { b->eof = 1; return 0; }
        int avail = 8 - (int)(b->pos & 7);
        int take = n < avail ? n : avail;
        uint32_t byte = b->d[b->pos >> 3];
        uint32_t chunk = (byte >> (avail - take)) & ((1u << take) - 1u);
        v = (v << take) | chunk;
        b->pos += (size_t)take; n -= take;
    }
    return (uint32_t)v;
}

static int32_t br_sbits(BitSrc *b, int n)
{
    if (n == 0) return 0;
    uint32_t v = br_bits(b, n);
    if (n < 32 && (v >> (n - 1))) v |= ~((1u << n) - 1u);
    return (int32_t)v;
}

static uint32_t br_unary(BitSrc *b)
{
    uint32_t q = 0;
    while (!b->eof && br_bit(b) == 0) q++;
    return q;
}

static const int dec_bs[16]   = {0,192,576,1152,2304,4608,0,0,256,512,1024,2048,4096,8192,16384,32768};
static const int dec_rate[12] = {0,88200,176400,192000,8000,16000,22050,24000,32000,44100,48000,96000};
static const int dec_bps[8]   = {0,8,12,0,16,20,24,0};

static int decode_residual(BitSrc *b, int32_t *out, int n, int order)
{
    int method = (int)br_bits(b, 2);
    if (method > 1) return 3;
    int pbits = method ? 5 : 4, esc = method ? 31 : 15;
    int porder = (int)br_bits(b, 4);
    int parts = 1 << porder;
    if ((n >> porder) << porder != n && porder) return 8;
    int psize = n >> porder, i = order;
    for (int p = 0; p < parts; p++) {
        int k = (int)br_bits(b, pbits);
        int cnt = psize - (p ? 0 : order);
        if (cnt < 0) return 8;
        if (k == esc) {
            int raw = (int)br_bits(b, 5);
            for (int j = 0; j < cnt; j++) out[i++] = br_sbits(b, raw);
        } else {
            for (int j = 0; j < cnt; j++) {
                uint32_t q = br_unary(b);
                uint32_t u = (q << k) | (k ? br_bits(b, k) : 0);
                out[i++] = (int32_t)(u >> 1) ^ -(int32_t)(u & 1);
            }
        }
        if (b->eof) return 6;
    }
    return 0;
}

static int decode_subframe(BitSrc *b, int32_t *out, int n, int bps)
{
    if (br_bit(b)) return 3;
    int type = (int)br_bits(b, 6);
    int wasted = 0;
    if (br_bit(b)) wasted = 1 + (int)br_unary(b);
    bps -= wasted;
    if (bps < 1) return 8;

    if (type == 0) {
        int32_t v = br_sbits(b, bps);
        for (int i = 0; i < n; i++) out[i] = v;
    } else if (type == 1) {
        for (int i = 0; i < n; i++) out[i] = br_sbits(b, bps);
    } else if (type >= 8 && type <= 12) {
        int order = type - 8;
        if (order > n) return 8;
        for (int i = 0; i < order; i++) out[i] = br_sbits(b, bps);
        int e = decode_residual(b, out, n, order);
        if (e) return e;
        for (int i = order; i < n; i++) {
            int64_t p;
            switch (order) {
            case 0: p = 0; break;
            case 1: p = out[i-1]; break;
            case 2: p = 2*(int64_t)out[i-1] - out[i-2]; break;
            case 3: p = 3*(int64_t)out[i-1] - 3*(int64_t)out[i-2] + out[i-3]; break;
            default: p = 4*(int64_t)out[i-1] - 6*(int64_t)out[i-2] + 4*(int64_t)out[i-3] - out[i-4]; break;
            }
            out[i] = (int32_t)((int64_t)out[i] + p);
        }
    } else if (type >= 32) {
        int order = type - 31;
        if (order > n) return 8;
        for (int i = 0; i < order; i++) out[i] = br_sbits(b, bps);
        int prec = (int)br_bits(b, 4) + 1;
        if (prec == 16) return 3;
        int shift = br_sbits(b, 5);
        if (shift < 0) return 8;
        int32_t coef[32];
        for (int i = 0; i < order; i++) coef[i] = br_sbits(b, prec);
        int e = decode_residual(b, out, n, order);
        if (e) return e;
        for (int i = order; i < n; i++) {
            int64_t p = 0;
            for (int j = 0; j < order; j++) p += (int64_t)coef[j] * out[i-1-j];
            out[i] = (int32_t)((int64_t)out[i] + (p >> shift));
        }
    } else {
        return 3;
    }
    if (wasted) for (int i = 0; i < n; i++) out[i] = (int32_t)((uint32_t)out[i] << wasted);
    return b->eof ? 6 : 0;
}

/* RFC-1321 over the decoded PCM, via the oracle's helper */
int64_t orc_flac_decode(const uint8_t *data, size_t len, int has_header,
                        int32_t *pcm, uint64_t pcm_cap, OrcDecInfo *info)
{
    OrcDecInfo local; if (!info) { memset(&local, 0, sizeof local); info = &local; }
    int channels = info->channels, bps = info->bps, rate = info->sample_rate;
    uint8_t want_md5[16]; int have_md5 = 0;
    size_t pos = 0;
    info->error = 0; info->error_pos = 0; info->nframes = 0; info->decoded_samples = 0;
    info->min_bs = 0xffffffffu; info->max_bs = 0; info->max_frame_bytes = 0; info->md5_ok = -1;
    info->total_samples = 0;
    uint32_t si_min_bs = 0, si_max_bs = 0, si_max_frame = 0;

    if (has_header) {
        if (len < 42 || memcmp(data, "fLaC", 4)) { info->error = 1; return -1; }
        pos = 4;
        int last = 0, first = 1;
        while (!last) {
            if (pos + 4 > len) { info->error = 1; return -1; }
            last = data[pos] >> 7;
            int type = data[pos] & 0x7f;
            size_t sz = ((size_t)data[pos+1] << 16) | ((size_t)data[pos+2] << 8) | data[pos+3];
            pos += 4;
            if (pos + sz > len) { info->error = 1; return -1; }
            if (first) {
                if (type != 0 || sz != 34) { info->error = 1; return -1; }
                const uint8_t *s = data + pos;
                si_min_bs = ((uint32_t)s[0] << 8) | s[1];
                si_max_bs = ((uint32_t)s[2] << 8) | s[3];
                si_max_frame = ((uint32_t)s[7] << 16) | ((uint32_t)s[8] << 8) | s[9];
                rate = (int)(((uint32_t)s[10] << 12) | ((uint32_t)s[11] << 4) | (s[12] >> 4));
                channels = ((s[12] >> 1) & 7) + 1;
                bps = (((s[12] & 1) << 4) | (s[13] >> 4)) + 1;
                info->total_samples = ((uint64_t)(s[13] & 15) << 32) | ((uint64_t)s[14] << 24) |
                                      ((uint64_t)s[15] << 16) | ((uint64_t)s[16] << 8) | s[17];
                memcpy(want_md5, s + 18, 16); have_md5 = 1;
                first = 0;
            }
            pos += sz;
        }
        info->channels = channels; info->bps = bps; info->sample_rate = rate;
    }
    if (channels < 1 || channels > 8 || bps < 4 || bps > 32) { info->error = 1; return -1; }

    int32_t *chbuf = (int32_t *)malloc(sizeof(int32_t) * 65536u * (size_t)channels);
    uint64_t done = 0, expect_number = 0;
    int variable = -1;

    while (pos < len) {
        size_t fstart = pos;
        BitSrc b = { data + fstart, len - fstart, 0, 0 };
        info->error_pos = fstart;
        uint32_t sync = br_bits(&b, 14);
        if (sync != 0x3ffe) { info->error = 2; goto fail; }
        if (br_bit(&b)) { info->error = 3; goto fail; }
        int vb = (int)br_bit(&b);
        if (variable < 0) variable = vb; else if (variable != vb) { info->error = 3; goto fail; }
        int bsc = (int)br_bits(&b, 4), src = (int)br_bits(&b, 4);
        int chc = (int)br_bits(&b, 4), bpc = (int)br_bits(&b, 3);
        if (br_bit(&b)) { info->error = 3; goto fail; }
        /* UTF-8 style number, up to 36 bits */
        uint64_t number;
        {
            uint32_t first = br_bits(&b, 8);
            int extra = 0;
            if (first < 0x80) { number = first; }
            else {
                uint32_t m = 0x80; while (first & m) { extra++; m >>= 1; }
                if (extra < 2 || extra > 7) { info->error = 3; goto fail; }
                extra -= 1;
                number = first & (m - 1);
                for (int i = 0; i < extra; i++) {
                    uint32_t c = br_bits(&b, 8);
                    if ((c & 0xc0) != 0x80) { info->error = 3; goto fail; }
                    number = (number << 6) | (c & 0x3f);
                }
            }
        }
        int n;
        if (bsc == 0) { info->error = 3; goto fail; }
        else if (bsc == 6) n = (int)br_bits(&b, 8) + 1;
        else if (bsc == 7) n = (int)br_bits(&b, 16) + 1;
        else n = dec_bs[bsc];
        if (src == 12) (void)br_bits(&b, 8);
        else if (src == 13 || src == 14) (void)br_bits(&b, 16);
        else if (src == 15) { info->error = 3; goto fail; }
        else if (src > 0 && src < 12 && rate && dec_rate[src] != rate) { info->error = 9; goto fail; }
        size_t hbytes = b.pos >> 3;
        uint32_t crc8 = br_bits(&b, 8);
        if (b.eof) { info->error = 6; goto fail; }
        if (orc_crc8(data + fstart, hbytes) != crc8) { info->error = 4; goto fail; }
        if (bpc && dec_bps[bpc] && dec_bps[bpc] != bps) { info->error = 9; goto fail; }

        int nch, mode = 0;
        if (chc < 8) nch = chc + 1;
        else if (chc <= 10) { nch = 2; mode = chc; }
        else { info->error = 3; goto fail; }
        if (nch != channels) { info->error = 9; goto fail; }

        if (vb) { if (number != (expect_number & 0xfffffffffull)) { /* sample number */ info->error = 9; goto fail; } }
        else    { if (number != (uint64_t)info->nframes) { info->error = 9; goto fail; } }

        for (int c = 0; c < nch; c++) {
            int sb = bps + ((mode == 8 && c == 1) || (mode == 9 && c == 0) || (mode == 10 && c == 1));
            int e = decode_subframe(&b, chbuf + (size_t)c * 65536u, n, sb);
            if (e) { info->error = e; goto fail; }
        }
        b.pos = (b.pos + 7) & ~(size_t)7;
        size_t body = b.pos >> 3;
        uint32_t crc16 = br_bits(&b, 16);
        if (b.eof) { info->error = 6; goto fail; }
        if (orc_crc16(data + fstart, body) != crc16) { info->error = 5; goto fail; }
        size_t fbytes = body + 2;

        if (done + (uint64_t)n > pcm_cap) { info->error = 7; goto fail; }
        int32_t *c0 = chbuf, *c1 = chbuf + 65536u;
        for (int i = 0; i < n; i++) {
            int32_t *dst = pcm + (done + (uint64_t)i) * (uint64_t)channels;
            if (mode == 8)       { dst[0] = c0[i]; dst[1] = c0[i] - c1[i]; }
            else if (mode == 9)  { dst[0] = c0[i] + c1[i]; dst[1] = c1[i]; }
            else if (mode == 10) {
                int32_t m = c0[i], s = c1[i];
                m = (int32_t)(((uint32_t)m << 1) | ((uint32_t)s & 1u));
                dst[0] = (m + s) >> 1; dst[1] = (m - s) >> 1;
            } else {
                for (int c = 0; c < channels; c++) dst[c] = chbuf[(size_t)c * 65536u + (size_t)i];
            }
        }
        done += (uint64_t)n;
        expect_number += (uint64_t)n;
        info->nframes++;
        if ((uint32_t)n < info->min_bs) info->min_bs = (uint32_t)n;
        if ((uint32_t)n > info->max_bs) info->max_bs = (uint32_t)n;
        if (fbytes > info->max_frame_bytes) info->max_frame_bytes = (uint32_t)fbytes;
        pos = fstart + fbytes;
    }
    free(chbuf);
    info->decoded_samples = done;
    if (have_md5) {
        uint8_t got[16];
        orc_md5_pcm(pcm, channels, bps, done, got);
        info->md5_ok = memcmp(got, want_md5, 16) == 0;
        if (si_max_bs && info->max_bs > si_max_bs) info->error = 9;
        if (si_min_bs && info->nframes > 1 && info->min_bs < si_min_bs && 0) info->error = 9;
        if (si_max_frame && info->max_frame_bytes > si_max_frame) info->error = 9;
    }
    return info->error ? -1 : (int64_t)done;
fail:
    free(chbuf);
    info->decoded_samples = done;
    return -1;
}
