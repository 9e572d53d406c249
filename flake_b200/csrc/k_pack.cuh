/*
 * k_pack.cuh -- parallel frame packer, one CTA per frame (encode.c:700-977,
 * bitio.h, crc.c).  Frames leave the kernel back to back in one contiguous byte
 * stream: their offsets come from a decoupled look-back over the frame lengths
 * (single-pass scan), so a frame is written once, straight to its final place.
 *
 * The reference writes a frame serially through a 32-bit accumulator.  Here
 * every thread owns a contiguous run of samples of a subframe: it sizes its
 * run's codes, a CTA-wide exclusive scan turns sizes into bit offsets, and the
 * run is then emitted MSB-first into a zeroed staging buffer (whole words with
 * plain stores, the two boundary words with atomicOr).  CRC-16 is computed in
 * parallel over 32-bit words with slicing tables and the chunk CRCs are
 * combined with x^(8*len) mod P multipliers, one per chunk (crc.c:24-92 defines
 * P, init 0, no reflection).  The size check that forces VERBATIM subframes
 * (encode.c:949-964) is evaluated on the exact byte count.
 */
#ifndef FLAKE_B200_K_PACK_CUH
#define FLAKE_B200_K_PACK_CUH

#include "dev_common.cuh"

#ifndef FB_PACK_EMIT_UNROLL
#define FB_PACK_EMIT_UNROLL 4     /* Rice codes per iteration of the emit loop */
#endif
#define FB_PACK_THREADS 256      /* upper bound; the engine launches 128 for small frames (>= 96 needed: the
                                  * preamble writers sit at threads 32.. and 64..) */
/* 59 registers: four CTAs per SM; register caps for 5, 6, 8 measured slower (spills) */

/* ---------------- CRC helpers ---------------------------------------- */
__device__ __forceinline__ uint32_t fb_crc16_byte(uint32_t crc, uint32_t byte)
{
    crc ^= byte << 8;
    for (int b = 0; b < 8; b++)
        crc = (crc & 0x8000u) ? ((crc << 1) ^ 0x8005u) & 0xffffu : (crc << 1) & 0xffffu;
    return crc;
}
__device__ __forceinline__ uint32_t fb_crc8_byte(uint32_t crc, uint32_t byte)
{
    crc ^= byte;
    for (int b = 0; b < 8; b++)
        crc = (crc & 0x80u) ? ((crc << 1) ^ 0x07u) & 0xffu : (crc << 1) & 0xffu;
    return crc;
}
/* a * b mod (x^16 + x^15 + x^2 + 1) over GF(2) */
__device__ __forceinline__ uint32_t fb_gf16_mul(uint32_t a, uint32_t b)
{
    uint32_t r = 0;
    for (int i = 15; i >= 0; i--) {
        r = (r & 0x8000u) ? ((r << 1) ^ 0x8005u) & 0xffffu : (r << 1) & 0xffffu;
        if ((b >> i) & 1u) r ^= a;
    }
    return r;
}
/* x^(8*nbytes) mod P */
__device__ __forceinline__ uint32_t fb_gf16_xpow8(uint32_t nbytes)
{
    uint32_t result = 1, base = 0x0100u;      /* x^8 */
    while (nbytes) {
        if (nbytes & 1u) result = fb_gf16_mul(result, base);
        base = fb_gf16_mul(base, base);
        nbytes >>= 1;
    }
    return result;
}

/* ---------------- MSB-first bit emitter -------------------------------- */
/* 64-bit left-aligned accumulator: `fill` (< 32 between calls) bits are pending in the top
 * of `acc`; whole 32-bit words are emitted as soon as they are complete.  The staging buffer
 * holds big-endian bytes and starts zeroed, so all-zero words are skipped.  The first word a
 * writer touches and its last partial word may be shared with the neighbouring writers and
 * are merged with atomicOr; every word in between is covered by this writer alone. */
struct FbBitPut {
    uint32_t *buf;
    uint32_t capw;      /* capacity in words */
    uint32_t widx;      /* current word */
    unsigned long long acc;
    uint32_t fill;
    bool boundary;
};

__device__ __forceinline__ void fb_bp_init(FbBitPut &b, uint32_t *buf, uint32_t capw, uint64_t bitpos)
{
    b.buf = buf; b.capw = capw;
    b.widx = (uint32_t)(bitpos >> 5);
    b.fill = (uint32_t)(bitpos & 31u);
    b.acc = 0;
    b.boundary = true;
}
__device__ __forceinline__ void fb_bp_emit(FbBitPut &b)
{
    const uint32_t w = (uint32_t)(b.acc >> 32);
    if (w && b.widx < b.capw) {
        const uint32_t be = __byte_perm(w, 0, 0x0123);
        if (b.boundary) atomicOr(&b.buf[b.widx], be);
        else b.buf[b.widx] = be;
    }
    b.acc <<= 32; b.widx++; b.fill -= 32; b.boundary = false;
}
/* nbits in 1..32, val < 2^nbits */
__device__ __forceinline__ void fb_bp_put(FbBitPut &b, uint32_t nbits, uint32_t val)
{
    b.acc |= (unsigned long long)val << (64u - b.fill - nbits);
    b.fill += nbits;
    if (b.fill >= 32u) fb_bp_emit(b);
}
__device__ __forceinline__ void fb_bp_skip(FbBitPut &b, uint32_t nzeros)
{
    const unsigned long long f = (unsigned long long)b.fill + nzeros;
    if (f >= 32u) {
        b.fill = 32u;                       /* close the current word ... */
        fb_bp_emit(b);                      /* (acc has no bits beyond it: fill was < 32) */
        b.widx += (uint32_t)(f >> 5) - 1u;  /* ... and jump over the all-zero ones */
        b.fill = (uint32_t)(f & 31u);
    } else {
        b.fill = (uint32_t)f;
    }
}
__device__ __forceinline__ void fb_bp_finish(FbBitPut &b)
{
    const uint32_t w = (uint32_t)(b.acc >> 32);
    if (b.fill && w && b.widx < b.capw) atomicOr(&b.buf[b.widx], __byte_perm(w, 0, 0x0123));
}
__device__ __forceinline__ void fb_bp_put_signed(FbBitPut &b, int nbits, int32_t v)
{
    if (nbits >= 32) { fb_bp_put(b, 32u, (uint32_t)v); return; }
    fb_bp_put(b, (uint32_t)nbits, (uint32_t)v & ((1u << nbits) - 1u));
}
/* Rice code of zig-zag value u with parameter k (bitio.h:120-141): u>>k zeros, a one, k low bits */
__device__ __forceinline__ void fb_bp_put_rice(FbBitPut &b, uint32_t u, uint32_t k)
{
    const uint32_t q = u >> k;
    const uint32_t low = (1u << k) | (u & ((1u << k) - 1u));
    if (q + k + 1u <= 32u) {
        fb_bp_put(b, q + k + 1u, low);      /* leading zeros ride along */
    } else {
        fb_bp_skip(b, q);
        fb_bp_put(b, k + 1u, low);
    }
}

/* ---------------- per-subframe geometry -------------------------------- */
struct FbSubLayout {
    int type, order, obits, wasted, shift, method, porder, psize, pbits;
    uint32_t preamble_bits;     /* header + warm-up + coefficients + method/porder/param[0] */
};

__device__ __forceinline__ FbSubLayout fb_sub_layout(const FbSub *sb, int n, bool force_verbatim)
{
    FbSubLayout L;
    L.type = force_verbatim ? 1 : sb->type;
    L.order = sb->order; L.obits = sb->obits; L.wasted = sb->wasted; L.shift = sb->shift;
    L.method = sb->method; L.porder = sb->porder;
    L.psize = n >> L.porder;
    L.pbits = 4 + L.method;
    uint32_t pre = 8u + (uint32_t)L.wasted;         /* 0, 6-bit type, wasted flag [+ unary] */
    if (L.type == 0) pre += (uint32_t)L.obits;
    else if (L.type == 8)  pre += (uint32_t)(L.order * L.obits) + 6u + (uint32_t)L.pbits;
    else if (L.type == 32) pre += (uint32_t)(L.order * L.obits) + 9u + (uint32_t)L.order * 15u + 6u + (uint32_t)L.pbits;
    L.preamble_bits = pre;
    return L;
}

/*
 * frames[f] -> staged bytes at slots + frames[f].slot, frame_len[f].
 * smem_words: capacity of the dynamic shared staging buffer (0: write the
 * global slot directly).  xpow32[j] = x^(32 j) mod P for the CRC-16 merge;
 * crc16_tables = the four 256-entry slicing tables (uint16, packed in words).
 */
#define FB_SCAN_AGG    (1ull << 62)      /* status = this frame's length */
#define FB_SCAN_PREFIX (2ull << 62)      /* status = sum of the lengths up to and including this frame */
#define FB_SCAN_VALUE  ((1ull << 62) - 1)

/*
 * ticket / status: zeroed before the launch.  Frames are taken in ticket order (not blockIdx
 * order), so every frame a CTA looks back at belongs to a CTA that is already running or done.
 */
__global__ void __launch_bounds__(FB_PACK_THREADS)
k_pack(FbConfig cfg, const FbFrame *frames, const uint32_t *nframes, const void *pcm, int fmt,
       const int32_t *planes, const int32_t *res, FbSub *subs, const uint8_t *ch_modes, uint8_t *slots,
       uint32_t *frame_len, uint32_t *frame_bs, int smem_words,
       const uint16_t *xpow32, const uint32_t *crc16_tables,
       uint32_t *ticket, unsigned long long *status, uint64_t *frame_off, uint8_t *out, FbSummary *summary)
{
    FB_DYN_SMEM(dyn);
    __shared__ uint32_t s_ticket;
    __shared__ unsigned long long s_off;
    __shared__ uint32_t scan_scratch[33];
    __shared__ uint32_t ch_scratch[FB_MAX_CH_UNROLL][32];   /* one scan row per subframe: one barrier per scan */
    __shared__ __align__(16) uint16_t crc_tab[4][256];
    __shared__ uint32_t s_hdr_len;
    __shared__ uint8_t s_hdr[24];
    __shared__ uint8_t s_params[FB_MAX_CH_UNROLL][256];

    const int tid = threadIdx.x, T = blockDim.x;
    if (tid == 0) s_ticket = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t f = s_ticket;
    const uint32_t nf = *nframes;
    if (f >= nf) return;
    const FbFrame fr = frames[f];
    const int n = (int)fr.n, C = cfg.channels;
    const int vsize = fb_verbatim_size(cfg, n);
    /* staging capacity: verbatim encoding plus header slack, see fb_slot_offset */
    const uint32_t cap_bytes = 64u + (uint32_t)(((uint64_t)n * (uint64_t)(C * cfg.bps + 1) + 7u) >> 3);
    const uint32_t capw = (cap_bytes + 3u) >> 2;
    uint32_t *gslot = (uint32_t *)(slots + fr.slot);
    const bool in_smem = capw <= (uint32_t)smem_words;
    uint32_t *wbuf = in_smem ? (uint32_t *)dyn : gslot;

    /* slicing tables: tab[t][b] = CRC-16 of byte b followed by t zero bytes (engine.cu) */
    for (int w = tid; w < 512; w += T) reinterpret_cast<uint32_t *>(&crc_tab[0][0])[w] = crc16_tables[w];
    /* zero the staging buffer while thread 0 builds the header */
    for (uint32_t w = tid; w < capw; w += T) wbuf[w] = 0;

    /* ---- frame header, encode.c:718-764 (thread 0, byte granular) -------- */
    if (tid == 0) {
        int bs0 = -1, bs1 = -1;
        const int tab[15] = {0, 192, 576, 1152, 2304, 4608, 0, 0, 256, 512, 1024, 2048, 4096, 8192, 16384};
        for (int i = 0; i < 15; i++) if (n == tab[i]) { bs0 = i; break; }
        if (bs0 < 0) { bs0 = n <= 256 ? 6 : 7; bs1 = n - 1; }
        const int mode = ch_modes[f];
        int h = 0;
        s_hdr[h++] = 0xff;
        s_hdr[h++] = (uint8_t)(0xf8 | (cfg.allow_vbs ? 1 : 0));
        s_hdr[h++] = (uint8_t)((bs0 << 4) | cfg.sr_code0);
        s_hdr[h++] = (uint8_t)(((mode == 0 ? C - 1 : mode) << 4) | (cfg.bps_code << 1));
        /* write_utf8, encode.c:700-716 */
        const uint32_t v = fr.number;
        if (v < 0x80u) {
            s_hdr[h++] = (uint8_t)v;
        } else {
            const int bytes = (fb_ilog2(v) + 4) / 5;
            int sh = (bytes - 1) * 6;
            s_hdr[h++] = (uint8_t)((256 - (256 >> bytes)) | (v >> sh));
            while (sh >= 6) { sh -= 6; s_hdr[h++] = (uint8_t)(0x80u | ((v >> sh) & 0x3fu)); }
        }
        if (bs1 >= 0) {
            if (bs1 < 256) s_hdr[h++] = (uint8_t)bs1;
            else { s_hdr[h++] = (uint8_t)(bs1 >> 8); s_hdr[h++] = (uint8_t)bs1; }
        }
        if (cfg.sr_code1 > 0) {
            if (cfg.sr_code1 < 256) s_hdr[h++] = (uint8_t)cfg.sr_code1;
            else { s_hdr[h++] = (uint8_t)(cfg.sr_code1 >> 8); s_hdr[h++] = (uint8_t)cfg.sr_code1; }
        }
        uint32_t crc = 0;
        for (int i = 0; i < h; i++) crc = fb_crc8_byte(crc, s_hdr[i]);
        s_hdr[h++] = (uint8_t)crc;
        s_hdr_len = (uint32_t)h;
    }
    __syncthreads();
    const uint32_t hdr_len = s_hdr_len;

    const int R = (n + T - 1) / T;                  /* samples per thread run */
    const int i0 = min(n, tid * R), i1 = min(n, i0 + R);

    /* ---- phase A: size every subframe -----------------------------------------------------
     * Per channel: bits of my run, CTA exclusive scan.  The frame length is known after this
     * phase (a second, trivial round if the frame has to fall back to VERBATIM subframes,
     * encode.c:949-964), so it is published for the look-back of the following frames before
     * a single bit is written. */
    uint32_t myoff[FB_MAX_CH_UNROLL];                /* bit offset of my run inside its subframe's tokens */
    uint64_t chbit[FB_MAX_CH_UNROLL];                /* first bit of each subframe */
    uint64_t total_bits = 0;
    bool verbatim = false;
    /* Rice parameters of every subframe into shared memory */
    for (int c = 0; c < C; c++) {
        const FbSub *sb = &subs[(size_t)f * C + c];
        if (sb->type == 8 || sb->type == 32)
            for (int j = tid; j < (1 << sb->porder); j += T) s_params[c][j] = sb->params[j];
    }
    __syncthreads();
    for (int pass = 0; pass < 2; pass++) {
        if (pass) __syncthreads();                   /* the scan rows of the first round are free again */
        uint64_t bitpos = (uint64_t)hdr_len * 8u;
        for (int c = 0; c < C; c++) {
            const FbSub *sb = &subs[(size_t)f * C + c];
            const FbSubLayout L = fb_sub_layout(sb, n, verbatim);
            const size_t off = (size_t)fr.start * C + (size_t)c * n;
            const int32_t *data = res + off;             /* read for Rice-coded subframes only */
            const bool rice = (L.type == 8 || L.type == 32);
            const uint8_t *kp = s_params[c];

            /* The partition index is tracked incrementally; 16 samples that lie in one
             * partition and start 16-byte aligned are taken with four 128-bit loads and one
             * Rice parameter */
            const int jbeg = rice ? max(i0, L.order) : i0;
            uint32_t mybits = 0;
            if (L.type == 1) {
                mybits = (uint32_t)(i1 - i0) * (uint32_t)L.obits;
            } else if (rice && jbeg < i1) {
                int i = jbeg;
                int p = i / L.psize;
                int nb = (p + 1) * L.psize;
                uint32_t k = kp[p];
                if (p > 0 && i == p * L.psize) mybits += (uint32_t)L.pbits;
                while (i < i1) {
                    if (i == nb) { p++; nb += L.psize; k = kp[p]; mybits += (uint32_t)L.pbits; }
                    if (i + 16 <= min(i1, nb) && (((size_t)(data + i)) & 15u) == 0) {
                        const int4 *src = reinterpret_cast<const int4 *>(data + i);
                        const int4 a = src[0], b = src[1], c4 = src[2], d = src[3];
                        uint32_t q = (fb_zigzag(a.x) >> k) + (fb_zigzag(a.y) >> k) + (fb_zigzag(a.z) >> k) + (fb_zigzag(a.w) >> k);
                        q += (fb_zigzag(b.x) >> k) + (fb_zigzag(b.y) >> k) + (fb_zigzag(b.z) >> k) + (fb_zigzag(b.w) >> k);
                        q += (fb_zigzag(c4.x) >> k) + (fb_zigzag(c4.y) >> k) + (fb_zigzag(c4.z) >> k) + (fb_zigzag(c4.w) >> k);
                        q += (fb_zigzag(d.x) >> k) + (fb_zigzag(d.y) >> k) + (fb_zigzag(d.z) >> k) + (fb_zigzag(d.w) >> k);
                        mybits += q + 16u * (k + 1u);
                        i += 16;
                    } else {
                        mybits += (fb_zigzag(data[i]) >> k) + 1u + k;
                        i++;
                    }
                }
            }
            uint32_t sub_tokens;
            myoff[c] = fb_block_exscan_u32_once(mybits, ch_scratch[c], &sub_tokens);
            chbit[c] = bitpos;
            bitpos += (uint64_t)L.preamble_bits + sub_tokens;
        }
        total_bits = bitpos;
        /* encode.c:949: eof (buffer = 3/2 verbatim size) or larger than the verbatim bound */
        if (pass == 0 && ((total_bits + 7u) >> 3) + 2u > (uint64_t)vsize) { verbatim = true; continue; }
        break;
    }

    uint32_t body = (uint32_t)((total_bits + 7u) >> 3);      /* bytes before the CRC-16 */
    if (body + 2u > cap_bytes) body = cap_bytes - 2u;         /* cannot happen for <= 24-bit input */
    const uint32_t nbytes = body + 2u;
    if (tid == 0)
        *(volatile unsigned long long *)&status[f] = (f ? FB_SCAN_AGG : FB_SCAN_PREFIX) | (unsigned long long)nbytes;

    /* ---- phase B: write the bits (no barrier between the subframes) -------------------------- */
    if (tid == 0) {
        FbBitPut b; fb_bp_init(b, wbuf, capw, 0);
        for (uint32_t i = 0; i < hdr_len; i++) fb_bp_put(b, 8, s_hdr[i]);
        fb_bp_finish(b);
    }
    for (int c = 0; c < C; c++) {
        const FbSub *sb = &subs[(size_t)f * C + c];
        const FbSubLayout L = fb_sub_layout(sb, n, verbatim);
        const size_t off = (size_t)fr.start * C + (size_t)c * n;
        const int32_t *data = res + off;                 /* warm-up samples and residuals of Rice-coded subframes */
        const bool rice = (L.type == 8 || L.type == 32);
        const uint8_t *kp = s_params[c];
        const uint64_t bitpos = chbit[c];
        const int jbeg = rice ? max(i0, L.order) : i0;

        /* preamble: fixed-width fields at known offsets, one writer per field
         * (every field is <= 32 bits, so both words it can touch are merged atomically) */
        {
            const uint64_t pre0 = bitpos + 8u + (uint32_t)L.wasted;          /* after the subframe header */
            if (tid == 0) {
                FbBitPut b; fb_bp_init(b, wbuf, capw, bitpos);
                int code = L.type;
                if (L.type == 8) code = 8 | L.order;
                if (L.type == 32) code = 32 | (L.order - 1);
                fb_bp_put(b, 7, (uint32_t)code);         /* leading 0 + 6-bit type */
                if (L.wasted) { fb_bp_put(b, 1, 1); fb_bp_skip(b, (uint32_t)(L.wasted - 1)); fb_bp_put(b, 1, 1); }
                else fb_bp_put(b, 1, 0);
                if (L.type == 0) fb_bp_put_signed(b, L.obits, sb->first);
                fb_bp_finish(b);
                if (rice) {
                    uint64_t at = pre0 + (uint64_t)(L.order * L.obits);
                    if (L.type == 32) {
                        fb_bp_init(b, wbuf, capw, at);
                        fb_bp_put(b, 4, 14);
                        fb_bp_put_signed(b, 5, L.shift);
                        fb_bp_finish(b);
                        at += 9u + (uint64_t)L.order * 15u;
                    }
                    fb_bp_init(b, wbuf, capw, at);
                    fb_bp_put(b, 2, (uint32_t)L.method);
                    fb_bp_put(b, 4, (uint32_t)L.porder);
                    fb_bp_put(b, (uint32_t)L.pbits, kp[0]);
                    fb_bp_finish(b);
                }
            } else if (rice && tid >= 32 && tid < 32 + L.order) {              /* warm-up samples */
                const int i = tid - 32;
                FbBitPut b; fb_bp_init(b, wbuf, capw, pre0 + (uint64_t)(i * L.obits));
                fb_bp_put_signed(b, L.obits, data[i]);
                fb_bp_finish(b);
            } else if (L.type == 32 && tid >= 64 && tid < 64 + L.order) {      /* LPC coefficients */
                const int i = tid - 64;
                FbBitPut b; fb_bp_init(b, wbuf, capw, pre0 + (uint64_t)(L.order * L.obits) + 9u + (uint64_t)i * 15u);
                fb_bp_put_signed(b, 15, sb->coefs[i]);
                fb_bp_finish(b);
            }
        }
        /* my tokens */
        if ((L.type == 1 && i0 < i1) || (rice && jbeg < i1)) {
            FbBitPut b; fb_bp_init(b, wbuf, capw, bitpos + L.preamble_bits + myoff[c]);
            if (L.type == 1) {
                /* VERBATIM (optimize.c:154-158, 278-289): the transformed samples themselves, taken
                 * from the packed PCM (mono, stereo) or from k_prep's planes (fb_uses_planes); rare */
                const int mode = ch_modes[f];
                for (int i = i0; i < i1; i++)
                    fb_bp_put_signed(b, L.obits, fb_uses_planes(C) ? planes[off + (size_t)i]
                                     : fb_pcm_sample(pcm, fmt, (size_t)fr.start * C, C, c, mode, L.wasted, i));
            } else {
                int i = jbeg;
                int p = i / L.psize;
                int nb = (p + 1) * L.psize;
                uint32_t k = kp[p];
                if (p > 0 && i == p * L.psize) fb_bp_put(b, (uint32_t)L.pbits, k);
                while (i < i1) {
                    if (i == nb) { p++; nb += L.psize; k = kp[p]; fb_bp_put(b, (uint32_t)L.pbits, k); }
                    if (i + 16 <= min(i1, nb) && (((size_t)(data + i)) & 15u) == 0) {
                        const int4 *src = reinterpret_cast<const int4 *>(data + i);
                        /* a short rolled body: the kernel is bound by instruction fetch (the fully
                         * unrolled 16 codes cost 1.68 ms per C2 stream against 1.41) */
#if FB_PACK_EMIT_UNROLL == 4
                        /* the next four samples are requested before the current four are coded: the
                         * emitter's stores may alias the source as far as the compiler knows */
                        int4 nxt = src[0];
#pragma unroll 1
                        for (int g = 0; g < 4; g++) {
                            const int4 v = nxt;
                            if (g < 3) nxt = src[g + 1];
                            fb_bp_put_rice(b, fb_zigzag(v.x), k); fb_bp_put_rice(b, fb_zigzag(v.y), k);
                            fb_bp_put_rice(b, fb_zigzag(v.z), k); fb_bp_put_rice(b, fb_zigzag(v.w), k);
                        }
#elif FB_PACK_EMIT_UNROLL == 2
#pragma unroll 1
                        for (int g = 0; g < 8; g++) {
                            const int2 v = reinterpret_cast<const int2 *>(src)[g];
                            fb_bp_put_rice(b, fb_zigzag(v.x), k); fb_bp_put_rice(b, fb_zigzag(v.y), k);
                        }
#else
#pragma unroll 1
                        for (int g = 0; g < 16; g++) fb_bp_put_rice(b, fb_zigzag(data[i + g]), k);
#endif
                        i += 16;
                    } else {
                        fb_bp_put_rice(b, fb_zigzag(data[i]), k);
                        i++;
                    }
                }
            }
            fb_bp_finish(b);
        }
    }
    __syncthreads();

    /* ---- CRC-16 over the body (crc.c:59-92) ---------------------------------------------
     * The body is cut into T right-aligned chunks of `per` words (virtual zero words in
     * front leave a zero-initialised CRC unchanged, and so do the zero bytes that pad the
     * first real word); each thread runs slicing-by-4 over its chunk and weighs the chunk CRC
     * with x^(32 * per * chunks behind it) mod P, read from the engine's table:
     * crc(A || B) = crc(A) * x^(8 |B|) + crc(B) over GF(2)[x]/P. */
    {
        const uint32_t nwords = (body + 3u) >> 2;
        const uint32_t padbytes = nwords * 4u - body;
        const uint32_t per = (nwords + (uint32_t)T - 1u) / (uint32_t)T;
        const uint32_t lead = per * (uint32_t)T - nwords;
        uint32_t crc = 0;
        for (uint32_t j = 0; j < per; j++) {
            const uint32_t vw = (uint32_t)tid * per + j;
            if (vw < lead) continue;
            const uint32_t wi = vw - lead;
            uint32_t word;
            if (padbytes == 0) {
                word = __byte_perm(wbuf[wi], 0, 0x0123);
            } else {
                const uint32_t hi = wi ? __byte_perm(wbuf[wi - 1], 0, 0x0123) : 0u;
                const uint32_t lo = __byte_perm(wbuf[wi], 0, 0x0123);
                word = (hi << (32u - 8u * padbytes)) | (lo >> (8u * padbytes));
            }
            const uint32_t x = word ^ (crc << 16);
            crc = (uint32_t)crc_tab[3][x >> 24] ^ (uint32_t)crc_tab[2][(x >> 16) & 255u] ^
                  (uint32_t)crc_tab[1][(x >> 8) & 255u] ^ (uint32_t)crc_tab[0][x & 255u];
        }
        /* crc(message) = XOR over chunks of crc(chunk) * x^(8 * bytes behind the chunk) */
        crc = fb_gf16_mul(crc, xpow32[per * (uint32_t)(T - 1 - tid)]);
        const int lane = tid & 31, warp = tid >> 5;
        for (int o = 16; o > 0; o >>= 1) crc ^= __shfl_xor_sync(FB_FULL_MASK, crc, o);
        if (lane == 0) scan_scratch[warp] = crc;
        __syncthreads();
        if (tid == 0) {
            uint32_t acc = 0;
            for (int w = 0; w < (T >> 5); w++) acc ^= scan_scratch[w];
            FbBitPut b; fb_bp_init(b, wbuf, capw, (uint64_t)body * 8u);
            fb_bp_put(b, 16, acc);
            fb_bp_finish(b);
        }
    }
    __syncthreads();

    /* ---- where the frame goes: exclusive prefix of the frame lengths ---------------------
     * Decoupled look-back (warp 0): walk back over the predecessors' published lengths, 32 at
     * a time, until one of them already knows its inclusive prefix.  Every predecessor
     * published its length before it started writing bits, so this rarely has to wait. */
    if (tid < 32) {
        const int lane = tid;
        unsigned long long excl = 0;
        if (f > 0) {
            long long base = (long long)f - 1;
            for (;;) {
                const long long idx = base - lane;
                unsigned long long v = FB_SCAN_PREFIX;              /* in front of frame 0: prefix 0 */
                if (idx >= 0) {
                    do { v = *(volatile unsigned long long *)&status[idx]; } while ((v >> 62) == 0);
                }
                const unsigned pm = __ballot_sync(FB_FULL_MASK, (v >> 62) == 2ull);
                const int stop = pm ? __ffs((int)pm) - 1 : 31;      /* nearest lane that holds a prefix */
                excl += fb_warp_sum_u64(lane <= stop ? (v & FB_SCAN_VALUE) : 0ull);
                if (pm) break;
                base -= 32;
            }
        }
        if (lane == 0) {
            if (f > 0) *(volatile unsigned long long *)&status[f] = FB_SCAN_PREFIX | (excl + (unsigned long long)nbytes);
            s_off = excl;
        }
    }
    __syncthreads();
    const unsigned long long off = s_off;

    /* ---- staged words (16-byte aligned) -> out + off (any alignment) ---------------------- */
    {
        const uint32_t *src = wbuf;
        uint8_t *dst = out + off;
        const uint32_t len = nbytes;
        const uint32_t mis = (uint32_t)((size_t)dst & 3u);
        const uint32_t head = mis ? min(len, 4u - mis) : 0u;        /* bytes up to the first aligned word */
        if ((uint32_t)tid < head) dst[tid] = (uint8_t)(src[0] >> (8u * (uint32_t)tid));
        if (len > head) {
            uint32_t *dw = (uint32_t *)(dst + head);
            const uint32_t rest = len - head, nwords = rest >> 2;
            /* destination word j holds source bytes [head + 4j, head + 4j + 4) */
            const uint32_t sel = head == 0 ? 0x3210u : head == 1 ? 0x4321u : head == 2 ? 0x5432u : 0x6543u;
            for (uint32_t j = tid; j < nwords; j += T) {
                const uint32_t a = src[j], b2 = head ? src[j + 1] : 0u;
                dw[j] = __byte_perm(a, b2, sel);
            }
            const uint32_t tail = rest & 3u;
            if ((uint32_t)tid < tail) {
                const uint32_t k = head + nwords * 4u + (uint32_t)tid;
                dst[k] = (uint8_t)(src[k >> 2] >> (8u * (k & 3u)));
            }
        }
    }
    if (tid == 0) {
        frame_len[f] = nbytes;
        frame_off[f] = off;
        if (frame_bs) frame_bs[f] = (uint32_t)n;
        atomicMax(&summary->max_frame_bytes, nbytes);
        atomicMax(&summary->min_frame_inv, ~nbytes);              /* min over the frames, kept as a max */
        if (verbatim) {
            atomicAdd(&summary->verbatim_frames, 1u);
            for (int c = 0; c < C; c++) subs[(size_t)f * C + c].type = 1;
        }
        if (f == nf - 1) {                                           /* the last frame closes the chunk summary */
            summary->nframes = nf;
            summary->total_bytes = off + nbytes;
        }
    }
}

#endif
