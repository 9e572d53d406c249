/*
 * batch_example.c -- a plain-C caller of flake_b200's libflake.so, in the shape of the
 * reference's util/api_example.c (init -> encode -> rewrite STREAMINFO -> close) but going
 * through BOTH entry points: the per-block flake_encode_frame loop of the reference API and
 * the flake_b200_encode_stream batch call.  The two files it writes must be identical.
 *
 *   cc batch_example.c -I../../include -L../../flake_b200/lib -lflake -o batch_example
 *   ./batch_example out_blocks.flac out_batch.flac [level]
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "flake.h"
#include "flake_b200.h"

static int32_t *make_pcm(unsigned n, int ch)
{
    int32_t *p = malloc(sizeof(int32_t) * (size_t)n * ch);
    uint32_t lfsr = 0xF1A4E000u;
    double ph = 0.0;
    for (unsigned i = 0; i < n; i++) {
        lfsr = lfsr * 1664525u + 1013904223u;
        ph += 0.031;
        /* triangle wave + noise, deterministic without libm */
        double t = ph - (long)ph;
        int32_t tone = (int32_t)((t < 0.5 ? t : 1.0 - t) * 40000.0) - 10000;
        int32_t noise = (int32_t)(lfsr >> 24) - 128;
        for (int c = 0; c < ch; c++) p[(size_t)i * ch + c] = tone / (c + 1) + noise * (c + 1);
    }
    return p;
}

static int write_file(const char *path, FlakeContext *s, int header_len, const unsigned char *frames, size_t nbytes)
{
    FlakeStreaminfo si;
    unsigned char sib[34];
    FILE *f = fopen(path, "wb");
    if (!f) return -1;
    if (flake_get_streaminfo(s, &si)) { fclose(f); return -1; }
    flake_write_streaminfo(&si, sib);
    memcpy(s->header + 8, sib, 34);             /* what flake/flake.c does with fseek(8) */
    fwrite(s->header, 1, (size_t)header_len, f);
    fwrite(frames, 1, nbytes, f);
    fclose(f);
    return 0;
}

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s blocks.flac batch.flac [level]\n", argv[0]); return 2; }
    const int level = argc > 3 ? atoi(argv[3]) : 8;
    const unsigned n = 4096u * 20u + 1234u;
    const int ch = 2;
    int32_t *pcm = make_pcm(n, ch);

    for (int pass = 0; pass < 2; pass++) {
        FlakeContext s;
        memset(&s, 0, sizeof s);
        s.channels = ch; s.sample_rate = 44100; s.bits_per_sample = 16; s.samples = n;
        s.params.compression = level;
        if (flake_set_defaults(&s.params)) return 1;
        if (flake_validate_params(&s) < 0) return 1;
        const int header_len = flake_encode_init(&s);
        if (header_len < 0) { fprintf(stderr, "flake_encode_init failed (no CUDA device?)\n"); return 1; }

        unsigned long long cap = flake_b200_max_encoded_size(&s, n);
        unsigned char *out = malloc(cap);
        size_t total = 0;
        if (pass == 0) {
            const unsigned bs = (unsigned)s.params.block_size;
            const unsigned char *buf = flake_get_buffer(&s);
            for (unsigned i = 0; i < n; i += bs) {
                const int nr = (int)(n - i < bs ? n - i : bs);
                const int fs = flake_encode_frame(&s, pcm + (size_t)i * ch, nr);
                if (fs < 0) { fprintf(stderr, "flake_encode_frame failed\n"); return 1; }
                memcpy(out + total, buf, (size_t)fs);
                total += (size_t)fs;
            }
        } else {
            unsigned nframes = 0;
            const long long rc = flake_b200_encode_stream(&s, pcm, FLAKE_B200_PCM_S32, n, out, cap,
                                                          NULL, NULL, 0, &nframes);
            if (rc < 0) { fprintf(stderr, "batch encode failed: %s\n", flake_b200_last_error(&s)); return 1; }
            total = (size_t)rc;
            printf("batch: %u frames, %lld bytes\n", nframes, rc);
        }
        if (write_file(argv[1 + pass], &s, header_len, out, total)) return 1;
        free(out);
        flake_encode_close(&s);
    }
    free(pcm);
    return 0;
}
