/*
 * md5.h -- RFC 1321 digest of the input PCM (host side; the reference keeps it
 * in libflake/md5.c and feeds it from flake_encode_frame, encode.c:1006).
 * MD5 is a serial chain over the byte stream, so it stays on a host thread and
 * overlaps the GPU work.
 */
#ifndef FLAKE_B200_MD5_H
#define FLAKE_B200_MD5_H

#include <stdint.h>
#include <stddef.h>

typedef struct FbMd5 {
    uint32_t h[4];
    uint64_t nbytes;
    uint8_t tail[64];
} FbMd5;

void fb_md5_init(FbMd5 *m);
void fb_md5_zero(FbMd5 *m);                       /* the all-zero state flake_encode_init sees */
void fb_md5_update(FbMd5 *m, const void *data, size_t n);
void fb_md5_final(const FbMd5 *m, uint8_t out[16]);   /* non-destructive */
/* interleaved int32 samples packed little-endian at ceil(bps/8) bytes (md5.c:281-320) */
void fb_md5_update_s32(FbMd5 *m, const int32_t *samples, size_t count, int bps);

#endif
