/*
 * md5_mb.h -- multi-buffer MD5 over independent streams (md5_mb.c): what lets one host
 * core hash the PCM of 16 files of a corpus at once (flake_b200_encode_corpus).
 */
#ifndef FLAKE_B200_MD5_MB_H
#define FLAKE_B200_MD5_MB_H

#include "md5.h"

#define FB_MD5_MB_MAX 32

/* streams advanced together by one call on this CPU: 32 with AVX-512, 16 with AVX2, else 1 (scalar) */
int fb_md5_mb_lanes(void);

/* fb_md5_update(ctx[i], data[i], len[i]) for i < n (n <= FB_MD5_MB_MAX), the common whole
 * 64-byte blocks of all streams in SIMD lanes, the rest per stream */
void fb_md5_mb_update(FbMd5 *const ctx[], const uint8_t *const data[], const size_t len[], int n);

#endif
