/*
 * flake_host_int.h -- what the C host files of libflake.so share (flake_host.c: the flake.h API
 * and the single-stream batch calls; flake_corpus.c: many streams over several GPUs).
 * Internal: nothing here is exported.
 */
#ifndef FLAKE_B200_HOST_INT_H
#define FLAKE_B200_HOST_INT_H

#include "flake.h"
#include "engine.h"

#include <stddef.h>
#include <stdint.h>

#define FB_CHUNK_DEVICE_INTS (320u << 20)     /* channel-samples per pass: device-resident API, stream length known */
#define FB_CHUNK_HOST_INTS (80u << 20)        /* host streaming paths (pipelined lanes), or length unknown */

/* frame-header codes and encoding parameters of a context's public fields (encode.c:400-434) */
void fb_config_from_context(const FlakeContext *s, FbConfig *g);
/* verbatim bound of a full block (encode.c:446-450) */
int fb_verbatim_bound(const FbConfig *g);
/* bytes per sample of a FLAKE_B200_PCM_* container */
size_t fb_pcm_container_bytes(int fmt);
/* int32 -> digest layout (md5.c:281-320); returns 1 when the packing is lossless */
int fb_pack_s32(const int32_t *src, size_t count, int bytes, uint8_t *dst);
double fb_now_ms(void);
/* blocks per engine pass: a multiple of the device's SM count near target_ints channel-samples */
int fb_chunk_blocks_for(int device, int block_size, int channels, uint64_t stream_samples, uint64_t target_ints);

#endif
