/*
 * engine.h -- internal C ABI between the C host layer (flake_host.c) and the
 * CUDA engine (engine.cu).  Plain pointers and sizes only.
 *
 * Vocabulary: a *block* is params.block_size inter-channel samples handed to
 * flake_encode_frame (libflake/encode.c:979); a block becomes one *frame*, or up
 * to 8 frames under variable block size (libflake/vbs.c:85); a frame holds one
 * *subframe* per channel.  A *chunk* is the run of blocks one engine pass
 * encodes at once.
 */
#ifndef FLAKE_B200_ENGINE_H
#define FLAKE_B200_ENGINE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* PCM layouts the ingest kernel reads (always channel-interleaved) */
enum {
    FB_PCM_S32   = 0,   /* int32, sign-extended: what flake_encode_frame takes */
    FB_PCM_S16LE = 1,   /* packed little-endian 16-bit (WAV data chunk)        */
    FB_PCM_S24LE = 2,   /* packed little-endian 24-bit, 3 bytes per sample     */
    FB_PCM_S8    = 3    /* signed 8-bit                                        */
};

typedef struct FbConfig {
    int channels, bps, block_size;
    int sr_code0, sr_code1, bps_code;        /* encode.c:399-434 */
    int order_method, stereo_method, prediction_type;
    int min_order, max_order, min_porder, max_porder;
    int variable_block_size, allow_vbs;
} FbConfig;

/* one frame of the chunk */
typedef struct FbFrame {
    uint32_t start;     /* first inter-channel sample, relative to the chunk */
    uint32_t n;         /* block size of this frame                          */
    uint32_t number;    /* value coded in the header (frame or sample index) */
    uint32_t slot;      /* byte offset of the staging slot                   */
} FbFrame;

/* per-subframe decisions (what FlacSubframe keeps, encode.h:52-63) */
typedef struct FbSub {
    int32_t type;       /* 0 constant, 1 verbatim, 8 fixed, 32 lpc */
    int32_t order;
    int32_t obits;
    int32_t wasted;
    int32_t shift;
    int32_t method;     /* 0 RICE, 1 RICE2 */
    int32_t porder;
    int32_t est_order;  /* lpc_calc_coefs return value */
    uint32_t est_bits;  /* encode_residual return value */
    int32_t first;      /* samples[0] after decorrelation and wasted shift */
    uint32_t maxabs;    /* max |sample| of the plane */
    int32_t is_const;
    int32_t coefs[32];
    uint8_t params[256];
} FbSub;

typedef struct FbSummary {
    uint32_t nframes;
    uint32_t max_frame_bytes;
    uint64_t total_bytes;
    uint32_t verbatim_frames;   /* frames that took the size fallback */
    uint32_t min_frame_inv;     /* ~(bytes of the smallest frame); 0 = no frame (the summary starts zeroed) */
} FbSummary;

/*
 * One node of the log-search plan (optimize.c:241-261).  The candidate orders the search costs
 * next depend only on (step, orders costed so far, best order so far), never on the samples, so
 * the whole decision tree is tabulated once per engine on the host (engine.cu,
 * fb_build_log_plan) and k_search walks it: cost `cnt` candidates (`ord`: 8 bits each, the
 * orders of `nsteps` consecutive steps that do not depend on each other's outcome), replay the
 * steps on the totals, continue at child[0] when the best order did not change or at
 * child[1 + s] when candidate s became the best.  0xffff ends the search.
 */
#define FB_PLAN_CHILDREN 5
#define FB_PLAN_END 0xffffu
typedef struct FbPlanNode {
    uint32_t ord;
    uint16_t cnt, nsteps;
    uint16_t child[FB_PLAN_CHILDREN];
    uint16_t start_order;       /* node 0 only: the 0-based order the search starts from */
} FbPlanNode;

typedef struct FbEngine FbEngine;

/* device < 0: current device.  max_blocks: chunk capacity in blocks. */
FbEngine *fb_engine_create(const FbConfig *cfg, int device, uint32_t max_blocks,
                           char *err, size_t errlen);
void fb_engine_destroy(FbEngine *e);
uint32_t fb_engine_max_blocks(const FbEngine *e);
uint32_t fb_engine_max_frames(const FbEngine *e);
/* upper bound of the encoded bytes of a full chunk */
uint64_t fb_engine_out_capacity(const FbEngine *e);

/*
 * Enqueue the encode of one chunk that is already in device memory.
 *   d_pcm      device pointer, `fmt` layout, nsamples inter-channel samples
 *   first_number  header number of the chunk's first frame (frame index, or
 *              sample index when allow_vbs)
 *   d_out      device buffer (>= fb_engine_out_capacity) for the compacted frames
 *   d_frame_len / d_frame_bs  device arrays (>= max_frames) or NULL
 *   d_summary  device FbSummary
 *   stream     cudaStream_t (as void*), NULL = engine's own stream
 * Asynchronous; returns 0 or a negative error.
 */
int fb_engine_encode_device(FbEngine *e, const void *d_pcm, int fmt, uint64_t nsamples,
                            uint32_t first_number, void *d_out, uint32_t *d_frame_len,
                            uint32_t *d_frame_bs, FbSummary *d_summary, void *stream);

/* debugging / stage-level parity: copy the per-subframe decisions of the last
 * chunk to the host (synchronises the stream). */
int fb_engine_read_subframes(FbEngine *e, FbSub *host, uint32_t max, void *stream);

/* per-stage CUDA-event timing (events recorded on the launching stream between
 * the kernels of a pass).  Stages: 0 frame table (+VBS split), 1 prep, 2 lpc,
 * 3 search, 4 pack (bit packing, offsets by look-back, frames written in place). */
#define FB_NUM_STAGES 5
int  fb_engine_set_timing(FbEngine *e, int on);
int  fb_engine_collect_timing(FbEngine *e, double *ms, uint64_t *launches);
void fb_engine_reset_timing(FbEngine *e);

/* kernels launched by this engine since creation (bench `gpu_launches`) */
uint64_t fb_engine_launch_count(const FbEngine *e);
const char *fb_engine_last_error(const FbEngine *e);

/* thin wrappers so that the C host never includes cuda_runtime.h */
void *fb_cuda_malloc(size_t n);
void  fb_cuda_free(void *p);
void *fb_cuda_malloc_host(size_t n);
void  fb_cuda_free_host(void *p);
void *fb_cuda_stream_create(void);
void  fb_cuda_stream_destroy(void *s);
int   fb_cuda_stream_sync(void *s);
int   fb_cuda_h2d(void *d, const void *h, size_t n, void *stream);
int   fb_cuda_d2h(void *h, const void *d, size_t n, void *stream);
void *fb_cuda_event_create(void);
void *fb_cuda_event_create_blocking(void);    /* waits yield the CPU; no timing */
int   fb_cuda_host_is_pinned(const void *p);  /* page-locked host memory? */
void  fb_cuda_event_destroy(void *ev);
int   fb_cuda_event_record(void *ev, void *stream);
int   fb_cuda_event_sync(void *ev);
int   fb_cuda_stream_wait_event(void *stream, void *ev);
float fb_cuda_event_elapsed_ms(void *a, void *b);
int   fb_cuda_device_count(void);
int   fb_cuda_sm_count(int device);            /* < 0: current device */
int   fb_cuda_current_device(void);           /* -1 on error */
int   fb_cuda_set_device(int dev);

#ifdef __cplusplus
}
#endif
#endif
