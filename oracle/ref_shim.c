/*
 * ref_shim.c -- TEST INFRASTRUCTURE.  A driver around the UNMODIFIED reference
 * libflake, compiled together with the reference's own sources into
 * oracle/_ref/libflake_ref.so (oracle/Makefile, target `ref`).  Nothing here is
 * linked into, or called by, the shipped library.
 *
 * Why it exists: the reference API is one synchronous flake_encode_frame call per
 * block on one thread (libflake/encode.c:979-1008).  Checking the CUDA path
 * byte-for-byte at BASELINE.json's full sizes (a 1-hour stream, a 10-minute -12
 * stream) would take one core minutes, and timing the reference through a Python
 * loop adds interpreter overhead per block.  This shim
 *   - runs the flake/flake.c:612-663 loop in C (refshim_encode_range), and
 *   - runs several such loops on threads over contiguous BLOCK RANGES of one
 *     stream (refshim_encode_parallel), each range on its own reference context
 *     whose frame counter is set to the value the serial encoder has at that
 *     block (encode.c:969-975: +1 per frame, or +blocksize when allow_vbs) --
 *     the only cross-frame state that reaches the frame bytes.
 * The frame counter lives in the reference's private FlacEncodeContext
 * (libflake/encode.h:77-93), which this file sees by including that header where
 * it lies; no reference source is copied.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "flake.h"
#include "encode.h"     /* /root/reference/libflake/encode.h: FlacEncodeContext */

#define SHIM_API __attribute__((visibility("default")))

typedef struct RefJob {
    /* stream description */
    int channels, sample_rate, bps;
    FlakeEncodeParams params;
    const int32_t *pcm;             /* whole stream, interleaved */
    uint64_t nsamples;              /* whole stream */
    /* this range */
    uint64_t b0, b1;                /* blocks [b0, b1) */
    uint8_t *out; uint64_t cap;     /* frames of the range, back to back */
    uint32_t *call_len; uint32_t call_cap;   /* bytes per flake_encode_frame call (may hold several VBS frames) */
    /* results */
    int64_t bytes;                  /* < 0: error */
    uint32_t ncalls;
    uint32_t max_frame_size;
} RefJob;

static void run_job(RefJob *j)
{
    FlakeContext s;
    memset(&s, 0, sizeof s);
    s.channels = j->channels;
    s.sample_rate = j->sample_rate;
    s.bits_per_sample = j->bps;
    s.samples = (unsigned int)j->nsamples;
    s.params = j->params;
    j->bytes = -1; j->ncalls = 0; j->max_frame_size = 0;
    if (flake_encode_init(&s) < 0) { flake_encode_close(&s); return; }
    FlacEncodeContext *ctx = (FlacEncodeContext *)s.private_ctx;
    const uint64_t B = (uint64_t)j->params.block_size;
    /* the counter the serial encoder would hold at block b0 */
    ctx->frame_count = (uint32_t)(j->params.allow_vbs ? j->b0 * B : j->b0);
    const uint8_t *frame = (const uint8_t *)flake_get_buffer(&s);
    uint64_t pos = 0;
    int ok = 1;
    for (uint64_t b = j->b0; b < j->b1 && ok; b++) {
        const uint64_t first = b * B;
        if (first >= j->nsamples) break;
        const int n = (int)((j->nsamples - first < B) ? (j->nsamples - first) : B);
        const int fs = flake_encode_frame(&s, (const int *)(j->pcm + first * (uint64_t)j->channels), n);
        if (fs < 0 || pos + (uint64_t)fs > j->cap) { ok = 0; break; }
        memcpy(j->out + pos, frame, (size_t)fs);
        if (j->call_len && j->ncalls < j->call_cap) j->call_len[j->ncalls] = (uint32_t)fs;
        j->ncalls++;
        pos += (uint64_t)fs;
    }
    FlakeStreaminfo si;
    if (ok && !flake_get_streaminfo(&s, &si)) j->max_frame_size = si.max_frame_size;
    flake_encode_close(&s);
    if (ok) j->bytes = (int64_t)pos;
}

static void *job_thread(void *arg) { run_job((RefJob *)arg); return NULL; }

/* The flake/flake.c loop over blocks [b0, b1) of the stream, in C, on the calling thread. */
SHIM_API int64_t refshim_encode_range(int channels, int sample_rate, int bps, const FlakeEncodeParams *params,
                                      const int32_t *pcm, uint64_t nsamples, uint64_t b0, uint64_t b1,
                                      uint8_t *out, uint64_t cap, uint32_t *call_len, uint32_t call_cap,
                                      uint32_t *ncalls, uint32_t *max_frame_size)
{
    RefJob j;
    memset(&j, 0, sizeof j);
    j.channels = channels; j.sample_rate = sample_rate; j.bps = bps; j.params = *params;
    j.pcm = pcm; j.nsamples = nsamples; j.b0 = b0; j.b1 = b1;
    j.out = out; j.cap = cap; j.call_len = call_len; j.call_cap = call_cap;
    run_job(&j);
    if (ncalls) *ncalls = j.ncalls;
    if (max_frame_size) *max_frame_size = j.max_frame_size;
    return j.bytes;
}

/*
 * The whole stream on `threads` threads, one contiguous block range and one reference
 * context each; the ranges' frames are concatenated in order, so the result equals the
 * serial encoder's frame bytes.  `out` must hold cap bytes >= the verbatim bound of the
 * stream; call_len (optional) gets one entry per block.
 */
SHIM_API int64_t refshim_encode_parallel(int channels, int sample_rate, int bps, const FlakeEncodeParams *params,
                                         const int32_t *pcm, uint64_t nsamples, int threads,
                                         uint8_t *out, uint64_t cap, uint32_t *call_len, uint32_t call_cap,
                                         uint32_t *ncalls, uint32_t *max_frame_size)
{
    const uint64_t B = (uint64_t)params->block_size;
    const uint64_t nblocks = (nsamples + B - 1) / B;
    if (threads < 1) threads = 1;
    if ((uint64_t)threads > nblocks) threads = (int)(nblocks ? nblocks : 1);
    RefJob *jobs = (RefJob *)calloc((size_t)threads, sizeof *jobs);
    pthread_t *th = (pthread_t *)calloc((size_t)threads, sizeof *th);
    uint8_t **tmp = (uint8_t **)calloc((size_t)threads, sizeof *tmp);
    if (!jobs || !th || !tmp) { free(jobs); free(th); free(tmp); return -1; }
    const uint64_t bytes_per_block = 64u + ((B * (uint64_t)(channels * bps + 1) + 7u) >> 3) +
                                     (params->variable_block_size ? 8u * 32u : 0u);
    int64_t total = 0;
    int bad = 0;
    for (int t = 0; t < threads; t++) {
        RefJob *j = &jobs[t];
        j->channels = channels; j->sample_rate = sample_rate; j->bps = bps; j->params = *params;
        j->pcm = pcm; j->nsamples = nsamples;
        j->b0 = nblocks * (uint64_t)t / (uint64_t)threads;
        j->b1 = nblocks * (uint64_t)(t + 1) / (uint64_t)threads;
        j->cap = (j->b1 - j->b0) * bytes_per_block * 3u / 2u + 4096u;
        tmp[t] = (uint8_t *)malloc((size_t)j->cap);
        j->out = tmp[t];
        if (call_len && j->b0 < call_cap) { j->call_len = call_len + j->b0; j->call_cap = call_cap - (uint32_t)j->b0; }
        if (!tmp[t]) bad = 1;
    }
    if (!bad) {
        for (int t = 0; t < threads; t++)
            if (pthread_create(&th[t], NULL, job_thread, &jobs[t])) { run_job(&jobs[t]); th[t] = 0; }
        for (int t = 0; t < threads; t++)
            if (th[t]) pthread_join(th[t], NULL);
    }
    uint32_t calls = 0, mx = 0;
    for (int t = 0; t < threads && !bad; t++) {
        if (jobs[t].bytes < 0 || (uint64_t)total + (uint64_t)jobs[t].bytes > cap) { bad = 1; break; }
        memcpy(out + total, tmp[t], (size_t)jobs[t].bytes);
        total += jobs[t].bytes;
        calls += jobs[t].ncalls;
        if (jobs[t].max_frame_size > mx) mx = jobs[t].max_frame_size;
    }
    for (int t = 0; t < threads; t++) free(tmp[t]);
    free(jobs); free(th); free(tmp);
    if (ncalls) *ncalls = calls;
    if (max_frame_size) *max_frame_size = mx;
    return bad ? -1 : total;
}
