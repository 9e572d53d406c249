/*
 * flake_corpus.c -- flake_b200_encode_corpus: many streams (or one long stream) over the GPUs
 * of one box, from ONE process, with library-owned threads.
 *
 * What it replaces: the reference encodes a file with one thread in one loop,
 *     read block -> flake_encode_frame -> fwrite            (flake/flake.c:612-678)
 * and a corpus by running that loop once per file.  Frames depend only on their own samples,
 * the stream parameters and their header number (encode.c:726-764, 969-975), so here the unit
 * of GPU work is a CHUNK -- a contiguous range of blocks of one stream -- and any GPU may
 * encode any chunk:
 *
 *   gpu workers   (threads_per_device per device)  take chunks in corpus order; per chunk:
 *                 DMA the PCM from the caller's buffer (no staging copy when it is page-locked),
 *                 one engine pass (engine.cu), read back the 24-byte summary, then place the
 *                 frames: the chunk's byte offset inside its stream is the prefix sum of the
 *                 sizes of the chunks before it, resolved on the host in chunk order, and the
 *                 frames are DMAed straight to out + offset.  Upload of chunk k+1, kernels of
 *                 chunk k and download of chunk k-1 overlap (three streams, two lanes).
 *   md5 workers   the MD5 of a stream's PCM (md5.c:281-320, encode.c:1006) is a serial chain
 *                 per stream but independent of the encoding, and chains of different streams
 *                 are independent of each other: each worker advances up to 32 streams at once
 *                 in the SIMD lanes of its core (md5_mb.c), refilling a lane when its stream
 *                 ends.  With no more streams than cores every stream gets a scalar thread.
 *
 * No exchange between GPUs: the only cross-chunk dependency is the host-side prefix sum
 * (SURVEY.md 8e).  There is no CPU encoding path; without a CUDA device the call fails.
 */
#define _GNU_SOURCE
#define FLAKE_BUILD_LIBRARY 1
#include "flake.h"
#include "flake_b200.h"
#include "flake_host_int.h"
#include "md5.h"
#include "md5_mb.h"

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#define CO_NONE 0xffffffffu
#define CO_MD5_SLICE (64u * 1024u)        /* bytes per lane and round of an MD5 worker */
#define CO_MAX_WORKERS 64

typedef struct CoStream {
    uint32_t nchunks, first_unit;
    uint32_t posted;            /* chunks [0, posted) are in the prefix sums below */
    uint64_t byte_prefix;
    uint32_t frame_prefix;
    uint32_t max_frame, min_frame, verbatim;
    int in_pinned, out_pinned;
    int err;
} CoStream;

struct CoWorker;
typedef struct FlakeB200Corpus Corpus;
struct FlakeB200Corpus {
    /* for the life of the handle */
    FlakeContext proto;         /* public fields only */
    FbConfig cfg;
    int fmt, digest_bytes, md5_direct;
    size_t in_bps;              /* input bytes per inter-channel sample */
    uint64_t chunk;             /* samples per chunk */
    int chunk_blocks;
    int devs[CO_MAX_WORKERS], ndev;
    int per_dev, md5_threads;
    struct CoWorker *gw;        /* ndev * per_dev; GPU resources are created by the worker thread on first use */
    int ngw;
    pthread_mutex_t mu;
    pthread_cond_t cv;
    /* of the call in progress */
    FlakeB200CorpusStream *streams;
    uint32_t nstreams;
    CoStream *st;
    uint32_t *unit_stream;
    uint32_t nunits;
    uint32_t next_unit, next_md5;
    int abort;
    char err[256];
};

typedef struct CoLane {
    void *d_in, *d_out;
    uint32_t *d_flen, *d_fbs, *h_flen, *h_fbs;
    FbSummary *d_sum, *h_sum;
    void *h_in, *h_out;         /* page-locked staging, only for pageable caller memory */
    void *ev_up, *ev_k, *ev_down;
    int state;                  /* 0 empty, 1 kernels in flight, 2 frames on their way to the host */
    int failed, staged_out;
    uint32_t stream, chunk;
    uint64_t ns, byte_off;
    uint32_t frame_off;
    FbSummary sum;
} CoLane;

typedef struct CoWorker {
    Corpus *co;
    int device, started, ready;
    pthread_t th;
    FbEngine *eng;
    void *st, *st_up, *st_down;
    CoLane lane[2];
    uint64_t units, samples, h2d, d2h, launches;
    double busy_ms;
} CoWorker;

typedef struct MdWorker {
    Corpus *co;
    int lanes, started;
    pthread_t th;
    double ms;
} MdWorker;

static int env_int(const char *name, int dflt)
{
    const char *v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

static void co_fail(Corpus *co, const char *msg)
{
    pthread_mutex_lock(&co->mu);
    if (!co->err[0]) snprintf(co->err, sizeof co->err, "%s", msg);
    co->abort = 1;
    pthread_mutex_unlock(&co->mu);
}

/* ------------------------------------------------------------------ */
/* gpu workers                                                          */
/* ------------------------------------------------------------------ */
static void lane_release(CoLane *l)
{
    fb_cuda_free(l->d_in); fb_cuda_free(l->d_out); fb_cuda_free(l->d_flen); fb_cuda_free(l->d_fbs);
    fb_cuda_free(l->d_sum);
    fb_cuda_free_host(l->h_flen); fb_cuda_free_host(l->h_fbs); fb_cuda_free_host(l->h_sum);
    fb_cuda_free_host(l->h_in); fb_cuda_free_host(l->h_out);
    fb_cuda_event_destroy(l->ev_up); fb_cuda_event_destroy(l->ev_k); fb_cuda_event_destroy(l->ev_down);
    memset(l, 0, sizeof *l);
}

static void worker_release(CoWorker *w)
{
    if (w->st) fb_cuda_stream_sync(w->st);
    if (w->st_up) fb_cuda_stream_sync(w->st_up);
    if (w->st_down) fb_cuda_stream_sync(w->st_down);
    lane_release(&w->lane[0]); lane_release(&w->lane[1]);
    if (w->eng) fb_engine_destroy(w->eng);
    fb_cuda_stream_destroy(w->st); fb_cuda_stream_destroy(w->st_up); fb_cuda_stream_destroy(w->st_down);
    w->eng = NULL; w->st = w->st_up = w->st_down = NULL;
    w->ready = 0;
}

static int worker_setup(CoWorker *w)
{
    Corpus *co = w->co;
    char err[200] = "";
    if (w->ready) return 0;
    w->eng = fb_engine_create(&co->cfg, w->device, (uint32_t)co->chunk_blocks, err, sizeof err);
    if (!w->eng) {
        char msg[256];
        snprintf(msg, sizeof msg, "device %d: cannot create the CUDA engine: %s", w->device, err);
        co_fail(co, msg);
        return -1;
    }
    w->st = fb_cuda_stream_create(); w->st_up = fb_cuda_stream_create(); w->st_down = fb_cuda_stream_create();
    const uint32_t mf = fb_engine_max_frames(w->eng);
    int ok = w->st && w->st_up && w->st_down;
    for (int i = 0; i < 2 && ok; i++) {
        CoLane *l = &w->lane[i];
        l->d_in = fb_cuda_malloc((size_t)co->chunk * co->in_bps);
        l->d_out = fb_cuda_malloc((size_t)fb_engine_out_capacity(w->eng));
        l->d_flen = (uint32_t *)fb_cuda_malloc(sizeof(uint32_t) * mf);
        l->d_fbs = (uint32_t *)fb_cuda_malloc(sizeof(uint32_t) * mf);
        l->d_sum = (FbSummary *)fb_cuda_malloc(sizeof(FbSummary));
        l->h_flen = (uint32_t *)fb_cuda_malloc_host(sizeof(uint32_t) * mf);
        l->h_fbs = (uint32_t *)fb_cuda_malloc_host(sizeof(uint32_t) * mf);
        l->h_sum = (FbSummary *)fb_cuda_malloc_host(sizeof(FbSummary));
        l->ev_up = fb_cuda_event_create_blocking();
        l->ev_k = fb_cuda_event_create_blocking();
        l->ev_down = fb_cuda_event_create_blocking();
        ok = l->d_in && l->d_out && l->d_flen && l->d_fbs && l->d_sum && l->h_flen && l->h_fbs && l->h_sum &&
             l->ev_up && l->ev_k && l->ev_down;
    }
    if (!ok) {
        char msg[256];
        snprintf(msg, sizeof msg, "device %d: CUDA allocation of the corpus lanes failed", w->device);
        co_fail(co, msg);
        worker_release(w);
        return -1;
    }
    w->ready = 1;
    return 0;
}

static uint32_t take_unit(Corpus *co)
{
    pthread_mutex_lock(&co->mu);
    const uint32_t u = (co->abort || co->next_unit >= co->nunits) ? CO_NONE : co->next_unit++;
    pthread_mutex_unlock(&co->mu);
    return u;
}

/* upload + engine pass + summary of unit u, all asynchronous */
static void lane_issue(CoWorker *w, CoLane *l, uint32_t u)
{
    Corpus *co = w->co;
    const uint32_t si = co->unit_stream[u];
    const FlakeB200CorpusStream *S = &co->streams[si];
    l->stream = si;
    l->chunk = u - co->st[si].first_unit;
    const uint64_t off = (uint64_t)l->chunk * co->chunk;
    l->ns = S->nsamples - off < co->chunk ? S->nsamples - off : co->chunk;
    l->state = 1; l->failed = 0; l->staged_out = 0;
    const size_t nbytes = (size_t)l->ns * co->in_bps;
    const uint8_t *src = (const uint8_t *)S->pcm + off * co->in_bps;
    if (!co->st[si].in_pinned) {
        /* pageable caller memory: through a page-locked buffer so that the copy stays asynchronous */
        if (!l->h_in) l->h_in = fb_cuda_malloc_host((size_t)co->chunk * co->in_bps);
        if (!l->h_in) { l->failed = 1; return; }
        memcpy(l->h_in, src, nbytes);
        src = (const uint8_t *)l->h_in;
    }
    const uint32_t first_number = co->cfg.allow_vbs ? (uint32_t)off : (uint32_t)(off / (uint64_t)co->cfg.block_size);
    const uint64_t before = fb_engine_launch_count(w->eng);
    if (fb_cuda_h2d(l->d_in, src, nbytes, w->st_up) || fb_cuda_event_record(l->ev_up, w->st_up) ||
        fb_cuda_stream_wait_event(w->st, l->ev_up) ||
        fb_engine_encode_device(w->eng, l->d_in, co->fmt, l->ns, first_number, l->d_out, l->d_flen, l->d_fbs,
                                l->d_sum, w->st) ||
        fb_cuda_d2h(l->h_sum, l->d_sum, sizeof(FbSummary), w->st) || fb_cuda_event_record(l->ev_k, w->st))
        l->failed = 1;
    w->launches += fb_engine_launch_count(w->eng) - before;
    w->h2d += nbytes;
    w->units++;
    w->samples += l->ns;
}

/* the pass is done: enter the chunk in its stream's prefix sums (in chunk order), then start the
 * download of its frames to their final place */
static void lane_collect(CoWorker *w, CoLane *l)
{
    Corpus *co = w->co;
    if (l->state != 1) return;
    CoStream *T = &co->st[l->stream];
    FlakeB200CorpusStream *S = &co->streams[l->stream];
    int bad = l->failed;
    if (!bad && fb_cuda_event_sync(l->ev_k)) bad = 1;
    FbSummary sm;
    memset(&sm, 0, sizeof sm);
    if (!bad) sm = *l->h_sum;
    const int want_len = S->frame_len != NULL, want_bs = S->frame_bs != NULL;

    pthread_mutex_lock(&co->mu);
    while (T->posted != l->chunk) pthread_cond_wait(&co->cv, &co->mu);
    int err = T->err;
    if (bad) err = -3;
    if (!err && T->byte_prefix + sm.total_bytes > S->out_cap) err = -2;
    if (!err && (want_len || want_bs) && (uint64_t)T->frame_prefix + sm.nframes > S->frame_cap) err = -2;
    if (err && !T->err) T->err = err;
    l->byte_off = T->byte_prefix; l->frame_off = T->frame_prefix;
    if (!T->err) {
        T->byte_prefix += sm.total_bytes;
        T->frame_prefix += sm.nframes;
        if (sm.max_frame_bytes > T->max_frame) T->max_frame = sm.max_frame_bytes;
        if (sm.nframes && ~sm.min_frame_inv < T->min_frame) T->min_frame = ~sm.min_frame_inv;
        T->verbatim += sm.verbatim_frames;
    }
    T->posted++;
    if (bad && !co->err[0]) {
        snprintf(co->err, sizeof co->err, "device %d: CUDA failure while encoding (%s)", w->device,
                 w->eng ? fb_engine_last_error(w->eng) : "no engine");
        co->abort = 1;
    }
    pthread_cond_broadcast(&co->cv);
    const int stream_err = T->err;
    pthread_mutex_unlock(&co->mu);

    l->sum = sm;
    if (stream_err) { l->state = 0; return; }
    void *dst = S->out + l->byte_off;
    if (!T->out_pinned) {
        if (!l->h_out) l->h_out = fb_cuda_malloc_host((size_t)fb_engine_out_capacity(w->eng));
        if (!l->h_out) { co_fail(co, "CUDA allocation of the output staging failed"); l->state = 0; return; }
        dst = l->h_out; l->staged_out = 1;
    }
    int rc = fb_cuda_d2h(dst, l->d_out, (size_t)sm.total_bytes, w->st_down);
    if (!rc && want_len) rc = fb_cuda_d2h(l->h_flen, l->d_flen, sizeof(uint32_t) * sm.nframes, w->st_down);
    if (!rc && want_bs) rc = fb_cuda_d2h(l->h_fbs, l->d_fbs, sizeof(uint32_t) * sm.nframes, w->st_down);
    if (!rc) rc = fb_cuda_event_record(l->ev_down, w->st_down);
    if (rc) { co_fail(co, "CUDA failure while copying frames"); l->state = 0; return; }
    w->d2h += sm.total_bytes + (uint64_t)sm.nframes * 4u * (uint64_t)(want_len + want_bs) + sizeof(FbSummary);
    l->state = 2;
}

/* the frames are on the host: last copies out of the staging buffers; the lane is free again */
static void lane_finish(CoWorker *w, CoLane *l)
{
    Corpus *co = w->co;
    if (l->state != 2) { l->state = 0; return; }
    FlakeB200CorpusStream *S = &co->streams[l->stream];
    if (fb_cuda_event_sync(l->ev_down)) {
        co_fail(co, "CUDA failure while copying frames");
        pthread_mutex_lock(&co->mu);
        if (!co->st[l->stream].err) co->st[l->stream].err = -3;
        pthread_mutex_unlock(&co->mu);
    } else {
        if (l->staged_out) memcpy(S->out + l->byte_off, l->h_out, (size_t)l->sum.total_bytes);
        if (S->frame_len) memcpy(S->frame_len + l->frame_off, l->h_flen, sizeof(uint32_t) * l->sum.nframes);
        if (S->frame_bs) memcpy(S->frame_bs + l->frame_off, l->h_fbs, sizeof(uint32_t) * l->sum.nframes);
    }
    l->state = 0;
}

static void *gpu_worker(void *arg)
{
    CoWorker *w = (CoWorker *)arg;
    Corpus *co = w->co;
    const double t0 = fb_now_ms();
    if (fb_cuda_set_device(w->device)) {
        char msg[96];
        snprintf(msg, sizeof msg, "cannot select CUDA device %d", w->device);
        co_fail(co, msg);
        return NULL;
    }
    if (worker_setup(w) == 0) {
        for (int k = 0;; k++) {
            const uint32_t u = take_unit(co);
            CoLane *l = &w->lane[k & 1], *p = &w->lane[(k + 1) & 1];
            lane_finish(w, l);                      /* unit k-2 leaves the lane ... */
            if (u != CO_NONE) lane_issue(w, l, u);  /* ... unit k enters it */
            lane_collect(w, p);                     /* unit k-1: sizes known, frames start to come back */
            if (u == CO_NONE) { lane_finish(w, p); break; }
        }
    }
    w->busy_ms = fb_now_ms() - t0;
    return NULL;
}

/* ------------------------------------------------------------------ */
/* md5 workers                                                          */
/* ------------------------------------------------------------------ */
static uint32_t take_md5(Corpus *co)
{
    pthread_mutex_lock(&co->mu);
    const uint32_t i = co->next_md5 >= co->nstreams ? CO_NONE : co->next_md5++;
    pthread_mutex_unlock(&co->mu);
    return i;
}

/* digest-layout bytes [pos, pos + len) of stream i: the caller's buffer itself, or packed from
 * int32 into `scratch` */
static const uint8_t *md5_fetch(const Corpus *co, uint32_t i, uint64_t pos, size_t len, uint8_t *scratch)
{
    const uint8_t *pcm = (const uint8_t *)co->streams[i].pcm;
    if (co->md5_direct) return pcm + pos;
    fb_pack_s32((const int32_t *)pcm + pos / (uint64_t)co->digest_bytes, len / (size_t)co->digest_bytes,
                co->digest_bytes, scratch);
    return scratch;
}

static void *md5_worker(void *arg)
{
    MdWorker *m = (MdWorker *)arg;
    Corpus *co = m->co;
    const int L = m->lanes;
    const double t0 = fb_now_ms();
    uint32_t sidx[FB_MD5_MB_MAX];
    uint64_t pos[FB_MD5_MB_MAX], total[FB_MD5_MB_MAX];
    FbMd5 ctx[FB_MD5_MB_MAX];
    uint8_t *scratch = co->md5_direct ? NULL : (uint8_t *)malloc((size_t)L * CO_MD5_SLICE);
    int drained = 0;
    /* lanes advance in steps that are whole MD5 blocks AND whole samples of the digest layout */
    const uint64_t step = co->md5_direct ? 64u : 64u * (uint64_t)co->digest_bytes;
    for (int l = 0; l < L; l++) sidx[l] = CO_NONE;
    if (!co->md5_direct && !scratch) { co_fail(co, "out of memory (MD5 scratch)"); return NULL; }
    for (;;) {
        int nact = 0;
        uint64_t minrem = ~0ull;
        for (int l = 0; l < L; l++) {
            if (sidx[l] == CO_NONE && !drained) {
                const uint32_t i = take_md5(co);
                if (i == CO_NONE) drained = 1;
                else {
                    sidx[l] = i; pos[l] = 0;
                    total[l] = co->streams[i].nsamples * (uint64_t)co->cfg.channels * (uint64_t)co->digest_bytes;
                    fb_md5_init(&ctx[l]);
                }
            }
            if (sidx[l] == CO_NONE) continue;
            nact++;
            if (total[l] - pos[l] < minrem) minrem = total[l] - pos[l];
        }
        if (!nact) break;
        if (minrem < step) {
            /* streams with less than a step left: the tail, then the digest; their lanes refill */
            for (int l = 0; l < L; l++) {
                if (sidx[l] == CO_NONE || total[l] - pos[l] >= step) continue;
                const size_t rem = (size_t)(total[l] - pos[l]);
                if (rem) fb_md5_update(&ctx[l], md5_fetch(co, sidx[l], pos[l], rem, scratch), rem);
                fb_md5_final(&ctx[l], co->streams[sidx[l]].md5sum);
                sidx[l] = CO_NONE;
            }
            continue;
        }
        size_t len = (size_t)(minrem < CO_MD5_SLICE ? minrem : CO_MD5_SLICE);
        len -= len % (size_t)step;
        FbMd5 *cp[FB_MD5_MB_MAX];
        const uint8_t *dp[FB_MD5_MB_MAX];
        size_t ln[FB_MD5_MB_MAX];
        int n = 0;
        for (int l = 0; l < L; l++) {
            if (sidx[l] == CO_NONE) continue;
            cp[n] = &ctx[l];
            dp[n] = md5_fetch(co, sidx[l], pos[l], len, scratch ? scratch + (size_t)l * CO_MD5_SLICE : NULL);
            ln[n] = len;
            pos[l] += len;
            n++;
        }
        fb_md5_mb_update(cp, dp, ln, n);
    }
    free(scratch);
    m->ms = fb_now_ms() - t0;
    return NULL;
}

/* ------------------------------------------------------------------ */
/* entry point                                                          */
/* ------------------------------------------------------------------ */
FlakeB200Corpus *flake_b200_corpus_open(const FlakeContext *proto, int pcm_format,
                                        const int *devices, int ndevices, const FlakeB200CorpusOptions *opt)
{
    if (!proto || flake_validate_params(proto) < 0) return NULL;
    if (pcm_format < FLAKE_B200_PCM_S32 || pcm_format > FLAKE_B200_PCM_S8) return NULL;
    const int digest_bytes = (proto->bits_per_sample + 7) >> 3;
    if (pcm_format != FLAKE_B200_PCM_S32 && fb_pcm_container_bytes(pcm_format) != (size_t)digest_bytes) return NULL;
    const int ndev_all = fb_cuda_device_count();
    if (ndev_all <= 0) {
        fprintf(stderr, "flake_b200: no CUDA device (there is no CPU encoding path)\n");
        return NULL;
    }
    Corpus *co = (Corpus *)calloc(1, sizeof *co);
    if (!co) return NULL;
    if (!devices || ndevices <= 0) {
        for (int d = 0; d < ndev_all && co->ndev < FLAKE_B200_MAX_DEVICES; d++) co->devs[co->ndev++] = d;
    } else {
        for (int d = 0; d < ndevices && co->ndev < FLAKE_B200_MAX_DEVICES; d++) {
            if (devices[d] < 0 || devices[d] >= ndev_all) { free(co); return NULL; }
            co->devs[co->ndev++] = devices[d];
        }
    }
    co->proto = *proto;
    co->proto.private_ctx = NULL; co->proto.header = NULL;
    fb_config_from_context(proto, &co->cfg);
    co->fmt = pcm_format;
    co->digest_bytes = digest_bytes;
    co->md5_direct = pcm_format != FLAKE_B200_PCM_S32;
    co->in_bps = fb_pcm_container_bytes(pcm_format) * (size_t)proto->channels;
    /* chunk size: a whole number of SM waves (fb_chunk_blocks_for), no longer than the longest
     * stream the caller announces (proto->samples, 0 = unknown) */
    co->chunk_blocks = opt && opt->chunk_blocks > 0 ? opt->chunk_blocks
                     : fb_chunk_blocks_for(co->devs[0], co->cfg.block_size, co->cfg.channels, proto->samples, FB_CHUNK_HOST_INTS);
    co->chunk = (uint64_t)co->chunk_blocks * (uint64_t)co->cfg.block_size;
    co->per_dev = opt && opt->threads_per_device > 0 ? opt->threads_per_device : env_int("FLAKE_B200_CORPUS_THREADS_PER_DEVICE", 2);
    if (co->per_dev > 8) co->per_dev = 8;
    long ncpu = sysconf(_SC_NPROCESSORS_ONLN);
    if (ncpu < 1) ncpu = 1;
    co->ngw = co->ndev * co->per_dev;
    if (co->ngw > CO_MAX_WORKERS) co->ngw = CO_MAX_WORKERS;
    /* cores the MD5 workers may take: all but one per GPU worker */
    const int spare = (int)ncpu - co->ngw > 1 ? (int)ncpu - co->ngw : 1;
    co->md5_threads = opt && opt->md5_threads > 0 ? opt->md5_threads : env_int("FLAKE_B200_CORPUS_MD5_THREADS", spare);
    if (co->md5_threads > CO_MAX_WORKERS) co->md5_threads = CO_MAX_WORKERS;
    co->gw = (CoWorker *)calloc((size_t)co->ngw, sizeof(CoWorker));
    if (!co->gw) { free(co); return NULL; }
    /* workers of one device are neighbours in corpus order: worker i -> device i % ndev */
    for (int i = 0; i < co->ngw; i++) { co->gw[i].co = co; co->gw[i].device = co->devs[i % co->ndev]; }
    pthread_mutex_init(&co->mu, NULL);
    pthread_cond_init(&co->cv, NULL);
    return co;
}

static void *release_thread(void *arg)
{
    CoWorker *w = (CoWorker *)arg;
    if (w->ready && fb_cuda_set_device(w->device) == 0) worker_release(w);
    return NULL;
}

void flake_b200_corpus_close(FlakeB200Corpus *co)
{
    if (!co) return;
    /* on a thread of its own so that the caller's current device stays what it was */
    for (int i = 0; i < co->ngw; i++) {
        pthread_t th;
        if (pthread_create(&th, NULL, release_thread, &co->gw[i]) == 0) pthread_join(th, NULL);
        else {
            const int dev = fb_cuda_current_device();
            release_thread(&co->gw[i]);
            if (dev >= 0) fb_cuda_set_device(dev);
        }
    }
    free(co->gw);
    pthread_mutex_destroy(&co->mu);
    pthread_cond_destroy(&co->cv);
    free(co);
}

const char *flake_b200_corpus_error(const FlakeB200Corpus *co) { return co ? co->err : "no corpus handle"; }

int flake_b200_corpus_encode(FlakeB200Corpus *co, FlakeB200CorpusStream *streams, unsigned int nstreams,
                             FlakeB200CorpusStats *stats)
{
    if (stats) memset(stats, 0, sizeof *stats);
    if (!co || (!streams && nstreams)) return -1;
    for (unsigned i = 0; i < nstreams; i++) {
        FlakeB200CorpusStream *S = &streams[i];
        S->bytes = -1; S->nframes = 0; S->max_frame_size = 0; S->min_frame_size = 0; S->verbatim_frames = 0;
        memset(S->md5sum, 0, sizeof S->md5sum);
        if ((!S->pcm && S->nsamples) || (!S->out && S->nsamples) || S->nsamples > 0xffffffffull) return -1;
    }
    const double t0 = fb_now_ms();
    co->streams = streams; co->nstreams = nstreams;
    co->nunits = 0; co->next_unit = 0; co->next_md5 = 0; co->abort = 0; co->err[0] = 0;
    uint64_t total_samples = 0;
    for (unsigned i = 0; i < nstreams; i++) total_samples += streams[i].nsamples;

    int rc = 0;
    co->st = (CoStream *)calloc(nstreams ? nstreams : 1, sizeof(CoStream));
    if (!co->st) rc = -3;
    for (unsigned i = 0; i < nstreams && !rc; i++) {
        co->st[i].nchunks = (uint32_t)((streams[i].nsamples + co->chunk - 1) / co->chunk);
        co->st[i].first_unit = co->nunits;
        co->nunits += co->st[i].nchunks;
        co->st[i].max_frame = (uint32_t)fb_verbatim_bound(&co->cfg);
        co->st[i].min_frame = 0xffffffffu;
        co->st[i].in_pinned = streams[i].pcm ? fb_cuda_host_is_pinned(streams[i].pcm) : 1;
        co->st[i].out_pinned = streams[i].out ? fb_cuda_host_is_pinned(streams[i].out) : 1;
    }
    co->unit_stream = rc ? NULL : (uint32_t *)malloc(sizeof(uint32_t) * (co->nunits ? co->nunits : 1));
    if (!co->unit_stream) rc = -3;
    for (unsigned i = 0; i < nstreams && !rc; i++)
        for (uint32_t j = 0; j < co->st[i].nchunks; j++) co->unit_stream[co->st[i].first_unit + j] = i;

    /* threads of this call */
    int nworkers = co->ngw;
    if ((uint32_t)nworkers > co->nunits) nworkers = (int)co->nunits;
    /* MD5 workers.  A stream hashed alone runs at the scalar rate (~0.8 GB/s); a stream in a SIMD
     * lane at ~0.6 of that, whatever the number of busy lanes up to 16.  So: with no more streams
     * than cores to spare, one scalar thread per stream; otherwise as few threads as keep 16 lanes
     * each (the cores left over drive the GPUs and take the DMA interrupts), and only a corpus of
     * more than 16 streams per core goes to 32 lanes. */
    int avail = co->md5_threads;
    if (avail < 1) avail = 1;
    int nmd5, lanes;
    const int simd = fb_md5_mb_lanes();
    if (nstreams <= (unsigned)avail || simd <= 1) {
        nmd5 = nstreams < (unsigned)avail ? (int)nstreams : avail;
        lanes = 1;
    } else {
        nmd5 = (int)((nstreams + 15u) / 16u);
        if (nmd5 > avail) nmd5 = avail;
        lanes = (int)((nstreams + (unsigned)nmd5 - 1) / (unsigned)nmd5);
        if (lanes > simd) lanes = simd;
        if (lanes > FB_MD5_MB_MAX) lanes = FB_MD5_MB_MAX;
    }
    MdWorker *mw = (MdWorker *)calloc((size_t)(nmd5 ? nmd5 : 1), sizeof(MdWorker));
    if (!mw) rc = -3;
    if (!rc) {
        for (int i = 0; i < co->ngw; i++) { CoWorker *w = &co->gw[i]; w->units = w->samples = w->h2d = w->d2h = w->launches = 0; w->busy_ms = 0; w->started = 0; }
        for (int i = 0; i < nmd5; i++) {
            mw[i].co = co; mw[i].lanes = lanes;
            mw[i].started = pthread_create(&mw[i].th, NULL, md5_worker, &mw[i]) == 0;
            if (!mw[i].started) md5_worker(&mw[i]);           /* no thread: do the share here */
        }
        int any = 0;
        for (int i = 0; i < nworkers; i++) {
            co->gw[i].started = pthread_create(&co->gw[i].th, NULL, gpu_worker, &co->gw[i]) == 0;
            any |= co->gw[i].started;
        }
        if (!any && nworkers > 0) {
            const int dev = fb_cuda_current_device();
            gpu_worker(&co->gw[0]);
            if (dev >= 0) fb_cuda_set_device(dev);
        }
        for (int i = 0; i < nworkers; i++) if (co->gw[i].started) pthread_join(co->gw[i].th, NULL);
        for (int i = 0; i < nmd5; i++) if (mw[i].started) pthread_join(mw[i].th, NULL);
    }

    /* results */
    for (unsigned i = 0; i < nstreams && co->st; i++) {
        FlakeB200CorpusStream *S = &streams[i];
        const CoStream *T = &co->st[i];
        if (rc) { S->bytes = rc; continue; }
        if (T->err) S->bytes = T->err;
        else if (T->posted != T->nchunks) S->bytes = -3;      /* the call was aborted before this stream was done */
        else {
            S->bytes = (long long)T->byte_prefix;
            S->nframes = T->frame_prefix;
            S->max_frame_size = T->max_frame;
            S->min_frame_size = T->min_frame == 0xffffffffu ? 0 : T->min_frame;
            S->verbatim_frames = T->verbatim;
        }
    }
    for (unsigned i = 0; i < nstreams && !rc; i++) if (streams[i].bytes < 0) rc = (int)streams[i].bytes;
    if (co->abort && !rc) rc = -3;
    if (co->err[0]) fprintf(stderr, "flake_b200: corpus: %s\n", co->err);
    if (stats) {
        stats->wall_ms = fb_now_ms() - t0;
        stats->streams = nstreams;
        stats->samples = total_samples;
        stats->chunks = co->nunits;
        stats->chunk_blocks = (unsigned)co->chunk_blocks;
        stats->devices = co->ndev;
        stats->gpu_threads = nworkers;
        stats->md5_threads = nmd5;
        stats->md5_lanes = lanes;
        for (int i = 0; i < nmd5 && mw; i++) if (mw[i].ms > stats->md5_ms) stats->md5_ms = mw[i].ms;
        for (int i = 0; i < co->ngw; i++) {
            const CoWorker *w = &co->gw[i];
            stats->h2d_bytes += w->h2d; stats->d2h_bytes += w->d2h; stats->kernel_launches += w->launches;
            for (int d = 0; d < co->ndev; d++)
                if (co->devs[d] == w->device) {
                    stats->device_samples[d] += w->samples;
                    if (w->busy_ms > stats->device_ms[d]) stats->device_ms[d] = w->busy_ms;
                    break;
                }
        }
        for (unsigned i = 0; i < nstreams; i++) if (streams[i].bytes > 0) stats->bytes += (unsigned long long)streams[i].bytes;
        snprintf(stats->error, sizeof stats->error, "%s", co->err);
    }
    free(mw); free(co->unit_stream); free(co->st);
    co->unit_stream = NULL; co->st = NULL; co->streams = NULL; co->nstreams = 0;
    return rc;
}

int flake_b200_encode_corpus(const FlakeContext *proto, int pcm_format,
                             FlakeB200CorpusStream *streams, unsigned int nstreams,
                             const int *devices, int ndevices, const FlakeB200CorpusOptions *opt,
                             FlakeB200CorpusStats *stats)
{
    if (stats) memset(stats, 0, sizeof *stats);
    if (!proto || (!streams && nstreams)) return -1;
    /* the longest stream bounds the chunk size (short corpora get small engines) */
    FlakeContext p = *proto;
    unsigned long long longest = 0;
    for (unsigned i = 0; i < nstreams; i++) if (streams[i].nsamples > longest) longest = streams[i].nsamples;
    p.samples = longest > 0xffffffffull ? 0u : (unsigned)longest;
    if (fb_cuda_device_count() <= 0) {
        fprintf(stderr, "flake_b200: no CUDA device (there is no CPU encoding path)\n");
        return flake_validate_params(proto) < 0 ? -1 : -3;
    }
    FlakeB200Corpus *co = flake_b200_corpus_open(&p, pcm_format, devices, ndevices, opt);
    if (!co) return -1;
    const int rc = flake_b200_corpus_encode(co, streams, nstreams, stats);
    flake_b200_corpus_close(co);
    return rc;
}

/* "fLaC" + STREAMINFO + VORBIS_COMMENT (vendor only) + PADDING for one stream of a corpus: what
 * flake_encode_init leaves in FlakeContext.header (encode.c:125-156) with the FINAL STREAMINFO
 * the CLI patches in at the end (flake/flake.c:665-673). */
int flake_b200_corpus_stream_header(const FlakeContext *proto, const FlakeB200CorpusStream *S,
                                    unsigned char *data, unsigned int cap)
{
    if (!proto || !S || flake_validate_params(proto) < 0) return -1;
    const FlakeEncodeParams *p = &proto->params;
    FlakeVorbisComment vc;
    flake_init_vorbiscomment(&vc);
    int vsz = flake_get_vorbiscomment_size(&vc);
    if (vsz < 8) vsz = 8;
    const unsigned need = 4u + 4u + 34u + 4u + (unsigned)vsz + (p->padding_size > 0 ? 4u + (unsigned)p->padding_size : 0u);
    if (!data) return (int)need;
    if (cap < need) return -1;
    memset(data, 0, need);
    FlakeStreaminfo si;
    memset(&si, 0, sizeof si);
    si.min_block_size = (p->variable_block_size || p->allow_vbs) ? 16u : (unsigned)p->block_size;
    si.max_block_size = (unsigned)p->block_size;
    si.min_frame_size = 0;
    si.max_frame_size = S->max_frame_size;
    si.sample_rate = (unsigned)proto->sample_rate;
    si.channels = (unsigned)proto->channels;
    si.bits_per_sample = (unsigned)proto->bits_per_sample;
    si.samples = (unsigned)S->nsamples;
    memcpy(si.md5sum, S->md5sum, 16);
    unsigned pos = 0;
    memcpy(data, "fLaC", 4); pos = 4;
    data[pos] = 0; data[pos + 1] = 0; data[pos + 2] = 0; data[pos + 3] = 34;
    flake_write_streaminfo(&si, data + pos + 4);
    pos += 38;
    const int last_vc = p->padding_size == 0;
    data[pos] = (unsigned char)((last_vc ? 0x80 : 0) | 4);
    data[pos + 1] = (unsigned char)(vsz >> 16); data[pos + 2] = (unsigned char)(vsz >> 8); data[pos + 3] = (unsigned char)vsz;
    if (flake_write_vorbiscomment(&vc, data + pos + 4)) memset(data + pos + 4, 0, (size_t)vsz);
    pos += 4u + (unsigned)vsz;
    if (p->padding_size > 0) {
        data[pos] = 0x80 | 1;
        data[pos + 1] = (unsigned char)(p->padding_size >> 16); data[pos + 2] = (unsigned char)(p->padding_size >> 8);
        data[pos + 3] = (unsigned char)p->padding_size;
        pos += 4u + (unsigned)p->padding_size;
    }
    return (int)pos;
}
