/*
 * flake_host.c -- the C host layer of flake_b200's libflake: the flake.h API
 * (libflake/flake.h) plus the batch extension (include/flake_b200.h).
 *
 * What stays on the host, once per stream or per call: parameter presets and
 * validation (encode.c:158-373), the stream header (encode.c:52-156,
 * metadata.c), bookkeeping of the frame counter / running maximum frame size
 * (encode.c:966-976), and the MD5 of the PCM (md5.c) which is a serial chain
 * and runs on a host thread while the GPU encodes.  Everything per frame --
 * flake_encode_frame's body, encode.c:919-977 and below -- runs in the CUDA
 * engine (engine.cu).  There is no CPU encoding path: without a usable CUDA
 * device flake_encode_init fails.
 */
#define _GNU_SOURCE
#define FLAKE_BUILD_LIBRARY 1
#include "flake.h"
#include "flake_b200.h"
#include "engine.h"
#include "md5.h"
#include "flake_host_int.h"

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define FB_VERSION "SVN"
#define FB_FALLBACK_CHUNK_BLOCKS 2048

typedef struct FbLane {
    void *h_in, *d_in;              /* pinned staging + device copy of the chunk's PCM */
    void *h_pack;                   /* pinned: the chunk packed at ceil(bps/8) bytes (digest layout) */
    size_t pack_bytes;              /* bytes of h_pack in use */
    void *d_out, *h_out;            /* compacted frames */
    uint32_t *d_flen, *d_fbs, *h_flen, *h_fbs;
    FbSummary *d_sum, *h_sum;
    void *ev_done;                  /* summary is on the host */
    uint64_t nsamples;
} FbLane;

typedef struct FbCtx {
    FlakeContext *parent;
    FlakeEncodeParams params;
    int channels, samplerate, bps;
    uint32_t sample_count, frame_count;
    int max_frame_size;
    uint32_t min_frame_size;        /* smallest frame so far, 0xffffffff = none yet */
    int report_min_frame;           /* flake_b200_set_streaminfo_sizes */
    int last_frame;
    FbMd5 md5;
    FbConfig cfg;
    int device;
    /* per-block path */
    FbEngine *eng1;
    uint8_t *frame_buffer;          /* pinned; what flake_get_buffer returns */
    size_t frame_buffer_size;
    FbLane one;
    /* batch path */
    FbEngine *engN;             /* host streaming path: chunk_blocks per pass, two lanes */
    FbEngine *engD;             /* device-resident API: dev_chunk_blocks per pass */
    int profiling;              /* flake_b200_set_profiling: applied to engines created later too */
    FbEngine *last_engine;      /* whichever ran the most recent pass */
    int chunk_blocks, dev_chunk_blocks;
    FbLane lane[2];
    int lanes_ready;
    void *st, *st_copy, *ev_a, *ev_b;
    FlakeB200Stats stats;
    char err[256];
} FbCtx;

static int g_device = -2;           /* process default; -2: not chosen yet.  Accessed with __atomic builtins */
static __thread int t_device = -1;  /* this thread's choice (flake_b200_set_thread_device), -1: none */
struct FbCtx;
static int default_chunk_blocks(const struct FbCtx *c, unsigned int stream_samples, uint64_t target_ints);

double fb_now_ms(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

/* ------------------------------------------------------------------ */
/* presets / validation                                                 */
/* ------------------------------------------------------------------ */
int flake_set_defaults(FlakeEncodeParams *p)
{
    /* one row per level: block, prediction, min/max order, order method,
     * max partition order, stereo, vbs (encode.c:171-263) */
    static const short lv[13][8] = {
        {1152, FLAKE_PREDICTION_FIXED,    2,  2, FLAKE_ORDER_METHOD_EST,    3, 0, 0},
        {1152, FLAKE_PREDICTION_FIXED,    2,  4, FLAKE_ORDER_METHOD_EST,    3, 1, 0},
        {1152, FLAKE_PREDICTION_FIXED,    0,  4, FLAKE_ORDER_METHOD_EST,    3, 1, 0},
        {4096, FLAKE_PREDICTION_LEVINSON, 1,  6, FLAKE_ORDER_METHOD_EST,    4, 0, 0},
        {4096, FLAKE_PREDICTION_LEVINSON, 1,  8, FLAKE_ORDER_METHOD_EST,    4, 1, 0},
        {4096, FLAKE_PREDICTION_LEVINSON, 1,  8, FLAKE_ORDER_METHOD_EST,    5, 1, 0},
        {4096, FLAKE_PREDICTION_LEVINSON, 1,  8, FLAKE_ORDER_METHOD_EST,    6, 1, 0},
        {4096, FLAKE_PREDICTION_LEVINSON, 1,  8, FLAKE_ORDER_METHOD_4LEVEL, 6, 1, 0},
        {4096, FLAKE_PREDICTION_LEVINSON, 1, 12, FLAKE_ORDER_METHOD_LOG,    6, 1, 0},
        {4096, FLAKE_PREDICTION_LEVINSON, 1, 12, FLAKE_ORDER_METHOD_LOG,    8, 1, 1},
        {4096, FLAKE_PREDICTION_LEVINSON, 1, 12, FLAKE_ORDER_METHOD_SEARCH, 8, 1, 1},
        {8192, FLAKE_PREDICTION_LEVINSON, 1, 32, FLAKE_ORDER_METHOD_LOG,    8, 1, 1},
        {8192, FLAKE_PREDICTION_LEVINSON, 1, 32, FLAKE_ORDER_METHOD_SEARCH, 8, 1, 1},
    };
    if (!p || p->compression < 0 || p->compression > 12) return -1;
    const short *r = lv[p->compression];
    p->block_size = r[0];
    p->prediction_type = r[1];
    p->min_prediction_order = r[2];
    p->max_prediction_order = r[3];
    p->order_method = r[4];
    p->min_partition_order = 0;
    p->max_partition_order = r[5];
    p->stereo_method = r[6] ? FLAKE_STEREO_METHOD_ESTIMATE : FLAKE_STEREO_METHOD_INDEPENDENT;
    p->variable_block_size = r[7];
    p->allow_vbs = r[7];
    p->padding_size = 8192;
    return 0;
}

int flake_validate_params(const FlakeContext *s)
{
    if (!s) return -1;
    const FlakeEncodeParams *p = &s->params;
    int non_subset = 0;
    if (s->channels < 1 || s->channels > 8) return -1;
    if (s->sample_rate < 1 || s->sample_rate > 655350) return -1;
    if (s->bits_per_sample < 4 || s->bits_per_sample > 32) return -1;
    if (s->bits_per_sample < 8 || s->bits_per_sample > 24 || s->bits_per_sample % 4) non_subset = 1;
    if (p->compression < 0 || p->compression > 12) return -1;
    if (p->order_method < 0 || p->order_method > 6) return -1;
    if (p->stereo_method < 0 || p->stereo_method > 1) return -1;
    if (p->block_size < 16 || p->block_size > 65535) return -1;
    if (s->sample_rate <= 48000 && p->block_size > 4608) non_subset = 1;
    if (p->prediction_type < 0 || p->prediction_type > 2) return -1;
    if (p->min_prediction_order > p->max_prediction_order) return -1;
    if (p->prediction_type == FLAKE_PREDICTION_FIXED) {
        if (p->min_prediction_order < 0 || p->min_prediction_order > 4) return -1;
        if (p->max_prediction_order < 0 || p->max_prediction_order > 4) return -1;
    } else {
        if (p->min_prediction_order < 1 || p->min_prediction_order > 32) return -1;
        if (p->max_prediction_order < 1 || p->max_prediction_order > 32) return -1;
        if (s->sample_rate <= 48000 && p->max_prediction_order > 12) non_subset = 1;
    }
    if (p->min_partition_order > p->max_partition_order) return -1;
    if (p->min_partition_order < 0 || p->min_partition_order > 8) return -1;
    if (p->max_partition_order < 0 || p->max_partition_order > 8) return -1;
    if (p->padding_size < 0 || p->padding_size >= (1 << 24)) return -1;
    if (p->variable_block_size < 0 || p->variable_block_size > 1) return -1;
    if (p->variable_block_size > 0 && !p->allow_vbs) return -1;
    if (p->block_size < 128 && p->allow_vbs) return -1;
    return non_subset;
}

const char *flake_get_version(void) { return FB_VERSION; }
const char *flake_b200_version(void) { return "flake_b200 0.1 (sm_100a)"; }

/* ------------------------------------------------------------------ */
/* metadata                                                             */
/* ------------------------------------------------------------------ */
static char g_vendor[32];

void flake_init_vorbiscomment(FlakeVorbisComment *vc)
{
    if (!g_vendor[0]) snprintf(g_vendor, sizeof g_vendor, "Flake %s", flake_get_version());
    vc->vendor_string = g_vendor;
    vc->num_entries = 0;
    memset(vc->entries, 0, sizeof vc->entries);
}

/* 0 = well formed "NAME=value" with NAME in 0x20..0x7D excluding '=' (metadata.c:101-125) */
static int vc_entry_invalid(const char *e)
{
    int seen_eq = 0;
    for (size_t i = 0;; i++) {
        const char ch = e[i];
        if (!seen_eq && ch == '=') seen_eq = 1;
        if (ch == '\0') return seen_eq ? 0 : 1;
        if (!seen_eq && (ch < ' ' || ch > '}' || ch == '=')) return 1;
    }
}

static int vc_invalid(const FlakeVorbisComment *vc)
{
    if (vc->num_entries > 1024) return 1;
    for (unsigned i = 0; i < vc->num_entries; i++)
        if (!vc->entries[i] || vc_entry_invalid(vc->entries[i])) return 1;
    return 0;
}

int flake_add_vorbiscomment_entry(FlakeVorbisComment *vc, char *entry)
{
    const int bad = vc_entry_invalid(entry);
    if (!bad) vc->entries[vc->num_entries++] = entry;
    return bad;
}

int flake_get_vorbiscomment_size(const FlakeVorbisComment *vc)
{
    if (vc_invalid(vc)) return -1;
    unsigned long long sz = 8;
    if (vc->vendor_string) sz += strlen(vc->vendor_string);
    for (unsigned i = 0; i < vc->num_entries; i++) sz += 4 + strlen(vc->entries[i]);
    return sz > 0x7fffffffull ? -1 : (int)sz;
}

static unsigned char *put_le32(unsigned char *p, unsigned v)
{
    p[0] = (unsigned char)v; p[1] = (unsigned char)(v >> 8);
    p[2] = (unsigned char)(v >> 16); p[3] = (unsigned char)(v >> 24);
    return p + 4;
}

int flake_write_vorbiscomment(const FlakeVorbisComment *vc, unsigned char *data)
{
    if (flake_get_vorbiscomment_size(vc) < 0) return -1;
    unsigned len = vc->vendor_string ? (unsigned)strlen(vc->vendor_string) : 0;
    data = put_le32(data, len);
    if (len) memcpy(data, vc->vendor_string, len);
    data += len;
    data = put_le32(data, vc->num_entries);
    for (unsigned i = 0; i < vc->num_entries; i++) {
        len = (unsigned)strlen(vc->entries[i]);
        data = put_le32(data, len);
        memcpy(data, vc->entries[i], len);
        data += len;
    }
    return 0;
}

void flake_write_streaminfo(const FlakeStreaminfo *si, unsigned char *d)
{
    /* 16+16+24+24+20+3+5+4+32 bits, then the digest (metadata.c:67-84) */
    memset(d, 0, 34);
    d[0] = (unsigned char)(si->min_block_size >> 8); d[1] = (unsigned char)si->min_block_size;
    d[2] = (unsigned char)(si->max_block_size >> 8); d[3] = (unsigned char)si->max_block_size;
    d[4] = (unsigned char)(si->min_frame_size >> 16); d[5] = (unsigned char)(si->min_frame_size >> 8);
    d[6] = (unsigned char)si->min_frame_size;
    d[7] = (unsigned char)(si->max_frame_size >> 16); d[8] = (unsigned char)(si->max_frame_size >> 8);
    d[9] = (unsigned char)si->max_frame_size;
    const unsigned sr = si->sample_rate & 0xfffffu, ch = (si->channels - 1) & 7u;
    const unsigned bp = (si->bits_per_sample - 1) & 31u;
    d[10] = (unsigned char)(sr >> 12); d[11] = (unsigned char)(sr >> 4);
    d[12] = (unsigned char)(((sr & 15u) << 4) | (ch << 1) | (bp >> 4));
    d[13] = (unsigned char)((bp & 15u) << 4);            /* low nibble: 4 zero bits of the 36-bit count */
    d[14] = (unsigned char)(si->samples >> 24); d[15] = (unsigned char)(si->samples >> 16);
    d[16] = (unsigned char)(si->samples >> 8);  d[17] = (unsigned char)si->samples;
    memcpy(d + 18, si->md5sum, 16);
}

int flake_get_streaminfo(const FlakeContext *s, FlakeStreaminfo *si)
{
    if (!s || !si) return -1;
    if (flake_validate_params(s) < 0) return -1;
    const FbCtx *c = (const FbCtx *)s->private_ctx;
    if (!c) return -1;
    si->min_block_size = (c->params.variable_block_size || c->params.allow_vbs)
                             ? 16u : (unsigned)c->params.block_size;
    si->max_block_size = (unsigned)c->params.block_size;
    /* the reference always leaves 0 = unknown (metadata.c:52); the real value is opt-in */
    si->min_frame_size = (c->report_min_frame && c->min_frame_size != 0xffffffffu) ? c->min_frame_size : 0;
    si->max_frame_size = (unsigned)c->max_frame_size;
    si->sample_rate = (unsigned)c->samplerate;
    si->channels = (unsigned)c->channels;
    si->bits_per_sample = (unsigned)c->bps;
    si->samples = c->sample_count;
    fb_md5_final(&c->md5, si->md5sum);
    return 0;
}

static void put_block_header(unsigned char *p, int last, int type, unsigned size)
{
    p[0] = (unsigned char)((last ? 0x80 : 0) | (type & 0x7f));
    p[1] = (unsigned char)(size >> 16); p[2] = (unsigned char)(size >> 8); p[3] = (unsigned char)size;
}

/* "fLaC", STREAMINFO, VORBIS_COMMENT (vendor only), PADDING -- encode.c:125-156 */
static int write_stream_header(FbCtx *c, unsigned char *h)
{
    int pos = 0;
    memcpy(h, "fLaC", 4); pos = 4;
    put_block_header(h + pos, 0, 0, 34);
    FlakeStreaminfo si;
    if (!flake_get_streaminfo(c->parent, &si)) flake_write_streaminfo(&si, h + pos + 4);
    pos += 38;
    const int last_vc = c->params.padding_size == 0;
    FlakeVorbisComment vc;
    flake_init_vorbiscomment(&vc);
    int vsz = flake_get_vorbiscomment_size(&vc);
    if (vsz < 0) vsz = 8;
    put_block_header(h + pos, last_vc, 4, (unsigned)vsz);
    if (vsz <= 8 || flake_write_vorbiscomment(&vc, h + pos + 4)) {
        vsz = 8; put_block_header(h + pos, last_vc, 4, 8); memset(h + pos + 4, 0, 8);
    }
    pos += 4 + vsz;
    if (c->params.padding_size > 0) {
        put_block_header(h + pos, 1, 1, (unsigned)c->params.padding_size);
        pos += 4 + c->params.padding_size;
    }
    return pos;
}

/* ------------------------------------------------------------------ */
/* lanes                                                                */
/* ------------------------------------------------------------------ */
static void lane_free(FbLane *l)
{
    fb_cuda_free_host(l->h_in); fb_cuda_free_host(l->h_pack); fb_cuda_free(l->d_in); fb_cuda_free(l->d_out);
    fb_cuda_free_host(l->h_out);
    fb_cuda_free(l->d_flen); fb_cuda_free(l->d_fbs); fb_cuda_free(l->d_sum);
    fb_cuda_free_host(l->h_flen); fb_cuda_free_host(l->h_fbs); fb_cuda_free_host(l->h_sum);
    fb_cuda_event_destroy(l->ev_done);
    memset(l, 0, sizeof *l);
}

/* h_out_external: use this (pinned) buffer as h_out instead of allocating one */
static int lane_alloc(FbLane *l, const FbEngine *e, const FbConfig *cfg, void *h_out_external)
{
    const uint64_t in_bytes = (uint64_t)fb_engine_max_blocks(e) * (uint64_t)cfg->block_size *
                              (uint64_t)cfg->channels * 4u;
    const uint64_t out_bytes = fb_engine_out_capacity(e);
    const uint32_t mf = fb_engine_max_frames(e);
    memset(l, 0, sizeof *l);
    l->h_in = fb_cuda_malloc_host(in_bytes);
    l->h_pack = fb_cuda_malloc_host(in_bytes / 4u * (uint64_t)((cfg->bps + 7) >> 3) + 64u);
    l->d_in = fb_cuda_malloc(in_bytes);
    l->d_out = fb_cuda_malloc(out_bytes);
    l->h_out = h_out_external ? NULL : fb_cuda_malloc_host(out_bytes);
    l->d_flen = (uint32_t *)fb_cuda_malloc(sizeof(uint32_t) * mf);
    l->d_fbs = (uint32_t *)fb_cuda_malloc(sizeof(uint32_t) * mf);
    l->d_sum = (FbSummary *)fb_cuda_malloc(sizeof(FbSummary));
    l->h_flen = (uint32_t *)fb_cuda_malloc_host(sizeof(uint32_t) * mf);
    l->h_fbs = (uint32_t *)fb_cuda_malloc_host(sizeof(uint32_t) * mf);
    l->h_sum = (FbSummary *)fb_cuda_malloc_host(sizeof(FbSummary));
    l->ev_done = fb_cuda_event_create();
    if (!l->h_in || !l->h_pack || !l->d_in || !l->d_out || (!l->h_out && !h_out_external) || !l->d_flen ||
        !l->d_fbs || !l->d_sum || !l->h_flen || !l->h_fbs || !l->h_sum || !l->ev_done) {
        lane_free(l);
        return -1;
    }
    return 0;
}

size_t fb_pcm_container_bytes(int fmt)
{
    switch (fmt) {
    case FLAKE_B200_PCM_S16LE: return 2;
    case FLAKE_B200_PCM_S24LE: return 3;
    case FLAKE_B200_PCM_S8:    return 1;
    default:                   return 4;
    }
}

/*
 * int32 samples -> the digest layout of md5.c:281-320 (little-endian, ceil(bps/8) bytes,
 * low bytes of each sample) in one pass.  Returns 1 when every sample is representable
 * in that many bytes, i.e. the packed buffer can also feed the GPU (sign extension on
 * ingest reproduces the int32 exactly); 0 when some sample is out of range and the
 * encoder must see the caller's int32 values untouched.
 */
int fb_pack_s32(const int32_t *src, size_t count, int bytes, uint8_t *dst)
{
    int32_t bad = 0;
    if (bytes == 2) {
        int16_t *d = (int16_t *)dst;
        for (size_t i = 0; i < count; i++) { const int32_t v = src[i]; const int16_t t = (int16_t)v; d[i] = t; bad |= v ^ (int32_t)t; }
    } else if (bytes == 3) {
        for (size_t i = 0; i < count; i++) {
            const int32_t v = src[i];
            dst[3 * i] = (uint8_t)v; dst[3 * i + 1] = (uint8_t)(v >> 8); dst[3 * i + 2] = (uint8_t)(v >> 16);
            bad |= v ^ ((v << 8) >> 8);
        }
    } else if (bytes == 1) {
        int8_t *d = (int8_t *)dst;
        for (size_t i = 0; i < count; i++) { const int32_t v = src[i]; const int8_t t = (int8_t)v; d[i] = t; bad |= v ^ (int32_t)t; }
    } else {
        memcpy(dst, src, count * 4);
    }
    return bad == 0;
}

static int packed_format(int bytes)
{
    return bytes == 2 ? FLAKE_B200_PCM_S16LE : bytes == 3 ? FLAKE_B200_PCM_S24LE
         : bytes == 1 ? FLAKE_B200_PCM_S8 : FLAKE_B200_PCM_S32;
}

/* Host staging of one chunk.  int32 input is packed to the digest layout (which the MD5
 * thread then hashes straight from the pinned buffer, and which is also what gets uploaded
 * when it is lossless: half or three quarters of the PCIe bytes); packed input is copied. */
static void lane_stage(FbCtx *c, FbLane *l, const void *pcm, int fmt, uint64_t nsamples,
                       const void **upload, size_t *upload_bytes, int *upload_fmt)
{
    const size_t count = (size_t)nsamples * (size_t)c->channels;
    l->nsamples = nsamples;
    l->pack_bytes = 0;
    if (fmt == FLAKE_B200_PCM_S32) {
        const int bytes = (c->bps + 7) >> 3;
        const int lossless = fb_pack_s32((const int32_t *)pcm, count, bytes, (uint8_t *)l->h_pack);
        l->pack_bytes = count * (size_t)bytes;
        if (lossless) {
            *upload = l->h_pack; *upload_bytes = l->pack_bytes; *upload_fmt = packed_format(bytes);
            return;
        }
    }
    const size_t bytes = count * fb_pcm_container_bytes(fmt);
    memcpy(l->h_in, pcm, bytes);
    *upload = l->h_in; *upload_bytes = bytes; *upload_fmt = fmt;
}

/* upload + enqueue the engine pass + fetch the summary (all async) */
static int lane_launch(FbCtx *c, FbEngine *e, FbLane *l, const void *upload, size_t bytes, int fmt,
                       uint32_t first_number)
{
    if (fb_cuda_h2d(l->d_in, upload, bytes, c->st)) return -3;
    c->stats.h2d_bytes += bytes;
    const uint64_t before = fb_engine_launch_count(e);
    const int rc = fb_engine_encode_device(e, l->d_in, fmt, l->nsamples, first_number, l->d_out,
                                           l->d_flen, l->d_fbs, l->d_sum, c->st);
    if (rc) { snprintf(c->err, sizeof c->err, "%s", fb_engine_last_error(e)); return -3; }
    c->stats.kernel_launches += fb_engine_launch_count(e) - before;
    c->last_engine = e;
    if (fb_cuda_d2h(l->h_sum, l->d_sum, sizeof(FbSummary), c->st)) return -3;
    c->stats.d2h_bytes += sizeof(FbSummary);
    if (fb_cuda_event_record(l->ev_done, c->st)) return -3;
    return 0;
}

/* wait for the pass, then pull frames + lengths (exact sizes) on the copy stream */
static int lane_collect(FbCtx *c, FbLane *l, uint8_t *h_dst, uint64_t dst_cap, int want_bs)
{
    if (fb_cuda_event_sync(l->ev_done)) { snprintf(c->err, sizeof c->err, "CUDA failure while encoding"); return -3; }
    const FbSummary *sm = l->h_sum;
    if (sm->total_bytes > dst_cap) return -2;
    if (fb_cuda_d2h(h_dst, l->d_out, (size_t)sm->total_bytes, c->st_copy)) return -3;
    if (fb_cuda_d2h(l->h_flen, l->d_flen, sizeof(uint32_t) * sm->nframes, c->st_copy)) return -3;
    if (want_bs && fb_cuda_d2h(l->h_fbs, l->d_fbs, sizeof(uint32_t) * sm->nframes, c->st_copy)) return -3;
    if (fb_cuda_stream_sync(c->st_copy)) { snprintf(c->err, sizeof c->err, "CUDA failure while copying frames"); return -3; }
    c->stats.d2h_bytes += sm->total_bytes + (want_bs ? 8u : 4u) * (uint64_t)sm->nframes;
    return 0;
}

/* ------------------------------------------------------------------ */
/* init / close                                                         */
/* ------------------------------------------------------------------ */
static void ctx_free(FbCtx *c)
{
    if (!c) return;
    if (c->st) fb_cuda_stream_sync(c->st);
    lane_free(&c->one);
    lane_free(&c->lane[0]); lane_free(&c->lane[1]);
    if (c->eng1) fb_engine_destroy(c->eng1);
    if (c->engN) fb_engine_destroy(c->engN);
    if (c->engD) fb_engine_destroy(c->engD);
    fb_cuda_free_host(c->frame_buffer);
    fb_cuda_event_destroy(c->ev_a); fb_cuda_event_destroy(c->ev_b);
    fb_cuda_stream_destroy(c->st); fb_cuda_stream_destroy(c->st_copy);
    free(c);
}

/* The CUDA current device is per host thread: every entry point that touches the device binds
 * the calling thread to the context's device first (callers may use one thread per stream) and
 * puts the caller's device back before it returns. */
static int ctx_bind_device(const FbCtx *c)
{
    if (c->device < 0) return -1;
    const int prev = fb_cuda_current_device();
    if (prev == c->device) return -1;
    fb_cuda_set_device(c->device);
    return prev;                    /* the caller's device, to be restored on the way out */
}

static void ctx_unbind_device(int prev)
{
    if (prev >= 0) fb_cuda_set_device(prev);
}

int flake_b200_set_device(int device)
{
    if (device < 0 || device >= fb_cuda_device_count()) return -1;
    __atomic_store_n(&g_device, device, __ATOMIC_RELEASE);
    return 0;
}

int flake_b200_set_thread_device(int device)
{
    if (device < -1 || device >= fb_cuda_device_count()) return -1;
    t_device = device;
    return 0;
}

/* FbConfig of a context's public fields: the frame-header codes of encode.c:400-434 and the
 * encoding parameters.  Shared by flake_encode_init and flake_b200_encode_corpus. */
void fb_config_from_context(const FlakeContext *s, FbConfig *g)
{
    static const int rates[16] = {0, 0, 0, 0, 8000, 16000, 22050, 24000, 32000, 44100, 48000, 96000, 0, 0, 0, 0};
    static const int depths[8] = {0, 8, 12, 0, 16, 20, 24, 0};
    const FlakeEncodeParams *p = &s->params;
    memset(g, 0, sizeof *g);
    g->channels = s->channels; g->bps = s->bits_per_sample; g->block_size = p->block_size;
    g->sr_code0 = 0; g->sr_code1 = 0;
    int i;
    for (i = 4; i < 12; i++) if (s->sample_rate == rates[i]) { g->sr_code0 = i; break; }
    if (i == 12) {                                          /* encode.c:410-422 */
        if (s->sample_rate % 1000 == 0 && s->sample_rate <= 255000) { g->sr_code0 = 12; g->sr_code1 = s->sample_rate / 1000; }
        else if (s->sample_rate % 10 == 0 && s->sample_rate <= 655350) { g->sr_code0 = 14; g->sr_code1 = s->sample_rate / 10; }
        else if (s->sample_rate < 65535) { g->sr_code0 = 13; g->sr_code1 = s->sample_rate; }
    }
    g->bps_code = 0;
    for (i = 1; i < 8; i++) if (s->bits_per_sample == depths[i]) { g->bps_code = i; break; }
    g->order_method = p->order_method;
    g->stereo_method = p->stereo_method;
    g->prediction_type = p->prediction_type;
    g->min_order = p->min_prediction_order;
    g->max_order = p->max_prediction_order;
    g->min_porder = p->min_partition_order;
    g->max_porder = p->max_partition_order;
    g->variable_block_size = p->variable_block_size;
    g->allow_vbs = p->allow_vbs;
}

/* verbatim bound of a full block, encode.c:446-450: the seed of STREAMINFO's maximum frame size */
int fb_verbatim_bound(const FbConfig *g)
{
    if (g->channels == 2) return 16 + ((g->block_size * (2 * g->bps + 1) + 7) >> 3);
    return 16 + ((g->block_size * g->channels * g->bps + 7) >> 3);
}

static int flake_encode_init_impl(FlakeContext *s)
{
    if (!s) return -1;
    s->header = NULL;
    s->private_ctx = NULL;
    if (flake_validate_params(s) < 0) return -1;

    FbCtx *c = (FbCtx *)calloc(1, sizeof *c);
    if (!c) return -1;
    s->private_ctx = c;
    c->parent = s;
    c->params = s->params;
    c->channels = s->channels;
    c->samplerate = s->sample_rate;
    c->bps = s->bits_per_sample;
    c->sample_count = s->samples;

    FbConfig *g = &c->cfg;
    fb_config_from_context(s, g);

    c->max_frame_size = fb_verbatim_bound(g);

    /* header first: the reference serialises STREAMINFO before md5_init() on a
     * zeroed context (encode.c:391, 458-469), so the provisional digest is
     * that of an all-zero MD5 state. */
    s->header = (unsigned char *)calloc((size_t)c->params.padding_size + 1024, 1);
    if (!s->header) { ctx_free(c); s->private_ctx = NULL; return -1; }
    fb_md5_zero(&c->md5);
    const int header_len = write_stream_header(c, s->header);
    fb_md5_init(&c->md5);

    /* GPU side: device, streams, the one-block engine */
    int gd = __atomic_load_n(&g_device, __ATOMIC_ACQUIRE);
    if (gd == -2) {
        const char *env = getenv("FLAKE_B200_DEVICE");
        int expect = -2;
        gd = env ? atoi(env) : -1;
        if (!__atomic_compare_exchange_n(&g_device, &expect, gd, 0, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE)) gd = expect;
    }
    c->device = t_device >= 0 ? t_device : (gd >= 0 ? gd : fb_cuda_current_device());
    c->chunk_blocks = default_chunk_blocks(c, s->samples, FB_CHUNK_HOST_INTS);
    c->dev_chunk_blocks = default_chunk_blocks(c, s->samples, s->samples ? FB_CHUNK_DEVICE_INTS : FB_CHUNK_HOST_INTS);
    c->eng1 = fb_engine_create(g, c->device, 1, c->err, sizeof c->err);
    if (!c->eng1) {
        fprintf(stderr, "flake_b200: cannot create the CUDA engine: %s\n", c->err);
        free(s->header); s->header = NULL;
        ctx_free(c); s->private_ctx = NULL;
        return -1;
    }
    c->st = fb_cuda_stream_create();
    c->st_copy = fb_cuda_stream_create();
    c->ev_a = fb_cuda_event_create();
    c->ev_b = fb_cuda_event_create();
    c->frame_buffer_size = (size_t)fb_engine_out_capacity(c->eng1);
    if (c->frame_buffer_size < (size_t)c->max_frame_size * 3 / 2)
        c->frame_buffer_size = (size_t)c->max_frame_size * 3 / 2;
    c->frame_buffer = (uint8_t *)fb_cuda_malloc_host(c->frame_buffer_size);
    if (!c->st || !c->st_copy || !c->ev_a || !c->ev_b || !c->frame_buffer ||
        lane_alloc(&c->one, c->eng1, g, c->frame_buffer)) {
        fprintf(stderr, "flake_b200: CUDA allocation failed\n");
        free(s->header); s->header = NULL;
        ctx_free(c); s->private_ctx = NULL;
        return -1;
    }
    memset(c->frame_buffer, 0, c->frame_buffer_size);
    c->frame_count = 0;
    c->last_frame = 0;
    c->min_frame_size = 0xffffffffu;
    return header_len;
}

void *flake_get_buffer(const FlakeContext *s)
{
    if (!s || !s->private_ctx) return NULL;
    return ((FbCtx *)s->private_ctx)->frame_buffer;
}

static void flake_encode_close_impl(FlakeContext *s)
{
    if (!s || !s->private_ctx) return;
    ctx_free((FbCtx *)s->private_ctx);
    free(s->header);
    s->header = NULL;
    s->private_ctx = NULL;
}

/* encode.c:966-976 applied to a whole pass */
static void account(FbCtx *c, const FbSummary *sm, uint64_t nsamples)
{
    if ((int)sm->max_frame_bytes > c->max_frame_size) c->max_frame_size = (int)sm->max_frame_bytes;
    if (sm->nframes && ~sm->min_frame_inv < c->min_frame_size) c->min_frame_size = ~sm->min_frame_inv;
    if (c->params.allow_vbs) c->frame_count += (uint32_t)nsamples;
    else c->frame_count += sm->nframes;
    c->stats.frames += sm->nframes;
    c->stats.bytes += sm->total_bytes;
    c->stats.samples += nsamples;
    c->stats.verbatim_frames += sm->verbatim_frames;
    c->stats.max_frame_size = (unsigned)c->max_frame_size;
}

/* ------------------------------------------------------------------ */
/* flake_encode_frame -- encode.c:979-1008                              */
/* ------------------------------------------------------------------ */
static int flake_encode_frame_impl(FlakeContext *s, const int *samples, int block_size)
{
    if (!s || !samples || !s->private_ctx) return -1;
    FbCtx *c = (FbCtx *)s->private_ctx;
    if (block_size < 1 || block_size > c->params.block_size) return -1;
    if (c->last_frame) return -1;
    if (!c->params.allow_vbs && block_size != c->params.block_size) c->last_frame = 1;

    /* One block, one wait: k_pack writes the frame straight into the page-locked frame buffer
     * (unified addressing: the device reaches it by its host address), so nothing is left to copy
     * after the pass but the 24-byte summary that travels with it. */
    FbLane *l = &c->one;
    const void *up; size_t nb; int ufmt;
    lane_stage(c, l, samples, FLAKE_B200_PCM_S32, (uint64_t)block_size, &up, &nb, &ufmt);
    if (fb_cuda_h2d(l->d_in, up, nb, c->st)) return -1;
    c->stats.h2d_bytes += nb;
    const uint64_t before = fb_engine_launch_count(c->eng1);
    if (fb_engine_encode_device(c->eng1, l->d_in, ufmt, (uint64_t)block_size, c->frame_count, c->frame_buffer,
                                l->d_flen, NULL, l->d_sum, c->st)) {
        snprintf(c->err, sizeof c->err, "%s", fb_engine_last_error(c->eng1));
        return -1;
    }
    c->stats.kernel_launches += fb_engine_launch_count(c->eng1) - before;
    c->last_engine = c->eng1;
    if (fb_cuda_d2h(l->h_sum, l->d_sum, sizeof(FbSummary), c->st) || fb_cuda_event_record(l->ev_done, c->st)) return -1;
    /* MD5 of this block while the GPU works (encode.c:1005-1006) */
    FbMd5 md5 = c->md5;
    fb_md5_update(&md5, l->h_pack, l->pack_bytes);             /* packed by lane_stage */
    if (fb_cuda_event_sync(l->ev_done)) { snprintf(c->err, sizeof c->err, "CUDA failure while encoding"); return -1; }
    const FbSummary *sm = l->h_sum;
    c->stats.d2h_bytes += sizeof(FbSummary) + sm->total_bytes;
    if (sm->total_bytes == 0 || sm->total_bytes > 0x7fffffffull) return -1;
    account(c, sm, (uint64_t)block_size);
    c->md5 = md5;
    return (int)sm->total_bytes;
}

/* ------------------------------------------------------------------ */
/* batch extension                                                      */
/* ------------------------------------------------------------------ */
/*
 * MD5 thread of a batch call.  Two sources:
 *   direct  the caller's buffer already has the digest layout (packed input whose container
 *           is ceil(bps/8) bytes): one fb_md5_update over it;
 *   piped   int32 (or odd-container) input: the main thread packs each chunk into the lane's
 *           pinned h_pack while staging it, this thread hashes chunk after chunk.  A lane is
 *           re-used two chunks later, so the main thread waits for `consumed` before packing.
 */
typedef struct Md5Pipe {
    FbMd5 *md5;
    const void *direct; size_t direct_bytes;
    pthread_mutex_t mu; pthread_cond_t cv;
    uint64_t produced, consumed, total;
    const void *ptr[2]; size_t len[2];
    double ms;
} Md5Pipe;

static void *md5_worker(void *arg)
{
    Md5Pipe *p = (Md5Pipe *)arg;
    if (p->direct) {
        const double t0 = fb_now_ms();
        fb_md5_update(p->md5, p->direct, p->direct_bytes);
        p->ms = fb_now_ms() - t0;
        return NULL;
    }
    for (uint64_t k = 0; k < p->total; k++) {
        pthread_mutex_lock(&p->mu);
        while (p->produced <= k) pthread_cond_wait(&p->cv, &p->mu);
        const void *ptr = p->ptr[k & 1]; const size_t len = p->len[k & 1];
        pthread_mutex_unlock(&p->mu);
        const double t0 = fb_now_ms();
        fb_md5_update(p->md5, ptr, len);
        p->ms += fb_now_ms() - t0;
        pthread_mutex_lock(&p->mu);
        p->consumed = k + 1;
        pthread_cond_broadcast(&p->cv);
        pthread_mutex_unlock(&p->mu);
    }
    return NULL;
}

/*
 * Blocks per engine pass.  Every kernel's grid is a multiple of the block count, so the
 * defaults are multiples of the device's SM count (whole waves for the per-subframe grids).
 *   device-resident API: up to 320 Mi channel-samples (B200: 148 x 276 = 40,848 blocks of 4096
 *     stereo samples, 1.25 GB per int32 plane buffer).  k_lpc is a lane pair per subframe,
 *     i.e. only subframes/16 warps: 8 warps per scheduler instead of 2 (measured on a 1-hour
 *     stream: 7.59 ms per pass with 10,212-block chunks, 7.16 ms in one pass);
 *   host streaming path: 80 Mi, so that upload, kernels, download and MD5 of different chunks
 *     overlap and the pinned lanes stay small; also when the stream length is not known.
 * Streams shorter than that get an engine of their own size.  FLAKE_B200_CHUNK_BLOCKS and
 * flake_b200_set_chunk_blocks() override both.
 */
int fb_chunk_blocks_for(int device, int block_size, int channels, uint64_t stream_samples, uint64_t target_ints)
{
    const char *cb = getenv("FLAKE_B200_CHUNK_BLOCKS");
    if (cb && atoi(cb) >= 1) return atoi(cb);
    const uint64_t per_block = (uint64_t)block_size * (uint64_t)channels;
    const int sms = fb_cuda_sm_count(device);
    uint64_t blocks = target_ints / (per_block ? per_block : 1);
    if (sms > 0 && blocks >= (uint64_t)sms) blocks -= blocks % (uint64_t)sms;
    if (blocks < 1) blocks = 1;
    if (sms <= 0 && blocks > FB_FALLBACK_CHUNK_BLOCKS) blocks = FB_FALLBACK_CHUNK_BLOCKS;
    if (stream_samples) {
        const uint64_t need = (stream_samples + (uint64_t)block_size - 1) / (uint64_t)block_size;
        if (need < blocks) blocks = need;
    }
    return (int)blocks;
}

static int default_chunk_blocks(const FbCtx *c, unsigned int stream_samples, uint64_t target_ints)
{
    return fb_chunk_blocks_for(c->device, c->params.block_size, c->channels, stream_samples, target_ints);
}

static int flake_b200_set_chunk_blocks_impl(FlakeContext *s, int blocks)
{
    if (!s || !s->private_ctx || blocks < 1) return -1;
    FbCtx *c = (FbCtx *)s->private_ctx;
    if (c->engN && blocks != c->chunk_blocks) {
        fb_cuda_stream_sync(c->st);
        lane_free(&c->lane[0]); lane_free(&c->lane[1]);
        if (c->last_engine == c->engN) c->last_engine = NULL;
        fb_engine_destroy(c->engN); c->engN = NULL; c->lanes_ready = 0;
    }
    if (c->engD && blocks != c->dev_chunk_blocks) {
        fb_cuda_stream_sync(c->st);
        if (c->last_engine == c->engD) c->last_engine = NULL;
        fb_engine_destroy(c->engD); c->engD = NULL;
    }
    c->chunk_blocks = blocks;
    c->dev_chunk_blocks = blocks;
    return 0;
}

static int ensure_batch_engine(FbCtx *c)
{
    if (c->engN && c->lanes_ready) return 0;
    if (!c->engN) {
        c->engN = fb_engine_create(&c->cfg, c->device, (uint32_t)c->chunk_blocks, c->err, sizeof c->err);
        if (!c->engN) return -3;
        if (c->profiling) fb_engine_set_timing(c->engN, 1);
    }
    /* lane_alloc starts from a zeroed lane: never called on a live one (lanes_ready guards) */
    if (lane_alloc(&c->lane[0], c->engN, &c->cfg, NULL) || lane_alloc(&c->lane[1], c->engN, &c->cfg, NULL)) {
        lane_free(&c->lane[0]); lane_free(&c->lane[1]);          /* whichever half was allocated */
        snprintf(c->err, sizeof c->err, "CUDA allocation of the batch lanes failed");
        return -3;
    }
    c->lanes_ready = 1;
    return 0;
}

unsigned long long flake_b200_max_encoded_size(const FlakeContext *s, unsigned long long nsamples)
{
    /* from the public fields: also valid for a context that was never initialised (the prototype
     * of flake_b200_encode_corpus) */
    if (!s || s->params.block_size < 1) return 0;
    const unsigned long long B = (unsigned long long)s->params.block_size;
    unsigned long long frames = (nsamples + B - 1) / B;
    if (s->params.variable_block_size) frames *= 8;
    return frames * 96ull + ((nsamples * (unsigned long long)(s->channels * s->bits_per_sample + 1) + 7) >> 3) + 64;
}

static long long flake_b200_encode_stream_impl(FlakeContext *s, const void *pcm, int fmt,
                                   unsigned long long nsamples, unsigned char *out,
                                   unsigned long long out_cap, unsigned int *frame_len,
                                   unsigned int *frame_bs, unsigned int frame_cap,
                                   unsigned int *nframes_out)
{
    if (!s || !s->private_ctx || !pcm || !out) return -1;
    FbCtx *c = (FbCtx *)s->private_ctx;
    if (fmt < FLAKE_B200_PCM_S32 || fmt > FLAKE_B200_PCM_S8) return -1;
    if (nframes_out) *nframes_out = 0;
    if (nsamples == 0) return 0;
    if (c->last_frame) return -1;
    int rc = ensure_batch_engine(c);
    if (rc) return rc;

    const double t0 = fb_now_ms();
    const uint64_t B = (uint64_t)c->params.block_size;
    const uint64_t chunk = (uint64_t)c->chunk_blocks * B;
    const size_t bps_in = fb_pcm_container_bytes(fmt) * (size_t)c->channels;   /* bytes per inter-channel sample */
    const uint64_t nchunks = (nsamples + chunk - 1) / chunk;

    /* MD5 runs beside the GPU for the whole call */
    const int digest_bytes = (c->bps + 7) >> 3;
    Md5Pipe pipe;
    memset(&pipe, 0, sizeof pipe);
    pipe.md5 = &c->md5;
    pipe.total = nchunks;
    pthread_mutex_init(&pipe.mu, NULL);
    pthread_cond_init(&pipe.cv, NULL);
    if (fmt != FLAKE_B200_PCM_S32 && fb_pcm_container_bytes(fmt) == (size_t)digest_bytes) {
        pipe.direct = pcm;
        pipe.direct_bytes = (size_t)nsamples * bps_in;
    }
    FbMd5 md5_backup = c->md5;
    pthread_t th;
    const int have_thread = pthread_create(&th, NULL, md5_worker, &pipe) == 0;

    fb_cuda_event_record(c->ev_a, c->st);
    uint64_t out_pos = 0;
    uint32_t nf = 0;
    uint32_t counter = c->frame_count;
    const uint32_t counter0 = c->frame_count;
    const int max0 = c->max_frame_size;
    const uint32_t min0 = c->min_frame_size;
    const FlakeB200Stats stats0 = c->stats;
    int err = 0;
    int32_t *widen = NULL;          /* only for packed input in a container != ceil(bps/8) */

    for (uint64_t k = 0; k <= nchunks && !err; k++) {
        if (k < nchunks) {
            const uint64_t off = k * chunk;
            const uint64_t ns = nsamples - off < chunk ? nsamples - off : chunk;
            FbLane *l = &c->lane[k & 1];
            const uint8_t *src = (const uint8_t *)pcm + off * bps_in;
            if (!pipe.direct && have_thread) {
                /* this lane's previous packed chunk (k-2) must have been hashed */
                pthread_mutex_lock(&pipe.mu);
                while (k >= 2 && pipe.consumed < k - 1) pthread_cond_wait(&pipe.cv, &pipe.mu);
                pthread_mutex_unlock(&pipe.mu);
            }
            const void *up; size_t nb; int ufmt;
            lane_stage(c, l, src, fmt, ns, &up, &nb, &ufmt);
            if (!pipe.direct) {
                if (fmt != FLAKE_B200_PCM_S32) {
                    /* container wider/narrower than the digest layout: widen, then pack */
                    const size_t cnt = (size_t)ns * (size_t)c->channels, have = fb_pcm_container_bytes(fmt);
                    if (!widen) widen = (int32_t *)malloc(sizeof(int32_t) * (size_t)chunk * (size_t)c->channels);
                    if (widen) {
                        const uint8_t *q = src;
                        for (size_t i = 0; i < cnt; i++, q += have)
                            widen[i] = have == 2 ? (int16_t)(q[0] | (q[1] << 8))
                                     : have == 3 ? (int32_t)((uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)(int8_t)q[2] << 16))
                                     : (int8_t)q[0];
                        fb_pack_s32(widen, cnt, digest_bytes, (uint8_t *)l->h_pack);
                        l->pack_bytes = cnt * (size_t)digest_bytes;
                    } else err = -3;
                }
                if (have_thread) {
                    pthread_mutex_lock(&pipe.mu);
                    pipe.ptr[k & 1] = l->h_pack; pipe.len[k & 1] = l->pack_bytes;
                    pipe.produced = k + 1;
                    pthread_cond_broadcast(&pipe.cv);
                    pthread_mutex_unlock(&pipe.mu);
                } else {
                    fb_md5_update(&c->md5, l->h_pack, l->pack_bytes);
                }
            }
            if (!err) err = lane_launch(c, c->engN, l, up, nb, ufmt, counter);
            /* header number of the next chunk: blocks, or samples when allow_vbs.
             * Under VBS the frame count of a chunk is data dependent, but then
             * allow_vbs is set and the counter advances by samples. */
            counter += c->params.allow_vbs ? (uint32_t)ns : (uint32_t)((ns + B - 1) / B);
            if (err) break;
        }
        if (k >= 1) {
            FbLane *l = &c->lane[(k - 1) & 1];
            if (fb_cuda_event_sync(l->ev_done)) { err = -3; snprintf(c->err, sizeof c->err, "CUDA failure while encoding"); break; }
            const FbSummary sm = *l->h_sum;
            if (out_pos + sm.total_bytes > out_cap) { err = -2; break; }
            if ((frame_len || frame_bs) && nf + sm.nframes > frame_cap) { err = -2; break; }
            err = lane_collect(c, l, (uint8_t *)l->h_out, fb_engine_out_capacity(c->engN), frame_bs != NULL);
            if (err) break;
            memcpy(out + out_pos, l->h_out, (size_t)sm.total_bytes);
            if (frame_len) memcpy(frame_len + nf, l->h_flen, sizeof(uint32_t) * sm.nframes);
            if (frame_bs) memcpy(frame_bs + nf, l->h_fbs, sizeof(uint32_t) * sm.nframes);
            out_pos += sm.total_bytes;
            nf += sm.nframes;
            account(c, &sm, l->nsamples);
        }
    }
    fb_cuda_event_record(c->ev_b, c->st);
    fb_cuda_stream_sync(c->st);
    if (have_thread) {
        if (err && !pipe.direct) {
            /* let the worker run out: mark every remaining chunk as an empty one */
            pthread_mutex_lock(&pipe.mu);
            pipe.len[0] = pipe.len[1] = 0;
            pipe.produced = pipe.total;
            pthread_cond_broadcast(&pipe.cv);
            pthread_mutex_unlock(&pipe.mu);
        }
        pthread_join(th, NULL);
    } else if (pipe.direct) {
        md5_worker(&pipe);
    }
    free(widen);
    pthread_mutex_destroy(&pipe.mu);
    pthread_cond_destroy(&pipe.cv);
    struct { double ms; } job = { pipe.ms };

    if (err) {
        /* leave the context as it was before the call */
        c->md5 = md5_backup;
        c->frame_count = counter0;
        c->max_frame_size = max0;
        c->min_frame_size = min0;
        c->stats = stats0;
        return err;
    }
    if (!c->params.allow_vbs && (nsamples % B) != 0) c->last_frame = 1;   /* encode.c:991-994 */
    if (nframes_out) *nframes_out = nf;
    const float ms = fb_cuda_event_elapsed_ms(c->ev_a, c->ev_b);
    c->stats.gpu_ms += ms > 0 ? ms : 0;
    c->stats.md5_ms += job.ms;
    c->stats.wall_ms += fb_now_ms() - t0;
    return (long long)out_pos;
}

int flake_b200_seek(FlakeContext *s, unsigned int frame_counter)
{
    if (!s || !s->private_ctx) return -1;
    ((FbCtx *)s->private_ctx)->frame_count = frame_counter;
    return 0;
}

int flake_b200_reset_stream(FlakeContext *s)
{
    if (!s || !s->private_ctx) return -1;
    FbCtx *c = (FbCtx *)s->private_ctx;
    c->frame_count = 0;
    c->last_frame = 0;
    c->min_frame_size = 0xffffffffu;
    fb_md5_init(&c->md5);
    c->max_frame_size = fb_verbatim_bound(&c->cfg);
    memset(&c->stats, 0, sizeof c->stats);
    return 0;
}

unsigned int flake_b200_tell(const FlakeContext *s)
{
    if (!s || !s->private_ctx) return 0;
    return ((const FbCtx *)s->private_ctx)->frame_count;
}

static int flake_b200_device_capacity_impl(FlakeContext *s, unsigned long long *max_samples,
                               unsigned long long *out_bytes, unsigned int *max_frames)
{
    if (!s || !s->private_ctx) return -1;
    FbCtx *c = (FbCtx *)s->private_ctx;
    if (!c->engD) {
        c->engD = fb_engine_create(&c->cfg, c->device, (uint32_t)c->dev_chunk_blocks, c->err, sizeof c->err);
        if (!c->engD) return -3;
        if (c->profiling) fb_engine_set_timing(c->engD, 1);
    }
    if (max_samples) *max_samples = (unsigned long long)fb_engine_max_blocks(c->engD) * (unsigned long long)c->params.block_size;
    if (out_bytes) *out_bytes = fb_engine_out_capacity(c->engD);
    if (max_frames) *max_frames = fb_engine_max_frames(c->engD);
    return 0;
}

int flake_b200_encode_device(FlakeContext *s, const void *d_pcm, int fmt, unsigned long long nsamples,
                             unsigned int first_number, void *d_out, unsigned int *d_frame_len,
                             unsigned int *d_frame_bs, void *d_summary, void *cuda_stream)
{
    if (!s || !s->private_ctx || !d_pcm || !d_out || !d_summary) return -1;
    FbCtx *c = (FbCtx *)s->private_ctx;
    if (flake_b200_device_capacity(s, NULL, NULL, NULL)) return -3;
    const uint64_t before = fb_engine_launch_count(c->engD);
    const int rc = fb_engine_encode_device(c->engD, d_pcm, fmt, nsamples, first_number, d_out,
                                           d_frame_len, d_frame_bs, (FbSummary *)d_summary,
                                           cuda_stream ? cuda_stream : c->st);
    c->stats.kernel_launches += fb_engine_launch_count(c->engD) - before;
    c->last_engine = c->engD;
    if (rc) snprintf(c->err, sizeof c->err, "%s", fb_engine_last_error(c->engD));
    return rc;
}

int flake_b200_set_profiling(FlakeContext *s, int on)
{
    if (!s || !s->private_ctx) return -1;
    FbCtx *c = (FbCtx *)s->private_ctx;
    /* every engine the context owns at this moment: the per-block one, the streaming one, the device one */
    FbEngine *engs[3] = { c->eng1, c->engN, c->engD };
    int rc = 0;
    c->profiling = on ? 1 : 0;
    for (int i = 0; i < 3; i++) {
        if (!engs[i]) continue;
        if (on) fb_engine_reset_timing(engs[i]);
        if (fb_engine_set_timing(engs[i], on)) rc = -1;
    }
    return rc;
}

int flake_b200_stage_times(FlakeContext *s, double *ms, unsigned long long *launches)
{
    if (!s || !s->private_ctx) return -1;
    FbCtx *c = (FbCtx *)s->private_ctx;
    FbEngine *engs[3] = { c->eng1, c->engN, c->engD };
    double tot[FB_NUM_STAGES] = { 0 };
    uint64_t cnt[FB_NUM_STAGES] = { 0 };
    for (int k = 0; k < 3; k++) {
        double m[FB_NUM_STAGES]; uint64_t l[FB_NUM_STAGES];
        if (!engs[k] || fb_engine_collect_timing(engs[k], m, l) < 0) continue;
        for (int i = 0; i < FB_NUM_STAGES; i++) { tot[i] += m[i]; cnt[i] += l[i]; }
    }
    for (int i = 0; i < FB_NUM_STAGES; i++) {
        if (ms) ms[i] = tot[i];
        if (launches) launches[i] = cnt[i];
    }
    return FB_NUM_STAGES;
}

int flake_b200_set_streaminfo_sizes(FlakeContext *s, int on)
{
    if (!s || !s->private_ctx) return -1;
    ((FbCtx *)s->private_ctx)->report_min_frame = on ? 1 : 0;
    return 0;
}

unsigned int flake_b200_subframe_record_size(void) { return (unsigned int)sizeof(FbSub); }

static int flake_b200_last_subframes_impl(FlakeContext *s, void *subs, unsigned int max)
{
    if (!s || !s->private_ctx || !subs) return -1;
    FbCtx *c = (FbCtx *)s->private_ctx;
    FbEngine *e = c->last_engine ? c->last_engine : c->eng1;
    return fb_engine_read_subframes(e, (FbSub *)subs, max, c->st);
}

int flake_b200_get_stats(const FlakeContext *s, FlakeB200Stats *st)
{
    if (!s || !s->private_ctx || !st) return -1;
    *st = ((const FbCtx *)s->private_ctx)->stats;
    return 0;
}

long long flake_b200_write_seektable(const unsigned int *frame_len, const unsigned int *frame_bs,
                                     unsigned int nframes, unsigned int interval_samples,
                                     unsigned char *data, unsigned long long cap)
{
    if (!frame_len || !frame_bs) return -1;
    unsigned long long sample = 0, offset = 0, next = 0, written = 0;
    for (unsigned int f = 0; f < nframes; f++) {
        if (sample >= next) {
            if (data) {
                if (written + 18u > cap) return -1;
                unsigned char *p = data + written;
                for (int i = 0; i < 8; i++) p[i] = (unsigned char)(sample >> (56 - 8 * i));
                for (int i = 0; i < 8; i++) p[8 + i] = (unsigned char)(offset >> (56 - 8 * i));
                p[16] = (unsigned char)(frame_bs[f] >> 8);
                p[17] = (unsigned char)frame_bs[f];
            }
            written += 18u;
            if (interval_samples) {
                while (next <= sample) next += interval_samples;
            } else {
                next = sample + 1;
            }
        }
        sample += frame_bs[f];
        offset += frame_len[f];
    }
    return (long long)written;
}

const char *flake_b200_last_error(const FlakeContext *s)
{
    if (!s || !s->private_ctx) return "no context";
    return ((const FbCtx *)s->private_ctx)->err;
}

/* ------------------------------------------------------------------ */
/* entry points that touch the device: bind the context's device, restore the caller's */
/* ------------------------------------------------------------------ */
void flake_encode_close(FlakeContext *s)
{
    if (!s || !s->private_ctx) return;
    const int prev = ctx_bind_device((const FbCtx *)s->private_ctx);
    flake_encode_close_impl(s);
    ctx_unbind_device(prev);
}

int flake_encode_frame(FlakeContext *s, const int *samples, int block_size)
{
    if (!s || !samples || !s->private_ctx) return -1;
    const int prev = ctx_bind_device((const FbCtx *)s->private_ctx);
    const int r = flake_encode_frame_impl(s, samples, block_size);
    ctx_unbind_device(prev);
    return r;
}

int flake_b200_set_chunk_blocks(FlakeContext *s, int blocks)
{
    if (!s || !s->private_ctx || blocks < 1) return -1;
    const int prev = ctx_bind_device((const FbCtx *)s->private_ctx);
    const int r = flake_b200_set_chunk_blocks_impl(s, blocks);
    ctx_unbind_device(prev);
    return r;
}

long long flake_b200_encode_stream(FlakeContext *s, const void *pcm, int fmt, unsigned long long nsamples, unsigned char *out,
                                   unsigned long long out_cap, unsigned int *frame_len, unsigned int *frame_bs,
                                   unsigned int frame_cap, unsigned int *nframes_out)
{
    if (!s || !s->private_ctx || !pcm || !out) return -1;
    const int prev = ctx_bind_device((const FbCtx *)s->private_ctx);
    const long long r = flake_b200_encode_stream_impl(s, pcm, fmt, nsamples, out, out_cap, frame_len, frame_bs, frame_cap, nframes_out);
    ctx_unbind_device(prev);
    return r;
}

int flake_b200_device_capacity(FlakeContext *s, unsigned long long *max_samples, unsigned long long *out_bytes,
                               unsigned int *max_frames)
{
    if (!s || !s->private_ctx) return -1;
    const int prev = ctx_bind_device((const FbCtx *)s->private_ctx);
    const int r = flake_b200_device_capacity_impl(s, max_samples, out_bytes, max_frames);
    ctx_unbind_device(prev);
    return r;
}

int flake_b200_last_subframes(FlakeContext *s, void *subs, unsigned int max)
{
    if (!s || !s->private_ctx || !subs) return -1;
    const int prev = ctx_bind_device((const FbCtx *)s->private_ctx);
    const int r = flake_b200_last_subframes_impl(s, subs, max);
    ctx_unbind_device(prev);
    return r;
}

int flake_encode_init(FlakeContext *s)
{
    const int prev = fb_cuda_current_device();      /* engine creation selects the context's device */
    const int r = flake_encode_init_impl(s);
    if (prev >= 0 && fb_cuda_current_device() != prev) fb_cuda_set_device(prev);
    return r;
}
