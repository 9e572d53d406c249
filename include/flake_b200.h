/*
 * flake_b200.h -- batch and device-resident entry points, exported by
 * libflake.so IN ADDITION to the unchanged flake.h API.
 *
 * Why they exist: flake_encode_frame (libflake/encode.c:979-1008) hands the
 * encoder ONE block per synchronous call, and the caller cannot supply block
 * N+1 before the call for block N returns (flake/flake.c:624-663).  One 4096
 * sample block is far too little work for a B200, so the per-block call is
 * latency bound.  flake_b200_encode_stream takes any number of consecutive
 * blocks and produces exactly the bytes the per-block loop would have
 * produced, with the same effect on the context (frame counter, running
 * maximum frame size, MD5 of the PCM).
 *
 * Each function states the reference interface it stands in for.
 */
#ifndef FLAKE_B200_H
#define FLAKE_B200_H

#include "flake.h"

#ifdef __cplusplus
extern "C" {
#endif

/* PCM container of the samples handed to the batch calls (always channel
 * interleaved).  S32 is the flake_encode_frame convention (flake.h:225-226);
 * the packed layouts are what libpcm_io reads from a WAV data chunk before it
 * widens to int32 (libpcm_io/pcm_io.c:155-277) -- the widening then happens on
 * the GPU. */
enum {
    FLAKE_B200_PCM_S32   = 0,
    FLAKE_B200_PCM_S16LE = 1,
    FLAKE_B200_PCM_S24LE = 2,
    FLAKE_B200_PCM_S8    = 3
};

typedef struct FlakeB200Stats {
    unsigned long long samples;         /* inter-channel samples encoded by the call      */
    unsigned long long frames;
    unsigned long long bytes;           /* frame bytes produced                            */
    unsigned long long h2d_bytes;       /* host->device bytes copied                       */
    unsigned long long d2h_bytes;       /* device->host bytes copied                       */
    unsigned long long kernel_launches; /* CUDA kernels launched by the call               */
    unsigned int max_frame_size;
    unsigned int verbatim_frames;       /* frames that took the size fallback (encode.c:949) */
    double gpu_ms;                      /* CUDA-event time of the call's kernels + copies  */
    double md5_ms;                      /* host MD5 thread busy time                       */
    double wall_ms;
} FlakeB200Stats;

/* Select the CUDA device used by contexts initialised afterwards (default:
 * $FLAKE_B200_DEVICE, else the current device).  Returns 0 or -1. */
FLAKE_API int flake_b200_set_device(int device);
/* The same for contexts initialised BY THE CALLING THREAD only (takes precedence over the process
 * default; -1 clears it): threads that each drive a GPU of their own need no shared state. */
FLAKE_API int flake_b200_set_thread_device(int device);

/* Blocks per engine pass of the batch calls (default: $FLAKE_B200_CHUNK_BLOCKS, else a multiple of
 * the device's SM count near 80 Mi channel-samples for host buffers, 320 Mi for device-resident input). */
FLAKE_API int flake_b200_set_chunk_blocks(FlakeContext *s, int blocks);

/*
 * Batch form of the flake/flake.c:624-663 loop over flake_encode_frame
 * (encode.c:979-1008): encodes `nsamples` inter-channel samples as consecutive
 * blocks of params.block_size (the last one may be short) from HOST memory and
 * writes the frames back to back into `out` (host).  frame_len / frame_bs
 * (optional, `frame_cap` entries) receive each frame's byte length and block
 * size; *nframes the number of frames.  Updates the context exactly as the
 * per-block calls would (frame counter, max frame size, MD5, last-block latch).
 * Returns the number of bytes written, or a negative value:
 *   -1 bad arguments / closed context / stream already ended with a short block
 *   -2 output or frame arrays too small      -3 CUDA failure (see flake_b200_last_error)
 */
FLAKE_API long long flake_b200_encode_stream(FlakeContext *s, const void *pcm, int pcm_format,
                                             unsigned long long nsamples,
                                             unsigned char *out, unsigned long long out_cap,
                                             unsigned int *frame_len, unsigned int *frame_bs,
                                             unsigned int frame_cap, unsigned int *nframes);

/* Upper bound of the bytes flake_b200_encode_stream can produce for nsamples. */
FLAKE_API unsigned long long flake_b200_max_encoded_size(const FlakeContext *s,
                                                         unsigned long long nsamples);

/*
 * Frame-range sharding: a frame's bytes depend only on its
 * samples, the stream parameters and its header number (encode.c:726-764,
 * 969-975), so a rank that encodes blocks [b0, b1) of a stream first seeks its
 * context to the counter the serial encoder would have had there:
 * b0 for fixed block size, b0*block_size samples when allow_vbs is set.
 */
FLAKE_API int flake_b200_seek(FlakeContext *s, unsigned int frame_counter);
FLAKE_API unsigned int flake_b200_tell(const FlakeContext *s);
/* Start a new stream with the same parameters on an initialised context: frame
 * counter 0, fresh MD5, max frame size back to the verbatim bound (the state
 * flake_encode_init leaves, encode.c:446-469), statistics cleared.  Saves
 * re-creating the GPU engine between files of one format. */
FLAKE_API int flake_b200_reset_stream(FlakeContext *s);

/*
 * Device-resident form: `d_pcm` and all outputs are DEVICE pointers, work is
 * enqueued on `cuda_stream` (a cudaStream_t; NULL = the context's own stream)
 * and the call returns without synchronising.  At most
 * flake_b200_device_capacity() samples per call.  `d_summary` receives
 * {uint32 nframes, uint32 max_frame_bytes, uint64 total_bytes, uint32
 * verbatim_frames, uint32 ~min_frame_bytes}.  Passes of one context share scratch buffers and
 * are therefore serial: a call on another stream than the previous one is made to wait for it
 * (stream-ordered, no host synchronisation).  Does not touch the context's MD5 or counters:
 * `first_number` is the header number of the first frame.
 * Returns 0 or a negative error.
 */
FLAKE_API int flake_b200_encode_device(FlakeContext *s, const void *d_pcm, int pcm_format,
                                       unsigned long long nsamples, unsigned int first_number,
                                       void *d_out, unsigned int *d_frame_len,
                                       unsigned int *d_frame_bs, void *d_summary,
                                       void *cuda_stream);
/* capacity of one device call: samples, output bytes, frames */
FLAKE_API int flake_b200_device_capacity(FlakeContext *s, unsigned long long *max_samples,
                                         unsigned long long *out_bytes, unsigned int *max_frames);

/* Per-subframe decisions of the most recent engine pass, for stage-level parity
 * tests (type, order, shift, coefficients, Rice parameters: what FlacSubframe
 * holds, encode.h:52-63).  `subs` is an array of `max` records of
 * flake_b200_subframe_record_size() bytes.  Returns the record count. */
FLAKE_API int flake_b200_last_subframes(FlakeContext *s, void *subs, unsigned int max);
FLAKE_API unsigned int flake_b200_subframe_record_size(void);

/*
 * Per-stage device timing of every pass of the context (flake_encode_frame, the streaming calls
 * and flake_b200_encode_device alike; the totals are sums over them), measured with CUDA events
 * recorded between the kernels on the launching stream.  Five stages:
 * 0 frame table (+VBS split), 1 prepare, 2 LPC analysis, 3 order/Rice search,
 * 4 pack (bits, CRCs, frame offsets, frames written back to back).
 * set_profiling(1) clears the totals; stage_times() synchronises the recorded
 * events and returns cumulative milliseconds and pass counts per stage (arrays of
 * FLAKE_B200_NUM_STAGES).
 */
#define FLAKE_B200_NUM_STAGES 5
FLAKE_API int flake_b200_set_profiling(FlakeContext *s, int on);
FLAKE_API int flake_b200_stage_times(FlakeContext *s, double *ms, unsigned long long *launches);

/*
 * SEEKTABLE from the per-frame lengths and block sizes that the batch calls return
 * (the reference writes none, so this is opt-in and not part of the
 * byte-identical stream).  One seek point for the first frame that starts at or after every
 * multiple of `interval_samples` (0: one per frame).  Each point is the FLAC triple
 * {sample number of the frame's first sample (u64 BE), byte offset from the first frame
 * (u64 BE), samples in the frame (u16 BE)} = 18 bytes; `data` receives the metadata block
 * BODY (the caller adds the 4-byte block header, type 3).  Returns the number of bytes
 * written, the number needed when data == NULL, or -1 when `cap` is too small.
 */
FLAKE_API long long flake_b200_write_seektable(const unsigned int *frame_len, const unsigned int *frame_bs,
                                               unsigned int nframes, unsigned int interval_samples,
                                               unsigned char *data, unsigned long long cap);

/*
 * Many streams over the GPUs of one box, from one process -- the batch form of running the
 * flake/flake.c:612-678 loop once per file of a corpus, and of SURVEY.md 8(e): a stream is cut
 * into chunks of consecutive blocks, any GPU of `devices` encodes any chunk (no exchange
 * between GPUs), and the host prefix-sums the chunk sizes into frame offsets.  One stream with
 * several devices is "one stream split by frame range over the GPUs".
 *
 * All streams share the format and parameters of `proto` (channels, sample_rate,
 * bits_per_sample, params: a FlakeContext as the caller would hand to flake_encode_init; it
 * need not be initialised and is not modified).  `pcm_format`: FLAKE_B200_PCM_S32, or the
 * packed layout whose container is ceil(bits_per_sample / 8) bytes (what a WAV data chunk
 * holds).  Buffers are HOST memory; page-locked buffers (cudaHostAlloc / cudaHostRegister)
 * are read and written by DMA directly, pageable ones go through a staging copy.
 *
 * Per stream the call produces exactly what flake_encode_init + the flake_encode_frame loop
 * would have: the frame bytes, their lengths and block sizes, STREAMINFO's maximum frame size
 * (seeded with the verbatim bound, encode.c:446-450, 967) and the MD5 of the PCM
 * (encode.c:1006), hashed by library threads that advance up to 32 streams per core at once.
 * flake_b200_corpus_stream_header() then writes the stream header with the final STREAMINFO.
 *
 * Returns 0, or the first negative stream result: -1 bad arguments, -2 a stream's `out` or
 * frame arrays too small, -3 CUDA failure / no device (FlakeB200CorpusStats.error says which).
 */
#define FLAKE_B200_MAX_DEVICES 16

typedef struct FlakeB200CorpusStream {
    /* in */
    const void *pcm;                    /* channel-interleaved samples, `pcm_format`            */
    unsigned long long nsamples;        /* inter-channel samples (< 2^32, STREAMINFO's field)   */
    unsigned char *out;                 /* frames, back to back                                  */
    unsigned long long out_cap;         /* >= flake_b200_max_encoded_size() is always enough     */
    unsigned int *frame_len, *frame_bs; /* optional, frame_cap entries each                      */
    unsigned int frame_cap;
    /* out */
    long long bytes;                    /* frame bytes written, or a negative error              */
    unsigned int nframes;
    unsigned int max_frame_size;
    unsigned int min_frame_size;        /* smallest frame (the reference's STREAMINFO keeps 0)   */
    unsigned int verbatim_frames;
    unsigned char md5sum[16];
} FlakeB200CorpusStream;

typedef struct FlakeB200CorpusOptions {
    int threads_per_device;             /* GPU worker threads per device, 0 = default (2)        */
    int md5_threads;                    /* most MD5 workers; 0 = online CPUs minus the GPU workers */
    int chunk_blocks;                   /* blocks per engine pass, 0 = default                   */
} FlakeB200CorpusOptions;

typedef struct FlakeB200CorpusStats {
    double wall_ms;
    double md5_ms;                      /* the MD5 worker that finished last                     */
    unsigned long long streams, samples, bytes, chunks;
    unsigned long long h2d_bytes, d2h_bytes, kernel_launches;
    unsigned int chunk_blocks;
    int devices, gpu_threads, md5_threads, md5_lanes;
    double device_ms[FLAKE_B200_MAX_DEVICES];                 /* per entry of `devices`: worker wall time */
    unsigned long long device_samples[FLAKE_B200_MAX_DEVICES]; /* samples encoded by that device   */
    char error[256];
} FlakeB200CorpusStats;

FLAKE_API int flake_b200_encode_corpus(const FlakeContext *proto, int pcm_format,
                                       FlakeB200CorpusStream *streams, unsigned int nstreams,
                                       const int *devices, int ndevices,      /* NULL / 0: every device */
                                       const FlakeB200CorpusOptions *options, /* NULL: defaults */
                                       FlakeB200CorpusStats *stats);          /* optional */
/* The same with a handle that keeps the per-device engines, lanes and streams between calls
 * (a corpus handed over in several batches pays the set-up once).  proto->samples, when not 0,
 * announces the longest stream and bounds the chunk size. */
typedef struct FlakeB200Corpus FlakeB200Corpus;
FLAKE_API FlakeB200Corpus *flake_b200_corpus_open(const FlakeContext *proto, int pcm_format,
                                                  const int *devices, int ndevices,
                                                  const FlakeB200CorpusOptions *options);
FLAKE_API int flake_b200_corpus_encode(FlakeB200Corpus *corpus, FlakeB200CorpusStream *streams,
                                       unsigned int nstreams, FlakeB200CorpusStats *stats);
FLAKE_API void flake_b200_corpus_close(FlakeB200Corpus *corpus);
FLAKE_API const char *flake_b200_corpus_error(const FlakeB200Corpus *corpus);
/* Stream header ("fLaC", final STREAMINFO, vendor comment, padding: what flake_encode_init
 * leaves in FlakeContext.header, encode.c:125-156, after the CLI's final STREAMINFO rewrite,
 * flake/flake.c:665-673) of one encoded stream of the corpus.  Returns the header length, the
 * length needed when data == NULL, or -1. */
FLAKE_API int flake_b200_corpus_stream_header(const FlakeContext *proto, const FlakeB200CorpusStream *stream,
                                              unsigned char *data, unsigned int cap);

/* STREAMINFO's minimum frame size: the reference always writes 0 = unknown (metadata.c:52), and
 * so does this library unless switched on here; then flake_get_streaminfo reports the smallest
 * frame encoded so far.  Not part of the byte-identical stream. */
FLAKE_API int flake_b200_set_streaminfo_sizes(FlakeContext *s, int on);

FLAKE_API int flake_b200_get_stats(const FlakeContext *s, FlakeB200Stats *stats);
FLAKE_API const char *flake_b200_last_error(const FlakeContext *s);
FLAKE_API const char *flake_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif
