/*
 * flake.h -- public C ABI of libflake as provided by flake_b200.
 *
 * This header is ABI-identical to the reference's libflake/flake.h (struct
 * layouts at flake.h:59-211, 239-249, 264-268; enums at flake.h:38-57; exported
 * functions at flake.h:217-234, 251-255, 274-295) so that the reference's own
 * callers -- flake/flake.c and util/api_example.c -- compile and link against
 * flake_b200's libflake.so unchanged.  It was written from that interface, not
 * copied; comments describe the behaviour of THIS implementation.
 *
 * The hot path behind flake_encode_frame runs on an NVIDIA B200 (sm_100a).
 * There is no CPU fallback: flake_encode_init fails (-1) when no CUDA device
 * is usable.  Batch / device-resident entry points live in flake_b200.h.
 */
#ifndef FLAKE_H
#define FLAKE_H

#if defined(FLAKE_BUILD_LIBRARY)
#  define FLAKE_API __attribute__((visibility("default")))
#else
#  define FLAKE_API extern
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* flake.h:38-46 -- how the LPC order of a subframe is picked (optimize.c:196-264) */
typedef enum {
    FLAKE_ORDER_METHOD_MAX,      /* always max_prediction_order                     */
    FLAKE_ORDER_METHOD_EST,      /* Schur reflection-coefficient estimate           */
    FLAKE_ORDER_METHOD_2LEVEL,   /* cost 2 evenly spread orders                     */
    FLAKE_ORDER_METHOD_4LEVEL,   /* ... 4                                           */
    FLAKE_ORDER_METHOD_8LEVEL,   /* ... 8                                           */
    FLAKE_ORDER_METHOD_SEARCH,   /* cost every order 1..max                         */
    FLAKE_ORDER_METHOD_LOG       /* coarse-to-fine search, steps 16,8,4,2,1         */
} FlakeOrderMethod;

/* flake.h:48-51 */
typedef enum {
    FLAKE_STEREO_METHOD_INDEPENDENT,
    FLAKE_STEREO_METHOD_ESTIMATE
} FlakeStereoMethod;

/* flake.h:53-57 */
typedef enum {
    FLAKE_PREDICTION_NONE,
    FLAKE_PREDICTION_FIXED,
    FLAKE_PREDICTION_LEVINSON
} FlakePrediction;

/* flake.h:59-161 -- twelve ints, in this order */
typedef struct FlakeEncodeParams {
    int compression;            /* preset 0..12; see flake_set_defaults                  */
    int order_method;           /* FlakeOrderMethod, 0..6                                */
    int stereo_method;          /* FlakeStereoMethod, 0..1                               */
    int block_size;             /* samples per block, 16..65535                          */
    int padding_size;           /* bytes of PADDING metadata after the Vorbis comment    */
    int min_prediction_order;   /* 0..4 fixed, 1..32 LPC                                 */
    int max_prediction_order;
    int prediction_type;        /* FlakePrediction                                       */
    int min_partition_order;    /* 0..8                                                  */
    int max_partition_order;
    int variable_block_size;    /* 1: library splits each block into up to 8 frames      */
    int allow_vbs;              /* 1: caller may pass blocks of varying size; frame
                                   headers then carry sample numbers                     */
} FlakeEncodeParams;

/* flake.h:163-211 */
typedef struct FlakeContext {
    int channels;               /* 1..8, set by the caller before flake_encode_init      */
    int sample_rate;            /* Hz                                                    */
    int bits_per_sample;        /* 4..32; parity with the reference is tested for 8..24  */
    unsigned int samples;       /* total inter-channel samples, 0 = unknown              */
    FlakeEncodeParams params;   /* copied into the private context at init               */
    unsigned char *header;      /* stream header bytes, owned by the library             */
    void *private_ctx;          /* opaque                                                */
} FlakeContext;

/* Fill *params from params->compression (encode.c:158-266).  0 or -1. */
FLAKE_API int flake_set_defaults(FlakeEncodeParams *params);

/* -1 invalid, 0 valid, 1 valid but outside the FLAC Subset (encode.c:268-373). */
FLAKE_API int flake_validate_params(const FlakeContext *s);

/* Allocates the private context, the header and the GPU engine.  Returns the
 * number of header bytes in s->header (write them first) or -1. */
FLAKE_API int flake_encode_init(FlakeContext *s);

/* Library-owned output buffer that flake_encode_frame fills; stable until close. */
FLAKE_API void *flake_get_buffer(const FlakeContext *s);

/* Encode one block of `block_size` channel-interleaved, sign-extended int32
 * samples.  Returns the bytes placed in the buffer (one frame, or several
 * back-to-back under variable_block_size) or -1.  Synchronous. */
FLAKE_API int flake_encode_frame(FlakeContext *s, const int *samples, int block_size);

FLAKE_API void flake_encode_close(FlakeContext *s);

/* "SVN", as the reference built without SVN_VERSION (encode.c:1028-1038). */
FLAKE_API const char *flake_get_version(void);

/* flake.h:239-249 */
typedef struct FlakeStreaminfo {
    unsigned int min_block_size;
    unsigned int max_block_size;
    unsigned int min_frame_size;
    unsigned int max_frame_size;
    unsigned int sample_rate;
    unsigned int channels;
    unsigned int bits_per_sample;
    unsigned int samples;
    unsigned char md5sum[16];
} FlakeStreaminfo;

/* Snapshot of the stream so far: running max frame size and the MD5 of all PCM
 * encoded up to now (metadata.c:32-65).  Does not disturb the running digest. */
FLAKE_API int flake_get_streaminfo(const FlakeContext *s, FlakeStreaminfo *strminfo);

/* Serialise to the 34-byte STREAMINFO body (metadata.c:67-84). */
FLAKE_API void flake_write_streaminfo(const FlakeStreaminfo *strminfo, unsigned char *data);

/* flake.h:264-268 */
typedef struct FlakeVorbisComment {
    char *vendor_string;
    unsigned int num_entries;
    char *entries[1024];
} FlakeVorbisComment;

FLAKE_API void flake_init_vorbiscomment(FlakeVorbisComment *vc);
FLAKE_API int  flake_add_vorbiscomment_entry(FlakeVorbisComment *vc, char *entry);
FLAKE_API int  flake_get_vorbiscomment_size(const FlakeVorbisComment *vc);
FLAKE_API int  flake_write_vorbiscomment(const FlakeVorbisComment *vc, unsigned char *data);

#ifdef __cplusplus
}
#endif
#endif /* FLAKE_H */
