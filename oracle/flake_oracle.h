/*
 * flake_oracle.h -- CPU restatement of Flake's FLAC encoding hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the shipped library (flake_b200/)
 * includes, links or calls this.  Allowed users: tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference leg.
 *
 * Parity status: PINNED.  The restatement is checked byte-for-byte against the
 * reference compiled from /root/reference (oracle/_ref/libflake_ref.so, see
 * oracle/Makefile) by tests/test_oracle_vs_ref.py and against the committed
 * golden vectors in tests/golden/ (generated from the compiled reference by
 * tests/golden/make_golden.py).
 *
 * Every function cites the reference file:line it restates (paths relative to
 * the reference tree root).
 */
#ifndef FLAKE_ORACLE_H
#define FLAKE_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_CH        8
#define ORC_MAX_ORDER     32
#define ORC_MAX_PARTS     256

typedef struct OrcParams {
    int channels;
    int sample_rate;
    int bps;
    int block_size;
    int order_method;          /* 0 MAX 1 EST 2 2LEVEL 3 4LEVEL 4 8LEVEL 5 SEARCH 6 LOG */
    int stereo_method;         /* 0 independent 1 estimate */
    int prediction_type;       /* 0 none 1 fixed 2 levinson */
    int min_order, max_order;
    int min_porder, max_porder;
    int variable_block_size;
    int allow_vbs;
    int padding_size;
    uint32_t total_samples;    /* what the caller put in FlakeContext.samples */
} OrcParams;

/* level presets: libflake/encode.c:158-266 */
int  orc_set_defaults(OrcParams *p, int level);
/* -1 invalid, 0 ok, 1 ok but non-Subset: libflake/encode.c:268-373 */
int  orc_validate(const OrcParams *p);

/* per-subframe decisions, exposed so that stage-level tests can compare them */
typedef struct OrcSubframe {
    int type;                  /* 0 constant 1 verbatim 8 fixed 32 lpc */
    int order;
    int obits;
    int wasted;
    int shift;
    int32_t coefs[ORC_MAX_ORDER];
    int method;                /* 0 RICE 1 RICE2 */
    int porder;
    int params[ORC_MAX_PARTS];
    uint32_t est_bits;         /* value returned by encode_residual() */
} OrcSubframe;

typedef struct OrcFrameInfo {
    int blocksize;
    int ch_mode;               /* 0 not stereo, 1 LR, 8 LS, 9 RS, 10 MS */
    int verbatim_fallback;     /* 1 if the size check forced VERBATIM */
    int nbytes;
    OrcSubframe sub[ORC_MAX_CH];
} OrcFrameInfo;

/*
 * Encode ONE frame (no VBS splitting) -- libflake/encode.c:919-977.
 * `number` is the value written with write_utf8 (frame index, or first sample
 * number when allow_vbs).  Returns the byte count or -1.
 */
int orc_encode_frame(const OrcParams *p, const int32_t *interleaved, int n,
                     uint32_t number, uint8_t *out, int out_cap,
                     OrcFrameInfo *info /* may be NULL */);

/* VBS split decision -- libflake/vbs.c:36-83.  Returns number of frames. */
int orc_vbs_split(const int32_t *interleaved, int channels, int block_size,
                  int sizes[8]);

/*
 * Encode a whole stream of `nsamples` inter-channel samples exactly like the
 * flake CLI loop does (flake/flake.c:612-663 calling flake_encode_frame,
 * libflake/encode.c:979-1008): block by block, VBS split when enabled.
 * Outputs the concatenated frames (NO stream header), per-frame byte lengths
 * and block sizes.  Returns total bytes or -1.
 */
int64_t orc_encode_stream(const OrcParams *p, const int32_t *interleaved,
                          uint64_t nsamples, uint8_t *out, size_t out_cap,
                          uint32_t *frame_len, uint32_t *frame_bs,
                          uint32_t frame_cap, uint32_t *nframes,
                          uint32_t *max_frame_size);

/* Stream header as flake_encode_init emits it (encode.c:125-156, 378-472) */
int orc_write_header(const OrcParams *p, uint8_t *hdr, int cap);
/* Final STREAMINFO body (34 bytes) -- metadata.c:32-84 */
void orc_streaminfo(const OrcParams *p, uint32_t max_frame_size,
                    const uint8_t md5[16], uint8_t out[34]);
/* initial max_frame_size (verbatim bound) -- encode.c:446-450 */
uint32_t orc_initial_max_frame_size(const OrcParams *p);

/* MD5 of PCM as md5_accumulate packs it -- md5.c:281-320 */
void orc_md5_pcm(const int32_t *interleaved, int channels, int bps,
                 uint64_t nsamples, uint8_t digest[16]);
/* the digest flake_encode_init writes into the provisional header:
 * md5_final on a zeroed (NOT md5_init'ed) context, encode.c:458-469 */
void orc_md5_zero_ctx(uint8_t digest[16]);

/* ---- stage-level entry points -------------------------------------- */
/* lpc.c:28-71 */
void orc_autocorr(const int32_t *smp, int n, int lag, double *autoc);
/* lpc.c:224-257; coefs is [32][32], shift[32]; returns order estimate */
int  orc_lpc_calc(const int32_t *smp, int n, int max_order, int omethod,
                  int32_t *coefs, int *shift);
/* rice.c:30-45 */
int  orc_rice_k(uint64_t sum, int n);
/* rice.c:157-187; returns estimated bits, fills method/porder/params */
uint32_t orc_rice_cost(const int32_t *res, int n, int pred_order, int obits,
                       int pmin, int pmax, int is_lpc,
                       int *method, int *porder, int *params);
/* optimize.c:34-122 */
void orc_residual_fixed(int32_t *res, const int32_t *smp, int n, int order);
void orc_residual_lpc(int32_t *res, const int32_t *smp, int n, int order,
                      const int32_t *coefs, int shift);
/* crc.c */
uint8_t  orc_crc8(const uint8_t *d, size_t n);
uint16_t orc_crc16(const uint8_t *d, size_t n);

/* ---- test-only FLAC decoder (flac_decode.c) ------------------------- */
typedef struct OrcDecInfo {
    int channels, bps, sample_rate;
    uint64_t total_samples;        /* from STREAMINFO (0 if no header) */
    uint64_t decoded_samples;
    uint32_t nframes;
    uint32_t min_bs, max_bs;
    uint32_t max_frame_bytes;
    int md5_ok;                    /* 1 match, 0 mismatch, -1 no header */
    int error;                     /* 0 ok, else code; see flac_decode.c */
    uint64_t error_pos;
} OrcDecInfo;

/*
 * Decode frames.  If has_header, parses "fLaC" + metadata first; otherwise
 * channels/bps must be supplied in info (sample_rate optional).
 * pcm_out receives interleaved int32 (cap in inter-channel samples).
 * Returns decoded inter-channel samples or -1.
 */
int64_t orc_flac_decode(const uint8_t *data, size_t len, int has_header,
                        int32_t *pcm_out, uint64_t pcm_cap, OrcDecInfo *info);

#ifdef __cplusplus
}
#endif
#endif
