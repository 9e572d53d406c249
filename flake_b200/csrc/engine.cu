/*
 * engine.cu -- host side of the CUDA engine: scratch management and the launch
 * sequence of one chunk.  Exposes the plain C ABI of engine.h.
 *
 * Launch sequence per chunk (all on one stream, no host synchronisation):
 *   k_frames_fixed | (k_vbs_split + k_frames_vbs)   frame table
 *   k_prep      per frame      ingest, stereo decision, wasted bits, constant test
 *   k_lpc       per subframe   FP64 window/autocorrelation/Levinson/quantise
 *   k_search    per subframe   order + Rice search, final residual
 *   k_pack      per frame      bit packing, CRC-8/16, verbatim size check, offsets by a
 *                              decoupled look-back over the frame lengths, frames written
 *                              back to back into the output, chunk summary
 */
#include "cuda_compat.h"
#include "engine.h"
#include "dev_common.cuh"
#include "k_prep.cuh"
#include "k_lpc.cuh"
#include "k_search.cuh"
#include "k_pack.cuh"

#include <new>
#include <map>
#include <tuple>
#include <vector>

#define FB_SMEM_BUDGET (200 * 1024)
#define FB_TIMING_RING 64

struct FbEngine {
    FbConfig cfg;
    int device;
    uint32_t max_blocks, max_frames, max_subs;
    uint64_t max_samples;
    uint64_t slot_bytes, out_bytes;
    cudaStream_t stream;
    /* scratch */
    FbFrame *d_frames;
    uint32_t *d_nframes;            /* [0] frame count, [2] k_pack's ticket counter */
    FbSub *d_subs;
    uint8_t *d_modes;
    int32_t *d_smp, *d_res, *d_coefs, *d_shifts;
    uint8_t *d_slots;
    uint32_t *d_frame_len;
    uint64_t *d_frame_off;
    uint32_t *d_vbs_sizes, *d_vbs_counts;
    uint16_t *d_xpow32;             /* x^(32 j) mod P, CRC-16 chunk merge (k_pack) */
    uint32_t *d_crc16tab;           /* CRC-16 slicing tables [4][256] uint16 (k_pack) */
    unsigned long long *d_status;   /* look-back status per frame (k_pack) */
    FbPlanNode *d_plan;             /* log-search decision tree (k_search), NULL for the other order methods */
    int search_smem_ints, pack_smem_words;
    uint64_t launches;
    /* optional per-kernel CUDA-event timing (bench.py roofline) */
    int timing_on, timing_passes;
    cudaEvent_t tev[FB_TIMING_RING][FB_NUM_STAGES + 1];
    double stage_ms[FB_NUM_STAGES];
    uint64_t stage_launches[FB_NUM_STAGES];
    cudaEvent_t ev_pass;            /* end of the most recent pass: the scratch buffers are shared between passes */
    cudaStream_t last_stream;       /* stream that pass ran on */
    int have_pass;
    uint32_t lpc_lat_max;           /* passes with no more subframes run k_lpc_lat (FLAKE_B200_LPC_LAT_MAX overrides) */
    int sync_launches;              /* FLAKE_B200_SYNC_LAUNCHES: name the kernel an asynchronous fault comes from */
    char err[256];
};

/* The decision tree of the log search for orders min_order..max_order with `group` candidates
 * costed at a time -- see FbPlanNode (engine.h).  Steps are merged into one node while the
 * candidate set of the next step is the same for every order that could be the best by then. */
static uint16_t fb_plan_node(std::vector<FbPlanNode> &nodes, std::map<std::tuple<int, uint32_t, int>, uint16_t> &index,
                             int lo, int hi, int group, int step, uint32_t done, int opt)
{
    if (step == 0) return (uint16_t)FB_PLAN_END;
    const auto key = std::make_tuple(step, done, opt);
    const auto it = index.find(key);
    if (it != index.end()) return it->second;
    const uint16_t idx = (uint16_t)nodes.size();
    index[key] = idx;
    nodes.push_back(FbPlanNode());
    const uint32_t range = (hi >= 31 ? 0xffffffffu : ((1u << (hi + 1)) - 1u)) & ~((1u << lo) - 1u);
    uint32_t ord = 0, gmask = 0, hyp = 1u << opt;
    int cnt = 0, nsteps = 0, members[FB_PLAN_CHILDREN];
    for (int s = step; s > 0; s >>= 1) {
        uint32_t cor = 0, cand = 0xffffffffu;
        for (int h = 0; h < 32; h++) {
            if (!((hyp >> h) & 1u)) continue;
            uint32_t cm = 1u << h;
            if (h - s >= 0) cm |= 1u << (h - s);
            if (h + s < 32) cm |= 1u << (h + s);
            cm &= range & ~(done | gmask);
            cor |= cm; cand &= cm;
        }
        if (cor != cand || cnt + __builtin_popcount(cor) > group) break;
        for (int b = 0; b < 32; b++)
            if ((cor >> b) & 1u) { ord |= (uint32_t)(b + 1) << (8 * cnt); members[cnt++] = b; }
        gmask |= cor; hyp |= cor; nsteps++;
    }
    FbPlanNode nd;
    memset(&nd, 0, sizeof nd);
    nd.ord = ord; nd.cnt = (uint16_t)cnt; nd.nsteps = (uint16_t)nsteps;
    for (int i = 0; i < FB_PLAN_CHILDREN; i++) nd.child[i] = (uint16_t)FB_PLAN_END;
    nd.child[0] = fb_plan_node(nodes, index, lo, hi, group, step >> nsteps, done | gmask, opt);
    for (int i = 0; i < cnt; i++)
        nd.child[1 + i] = fb_plan_node(nodes, index, lo, hi, group, step >> nsteps, done | gmask, members[i]);
    nodes[idx] = nd;
    return idx;
}

static std::vector<FbPlanNode> fb_build_log_plan(int min_order, int max_order, int group)
{
    std::vector<FbPlanNode> nodes;
    std::map<std::tuple<int, uint32_t, int>, uint16_t> index;
    const int lo = min_order - 1, hi = max_order - 1;
    const int start = lo + (max_order - min_order) / 3;
    fb_plan_node(nodes, index, lo, hi, group, 16, 0u, start);
    nodes[0].start_order = (uint16_t)start;
    return nodes;
}

static void set_err(char *dst, size_t n, const char *msg, cudaError_t ce)
{
    if (!dst || !n) return;
    if (ce != cudaSuccess) snprintf(dst, n, "%s: %s", msg, cudaGetErrorString(ce));
    else snprintf(dst, n, "%s", msg);
}

#define FB_TRY_ALLOC(ptr, bytes)                                                   \
    do {                                                                           \
        cudaError_t ce_ = cudaMalloc((void **)&(ptr), (bytes));                    \
        if (ce_ != cudaSuccess) { set_err(err, errlen, "cudaMalloc failed", ce_);  \
                                  fb_engine_destroy(e); return nullptr; }          \
    } while (0)

extern "C" FbEngine *fb_engine_create(const FbConfig *cfg, int device, uint32_t max_blocks,
                                      char *err, size_t errlen)
{
    if (!cfg || max_blocks == 0) { set_err(err, errlen, "bad arguments", cudaSuccess); return nullptr; }
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev < 1) {
        set_err(err, errlen, "no CUDA device (flake_b200 has no CPU fallback)", ce);
        return nullptr;
    }
    if (device >= 0) {
        ce = cudaSetDevice(device);
        if (ce != cudaSuccess) { set_err(err, errlen, "cudaSetDevice failed", ce); return nullptr; }
    } else {
        cudaGetDevice(&device);
    }
    FbEngine *e = new (std::nothrow) FbEngine();
    if (!e) { set_err(err, errlen, "out of host memory", cudaSuccess); return nullptr; }
    memset(e, 0, sizeof *e);
    e->cfg = *cfg;
    e->device = device;
    const int C = cfg->channels, B = cfg->block_size;
    e->max_blocks = max_blocks;
    e->max_frames = cfg->variable_block_size ? max_blocks * 8u : max_blocks;
    e->max_subs = e->max_frames * (uint32_t)C;
    e->max_samples = (uint64_t)max_blocks * (uint64_t)B;
    if (e->max_samples * (uint64_t)(C * cfg->bps + 1) / 8u + (uint64_t)e->max_frames * 96u > 0xfff00000ull) {
        set_err(err, errlen, "chunk too large for 32-bit slot offsets", cudaSuccess);
        delete e; return nullptr;
    }
    e->slot_bytes = (uint64_t)fb_slot_offset(e->max_frames, (uint32_t)e->max_samples, C, cfg->bps) + 256u;
    e->out_bytes = e->slot_bytes;

    if (cudaEventCreateWithFlags(&e->ev_pass, cudaEventDisableTiming) != cudaSuccess) e->ev_pass = nullptr;
    { const char *lm = getenv("FLAKE_B200_LPC_LAT_MAX"); e->lpc_lat_max = lm && *lm ? (uint32_t)atoi(lm) : FB_LPC_LAT_MAX_SUBFRAMES; }
    { const char *sl = getenv("FLAKE_B200_SYNC_LAUNCHES"); e->sync_launches = sl && *sl == '1'; }
    ce = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
    if (ce != cudaSuccess) { set_err(err, errlen, "cudaStreamCreate failed", ce); delete e; return nullptr; }

    const uint64_t nint = e->max_samples * (uint64_t)C;
    FB_TRY_ALLOC(e->d_frames, sizeof(FbFrame) * (size_t)e->max_frames);
    FB_TRY_ALLOC(e->d_nframes, sizeof(uint32_t) * 4);
    FB_TRY_ALLOC(e->d_subs, sizeof(FbSub) * (size_t)e->max_subs);
    FB_TRY_ALLOC(e->d_modes, (size_t)e->max_frames);
    FB_TRY_ALLOC(e->d_res, sizeof(int32_t) * nint);
    FB_TRY_ALLOC(e->d_slots, e->slot_bytes);
    FB_TRY_ALLOC(e->d_frame_len, sizeof(uint32_t) * (size_t)e->max_frames);
    FB_TRY_ALLOC(e->d_frame_off, sizeof(uint64_t) * (size_t)e->max_frames);
    FB_TRY_ALLOC(e->d_status, sizeof(unsigned long long) * (size_t)e->max_frames);
    if (cfg->prediction_type == 2) {
        FB_TRY_ALLOC(e->d_coefs, sizeof(int32_t) * FB_MAX_ORDER * FB_MAX_ORDER * (size_t)e->max_subs);
        FB_TRY_ALLOC(e->d_shifts, sizeof(int32_t) * FB_MAX_ORDER * (size_t)e->max_subs);
    }
    if (cfg->variable_block_size) {
        FB_TRY_ALLOC(e->d_vbs_sizes, sizeof(uint32_t) * 8 * (size_t)max_blocks);
        FB_TRY_ALLOC(e->d_vbs_counts, sizeof(uint32_t) * (size_t)max_blocks);
    }

    /* shared-memory staging where a whole block fits */
    {
        /* the kernel this configuration launches (dispatch in fb_engine_encode_device) */
        const int group = (cfg->prediction_type == 2 && cfg->max_order > 12) ? FB_GROUP_OF(32) : FB_GROUP_OF(12);
        const int words = fb_search_smem_words(B, group);
        e->search_smem_ints = ((size_t)words * 4 <= FB_SMEM_BUDGET) ? words : 0;
        /* int32 planes exist in global memory only for more than two channels (k_prep deinterleaves
         * once, fb_uses_planes) and as k_search's scratch for blocks that do not fit shared memory */
        if (!e->search_smem_ints || fb_uses_planes(C)) FB_TRY_ALLOC(e->d_smp, sizeof(int32_t) * nint);
    }
    {
        const uint64_t capb = 64u + (((uint64_t)B * (uint64_t)(C * cfg->bps + 1) + 7u) >> 3);
        const uint64_t capw = (capb + 3u) >> 2;
        e->pack_smem_words = (capw * 4u <= FB_SMEM_BUDGET) ? (int)capw : 0;
    }
    {
        /* x^(32 j) mod (x^16 + x^15 + x^2 + 1): j zero words appended to a CRC (crc.c:24-46) */
        const uint64_t capb = 64u + (((uint64_t)B * (uint64_t)(C * cfg->bps + 1) + 7u) >> 3);
        const size_t nent = (size_t)((capb + 3u) >> 2) + 1024 + 64;   /* per * (threads - 1) < words + threads */
        uint16_t *tab = (uint16_t *)malloc(nent * sizeof(uint16_t));
        if (!tab) { set_err(err, errlen, "out of host memory", cudaSuccess); fb_engine_destroy(e); return nullptr; }
        uint32_t r = 1;
        for (size_t j = 0; j < nent; j++) {
            tab[j] = (uint16_t)r;
            for (int b = 0; b < 32; b++) r = (r & 0x8000u) ? ((r << 1) ^ 0x8005u) & 0xffffu : (r << 1) & 0xffffu;
        }
        cudaError_t ce_ = cudaMalloc((void **)&e->d_xpow32, nent * sizeof(uint16_t));
        if (ce_ == cudaSuccess) ce_ = cudaMemcpy(e->d_xpow32, tab, nent * sizeof(uint16_t), cudaMemcpyHostToDevice);
        free(tab);
        if (ce_ != cudaSuccess) { set_err(err, errlen, "CRC table upload failed", ce_); fb_engine_destroy(e); return nullptr; }
    }
    {
        /* slicing tables of crc.c's CRC-16 (poly 0x8005, init 0, MSB first):
         * tab[t][b] = CRC of byte b followed by t zero bytes */
        uint16_t tab[4][256];
        for (int b = 0; b < 256; b++) {
            uint32_t c = 0;
            for (int t = 0; t < 4; t++) {
                c ^= (t == 0 ? (uint32_t)b : 0u) << 8;
                for (int i = 0; i < 8; i++) c = (c & 0x8000u) ? ((c << 1) ^ 0x8005u) & 0xffffu : (c << 1) & 0xffffu;
                tab[t][b] = (uint16_t)c;
            }
        }
        cudaError_t ce_ = cudaMalloc((void **)&e->d_crc16tab, sizeof tab);
        if (ce_ == cudaSuccess) ce_ = cudaMemcpy(e->d_crc16tab, tab, sizeof tab, cudaMemcpyHostToDevice);
        if (ce_ != cudaSuccess) { set_err(err, errlen, "CRC table upload failed", ce_); fb_engine_destroy(e); return nullptr; }
    }
    if (cfg->prediction_type == 2 && cfg->order_method == 6) {
        const int group = cfg->max_order > 12 ? FB_GROUP_OF(32) : FB_GROUP_OF(12);
        std::vector<FbPlanNode> plan = fb_build_log_plan(cfg->min_order, cfg->max_order, group);
        if (plan.size() < FB_PLAN_SMEM_NODES) plan.resize(FB_PLAN_SMEM_NODES);      /* k_search stages that many */
        cudaError_t ce_ = cudaMalloc((void **)&e->d_plan, plan.size() * sizeof(FbPlanNode));
        if (ce_ == cudaSuccess) ce_ = cudaMemcpy(e->d_plan, plan.data(), plan.size() * sizeof(FbPlanNode), cudaMemcpyHostToDevice);
        if (ce_ != cudaSuccess) { set_err(err, errlen, "search plan upload failed", ce_); fb_engine_destroy(e); return nullptr; }
    }
    /* The attribute is per function, per device and per PROCESS, not per engine: every engine sets
     * the same constant (the staging budget), so a later engine with smaller blocks can never
     * lower it under a live engine's launch size. */
    cudaFuncSetAttribute(k_lpc_lat<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM_BUDGET);
    cudaFuncSetAttribute(k_lpc_lat<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM_BUDGET);
    cudaFuncSetAttribute(k_lpc_lat<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM_BUDGET);
    cudaFuncSetAttribute(k_lpc_lat<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM_BUDGET);
    cudaFuncSetAttribute(k_lpc_lat<24>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM_BUDGET);
    cudaFuncSetAttribute(k_lpc_lat<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM_BUDGET);
    cudaFuncSetAttribute(k_search<12, FB_SEARCH_ANY>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM_BUDGET);
    cudaFuncSetAttribute(k_search<12, FB_SEARCH_CD>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM_BUDGET);
    cudaFuncSetAttribute(k_search<32, FB_SEARCH_ANY>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM_BUDGET);
    cudaFuncSetAttribute(k_pack, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM_BUDGET);
    ce = cudaGetLastError();
    if (ce != cudaSuccess) { set_err(err, errlen, "cudaFuncSetAttribute failed", ce); fb_engine_destroy(e); return nullptr; }
    return e;
}

extern "C" void fb_engine_destroy(FbEngine *e)
{
    if (!e) return;
    if (e->stream) cudaStreamSynchronize(e->stream);
    cudaFree(e->d_frames); cudaFree(e->d_nframes); cudaFree(e->d_subs); cudaFree(e->d_modes);
    cudaFree(e->d_smp); cudaFree(e->d_res); cudaFree(e->d_coefs); cudaFree(e->d_shifts);
    cudaFree(e->d_slots); cudaFree(e->d_frame_len); cudaFree(e->d_frame_off);
    if (e->ev_pass) cudaEventDestroy(e->ev_pass);
    cudaFree(e->d_vbs_sizes); cudaFree(e->d_vbs_counts); cudaFree(e->d_xpow32); cudaFree(e->d_crc16tab); cudaFree(e->d_status); cudaFree(e->d_plan);
    if (e->tev[0][0])
        for (int p = 0; p < FB_TIMING_RING; p++)
            for (int i = 0; i <= FB_NUM_STAGES; i++) cudaEventDestroy(e->tev[p][i]);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

extern "C" uint32_t fb_engine_max_blocks(const FbEngine *e) { return e ? e->max_blocks : 0; }
extern "C" uint32_t fb_engine_max_frames(const FbEngine *e) { return e ? e->max_frames : 0; }
extern "C" uint64_t fb_engine_out_capacity(const FbEngine *e) { return e ? e->out_bytes : 0; }
extern "C" uint64_t fb_engine_launch_count(const FbEngine *e) { return e ? e->launches : 0; }
extern "C" const char *fb_engine_last_error(const FbEngine *e) { return e ? e->err : "no engine"; }

/* After every launch / memset of a pass: a failure is reported with the name of the step that
 * caused it.  Launch-configuration errors show up here at once; a fault INSIDE a kernel is
 * asynchronous and surfaces at a later call -- FLAKE_B200_SYNC_LAUNCHES=1 synchronises after
 * every step so that the failing kernel is named (diagnosis only: it serialises the pass). */
static int fb_step_failed(FbEngine *e, cudaStream_t st, cudaError_t ce, const char *what)
{
    if (ce == cudaSuccess) ce = cudaGetLastError();
    if (ce == cudaSuccess && e->sync_launches) ce = cudaStreamSynchronize(st);
    if (ce == cudaSuccess) return 0;
    snprintf(e->err, sizeof e->err, "%s failed: %s", what, cudaGetErrorString(ce));
    return 1;
}

extern "C" int fb_engine_encode_device(FbEngine *e, const void *d_pcm, int fmt, uint64_t nsamples,
                                       uint32_t first_number, void *d_out, uint32_t *d_frame_len,
                                       uint32_t *d_frame_bs, FbSummary *d_summary, void *stream_v)
{
    if (!e || !d_pcm || !d_out || !d_summary) return -1;
    if (nsamples == 0 || nsamples > e->max_samples) {
        snprintf(e->err, sizeof e->err, "chunk of %llu samples exceeds capacity %llu",
                 (unsigned long long)nsamples, (unsigned long long)e->max_samples);
        return -2;
    }
    if (fmt < FB_PCM_S32 || fmt > FB_PCM_S8) return -3;
    const FbConfig cfg = e->cfg;
    cudaStream_t st = stream_v ? (cudaStream_t)stream_v : e->stream;
    const uint32_t ns = (uint32_t)nsamples;
    const uint32_t B = (uint32_t)cfg.block_size;
    const uint32_t nblocks = (ns + B - 1) / B;
    const uint32_t grid_frames = cfg.variable_block_size ? nblocks * 8u : nblocks;
    const uint32_t grid_subs = grid_frames * (uint32_t)cfg.channels;
    uint32_t *flen = d_frame_len ? d_frame_len : e->d_frame_len;

    cudaEvent_t *tev = nullptr;
    if (e->timing_on) {
        if (e->timing_passes >= FB_TIMING_RING) fb_engine_collect_timing(e, nullptr, nullptr);
        tev = e->tev[e->timing_passes++];
    }
#define FB_MARK(i) do { if (tev) cudaEventRecord(tev[(i)], st); } while (0)
    /* One engine = one set of scratch buffers (frame table, subframe records, residuals, look-back
     * status): passes are serial.  A pass issued on another stream than the previous one waits for
     * it, so callers may alternate streams without corrupting frames. */
    if (e->have_pass && e->ev_pass && st != e->last_stream && cudaStreamWaitEvent(st, e->ev_pass, 0) != cudaSuccess) {
        snprintf(e->err, sizeof e->err, "cudaStreamWaitEvent on the previous pass failed");
        return -4;
    }
#define FB_STEP(call_, what_) do { if (fb_step_failed(e, st, (call_), (what_))) return -4; } while (0)
#define FB_LAUNCHED(what_) FB_STEP(cudaSuccess, what_)
    /* frame count, -, pack ticket, -; look-back status; chunk summary.  With fixed blocks k_frames_fixed
     * clears them itself (three launches less per pass: 7 us of a one-block flake_encode_frame call) */
    if (cfg.variable_block_size) {
        FB_STEP(cudaMemsetAsync(e->d_nframes, 0, sizeof(uint32_t) * 4, st), "clearing the frame count");
        FB_STEP(cudaMemsetAsync(e->d_status, 0, sizeof(unsigned long long) * (size_t)grid_frames, st), "clearing the look-back status");
        FB_STEP(cudaMemsetAsync(d_summary, 0, sizeof(FbSummary), st), "clearing the chunk summary");
    }
    FB_MARK(0);
    if (cfg.variable_block_size) {
        FB_LAUNCH(k_vbs_split, dim3(nblocks), dim3(FB_PREP_THREADS), 0, st,
                  cfg, d_pcm, fmt, ns, e->d_vbs_sizes, e->d_vbs_counts);
        FB_LAUNCHED("k_vbs_split");
        FB_LAUNCH(k_frames_vbs, dim3(1), dim3(1024), 0, st,
                  cfg, ns, first_number, e->d_vbs_sizes, e->d_vbs_counts, e->d_frames, e->d_nframes);
        FB_LAUNCHED("k_frames_vbs");
        e->launches += 2;
    } else {
        FB_LAUNCH(k_frames_fixed, dim3((nblocks + 255) / 256), dim3(256), 0, st,
                  cfg, ns, first_number, e->d_frames, e->d_nframes, e->d_status, (FbSummary *)d_summary);
        FB_LAUNCHED("k_frames_fixed");
        e->launches += 1;
    }
    FB_MARK(1);
    const unsigned long long pcm_bytes = (unsigned long long)ns * (unsigned long long)cfg.channels *
                                         (unsigned long long)(fmt == FB_PCM_S16LE ? 2 : fmt == FB_PCM_S24LE ? 3 : fmt == FB_PCM_S8 ? 1 : 4);
    FB_LAUNCH(k_prep, dim3(grid_frames), dim3(FB_PREP_THREADS), 0, st,
              cfg, d_pcm, fmt, pcm_bytes, e->d_frames, e->d_nframes, e->d_smp, e->d_subs, e->d_modes);
    FB_LAUNCHED("k_prep");
    e->launches += 1;
    FB_MARK(2);
    if (cfg.prediction_type == 2) {
        /* a lane pair per subframe; the register ring is sized by the template (lags 0..ML) */
        const dim3 lpc_grid((grid_subs + FB_LPC_SUBS_PER_CTA - 1) / FB_LPC_SUBS_PER_CTA);
        /* stereo 16/24-bit layouts have a loader of their own (one load per sample pair, branch
         * free); specialised for the orders the presets use */
        int lk = FB_LPC_LK_GENERIC;
        if (cfg.channels == 2 && fmt == FB_PCM_S16LE && (((size_t)d_pcm) & 3u) == 0) lk = FB_LPC_LK_S16_STEREO;
        if (cfg.channels == 2 && fmt == FB_PCM_S24LE && (((size_t)d_pcm) & 1u) == 0) lk = FB_LPC_LK_S24_STEREO;
        if (fb_uses_planes(cfg.channels)) lk = FB_LPC_LK_PLANES;
#define FB_LPC_GO2(ML_, LK_)                                                                      \
        FB_LAUNCH((k_lpc<ML_, LK_>), lpc_grid, dim3(FB_LPC_THREADS), 0, st,                       \
                  cfg, e->d_frames, e->d_nframes, d_pcm, fmt, e->d_smp, e->d_modes, e->d_subs, e->d_coefs, e->d_shifts)
#define FB_LPC_GO(ML_)                                                                            \
        do {                                                                                      \
            if (lk == FB_LPC_LK_PLANES) FB_LPC_GO2(ML_, FB_LPC_LK_PLANES);                        \
            else FB_LPC_GO2(ML_, FB_LPC_LK_GENERIC);                                              \
        } while (0)
#define FB_LPC_GO3(ML_)                                                                           \
        do {                                                                                      \
            if (lk == FB_LPC_LK_S16_STEREO) FB_LPC_GO2(ML_, FB_LPC_LK_S16_STEREO);                \
            else if (lk == FB_LPC_LK_S24_STEREO) FB_LPC_GO2(ML_, FB_LPC_LK_S24_STEREO);           \
            else FB_LPC_GO(ML_);                                                                  \
        } while (0)
        /* a handful of blocks (flake_encode_frame: one): the latency form, a warp per subframe */
        const size_t lat_smem = ((size_t)B + 2u) * sizeof(double);
        const bool lat = grid_subs <= e->lpc_lat_max && lat_smem <= FB_SMEM_BUDGET;
#define FB_LPC_LAT(ML_)                                                                           \
        FB_LAUNCH((k_lpc_lat<ML_>), dim3(grid_subs), dim3(32), lat_smem, st,                      \
                  cfg, e->d_frames, e->d_nframes, d_pcm, fmt, e->d_smp, e->d_modes, e->d_subs, e->d_coefs, e->d_shifts)
        if (lat) {
            if (cfg.max_order <= 4) FB_LPC_LAT(4);
            else if (cfg.max_order <= 8) FB_LPC_LAT(8);
            else if (cfg.max_order <= 12) FB_LPC_LAT(12);
            else if (cfg.max_order <= 16) FB_LPC_LAT(16);
            else if (cfg.max_order <= 24) FB_LPC_LAT(24);
            else FB_LPC_LAT(32);
        }
        else if (cfg.max_order <= 4) FB_LPC_GO(4);
        else if (cfg.max_order <= 8) FB_LPC_GO3(8);
        else if (cfg.max_order <= 12) FB_LPC_GO3(12);
        else if (cfg.max_order <= 16) FB_LPC_GO(16);
        else if (cfg.max_order <= 24) FB_LPC_GO(24);
        else FB_LPC_GO3(32);
#undef FB_LPC_LAT
#undef FB_LPC_GO3
#undef FB_LPC_GO2
#undef FB_LPC_GO
        FB_LAUNCHED("k_lpc");
        e->launches += 1;
    }
    FB_MARK(3);
    /* the CD-audio shape of presets 8, 9 (and 11 with longer predictors) has an instantiation of
     * its own: less code to fetch (k_search.cuh, FB_SEARCH_CD) */
    const bool cd_shape = cfg.channels == 2 && fmt == FB_PCM_S16LE && cfg.prediction_type == 2 && cfg.order_method == 6;
    const char *search_name;
#define FB_SEARCH_GO(MAXP_, SPEC_)                                                                              \
        FB_LAUNCH((k_search<MAXP_, SPEC_>), dim3(grid_subs), dim3(FB_SEARCH_THREADS), (size_t)e->search_smem_ints * 4, st, \
                  cfg, e->d_frames, e->d_nframes, d_pcm, fmt, pcm_bytes, e->d_modes, e->d_smp, e->d_res, e->d_subs,     \
                  e->d_coefs, e->d_shifts, e->d_plan, e->search_smem_ints)
    if (cfg.prediction_type == 2 && cfg.max_order > 12) { search_name = "k_search<32>"; FB_SEARCH_GO(32, FB_SEARCH_ANY); }
    else if (cd_shape) { search_name = "k_search<12, CD>"; FB_SEARCH_GO(12, FB_SEARCH_CD); }
    else { search_name = "k_search<12>"; FB_SEARCH_GO(12, FB_SEARCH_ANY); }
#undef FB_SEARCH_GO
    FB_LAUNCHED(search_name);
    FB_MARK(4);
    /* threads per frame by the work in it (measured: 8192 samples per frame 1.75 ms with 128 threads vs
     * 1.88 with 256 per C2 stream; 32768 samples per frame 2.79 vs 2.08) */
    /* a pass of a few blocks (flake_encode_frame: one) is a latency chain: shorter runs per thread */
    const int pack_threads = ((uint64_t)B * (uint64_t)cfg.channels >= 16384u || nblocks <= 32u) ? FB_PACK_THREADS : FB_PACK_THREADS / 2;
    FB_LAUNCH(k_pack, dim3(grid_frames), dim3(pack_threads), (size_t)e->pack_smem_words * 4, st,
              cfg, e->d_frames, e->d_nframes, d_pcm, fmt, e->d_smp, e->d_res, e->d_subs, e->d_modes, e->d_slots,
              flen, d_frame_bs, e->pack_smem_words, e->d_xpow32, e->d_crc16tab,
              e->d_nframes + 2, e->d_status, e->d_frame_off, (uint8_t *)d_out, d_summary);
    FB_LAUNCHED("k_pack");
    FB_MARK(5);
#undef FB_MARK
#undef FB_LAUNCHED
#undef FB_STEP
    e->launches += 2;
    if (e->ev_pass && cudaEventRecord(e->ev_pass, st) == cudaSuccess) { e->last_stream = st; e->have_pass = 1; }
    return 0;
}

extern "C" int fb_engine_set_timing(FbEngine *e, int on)
{
    if (!e) return -1;
    if (on && !e->tev[0][0]) {
        for (int p = 0; p < FB_TIMING_RING; p++)
            for (int i = 0; i <= FB_NUM_STAGES; i++)
                if (cudaEventCreate(&e->tev[p][i]) != cudaSuccess) return -2;
    }
    e->timing_on = on ? 1 : 0;
    return 0;
}

/* Fold the recorded passes into the per-stage totals; optionally copy them out.
 * Stages: 0 frame table (+VBS split), 1 prep, 2 lpc, 3 search, 4 pack. */
extern "C" int fb_engine_collect_timing(FbEngine *e, double *ms, uint64_t *launches)
{
    if (!e) return -1;
    for (int p = 0; p < e->timing_passes; p++) {
        cudaEventSynchronize(e->tev[p][FB_NUM_STAGES]);
        for (int i = 0; i < FB_NUM_STAGES; i++) {
            float t = 0.f;
            if (cudaEventElapsedTime(&t, e->tev[p][i], e->tev[p][i + 1]) == cudaSuccess) {
                e->stage_ms[i] += t;
                e->stage_launches[i] += 1;
            }
        }
    }
    e->timing_passes = 0;
    if (ms) for (int i = 0; i < FB_NUM_STAGES; i++) ms[i] = e->stage_ms[i];
    if (launches) for (int i = 0; i < FB_NUM_STAGES; i++) launches[i] = e->stage_launches[i];
    return FB_NUM_STAGES;
}

extern "C" void fb_engine_reset_timing(FbEngine *e)
{
    if (!e) return;
    e->timing_passes = 0;
    for (int i = 0; i < FB_NUM_STAGES; i++) { e->stage_ms[i] = 0; e->stage_launches[i] = 0; }
}

extern "C" int fb_engine_read_subframes(FbEngine *e, FbSub *host, uint32_t max, void *stream_v)
{
    if (!e || !host) return -1;
    cudaStream_t st = stream_v ? (cudaStream_t)stream_v : e->stream;
    if (cudaStreamSynchronize(st) != cudaSuccess) return -2;
    uint32_t nf = 0;
    if (cudaMemcpy(&nf, e->d_nframes, sizeof nf, cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
    uint32_t ns = nf * (uint32_t)e->cfg.channels;
    if (ns > max) ns = max;
    if (cudaMemcpy(host, e->d_subs, sizeof(FbSub) * (size_t)ns, cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
    return (int)ns;
}

/* ---- thin runtime wrappers for the C host layer ------------------------- */
extern "C" void *fb_cuda_malloc(size_t n) { void *p = nullptr; return cudaMalloc(&p, n) == cudaSuccess ? p : nullptr; }
extern "C" void fb_cuda_free(void *p) { if (p) cudaFree(p); }
extern "C" void *fb_cuda_malloc_host(size_t n) { void *p = nullptr; return cudaMallocHost(&p, n) == cudaSuccess ? p : nullptr; }
extern "C" void fb_cuda_free_host(void *p) { if (p) cudaFreeHost(p); }
extern "C" void *fb_cuda_stream_create(void)
{
    cudaStream_t s = nullptr;
    return cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) == cudaSuccess ? (void *)s : nullptr;
}
extern "C" void fb_cuda_stream_destroy(void *s) { if (s) cudaStreamDestroy((cudaStream_t)s); }
extern "C" int fb_cuda_stream_sync(void *s) { return cudaStreamSynchronize((cudaStream_t)s) == cudaSuccess ? 0 : -1; }
extern "C" int fb_cuda_h2d(void *d, const void *h, size_t n, void *s)
{
    return cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, (cudaStream_t)s) == cudaSuccess ? 0 : -1;
}
extern "C" int fb_cuda_d2h(void *h, const void *d, size_t n, void *s)
{
    return cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, (cudaStream_t)s) == cudaSuccess ? 0 : -1;
}
extern "C" void *fb_cuda_event_create(void)
{
    cudaEvent_t ev = nullptr;
    return cudaEventCreate(&ev) == cudaSuccess ? (void *)ev : nullptr;
}
/* for host threads that wait on the GPU while other threads need the cores (MD5 of a corpus):
 * cudaEventSynchronize on such an event yields the CPU instead of spinning */
extern "C" void *fb_cuda_event_create_blocking(void)
{
    cudaEvent_t ev = nullptr;
    return cudaEventCreateWithFlags(&ev, cudaEventBlockingSync | cudaEventDisableTiming) == cudaSuccess ? (void *)ev : nullptr;
}
/* 1 when `p` is page-locked host memory the copy engines can reach directly (cudaMallocHost /
 * cudaHostRegister), 0 for pageable memory */
extern "C" int fb_cuda_host_is_pinned(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return 0; }
    return at.type == cudaMemoryTypeHost ? 1 : 0;
}
extern "C" void fb_cuda_event_destroy(void *ev) { if (ev) cudaEventDestroy((cudaEvent_t)ev); }
extern "C" int fb_cuda_event_record(void *ev, void *s) { return cudaEventRecord((cudaEvent_t)ev, (cudaStream_t)s) == cudaSuccess ? 0 : -1; }
extern "C" int fb_cuda_event_sync(void *ev) { return cudaEventSynchronize((cudaEvent_t)ev) == cudaSuccess ? 0 : -1; }
extern "C" int fb_cuda_stream_wait_event(void *s, void *ev) { return cudaStreamWaitEvent((cudaStream_t)s, (cudaEvent_t)ev, 0) == cudaSuccess ? 0 : -1; }
extern "C" float fb_cuda_event_elapsed_ms(void *a, void *b)
{
    float ms = -1.f;
    if (cudaEventElapsedTime(&ms, (cudaEvent_t)a, (cudaEvent_t)b) != cudaSuccess) return -1.f;
    return ms;
}
extern "C" int fb_cuda_device_count(void) { int n = 0; return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0; }
extern "C" int fb_cuda_sm_count(int device)
{
    int n = 0;
    if (device < 0 && cudaGetDevice(&device) != cudaSuccess) return 0;
    return cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) == cudaSuccess ? n : 0;
}
extern "C" int fb_cuda_current_device(void) { int d = -1; return cudaGetDevice(&d) == cudaSuccess ? d : -1; }
extern "C" int fb_cuda_set_device(int dev) { return cudaSetDevice(dev) == cudaSuccess ? 0 : -1; }

#ifdef FB_SEARCH_PROF
extern "C" __attribute__((visibility("default"))) int flake_b200_debug_search_prof(unsigned long long *out, int reset)
{
    if (out && cudaMemcpyFromSymbol(out, g_sprof, sizeof(unsigned long long) * 16) != cudaSuccess) return -1;
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_sprof, z, sizeof z); }
    return 0;
}
#endif

#ifdef FLAKE_B200_CUDA_EMU
/* hooks for the host-side unit tests of device helpers (emulated build only) */
extern "C" int fb_test_rice_k(uint64_t sum, int n) { return fb_rice_k(sum, n); }
extern "C" int fb_test_limit_porder(int p, int n, int order) { return fb_limit_porder(p, n, order); }
extern "C" uint32_t fb_test_crc16_words(const uint8_t *d, uint32_t n)
{
    uint32_t c = 0;
    for (uint32_t i = 0; i < n; i++) c = fb_crc16_byte(c, d[i]);
    return c;
}
extern "C" uint32_t fb_test_gf16_xpow8(uint32_t nbytes) { return fb_gf16_xpow8(nbytes); }
extern "C" uint32_t fb_test_gf16_mul(uint32_t a, uint32_t b) { return fb_gf16_mul(a, b); }
#endif
