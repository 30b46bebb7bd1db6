// Host layer + C ABI (include/nq_celt_synth.h) of the CELT synthesis stage.
//
// Mirrors, for this one path, the reference's interfaces:
//   compute_inv_mdcts / clt_mdct_backward(_B1_C2)   celt_decoder_clean.c:264, mdct.c:258,267
//   the fork's GPU seam processMDCTCuda*             cuda/mdct_cuda.hpp:79-103
// and adds the batched phase-2 entry the restructured decoder calls once per
// batch of frames.  No CPU fallback anywhere: every entry needs a CUDA device.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/nq_celt_synth.h"
#include "celt_synth_kernels.cuh"

using namespace nq;

// ------------------------------------------------------------ tables -------
namespace {

struct HostTables {
    FastTables fast;
    GenericTables gen;
};

// The reference's sin(x)~x shortcut: sine = 2*PI*(.125f)/N in float, PI = 3.141592653f
// (mdct.c:292, mathops.h:83).  Each of the two rotations is multiplied by (1 + j*sine).
double ref_sine(int N) { return (double)((float)2 * 3.141592653f * (.125f) / N); }

const HostTables &host_tables()
{
    static HostTables *T = [] {
        HostTables *t = new HostTables();
        const double pi = 3.14159265358979323846264338327;
        memset(t, 0, sizeof *t);
        // window120: modes.c:374
        for (int i = 0; i < kOverlap; i++) {
            const double s = sin(.5 * pi * (i + .5) / kOverlap);
            t->gen.window[i] = (float)sin(.5 * pi * s * s);
        }
        for (int h = 0; h < 2; h++)
            for (int i = 0; i < 30; i++) {
                const int m = 2 * i + h;
                t->fast.wpair[h * 30 + i] = make_float2(t->gen.window[59 - m], t->gen.window[60 + m]);
            }
        // mdct trig: mdct.c:99, argument evaluated in float (PI is a float macro)
        for (int i = 0; i <= 480; i++) t->gen.trig[i] = (float)cos(2 * 3.141592653f * i / kMdctN);
        // inter-stage twiddles with both MDCT rotations folded in
        {
            const double s = ref_sine(1920);
            const double gr = 1.0 - s * s, gi = 2.0 * s;   // (1 + js)^2
            for (int n2 = 0; n2 < 16; n2++)
                for (int k1 = 0; k1 < 30; k1++) {
                    const double ph = 2 * pi * ((n2 + k1) / 1920.0 + (double)(n2 * k1) / 480.0);
                    const double cr = cos(ph), ci = sin(ph);
                    t->fast.t_long[n2 * kXRowF2 + k1] =
                        make_float2((float)-(gr * cr - gi * ci), (float)-(gr * ci + gi * cr));
                }
        }
        {
            const double s = ref_sine(240);
            const double gr = 1.0 - s * s, gi = 2.0 * s;
            for (int h = 0; h < 2; h++)
                for (int k1 = 0; k1 < 30; k1++) {
                    const double ph = 2 * pi * ((h + k1) / 240.0 + (double)(h * k1) / 60.0);
                    const double cr = cos(ph), ci = sin(ph);
                    t->fast.t_short[h * 30 + k1] =
                        make_float2((float)-(gr * cr - gi * ci), (float)-(gr * ci + gi * cr));
                }
        }
        return t;
    }();
    return *T;
}

}  // namespace

// ------------------------------------------------------------- context -----
struct nq_celt_ctx {
    int device = -1;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    static constexpr int kSlots = 3;
    cudaStream_t slot_stream[kSlots] = {};
    cudaEvent_t kernel_done[kSlots] = {};
    FastTables *d_fast = nullptr;
    GenericTables *d_gen = nullptr;
    // work counters of the dynamically scheduled launches: a ring of kWorkSlots, each guarded by an
    // event recorded behind the launch that used it last -- a launch that reuses the slot (possibly on
    // another stream) first makes its stream wait for that event, so a counter is never zeroed while a
    // kernel is still claiming runs from it
    unsigned long long *d_work = nullptr;
    static constexpr int kWorkSlots = 64;
    cudaEvent_t work_done[kWorkSlots] = {};
    bool work_used[kWorkSlots] = {};
    int work_slot = 0;
    // scratch for the host-pointer batch entry
    float *d_in[kSlots] = {};
    float *d_out[kSlots] = {};
    uint8_t *d_flags[kSlots] = {};
    size_t slot_frames_cap = 0;
    int slot_C_cap = 0;
    size_t slot_in_cap = 0, slot_out_cap = 0, slot_flag_cap = 0;   // bytes
    float *d_tail[2] = {};
    float *d_halo = nullptr;
    int tail_C_cap = 0;
    // post stage: side info + jobs per slot, filter state ping-pong between chunks
    PostFrame *d_pframes[kSlots] = {};
    size_t pframes_cap[kSlots] = {};
    PostJob *d_pjobs[kSlots] = {};
    int pjobs_cap[kSlots] = {};
    long long *d_offs[kSlots] = {};
    size_t offs_cap[kSlots] = {};
    float *d_hist[2] = {};
    float *d_mem[2] = {};
    int post_rows_cap = 0;
    // ... and for the device-pointer post entry: side information and jobs travel on a stream of
    // their own, so that they do not queue behind the synthesis kernel the caller has just enqueued
    cudaStream_t side_stream = nullptr;
    cudaEvent_t side_ready = nullptr, side_free = nullptr;   // upload done / the post kernel that read it is done
    PostFrame *d_pframes_dev = nullptr;
    size_t pframes_dev_cap = 0;
    PostJob *d_pjobs_dev = nullptr;
    int pjobs_dev_cap = 0;
    // scratch for the single-call entries
    float *d_call_buf = nullptr;
    size_t call_buf_cap = 0;
    MdctCall *d_calls = nullptr;
    int calls_cap = 0;
    long long launches = 0;
    char err[512] = {0};
};

namespace {

int fail(nq_celt_ctx *ctx, int code, const char *fmt, ...)
{
    if (ctx) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(ctx->err, sizeof ctx->err, fmt, ap);
        va_end(ap);
    }
    return code;
}

#define NQ_CUDA(ctx, call)                                                                          \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return fail(ctx, NQ_INTERNAL_ERROR, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                        \
    } while (0)

// Splits nframes into runs of consecutive frames, one run per warp-item.
// Every run after the first re-computes one extra frame (its predecessor) to
// obtain the raw tail, so runs are kept long: >= 8 frames when the batch
// allows, and otherwise just long enough to give every resident item one run.
// Runs are kept short and claimed dynamically: stereo measured 383 -> 411 M frames/s at 4 M frames for
// any run length between 32 and 96 (the warm-up frame of every run costs 1/64 extra coefficient reads).
// Mono runs start with a single (half-efficient) frame before pairing up: twice as long (+2 %); the
// issue-bound group variants gain 0-2 %.
constexpr long long kDynamicRun = 64;

// Dynamically claimed batches end on SHORT runs: when the work counter runs dry every warp is somewhere
// inside its last run, so the launch's tail is half a run long on average (64 frames of one warp = 0.3 ms:
// 1 % of a 10 M-frame launch, but 10 % of the 1.25 M frames a GPU gets when 10 M are split over eight).
// The last half wave's worth of frames is therefore cut into runs a quarter as long.

void plan_runs(long long nframes, long long resident_items, long long max_run, SynthParams *p)
{
    long long target_runs = resident_items;
    if (target_runs < 1) target_runs = 1;
    long long K = (nframes + target_runs - 1) / target_runs;
    if (K < 8) K = 8;
    if (max_run > 0 && K > max_run) K = max_run;
    if (K > nframes) K = nframes > 0 ? nframes : 1;
    p->frames_per_run = K;
    p->nruns = p->big_runs = (nframes + K - 1) / K;
    p->small_run = K;
    // more than ~3 waves of full-size runs (i.e. dynamically claimed): the frames of the last wave go in small runs
    if (max_run > 0 && K == max_run && K >= 32 && nframes >= 3 * target_runs * K) {
        const long long small_frames = target_runs * K / 2, small = K / 4;
        p->big_runs = (nframes - small_frames) / K;
        p->small_run = small;
        const long long rest = nframes - p->big_runs * K;
        p->nruns = p->big_runs + (rest + small - 1) / small;
    }
}

// Channel layout of a batch: the arguments of opus_multistream_decoder_create
// (opus_multistream_decoder.c:110).  Decoded channel d is row d of a frame's coefficients:
// coupled stream s -> rows 2s, 2s+1; mono stream s -> row s + coupled (opus_multistream.c:57-91).
struct Layout {
    int C = 0;          // output channels
    int streams = 0;
    int coupled = 0;
    int D = 0;          // streams + coupled
    bool per_stream_flags = false;
    bool identity = true;
    unsigned char mapping[kMaxChannels + 1] = {};
};

// Plain C-channel batch (compute_inv_mdcts with C channels sharing one transient flag):
// pairs (0,1), (2,3), ... plus a trailing mono channel; identity mapping.
Layout plain_layout(int C)
{
    Layout L;
    L.C = L.D = C;
    L.coupled = C / 2;
    L.streams = (C + 1) / 2;
    for (int c = 0; c < C; c++) L.mapping[c] = (unsigned char)c;
    return L;
}

void build_post_jobs(const Layout &L, long long sample0, int frame0, int nframes, bool reset, bool write_state,
                     std::vector<PostJob> *jobs);

// Everything about a launch that follows from the channel layout and the batch size alone: kernel
// variant, warps per group, the streams each warp synthesises, the store pass, the runs.  Pure host
// code (also behind nq_celt_debug_plan, so the CPU tests can check it without a device).
int plan_layout(const Layout &L, int num_sms, long long nframes, SynthParams *pp, int *mode_out)
{
    SynthParams &p = *pp;
    p.D = L.D;
    p.C = L.C;
    p.npairs = (L.D + 1) / 2;
    p.flag_stride = L.per_stream_flags ? L.streams : 1;
    p.flag_per_stream = L.per_stream_flags ? 1 : 0;
    // warps per group: one per coupled stream, one per PAIR of mono streams (two mono streams with
    // their own transient flags share a warp like the two channels of a coupled stream)
    const int nmono = L.streams - L.coupled;
    const int nslots = L.coupled + (nmono + 1) / 2;
    p.nstreams = nslots;
    const int mode = *mode_out = synth_mode(L.D, L.C, nslots, L.identity, L.streams == 1);
    if (mode == kModeDirect && (!L.identity || L.per_stream_flags))
        return NQ_UNIMPLEMENTED;
    if (L.per_stream_flags && L.streams > 30) return NQ_UNIMPLEMENTED;
    long long resident = (long long)num_sms * kWarpsPerCta / p.npairs;
    if (mode == kModeGroup) {
        resident = (long long)num_sms * groups_per_cta(nslots);
        p.groups_per_cta = groups_per_cta(nslots);
        p.store_warps = group_store_warps(nslots);
        p.store_warps_cta = group_store_warps_cta(nslots);
        p.store_threads = group_store_threads(L.C, nslots);
        for (int s = 0; s < L.coupled; s++) {
            p.streams[s].nch = 2;
            p.streams[s].row = (uint8_t)(2 * s);
            p.streams[s].flag_col = p.streams[s].flag_col1 = (uint8_t)(L.per_stream_flags ? s : 0);
        }
        for (int j = 0; j < nmono; j += 2) {
            StreamDesc &sd = p.streams[L.coupled + j / 2];
            const int s0 = L.coupled + j;
            sd.nch = j + 1 < nmono ? 2 : 1;
            sd.row = (uint8_t)(2 * L.coupled + j);
            sd.flag_col = (uint8_t)(L.per_stream_flags ? s0 : 0);
            sd.flag_col1 = (uint8_t)(L.per_stream_flags && sd.nch == 2 ? s0 + 1 : sd.flag_col);
        }
        // store-pass loop shape (celt_synth_kernels.cu group_store_frame): every output channel pair
        // (2k, 2k+1) = one coupled stream's (L, R) in order, so that either half of any float4 of the
        // output is one 8-byte element of a plane -> vector loads; any silent channel -> masked loop
        bool vec = L.C % 2 == 0, muted = false;
        for (int c = 0; c < L.C; c++) muted = muted || L.mapping[c] == 255;
        for (int c = 0; c + 1 < L.C && vec; c += 2) {
            const int d = L.mapping[c];
            vec = d != 255 && d < 2 * L.coupled && (d & 1) == 0 && L.mapping[c + 1] == d + 1;
        }
        p.store_shape = p.store_threads == 0 ? 3 : (muted ? 2 : (vec ? 0 : 1));
        // a lone mono stream (odd number of mono streams: 5.0, 6.1, ...): its warp synthesises frame PAIRS, the
        // time it saves goes to the group's other warps.  Groups of two warps (3.0) do not gain: with one coupled
        // warp per group the coupled stream's own latency sets the pace (measured 0.64 -> 0.62; 5.0: 0.68 -> 0.73)
        p.lone_pairs = (nmono & 1) && nslots >= 3 ? 1 : 0;
        if (const char *e = getenv("NQ_LONE_PAIRS")) p.lone_pairs = (nmono & 1) && atoi(e) ? 1 : 0;
        for (int c = 0; c < L.C; c++) {
            const int d = L.mapping[c];
            if (d == 255) p.chan_src[c] = 0xffffu;   // muted channel, opus_multistream_decoder.c:291-299
            else if (d < 2 * L.coupled) p.chan_src[c] = (uint16_t)(((d >> 1) << 1) | (d & 1));
            else p.chan_src[c] = (uint16_t)(((L.coupled + (d - 2 * L.coupled) / 2) << 1) | ((d - 2 * L.coupled) & 1));
        }
    }
    // stereo / mono: short runs claimed dynamically (see celt_synth_kernel); the warm-up frame of every
    // run costs 1/kDynamicRun extra coefficient reads
    long long max_run = mode == kModeMono ? 2 * kDynamicRun : (mode == kModeDirect ? 0 : kDynamicRun);
    if (const char *e = getenv("NQ_FRAMES_PER_RUN")) max_run = atoll(e);   // tuning knob (0 = one run per resident warp)
    plan_runs(nframes, resident, max_run, &p);
    return NQ_OK;
}

long long resident_items(const SynthParams &p, int mode, int num_sms)
{
    return mode == kModeGroup ? (long long)num_sms * groups_per_cta(p.nstreams) : (long long)num_sms * kWarpsPerCta / p.npairs;
}

// halo_flags: bit s = transient flag of stream s of the halo frame (bit 0 for everybody without
// per-stream flags), bits 30-31 = 3 - LM of the halo frame.
int enqueue_synth(nq_celt_ctx *ctx, const Layout &L, const float *coef, const uint8_t *transient, const float *tail_in,
                  const float *halo_coef, unsigned halo_flags, float *pcm, float *tail_out, long long nframes,
                  cudaStream_t stream, const long long *frame_offset = nullptr)
{
    const unsigned halo_transient_bits = halo_flags & 0x3fffffffu;
    SynthParams p;
    memset(&p, 0, sizeof p);
    p.coef = coef;
    p.transient = transient;
    p.tail_in = tail_in;
    p.halo_coef = tail_in ? nullptr : halo_coef;
    p.halo_transient = (int)halo_transient_bits;
    p.pcm = pcm;
    p.tail_out = tail_out;
    p.tables = ctx->d_fast;
    p.gen = ctx->d_gen;
    p.frame_offset = frame_offset;
    p.halo_lm_shift = (int)(halo_flags >> 30);
    p.nframes = nframes;
    int mode = 0;
    const int rc = plan_layout(L, ctx->num_sms, nframes, &p, &mode);
    if (rc == NQ_UNIMPLEMENTED)
        return fail(ctx, rc, "a channel layout needs at most %d warps (coupled streams + pairs of mono streams) and at most 30 streams with their own flags; got %d streams, %d coupled",
                    kMaxGroupStreams, L.streams, L.coupled);
    int wslot = -1;
    if (p.nruns > resident_items(p, mode, ctx->num_sms)) {
        // one counter per launch in flight (launches on different streams may overlap)
        wslot = ctx->work_slot++ % nq_celt_ctx::kWorkSlots;
        p.work_counter = ctx->d_work + wslot;
        if (ctx->work_used[wslot]) NQ_CUDA(ctx, cudaStreamWaitEvent(stream, ctx->work_done[wslot], 0));
        NQ_CUDA(ctx, cudaMemsetAsync(p.work_counter, 0, sizeof(unsigned long long), stream));
    }
    NQ_CUDA(ctx, launch_synth(p, mode, ctx->num_sms, stream, nullptr));
    if (wslot >= 0) {
        NQ_CUDA(ctx, cudaEventRecord(ctx->work_done[wslot], stream));
        ctx->work_used[wslot] = true;
    }
    ctx->launches++;
    return NQ_OK;
}

int check_layout(nq_celt_ctx *ctx, int channels, int streams, int coupled, const unsigned char *mapping, Layout *L)
{
    // opus_multistream_decoder_init, opus_multistream_decoder.c:63-108 (validate_layout, opus_multistream.c:40-55)
    if (channels < 1 || channels > 255 || streams < 1 || coupled < 0 || coupled > streams || streams + coupled > 255 || !mapping)
        return fail(ctx, NQ_BAD_ARG, "channels=%d streams=%d coupled_streams=%d", channels, streams, coupled);
    L->C = channels;
    L->streams = streams;
    L->coupled = coupled;
    L->D = streams + coupled;
    L->per_stream_flags = true;
    L->identity = channels == L->D;
    for (int c = 0; c < channels; c++) {
        if (mapping[c] != 255 && mapping[c] >= L->D)
            return fail(ctx, NQ_BAD_ARG, "mapping[%d]=%d names no decoded channel (streams+coupled=%d)", c, mapping[c], L->D);
        L->mapping[c] = mapping[c];
        if (mapping[c] != c) L->identity = false;
    }
    return NQ_OK;
}

std::mutex g_mu;
nq_celt_ctx *g_ctx = nullptr;   // process-global context of the reference-shaped void entries

nq_celt_ctx *global_ctx()
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_ctx) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
        int rc = nq_celt_ctx_create(dev, &g_ctx);
        if (rc != NQ_OK) {
            fprintf(stderr, "libnq_celt_b200: cannot create CUDA context (%s); there is no CPU fallback\n",
                    nq_celt_strerror(rc));
            abort();
        }
    }
    return g_ctx;
}

[[noreturn]] void die(nq_celt_ctx *ctx, const char *where, int rc)
{
    fprintf(stderr, "libnq_celt_b200: %s failed: %s: %s\n", where, nq_celt_strerror(rc), ctx ? ctx->err : "");
    abort();
}

}  // namespace

extern "C" {

const char *nq_celt_strerror(int code)
{
    switch (code) {
    case NQ_OK: return "success";
    case NQ_BAD_ARG: return "invalid argument";
    case NQ_INTERNAL_ERROR: return "CUDA error";
    case NQ_UNIMPLEMENTED: return "unimplemented";
    case NQ_INVALID_STATE: return "invalid state";
    case NQ_ALLOC_FAIL: return "allocation failed";
    default: return "unknown error";
    }
}

const char *nq_celt_last_error(const nq_celt_ctx *ctx) { return ctx ? ctx->err : ""; }

long long nq_celt_launch_count(const nq_celt_ctx *ctx) { return ctx ? ctx->launches : 0; }

int nq_celt_ctx_device(const nq_celt_ctx *ctx) { return ctx ? ctx->device : -1; }

void *nq_celt_ctx_stream(const nq_celt_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int nq_celt_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

void *nq_celt_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
    return p;
}

void nq_celt_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

int nq_celt_debug_plan(int channels, int streams, int coupled_streams, const unsigned char *mapping, int64_t nframes,
                       int num_sms, int64_t out[12])
{
    if (!out || nframes < 0 || num_sms < 1) return NQ_BAD_ARG;
    Layout L;
    if (mapping) {
        const int rc = check_layout(nullptr, channels, streams, coupled_streams, mapping, &L);
        if (rc != NQ_OK) return rc;
    } else {
        if (channels < 1 || channels > 255) return NQ_BAD_ARG;
        L = plain_layout(channels);
    }
    SynthParams p;
    memset(&p, 0, sizeof p);
    int mode = 0;
    const int rc = plan_layout(L, num_sms, nframes, &p, &mode);
    if (rc != NQ_OK) return rc;
    bool paired = false;
    for (int s = 0; s < p.nstreams; s++) paired = paired || p.streams[s].flag_col1 != p.streams[s].flag_col;
    std::vector<PostJob> jobs;
    build_post_jobs(L, 0, 0, (int)(nframes > 0x7fffffff ? 0x7fffffff : nframes), false, true, &jobs);
    int jobs2 = 0;
    for (const PostJob &j : jobs) jobs2 += j.nch == 2;
    out[0] = mode;
    out[1] = mode == kModeGroup ? p.nstreams : 0;                   // warps per group
    out[2] = mode == kModeGroup ? groups_per_cta(p.nstreams) : 0;   // groups per CTA
    out[3] = mode == kModeGroup ? p.store_threads : 0;
    out[4] = mode == kModeGroup ? p.store_shape : 0;
    out[5] = paired ? 1 : 0;
    out[6] = p.frames_per_run;
    out[7] = p.nruns;
    out[8] = (int64_t)jobs.size();                                  // post-stage CTAs
    out[9] = jobs2;                                                 // ... of which two-channel
    out[10] = L.D;
    out[11] = L.identity ? 1 : 0;
    return NQ_OK;
}

int nq_celt_debug_runs(int channels, int64_t nframes, int num_sms, int64_t *run_first, int64_t capacity, int64_t *nruns)
{
    if (channels < 1 || channels > 255 || nframes < 0 || num_sms < 1 || !nruns) return NQ_BAD_ARG;
    SynthParams p;
    memset(&p, 0, sizeof p);
    p.nframes = nframes;
    int mode = 0;
    const int rc = plan_layout(plain_layout(channels), num_sms, nframes, &p, &mode);
    if (rc != NQ_OK) return rc;
    *nruns = p.nruns;
    for (long long r = 0; r <= p.nruns && run_first && r < capacity; r++) {
        long long f0 = nframes, f1 = nframes;
        if (r < p.nruns) run_range(p, r, &f0, &f1);
        run_first[r] = f0;
    }
    return NQ_OK;
}

void nq_celt_debug_tables(float *t_long, float *t_short, float *window, float *trig)
{
    const HostTables &t = host_tables();
    if (t_long) memcpy(t_long, t.fast.t_long, sizeof t.fast.t_long);
    if (t_short) memcpy(t_short, t.fast.t_short, sizeof t.fast.t_short);
    if (window) memcpy(window, t.gen.window, sizeof t.gen.window);
    if (trig) memcpy(trig, t.gen.trig, sizeof t.gen.trig);
}

int nq_celt_ctx_create(int device, nq_celt_ctx **out)
{
    if (!out) return NQ_BAD_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return NQ_INTERNAL_ERROR;
    if (device < 0 || device >= ndev) return NQ_BAD_ARG;
    nq_celt_ctx *ctx = new (std::nothrow) nq_celt_ctx();
    if (!ctx) return NQ_ALLOC_FAIL;
    ctx->device = device;
    auto bail = [&](int rc) {
        nq_celt_ctx_destroy(ctx);
        return rc;
    };
    if (cudaSetDevice(device) != cudaSuccess) return bail(NQ_INTERNAL_ERROR);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return bail(NQ_INTERNAL_ERROR);
    if (prop.major != 10) {
        fprintf(stderr, "libnq_celt_b200: device %d is sm_%d%d; this library ships sm_100a code only\n", device,
                prop.major, prop.minor);
        return bail(NQ_INTERNAL_ERROR);
    }
    ctx->num_sms = prop.multiProcessorCount;
    if (prepare_kernels() != cudaSuccess || prepare_post_kernel() != cudaSuccess) return bail(NQ_INTERNAL_ERROR);
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(NQ_INTERNAL_ERROR);
    if (cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking) != cudaSuccess) return bail(NQ_INTERNAL_ERROR);
    if (cudaEventCreateWithFlags(&ctx->side_ready, cudaEventDisableTiming) != cudaSuccess) return bail(NQ_INTERNAL_ERROR);
    if (cudaEventCreateWithFlags(&ctx->side_free, cudaEventDisableTiming) != cudaSuccess) return bail(NQ_INTERNAL_ERROR);
    for (int s = 0; s < nq_celt_ctx::kSlots; s++) {
        if (cudaStreamCreateWithFlags(&ctx->slot_stream[s], cudaStreamNonBlocking) != cudaSuccess) return bail(NQ_INTERNAL_ERROR);
        if (cudaEventCreateWithFlags(&ctx->kernel_done[s], cudaEventDisableTiming) != cudaSuccess) return bail(NQ_INTERNAL_ERROR);
    }
    for (int s = 0; s < nq_celt_ctx::kWorkSlots; s++)
        if (cudaEventCreateWithFlags(&ctx->work_done[s], cudaEventDisableTiming) != cudaSuccess) return bail(NQ_INTERNAL_ERROR);
    const HostTables &t = host_tables();
    if (cudaMalloc(&ctx->d_fast, sizeof(FastTables)) != cudaSuccess) return bail(NQ_ALLOC_FAIL);
    if (cudaMalloc(&ctx->d_gen, sizeof(GenericTables)) != cudaSuccess) return bail(NQ_ALLOC_FAIL);
    if (cudaMalloc(&ctx->d_work, sizeof(unsigned long long) * nq_celt_ctx::kWorkSlots) != cudaSuccess) return bail(NQ_ALLOC_FAIL);
    if (cudaMemcpy(ctx->d_fast, &t.fast, sizeof(FastTables), cudaMemcpyHostToDevice) != cudaSuccess) return bail(NQ_INTERNAL_ERROR);
    if (cudaMemcpy(ctx->d_gen, &t.gen, sizeof(GenericTables), cudaMemcpyHostToDevice) != cudaSuccess) return bail(NQ_INTERNAL_ERROR);
    *out = ctx;
    return NQ_OK;
}

void nq_celt_ctx_destroy(nq_celt_ctx *ctx)
{
    if (!ctx) return;
    if (ctx->device >= 0) cudaSetDevice(ctx->device);
    for (int s = 0; s < nq_celt_ctx::kSlots; s++) {
        if (ctx->slot_stream[s]) { cudaStreamSynchronize(ctx->slot_stream[s]); cudaStreamDestroy(ctx->slot_stream[s]); }
        if (ctx->kernel_done[s]) cudaEventDestroy(ctx->kernel_done[s]);
        cudaFree(ctx->d_in[s]);
        cudaFree(ctx->d_out[s]);
        cudaFree(ctx->d_flags[s]);
    }
    for (int s = 0; s < nq_celt_ctx::kWorkSlots; s++)
        if (ctx->work_done[s]) cudaEventDestroy(ctx->work_done[s]);
    if (ctx->side_stream) { cudaStreamSynchronize(ctx->side_stream); cudaStreamDestroy(ctx->side_stream); }
    if (ctx->side_ready) cudaEventDestroy(ctx->side_ready);
    if (ctx->side_free) cudaEventDestroy(ctx->side_free);
    if (ctx->stream) { cudaStreamSynchronize(ctx->stream); cudaStreamDestroy(ctx->stream); }
    cudaFree(ctx->d_tail[0]);
    cudaFree(ctx->d_tail[1]);
    cudaFree(ctx->d_halo);
    for (int s = 0; s < nq_celt_ctx::kSlots; s++) { cudaFree(ctx->d_pframes[s]); cudaFree(ctx->d_pjobs[s]); cudaFree(ctx->d_offs[s]); }
    for (int i = 0; i < 2; i++) { cudaFree(ctx->d_hist[i]); cudaFree(ctx->d_mem[i]); }
    cudaFree(ctx->d_pframes_dev);
    cudaFree(ctx->d_pjobs_dev);
    cudaFree(ctx->d_call_buf);
    cudaFree(ctx->d_calls);
    cudaFree(ctx->d_fast);
    cudaFree(ctx->d_gen);
    cudaFree(ctx->d_work);
    delete ctx;
}

int nq_celt_synth_batch_device(nq_celt_ctx *ctx, const float *coef, const uint8_t *transient, const float *tail_in,
                               const float *halo_coef, int halo_transient, float *pcm_out, float *tail_out,
                               int64_t nframes, int C, void *stream)
{
    if (!ctx) return NQ_BAD_ARG;
    if (nframes < 0 || C < 1 || C > 255) return fail(ctx, NQ_BAD_ARG, "nframes=%lld C=%d out of range", (long long)nframes, C);
    if (nframes == 0) {
        // nothing to synthesise: the tail passes through unchanged
        if (tail_out) {
            NQ_CUDA(ctx, cudaSetDevice(ctx->device));
            cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
            if (tail_in) NQ_CUDA(ctx, cudaMemcpyAsync(tail_out, tail_in, sizeof(float) * C * kHalfOvl, cudaMemcpyDeviceToDevice, st));
            else NQ_CUDA(ctx, cudaMemsetAsync(tail_out, 0, sizeof(float) * C * kHalfOvl, st));
        }
        return NQ_OK;
    }
    if (!coef || !transient || !pcm_out) return fail(ctx, NQ_BAD_ARG, "null coef/transient/pcm_out");
    if ((reinterpret_cast<uintptr_t>(coef) & 15) || (reinterpret_cast<uintptr_t>(pcm_out) & 15) ||
        (halo_coef && (reinterpret_cast<uintptr_t>(halo_coef) & 15)))
        return fail(ctx, NQ_BAD_ARG, "coef, halo_coef and pcm_out must be 16-byte aligned device pointers");
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    const unsigned hf = (unsigned)halo_transient;   // a flag byte: bit 0 transient, bits 1-2 = 3 - LM
    return enqueue_synth(ctx, plain_layout(C), coef, transient, tail_in, halo_coef, (hf & 1u) | ((hf >> 1 & 3u) << 30),
                         pcm_out, tail_out, nframes, stream ? (cudaStream_t)stream : ctx->stream);
}

int nq_celt_synth_batch_device_ms(nq_celt_ctx *ctx, const float *coef, const uint8_t *transient, const float *tail_in,
                                  const float *halo_coef, const uint8_t *halo_transient, float *pcm_out,
                                  float *tail_out, const int64_t *frame_offset, int64_t nframes, int channels,
                                  int streams, int coupled_streams, const unsigned char *mapping, void *stream)
{
    if (!ctx) return NQ_BAD_ARG;
    Layout L;
    if (mapping) {
        int rc = check_layout(ctx, channels, streams, coupled_streams, mapping, &L);
        if (rc != NQ_OK) return rc;
    } else {
        if (channels < 1 || channels > 2 || streams != 1)
            return fail(ctx, NQ_BAD_ARG, "without a mapping: one CELT decoder, channels 1 or 2 (got %d channels, %d streams)", channels, streams);
        L = plain_layout(channels);
    }
    if (nframes < 0) return fail(ctx, NQ_BAD_ARG, "nframes=%lld", (long long)nframes);
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    if (nframes == 0) {
        if (tail_out) {
            if (tail_in) NQ_CUDA(ctx, cudaMemcpyAsync(tail_out, tail_in, sizeof(float) * L.D * kHalfOvl, cudaMemcpyDeviceToDevice, st));
            else NQ_CUDA(ctx, cudaMemsetAsync(tail_out, 0, sizeof(float) * L.D * kHalfOvl, st));
        }
        return NQ_OK;
    }
    if (!coef || !transient || !pcm_out) return fail(ctx, NQ_BAD_ARG, "null coef/transient/pcm_out");
    if ((reinterpret_cast<uintptr_t>(coef) & 15) || (reinterpret_cast<uintptr_t>(pcm_out) & 15) ||
        (halo_coef && (reinterpret_cast<uintptr_t>(halo_coef) & 15)))
        return fail(ctx, NQ_BAD_ARG, "coef, halo_coef and pcm_out must be 16-byte aligned device pointers");
    unsigned halo_bits = 0;
    if (halo_coef && !tail_in) {
        if (!halo_transient) return fail(ctx, NQ_BAD_ARG, "halo_coef needs halo_transient[streams] (host pointer)");
        for (int s = 0; s < streams; s++) halo_bits |= (unsigned)(halo_transient[s] & 1) << s;
        halo_bits |= (unsigned)(halo_transient[0] >> 1 & 3) << 30;
    }
    static_assert(sizeof(long long) == sizeof(int64_t), "frame_offset element type");
    return enqueue_synth(ctx, L, coef, transient, tail_in, halo_coef, halo_bits, pcm_out, tail_out, nframes, st,
                         reinterpret_cast<const long long *>(frame_offset));
}

}  // extern "C"

namespace {

static_assert(sizeof(PostFrame) == sizeof(nq_celt_post_frame), "side-info record layout");

// Post jobs of one contiguous frame range: adjacent output channels fed by the two channels of
// one coupled stream share a warp; every other live channel gets its own; silent channels none.
void build_post_jobs(const Layout &L, long long sample0, int frame0, int nframes, bool reset, bool write_state,
                     std::vector<PostJob> *jobs)
{
    for (int c = 0; c < L.C;) {
        const int d = L.mapping[c];
        if (d == 255) { c++; continue; }
        PostJob j;
        memset(&j, 0, sizeof j);
        j.sample0 = sample0;
        j.frame0 = frame0;
        j.nframes = nframes;
        j.ch0 = c;
        j.state_row = d;
        j.reset = reset ? 1 : 0;
        j.write_state = write_state ? 1 : 0;
        if (d < 2 * L.coupled && (d & 1) == 0 && c + 1 < L.C && L.mapping[c + 1] == d + 1) {
            j.nch = 2;
            j.stream_col = d >> 1;
            c += 2;
        } else {
            j.nch = 1;
            j.stream_col = d < 2 * L.coupled ? d >> 1 : d - L.coupled;
            c += 1;
        }
        jobs->push_back(j);
    }
}

// Post jobs of frames [0, n) of a (chunk of a) batch whose flag bytes `flags` (host, one record of
// `fcols` bytes per frame) may carry kFlagReset: a reset starts a new piece with its own CTA and a
// zeroed filter state.  Every stream is a decoder of its own and is reset on its own, so each
// channel job is cut at the resets of ITS stream (column stream_col; column 0 without per-stream
// flags).  The first piece continues from the incoming state unless it opens on a reset; only the
// last piece leaves its state behind.
void build_post_jobs_with_resets(const Layout &L, const uint8_t *flags, int fcols, const nq_celt_post_frame *frames,
                                 long long n, std::vector<PostJob> *jobs)
{
    std::vector<PostJob> whole;
    build_post_jobs(L, 0, 0, (int)n, false, true, &whole);
    bool any_reset = false;
    for (long long f = 0; f < n && !any_reset; f++)
        for (int c = 0; c < fcols && !any_reset; c++) any_reset = (flags[f * fcols + c] & kFlagReset) != 0;
    if (!any_reset) {
        jobs->insert(jobs->end(), whole.begin(), whole.end());
        return;
    }
    std::vector<long long> first(n + 1, 0);   // first sample of every frame
    for (long long f = 0; f < n; f++) first[f + 1] = first[f] + frames[f * L.streams].N;
    for (const PostJob &w : whole) {
        const int col = fcols > 1 ? w.stream_col : 0;
        long long start = 0;
        for (long long f = 1; f <= n; f++) {
            if (f == n || (flags[f * fcols + col] & kFlagReset)) {
                PostJob j = w;
                j.sample0 = first[start];
                j.frame0 = (int)start;
                j.nframes = (int)(f - start);
                j.reset = (flags[start * fcols + col] & kFlagReset) ? 1 : 0;
                j.write_state = f == n ? 1 : 0;
                jobs->push_back(j);
                start = f;
            }
        }
    }
}

int enqueue_post(nq_celt_ctx *ctx, const Layout &L, float *pcm, const PostFrame *d_frames, const PostJob *d_jobs, int njobs,
                 const float *hist_in, const float *mem_in, float *hist_out, float *mem_out, cudaStream_t stream)
{
    PostParams p;
    memset(&p, 0, sizeof p);
    p.pcm = pcm;
    p.frames = d_frames;
    p.jobs = d_jobs;
    p.window = ctx->d_gen->window;
    p.hist_in = hist_in;
    p.mem_in = mem_in;
    p.hist_out = hist_out;
    p.mem_out = mem_out;
    p.C = L.C;
    p.frame_stride = L.streams;
    NQ_CUDA(ctx, launch_post(p, njobs, stream));
    ctx->launches++;
    return NQ_OK;
}

template <class T>
int grow(nq_celt_ctx *ctx, T **buf, size_t *cap, size_t need_bytes, const char *what)
{
    if (need_bytes <= *cap) return NQ_OK;
    cudaFree(*buf);
    *buf = nullptr;
    *cap = 0;
    if (cudaMalloc(buf, need_bytes) != cudaSuccess) return fail(ctx, NQ_ALLOC_FAIL, "%s (%zu bytes)", what, need_bytes);
    *cap = need_bytes;
    return NQ_OK;
}

// Entries without side information synthesise 20 ms frames only (rows and output slots of 960
// samples): a flag byte that asks for a shorter frame (bits 1-2) or carries unknown bits is an error
// there, not a silently mis-sized frame.  Returns the first offending frame or -1.
long long first_non_20ms_flag(const uint8_t *flags, long long n)
{
    for (long long f = 0; f < n; f++)
        if (flags[f] & ~(kFlagTransient | kFlagReset)) return f;
    return -1;
}

// Tuning knobs of the host-buffer pipeline (read per call, so a probe can sweep them in one
// process): NQ_HOST_SLOTS = chunks in flight (1..3), NQ_HOST_CHUNK_MB = megabytes of coefficients
// per chunk.
int host_slots()
{
    const char *e = getenv("NQ_HOST_SLOTS");
    const int n = e ? atoi(e) : nq_celt_ctx::kSlots;
    return n < 1 ? 1 : (n > nq_celt_ctx::kSlots ? nq_celt_ctx::kSlots : n);
}
size_t host_chunk_bytes()
{
    const char *e = getenv("NQ_HOST_CHUNK_MB");
    const long mb = e ? atol(e) : 64;
    return (size_t)(mb < 1 ? 1 : (mb > 1024 ? 1024 : mb)) << 20;
}

// Host-pointer batch over [0, nframes): chunks are pipelined H2D / synthesis (/ post stage) / D2H
// on the context's slot streams.  `halo_coef` (+ flags): optional host halo frame of a shard that
// starts mid-stream.  `pframes` != NULL adds the post stage (all frames must be 20 ms frames).
struct HostState {
    const float *tail_in = nullptr;    // [D][60]
    float *tail_out = nullptr;
    const float *hist_in = nullptr;    // [D][1026]
    const float *mem_in = nullptr;     // [D]
    float *hist_out = nullptr;
    float *mem_out = nullptr;
};

int host_range(nq_celt_ctx *ctx, const Layout &L, const float *coef, const uint8_t *transient,
               const nq_celt_post_frame *pframes, const HostState &st, const float *halo_coef, unsigned halo_bits,
               float *pcm_out, long long nframes)
{
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    const int S = host_slots();
    const size_t in_row = (size_t)L.D * kFrame, out_row = (size_t)L.C * kFrame;
    const size_t flag_row = L.per_stream_flags ? (size_t)L.streams : 1;
    // chunk size: ~64 MB of coefficients per slot, enough frames to fill the GPU
    long long chunk = (long long)(host_chunk_bytes() / (in_row * sizeof(float)));
    if (chunk < 1024) chunk = 1024;
    if (chunk > nframes) chunk = nframes;
    {
        bool need = false;
        for (int s = 0; s < S; s++)
            need = need || !ctx->d_in[s] || !ctx->d_out[s] || !ctx->d_flags[s];
        need = need || chunk * in_row * sizeof(float) > ctx->slot_in_cap || chunk * out_row * sizeof(float) > ctx->slot_out_cap ||
               chunk * flag_row > ctx->slot_flag_cap;
        if (need) {
            for (int s = 0; s < nq_celt_ctx::kSlots; s++) {   // (all of them: the slots beyond S must not keep a smaller buffer)
                NQ_CUDA(ctx, cudaStreamSynchronize(ctx->slot_stream[s]));
                cudaFree(ctx->d_in[s]); cudaFree(ctx->d_out[s]); cudaFree(ctx->d_flags[s]);
                ctx->d_in[s] = ctx->d_out[s] = nullptr; ctx->d_flags[s] = nullptr;
            }
            ctx->slot_in_cap = ctx->slot_out_cap = ctx->slot_flag_cap = 0;
            for (int s = 0; s < S; s++) {
                if (cudaMalloc(&ctx->d_in[s], chunk * in_row * sizeof(float)) != cudaSuccess ||
                    cudaMalloc(&ctx->d_out[s], chunk * out_row * sizeof(float)) != cudaSuccess ||
                    cudaMalloc(&ctx->d_flags[s], chunk * flag_row) != cudaSuccess)
                    return fail(ctx, NQ_ALLOC_FAIL, "device scratch for %lld frames x %d channels", chunk, L.D);
            }
            ctx->slot_in_cap = chunk * in_row * sizeof(float);
            ctx->slot_out_cap = chunk * out_row * sizeof(float);
            ctx->slot_flag_cap = chunk * flag_row;
        }
    }
    if (L.D > ctx->tail_C_cap) {
        cudaFree(ctx->d_tail[0]); cudaFree(ctx->d_tail[1]); cudaFree(ctx->d_halo);
        ctx->d_tail[0] = ctx->d_tail[1] = ctx->d_halo = nullptr;
        ctx->tail_C_cap = 0;
        if (cudaMalloc(&ctx->d_tail[0], sizeof(float) * L.D * kHalfOvl) != cudaSuccess ||
            cudaMalloc(&ctx->d_tail[1], sizeof(float) * L.D * kHalfOvl) != cudaSuccess ||
            cudaMalloc(&ctx->d_halo, sizeof(float) * in_row) != cudaSuccess)
            return fail(ctx, NQ_ALLOC_FAIL, "device tail buffers");
        ctx->tail_C_cap = L.D;
    }
    if (pframes && L.D > ctx->post_rows_cap) {
        for (int i = 0; i < 2; i++) {
            cudaFree(ctx->d_hist[i]); cudaFree(ctx->d_mem[i]);
            ctx->d_hist[i] = ctx->d_mem[i] = nullptr;
        }
        ctx->post_rows_cap = 0;
        for (int i = 0; i < 2; i++)
            if (cudaMalloc(&ctx->d_hist[i], sizeof(float) * L.D * kPostHist) != cudaSuccess ||
                cudaMalloc(&ctx->d_mem[i], sizeof(float) * L.D) != cudaSuccess)
                return fail(ctx, NQ_ALLOC_FAIL, "device post-filter state");
        ctx->post_rows_cap = L.D;
    }
    const size_t tail_bytes = sizeof(float) * L.D * kHalfOvl;
    const size_t hist_bytes = sizeof(float) * L.D * kPostHist, mem_bytes = sizeof(float) * L.D;
    // chunk i reads its initial state from buffer (i+1)&1 and leaves its final state in buffer i&1
    const bool use_halo = !st.tail_in && halo_coef;
    cudaStream_t s0 = ctx->slot_stream[0];
    if (st.tail_in) NQ_CUDA(ctx, cudaMemcpyAsync(ctx->d_tail[1], st.tail_in, tail_bytes, cudaMemcpyHostToDevice, s0));
    else NQ_CUDA(ctx, cudaMemsetAsync(ctx->d_tail[1], 0, tail_bytes, s0));
    if (use_halo) NQ_CUDA(ctx, cudaMemcpyAsync(ctx->d_halo, halo_coef, in_row * sizeof(float), cudaMemcpyHostToDevice, s0));
    if (pframes) {
        if (st.hist_in) NQ_CUDA(ctx, cudaMemcpyAsync(ctx->d_hist[1], st.hist_in, hist_bytes, cudaMemcpyHostToDevice, s0));
        else NQ_CUDA(ctx, cudaMemsetAsync(ctx->d_hist[1], 0, hist_bytes, s0));
        if (st.mem_in) NQ_CUDA(ctx, cudaMemcpyAsync(ctx->d_mem[1], st.mem_in, mem_bytes, cudaMemcpyHostToDevice, s0));
        else NQ_CUDA(ctx, cudaMemsetAsync(ctx->d_mem[1], 0, mem_bytes, s0));
    }

    // frames shorter than 20 ms: output offsets come from the side info (prefix sums of N)
    std::vector<long long> offs, rel;
    if (pframes) {
        bool any = false;
        for (long long f = 0; f < nframes && !any; f++) any = pframes[f * L.streams].N != kFrame;
        if (any) {
            offs.resize(nframes + 1);
            offs[0] = 0;
            for (long long f = 0; f < nframes; f++) offs[f + 1] = offs[f] + pframes[f * L.streams].N;
        }
    }
    std::vector<PostJob> jobs;
    const long long nchunks = (nframes + chunk - 1) / chunk;
    for (long long i = 0; i < nchunks; i++) {
        const int s = (int)(i % S);
        cudaStream_t sst = ctx->slot_stream[s];
        const long long f0 = i * chunk, n = (f0 + chunk <= nframes) ? chunk : nframes - f0;
        NQ_CUDA(ctx, cudaMemcpyAsync(ctx->d_in[s], coef + f0 * in_row, n * in_row * sizeof(float), cudaMemcpyHostToDevice, sst));
        NQ_CUDA(ctx, cudaMemcpyAsync(ctx->d_flags[s], transient + f0 * flag_row, (size_t)n * flag_row, cudaMemcpyHostToDevice, sst));
        if (pframes) {
            size_t cap = ctx->pframes_cap[s];
            int rc = grow(ctx, &ctx->d_pframes[s], &cap, (size_t)n * L.streams * sizeof(PostFrame), "post side info");
            ctx->pframes_cap[s] = cap;
            if (rc != NQ_OK) return rc;
            NQ_CUDA(ctx, cudaMemcpyAsync(ctx->d_pframes[s], pframes + f0 * L.streams, (size_t)n * L.streams * sizeof(PostFrame),
                                         cudaMemcpyHostToDevice, sst));
        }
        if (i > 0) NQ_CUDA(ctx, cudaStreamWaitEvent(sst, ctx->kernel_done[(i - 1) % S], 0));
        const bool first_with_halo = (i == 0 && use_halo);
        const long long *d_off = nullptr;
        long long out_first = f0 * kFrame, out_count = n * kFrame;   // samples per channel
        if (!offs.empty()) {
            rel.resize(n);
            for (long long k = 0; k < n; k++) rel[k] = offs[f0 + k] - offs[f0];
            size_t cap = ctx->offs_cap[s];
            int rc0 = grow(ctx, &ctx->d_offs[s], &cap, (size_t)n * sizeof(long long), "frame offsets");
            ctx->offs_cap[s] = cap;
            if (rc0 != NQ_OK) return rc0;
            NQ_CUDA(ctx, cudaMemcpyAsync(ctx->d_offs[s], rel.data(), (size_t)n * sizeof(long long), cudaMemcpyHostToDevice, sst));
            d_off = ctx->d_offs[s];
            out_first = offs[f0];
            out_count = offs[f0 + n] - offs[f0];
        }
        int rc = enqueue_synth(ctx, L, ctx->d_in[s], ctx->d_flags[s], first_with_halo ? nullptr : ctx->d_tail[(i + 1) & 1],
                               first_with_halo ? ctx->d_halo : nullptr, halo_bits, ctx->d_out[s], ctx->d_tail[i & 1], n, sst,
                               d_off);
        if (rc != NQ_OK) return rc;
        if (pframes) {
            jobs.clear();
            build_post_jobs_with_resets(L, transient + f0 * flag_row, (int)flag_row, pframes + f0 * L.streams, n, &jobs);
            if ((int)jobs.size() > ctx->pjobs_cap[s]) {
                cudaFree(ctx->d_pjobs[s]);
                ctx->d_pjobs[s] = nullptr;
                ctx->pjobs_cap[s] = 0;
                if (cudaMalloc(&ctx->d_pjobs[s], jobs.size() * sizeof(PostJob)) != cudaSuccess) return fail(ctx, NQ_ALLOC_FAIL, "post jobs");
                ctx->pjobs_cap[s] = (int)jobs.size();
            }
            NQ_CUDA(ctx, cudaMemcpyAsync(ctx->d_pjobs[s], jobs.data(), jobs.size() * sizeof(PostJob), cudaMemcpyHostToDevice, sst));
            rc = enqueue_post(ctx, L, ctx->d_out[s], ctx->d_pframes[s], ctx->d_pjobs[s], (int)jobs.size(), ctx->d_hist[(i + 1) & 1],
                              ctx->d_mem[(i + 1) & 1], ctx->d_hist[i & 1], ctx->d_mem[i & 1], sst);
            if (rc != NQ_OK) return rc;
        }
        NQ_CUDA(ctx, cudaEventRecord(ctx->kernel_done[s], sst));
        NQ_CUDA(ctx, cudaMemcpyAsync(pcm_out + out_first * L.C, ctx->d_out[s], (size_t)out_count * L.C * sizeof(float),
                                     cudaMemcpyDeviceToHost, sst));
        if (i == nchunks - 1) {
            if (st.tail_out) NQ_CUDA(ctx, cudaMemcpyAsync(st.tail_out, ctx->d_tail[i & 1], tail_bytes, cudaMemcpyDeviceToHost, sst));
            if (pframes && st.hist_out) NQ_CUDA(ctx, cudaMemcpyAsync(st.hist_out, ctx->d_hist[i & 1], hist_bytes, cudaMemcpyDeviceToHost, sst));
            if (pframes && st.mem_out) NQ_CUDA(ctx, cudaMemcpyAsync(st.mem_out, ctx->d_mem[i & 1], mem_bytes, cudaMemcpyDeviceToHost, sst));
        }
    }
    for (int s = 0; s < S; s++) NQ_CUDA(ctx, cudaStreamSynchronize(ctx->slot_stream[s]));
    return NQ_OK;
}

int synth_host_range(nq_celt_ctx *ctx, const float *coef, const uint8_t *transient, const float *tail_in,
                     const float *halo_coef, int halo_transient, float *pcm_out, float *tail_out,
                     long long nframes, int C)
{
    HostState st;
    st.tail_in = tail_in;
    st.tail_out = tail_out;
    const unsigned hf = (unsigned)halo_transient;
    return host_range(ctx, plain_layout(C), coef, transient, nullptr, st, halo_coef, (hf & 1u) | ((hf >> 1 & 3u) << 30), pcm_out, nframes);
}

}  // namespace

extern "C" {

int nq_celt_synth_batch_host(nq_celt_ctx *ctx, const float *coef, const uint8_t *transient, const float *tail_in,
                             float *pcm_out, float *tail_out, int64_t nframes, int C)
{
    if (!ctx) return NQ_BAD_ARG;
    if (nframes < 0 || C < 1 || C > 255) return fail(ctx, NQ_BAD_ARG, "nframes=%lld C=%d out of range", (long long)nframes, C);
    if (nframes == 0) {
        if (tail_out) {
            if (tail_in) memcpy(tail_out, tail_in, sizeof(float) * C * kHalfOvl);
            else memset(tail_out, 0, sizeof(float) * C * kHalfOvl);
        }
        return NQ_OK;
    }
    if (!coef || !transient || !pcm_out) return fail(ctx, NQ_BAD_ARG, "null coef/transient/pcm_out");
    if (const long long bad = first_non_20ms_flag(transient, nframes); bad >= 0)
        return fail(ctx, NQ_BAD_ARG, "frame %lld: flag byte 0x%02x; this entry takes 20 ms frames only (bit 0 transient, bit 3 reset); "
                    "shorter frames need side information: nq_celt_decode_batch_host", bad, transient[bad]);
    return synth_host_range(ctx, coef, transient, tail_in, nullptr, 0, pcm_out, tail_out, nframes, C);
}

// ---- post stage + whole phase 2 ------------------------------------------------------------
static int post_device(nq_celt_ctx *ctx, float *pcm, const nq_celt_post_frame *frames, const int64_t *seg_start, int nseg,
                       const float *hist_in, const float *mem_in, float *hist_out, float *mem_out, int64_t nframes,
                       int channels, int streams, int coupled_streams, const unsigned char *mapping, void *stream)
{
    if (!ctx) return NQ_BAD_ARG;
    Layout L;
    if (mapping) {
        int rc = check_layout(ctx, channels, streams, coupled_streams, mapping, &L);
        if (rc != NQ_OK) return rc;
    } else {
        if (channels < 1 || channels > 2) return fail(ctx, NQ_BAD_ARG, "without a mapping: one CELT decoder, channels 1 or 2 (got %d)", channels);
        L = plain_layout(channels);
    }
    if (nframes < 0 || nframes > 0x7fffffff) return fail(ctx, NQ_BAD_ARG, "nframes=%lld", (long long)nframes);
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    const size_t hist_bytes = sizeof(float) * L.D * kPostHist, mem_bytes = sizeof(float) * L.D;
    if (nframes == 0) {
        if (hist_out) {
            if (hist_in) NQ_CUDA(ctx, cudaMemcpyAsync(hist_out, hist_in, hist_bytes, cudaMemcpyDeviceToDevice, st));
            else NQ_CUDA(ctx, cudaMemsetAsync(hist_out, 0, hist_bytes, st));
        }
        if (mem_out) {
            if (mem_in) NQ_CUDA(ctx, cudaMemcpyAsync(mem_out, mem_in, mem_bytes, cudaMemcpyDeviceToDevice, st));
            else NQ_CUDA(ctx, cudaMemsetAsync(mem_out, 0, mem_bytes, st));
        }
        return NQ_OK;
    }
    if (!pcm || !frames) return fail(ctx, NQ_BAD_ARG, "null pcm/frames");
    if (reinterpret_cast<uintptr_t>(pcm) & 15) return fail(ctx, NQ_BAD_ARG, "pcm must be a 16-byte aligned device pointer");
    // One pass over the side information: validate it and note the first sample of every segment
    // (this loop runs while the synthesis kernel of the same batch is still busy, so it should not
    // take longer than that: no per-frame tables, nothing allocated per frame).
    std::vector<long long> seg_first;
    if (seg_start) {
        if (nseg < 1 || seg_start[0] != 0 || seg_start[nseg] != nframes) return fail(ctx, NQ_BAD_ARG, "seg_start must run from 0 to nframes");
        for (int k = 0; k < nseg; k++)
            if (seg_start[k + 1] < seg_start[k]) return fail(ctx, NQ_BAD_ARG, "seg_start must be non-decreasing");
        seg_first.assign(nseg, 0);
    }
    {
        // frames [f0, f1): returns the first bad frame (or -1) and the samples they hold; segments
        // that start inside the range get their first sample RELATIVE to the range's first sample
        const int S = L.streams;
        auto scan = [&](int64_t f0, int64_t f1, long long *samples) -> int64_t {
            long long pos = 0;
            int k = 0;
            if (seg_start) k = (int)(std::lower_bound(seg_start, seg_start + nseg, f0) - seg_start);
            for (int64_t fi = f0; fi < f1; fi++) {
                while (seg_start && k < nseg && seg_start[k] == fi) seg_first[k++] = pos;
                const nq_celt_post_frame *row = frames + fi * S;
                const int N0 = row[0].N;
                bool ok = N0 == 120 || N0 == 240 || N0 == 480 || N0 == 960;
                for (int sidx = 0; sidx < S && ok; sidx++) {
                    const nq_celt_post_frame &f = row[sidx];
                    ok = f.N == N0;
                    for (int j = 0; j < 3 && ok; j++) {
                        // a zero-gain filter is never evaluated, whatever its period (celt_decoder_clean.c passes pitch 0 then)
                        if (f.gain[j] != 0.f && (f.pitch[j] < 15 || f.pitch[j] > 1022)) ok = false;   // COMBFILTER_MINPERIOD .. 1022
                        if (f.tapset[j] < 0 || f.tapset[j] > 2) ok = false;
                    }
                }
                if (!ok) return fi;
                pos += N0;
            }
            *samples = pos;
            return -1;
        };
        // big batches: a few host threads, so that the scan hides behind the synthesis kernel
        const int nthreads = nframes * S >= 200000 ? 4 : 1;
        std::vector<long long> part_samples(nthreads, 0);
        std::vector<int64_t> part_bad(nthreads, -1);
        if (nthreads == 1) {
            part_bad[0] = scan(0, nframes, &part_samples[0]);
        } else {
            std::vector<std::thread> th;
            for (int t = 0; t < nthreads; t++)
                th.emplace_back([&, t] { part_bad[t] = scan(nframes * t / nthreads, nframes * (t + 1) / nthreads, &part_samples[t]); });
            for (std::thread &x : th) x.join();
        }
        for (int t = 0; t < nthreads; t++)
            if (part_bad[t] >= 0)
                return fail(ctx, NQ_BAD_ARG, "frame %lld: bad side info (N=%d; every stream of a frame must carry the same N in {120,240,480,960}, "
                            "periods 15..1022 where the gain is not zero, tapsets 0..2)", (long long)part_bad[t], frames[part_bad[t] * S].N);
        if (seg_start) {   // relative -> absolute; segments that start at nframes (empty, at the end) hold everything
            long long part_first = 0;
            int k = 0;
            for (int t = 0; t < nthreads; t++) {
                const int64_t f1 = nframes * (t + 1) / nthreads;
                for (; k < nseg && seg_start[k] < f1; k++) seg_first[k] += part_first;
                part_first += part_samples[t];
            }
            for (; k < nseg; k++) seg_first[k] = part_first;
        }
    }
    int rc = grow(ctx, &ctx->d_pframes_dev, &ctx->pframes_dev_cap, (size_t)nframes * L.streams * sizeof(PostFrame), "post side info");
    if (rc != NQ_OK) return rc;
    // (the previous post kernel of this context may still be reading the buffers)
    NQ_CUDA(ctx, cudaStreamWaitEvent(ctx->side_stream, ctx->side_free, 0));
    NQ_CUDA(ctx, cudaMemcpyAsync(ctx->d_pframes_dev, frames, (size_t)nframes * L.streams * sizeof(PostFrame), cudaMemcpyHostToDevice, ctx->side_stream));
    std::vector<PostJob> jobs;
    if (seg_start) {
        for (int k = 0; k < nseg; k++)
            if (seg_start[k + 1] > seg_start[k])
                build_post_jobs(L, seg_first[k], (int)seg_start[k], (int)(seg_start[k + 1] - seg_start[k]), true, false, &jobs);
    } else {
        build_post_jobs(L, 0, 0, (int)nframes, false, true, &jobs);
    }
    if (jobs.empty()) return NQ_OK;
    if ((int)jobs.size() > ctx->pjobs_dev_cap) {
        cudaFree(ctx->d_pjobs_dev);
        ctx->d_pjobs_dev = nullptr;
        ctx->pjobs_dev_cap = 0;
        if (cudaMalloc(&ctx->d_pjobs_dev, jobs.size() * sizeof(PostJob)) != cudaSuccess) return fail(ctx, NQ_ALLOC_FAIL, "post jobs");
        ctx->pjobs_dev_cap = (int)jobs.size();
    }
    NQ_CUDA(ctx, cudaMemcpyAsync(ctx->d_pjobs_dev, jobs.data(), jobs.size() * sizeof(PostJob), cudaMemcpyHostToDevice, ctx->side_stream));
    NQ_CUDA(ctx, cudaEventRecord(ctx->side_ready, ctx->side_stream));
    NQ_CUDA(ctx, cudaStreamWaitEvent(st, ctx->side_ready, 0));
    rc = enqueue_post(ctx, L, pcm, ctx->d_pframes_dev, ctx->d_pjobs_dev, (int)jobs.size(), hist_in, mem_in, hist_out, mem_out, st);
    if (rc == NQ_OK) NQ_CUDA(ctx, cudaEventRecord(ctx->side_free, st));
    return rc;
}

int nq_celt_post_batch_device(nq_celt_ctx *ctx, float *pcm, const nq_celt_post_frame *frames, const float *hist_in,
                              const float *mem_in, float *hist_out, float *mem_out, int64_t nframes, int channels,
                              int streams, int coupled_streams, const unsigned char *mapping, void *stream)
{
    return post_device(ctx, pcm, frames, nullptr, 0, hist_in, mem_in, hist_out, mem_out, nframes, channels, streams,
                       coupled_streams, mapping, stream);
}

int nq_celt_post_segments_device(nq_celt_ctx *ctx, float *pcm, const nq_celt_post_frame *frames, const int64_t *seg_start,
                                 int nseg, int64_t nframes, int channels, int streams, int coupled_streams,
                                 const unsigned char *mapping, void *stream)
{
    if (!seg_start) return NQ_BAD_ARG;
    return post_device(ctx, pcm, frames, seg_start, nseg, nullptr, nullptr, nullptr, nullptr, nframes, channels, streams,
                       coupled_streams, mapping, stream);
}

int nq_celt_decode_batch_host(nq_celt_ctx *ctx, const float *coef, const uint8_t *transient, const nq_celt_post_frame *frames,
                              const float *tail_in, const float *hist_in, const float *mem_in, float *pcm_out,
                              float *tail_out, float *hist_out, float *mem_out, int64_t nframes, int channels, int streams,
                              int coupled_streams, const unsigned char *mapping)
{
    if (!ctx) return NQ_BAD_ARG;
    Layout L;
    if (mapping) {
        int rc = check_layout(ctx, channels, streams, coupled_streams, mapping, &L);
        if (rc != NQ_OK) return rc;
    } else {
        if (channels < 1 || channels > 2) return fail(ctx, NQ_BAD_ARG, "without a mapping: one CELT decoder, channels 1 or 2 (got %d)", channels);
        L = plain_layout(channels);
    }
    if (nframes < 0) return fail(ctx, NQ_BAD_ARG, "nframes=%lld", (long long)nframes);
    const size_t tail_n = (size_t)L.D * kHalfOvl, hist_n = (size_t)L.D * kPostHist;
    if (nframes == 0) {
        if (tail_out) { if (tail_in) memcpy(tail_out, tail_in, 4 * tail_n); else memset(tail_out, 0, 4 * tail_n); }
        if (hist_out) { if (hist_in) memcpy(hist_out, hist_in, 4 * hist_n); else memset(hist_out, 0, 4 * hist_n); }
        if (mem_out) { if (mem_in) memcpy(mem_out, mem_in, 4 * (size_t)L.D); else memset(mem_out, 0, 4 * (size_t)L.D); }
        return NQ_OK;
    }
    if (!coef || !transient || !frames || !pcm_out) return fail(ctx, NQ_BAD_ARG, "null coef/transient/frames/pcm_out");
    const int fcols = L.per_stream_flags ? L.streams : 1;
    for (int64_t f = 0; f < nframes; f++)
        for (int sidx = 0; sidx < L.streams; sidx++) {
            const int N = frames[f * L.streams + sidx].N, flag = transient[f * fcols + (L.per_stream_flags ? sidx : 0)];
            if (N != (kFrame >> ((flag >> 1) & 3)) || (flag >> 4))
                return fail(ctx, NQ_BAD_ARG, "frame %lld stream %d: flag byte 0x%02x does not say N=%d (bit 0 transient, bits 1-2 = 3-LM, bit 3 reset)",
                            (long long)f, sidx, flag, N);
            // offsets and sizes of a frame are taken from its first stream: every stream of a frame must agree
            if (N != frames[f * L.streams].N)
                return fail(ctx, NQ_BAD_ARG, "frame %lld: stream %d carries N=%d, stream 0 N=%d (a multistream packet holds one frame size)",
                            (long long)f, sidx, N, frames[f * L.streams].N);
        }
    HostState st;
    st.tail_in = tail_in; st.tail_out = tail_out;
    st.hist_in = hist_in; st.hist_out = hist_out;
    st.mem_in = mem_in; st.mem_out = mem_out;
    return host_range(ctx, L, coef, transient, frames, st, nullptr, 0, pcm_out, nframes);
}

// Contexts of nq_celt_synth_batch_host_multi: one per device, created on first use and kept (a
// context = streams, events, tables and ~400 MB of staging buffers; creating it per call cost more
// than a small batch).  Calls are serialised by g_multi_mu; nq_celt_multi_release / cleanupCudaBuffers free them.
namespace {
std::mutex g_multi_mu;
std::vector<nq_celt_ctx *> g_multi_ctx;   // index = position in the caller's device list (a device named twice gets two contexts)
}

void nq_celt_multi_release(void)
{
    std::lock_guard<std::mutex> lk(g_multi_mu);
    for (nq_celt_ctx *c : g_multi_ctx) nq_celt_ctx_destroy(c);
    g_multi_ctx.clear();
}

int nq_celt_synth_batch_host_multi(const int *devices, int ndev, const float *coef, const uint8_t *transient,
                                   const float *tail_in, float *pcm_out, float *tail_out, int64_t nframes, int C)
{
    if (ndev < 1 || nframes < 0 || C < 1 || C > 255) return NQ_BAD_ARG;
    if (nframes == 0) {
        if (tail_out) {
            if (tail_in) memcpy(tail_out, tail_in, sizeof(float) * C * kHalfOvl);
            else memset(tail_out, 0, sizeof(float) * C * kHalfOvl);
        }
        return NQ_OK;
    }
    if (!coef || !transient || !pcm_out) return NQ_BAD_ARG;
    if (const long long bad = first_non_20ms_flag(transient, nframes); bad >= 0) {
        fprintf(stderr, "libnq_celt_b200: nq_celt_synth_batch_host_multi: frame %lld: flag byte 0x%02x; 20 ms frames only\n", bad, transient[bad]);
        return NQ_BAD_ARG;
    }
    if (ndev > nframes) ndev = (int)nframes;
    std::lock_guard<std::mutex> lk(g_multi_mu);
    std::vector<int> rcs(ndev, NQ_OK);
    std::vector<std::thread> th;
    const size_t row = (size_t)C * kFrame;
    if ((int)g_multi_ctx.size() < ndev) g_multi_ctx.resize(ndev, nullptr);
    for (int d = 0; d < ndev; d++) {
        th.emplace_back([&, d] {
            const int dev = devices ? devices[d] : d;
            nq_celt_ctx *&ctx = g_multi_ctx[d];   // (distinct elements: no two threads share one)
            if (ctx && ctx->device != dev) {      // another device list than last time
                nq_celt_ctx_destroy(ctx);
                ctx = nullptr;
            }
            int rc = ctx ? NQ_OK : nq_celt_ctx_create(dev, &ctx);
            if (rc == NQ_OK) {
                // contiguous, disjoint frame ranges; no collective, no peer traffic
                const long long f0 = nframes * d / ndev, f1 = nframes * (d + 1) / ndev;
                rc = synth_host_range(ctx, coef + f0 * row, transient + f0, f0 == 0 ? tail_in : nullptr,
                                      f0 > 0 ? coef + (f0 - 1) * row : nullptr, f0 > 0 ? (transient[f0 - 1] & kFlagTransient) : 0,
                                      pcm_out + f0 * row, d == ndev - 1 ? tail_out : nullptr, f1 - f0, C);
                if (rc != NQ_OK) fprintf(stderr, "libnq_celt_b200: device %d: %s\n", dev, ctx->err);
            }
            rcs[d] = rc;
        });
    }
    for (auto &t : th) t.join();
    for (int rc : rcs)
        if (rc != NQ_OK) return rc;
    return NQ_OK;
}

// ------------------------------------------------- single-call entries -----
int nq_celt_mdct_backward_host(nq_celt_ctx *ctx, const float *const *in, float *const *out, int ncalls, int shift,
                               int stride)
{
    if (!ctx) ctx = global_ctx();
    if (ncalls < 1 || shift < 0 || shift > 3 || stride < 1 || !in || !out)
        return fail(ctx, NQ_BAD_ARG, "ncalls=%d shift=%d stride=%d", ncalls, shift, stride);
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    const int N2 = (kMdctN >> shift) >> 1;
    const size_t in_span = (size_t)(N2 - 1) * stride + 1;
    const size_t in_pad = (in_span + 3) & ~(size_t)3, out_len = (size_t)N2 + kHalfOvl;
    const size_t per_call = in_pad + out_len;
    if (per_call * ncalls > ctx->call_buf_cap) {
        cudaFree(ctx->d_call_buf);
        ctx->d_call_buf = nullptr;
        ctx->call_buf_cap = 0;
        if (cudaMalloc(&ctx->d_call_buf, per_call * ncalls * sizeof(float)) != cudaSuccess)
            return fail(ctx, NQ_ALLOC_FAIL, "call scratch");
        ctx->call_buf_cap = per_call * ncalls;
    }
    if (ncalls > ctx->calls_cap) {
        cudaFree(ctx->d_calls);
        ctx->d_calls = nullptr;
        ctx->calls_cap = 0;
        if (cudaMalloc(&ctx->d_calls, sizeof(MdctCall) * ncalls) != cudaSuccess) return fail(ctx, NQ_ALLOC_FAIL, "call table");
        ctx->calls_cap = ncalls;
    }
    std::vector<MdctCall> calls(ncalls);
    cudaStream_t st = ctx->stream;
    for (int i = 0; i < ncalls; i++) {
        float *d_in = ctx->d_call_buf + per_call * i, *d_out = d_in + in_pad;
        calls[i] = MdctCall{d_in, d_out, shift, stride, 0};
        NQ_CUDA(ctx, cudaMemcpyAsync(d_in, in[i], in_span * sizeof(float), cudaMemcpyHostToDevice, st));
        NQ_CUDA(ctx, cudaMemcpyAsync(d_out, out[i], kHalfOvl * sizeof(float), cudaMemcpyHostToDevice, st));
    }
    NQ_CUDA(ctx, cudaMemcpyAsync(ctx->d_calls, calls.data(), sizeof(MdctCall) * ncalls, cudaMemcpyHostToDevice, st));
    NQ_CUDA(ctx, launch_mdct_generic(ctx->d_calls, ncalls, ctx->d_gen, st));
    ctx->launches++;
    for (int i = 0; i < ncalls; i++)
        NQ_CUDA(ctx, cudaMemcpyAsync(out[i], calls[i].out, out_len * sizeof(float), cudaMemcpyDeviceToHost, st));
    NQ_CUDA(ctx, cudaStreamSynchronize(st));
    return NQ_OK;
}

// opus_ifft (kiss_fft.c:696) with the static state kfft[shift]: `count` independent
// N4 = 480 >> shift point unnormalised inverse DFTs, interleaved re/im, host pointers.
int nq_opus_ifft_host(nq_celt_ctx *ctx, int shift, const float *in_ri, float *out_ri, int count)
{
    if (!ctx) ctx = global_ctx();
    if (shift < 0 || shift > 3 || count < 1 || !in_ri || !out_ri || in_ri == out_ri)
        return fail(ctx, NQ_BAD_ARG, "shift=%d count=%d (in-place not supported, kiss_fft.c:707)", shift, count);
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n = (size_t)2 * (480 >> shift);
    if (2 * n * count > ctx->call_buf_cap) {
        cudaFree(ctx->d_call_buf);
        ctx->d_call_buf = nullptr;
        ctx->call_buf_cap = 0;
        if (cudaMalloc(&ctx->d_call_buf, 2 * n * count * sizeof(float)) != cudaSuccess) return fail(ctx, NQ_ALLOC_FAIL, "call scratch");
        ctx->call_buf_cap = 2 * n * count;
    }
    if (count > ctx->calls_cap) {
        cudaFree(ctx->d_calls);
        ctx->d_calls = nullptr;
        ctx->calls_cap = 0;
        if (cudaMalloc(&ctx->d_calls, sizeof(MdctCall) * count) != cudaSuccess) return fail(ctx, NQ_ALLOC_FAIL, "call table");
        ctx->calls_cap = count;
    }
    cudaStream_t st = ctx->stream;
    float *d_in = ctx->d_call_buf, *d_out = ctx->d_call_buf + n * count;
    std::vector<MdctCall> calls(count);
    for (int i = 0; i < count; i++) calls[i] = MdctCall{d_in + n * i, d_out + n * i, shift, 1, 1};
    NQ_CUDA(ctx, cudaMemcpyAsync(d_in, in_ri, n * count * sizeof(float), cudaMemcpyHostToDevice, st));
    NQ_CUDA(ctx, cudaMemcpyAsync(ctx->d_calls, calls.data(), sizeof(MdctCall) * count, cudaMemcpyHostToDevice, st));
    NQ_CUDA(ctx, launch_mdct_generic(ctx->d_calls, count, ctx->d_gen, st));
    ctx->launches++;
    NQ_CUDA(ctx, cudaMemcpyAsync(out_ri, d_out, n * count * sizeof(float), cudaMemcpyDeviceToHost, st));
    NQ_CUDA(ctx, cudaStreamSynchronize(st));
    return NQ_OK;
}

static void check_static_mode(const nq_mdct_lookup *l, int overlap, int shift)
{
    if ((l && l->n != kMdctN) || overlap != kOverlap || shift < 0 || shift > 3) {
        fprintf(stderr, "libnq_celt_b200: only the static 48 kHz mode is supported (mdct n=1920, overlap=120, shift 0..3); "
                        "got n=%d overlap=%d shift=%d\n", l ? l->n : kMdctN, overlap, shift);
        abort();
    }
}

void nq_clt_mdct_backward(const nq_mdct_lookup *l, float *in, float *out, const float *window, int overlap, int shift,
                          int stride)
{
    (void)window;
    check_static_mode(l, overlap, shift);
    nq_celt_ctx *ctx = global_ctx();
    const float *ins[1] = {in};
    float *outs[1] = {out};
    int rc = nq_celt_mdct_backward_host(ctx, ins, outs, 1, shift, stride);
    if (rc != NQ_OK) die(ctx, "clt_mdct_backward", rc);
}

void nq_clt_mdct_backward_B1_C2(const nq_mdct_lookup *l, float *in[2], float *out[2], const float *window, int overlap,
                                int shift, int stride)
{
    (void)window;
    check_static_mode(l, overlap, shift);
    nq_celt_ctx *ctx = global_ctx();
    int rc = nq_celt_mdct_backward_host(ctx, in, out, 2, shift, stride);
    if (rc != NQ_OK) die(ctx, "clt_mdct_backward_B1_C2", rc);
}

int nq_compute_inv_mdcts(nq_celt_ctx *ctx, int shortBlocks, const float *X, float *const *out_mem, int C, int LM)
{
    if (!ctx) ctx = global_ctx();
    if (!X || !out_mem || C < 1 || C > 255 || LM < 0 || LM > 3 || (shortBlocks != 0 && shortBlocks != (1 << LM)))
        return fail(ctx, NQ_BAD_ARG, "shortBlocks=%d C=%d LM=%d", shortBlocks, C, LM);
    // celt_decoder_clean.c:273-284
    int B, N, shift;
    if (shortBlocks) { B = shortBlocks; N = 120; shift = 3; }
    else { B = 1; N = 120 << LM; shift = 3 - LM; }
    NQ_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t xlen = (size_t)N * B, olen = xlen + kHalfOvl;
    const size_t xpad = (xlen + 3) & ~(size_t)3, opad = (olen + 3) & ~(size_t)3;
    const size_t need = (xpad + opad) * C;
    if (need > ctx->call_buf_cap) {
        cudaFree(ctx->d_call_buf);
        ctx->d_call_buf = nullptr;
        ctx->call_buf_cap = 0;
        if (cudaMalloc(&ctx->d_call_buf, need * sizeof(float)) != cudaSuccess) return fail(ctx, NQ_ALLOC_FAIL, "call scratch");
        ctx->call_buf_cap = need;
    }
    if (C * B > ctx->calls_cap) {
        cudaFree(ctx->d_calls);
        ctx->d_calls = nullptr;
        ctx->calls_cap = 0;
        if (cudaMalloc(&ctx->d_calls, sizeof(MdctCall) * C * B) != cudaSuccess) return fail(ctx, NQ_ALLOC_FAIL, "call table");
        ctx->calls_cap = C * B;
    }
    cudaStream_t st = ctx->stream;
    float *d_x = ctx->d_call_buf, *d_o = ctx->d_call_buf + xpad * C;
    NQ_CUDA(ctx, cudaMemcpyAsync(d_x, X, xlen * C * sizeof(float), cudaMemcpyHostToDevice, st));   // X is [C][N*B] contiguous
    std::vector<MdctCall> calls((size_t)C * B);
    for (int c = 0; c < C; c++) {
        NQ_CUDA(ctx, cudaMemcpyAsync(d_o + opad * c, out_mem[c], kHalfOvl * sizeof(float), cudaMemcpyHostToDevice, st));
        for (int b = 0; b < B; b++)   // celt_decoder_clean.c:296-298,309: in = X + b + c*N*B, out = out_mem[c] + N*b, stride B
            calls[(size_t)b * C + c] = MdctCall{d_x + xlen * c + b, d_o + opad * c + (size_t)N * b, shift, B, 0};
    }
    NQ_CUDA(ctx, cudaMemcpyAsync(ctx->d_calls, calls.data(), sizeof(MdctCall) * calls.size(), cudaMemcpyHostToDevice, st));
    // sub-block b+1 consumes the raw tail sub-block b leaves behind: B launches in stream order, C calls each
    for (int b = 0; b < B; b++) {
        NQ_CUDA(ctx, launch_mdct_generic(ctx->d_calls + (size_t)b * C, C, ctx->d_gen, st));
        ctx->launches++;
    }
    for (int c = 0; c < C; c++)
        NQ_CUDA(ctx, cudaMemcpyAsync(out_mem[c], d_o + opad * c, olen * sizeof(float), cudaMemcpyDeviceToHost, st));
    NQ_CUDA(ctx, cudaStreamSynchronize(st));
    return NQ_OK;
}

// ------------------------------------------ fork seam, cuda/mdct_cuda.hpp --
void processMDCTCuda(const float *input, float *output, const float *trig, int N, int shift, int stride, float sine,
                     int overlap, const float *window)
{
    (void)trig; (void)sine; (void)window;
    if (N != (kMdctN >> shift)) {
        fprintf(stderr, "libnq_celt_b200: processMDCTCuda: N=%d does not match 1920 >> shift (%d)\n", N, shift);
        abort();
    }
    check_static_mode(nullptr, overlap, shift);
    nq_celt_ctx *ctx = global_ctx();
    const float *ins[1] = {input};
    float *outs[1] = {output};
    int rc = nq_celt_mdct_backward_host(ctx, ins, outs, 1, shift, stride);
    if (rc != NQ_OK) die(ctx, "processMDCTCuda", rc);
}

void processMDCTCudaB1C2(const float *input[2], float *output[2], const float *trig, int N, int shift, int stride,
                         float sine, int overlap, const float *window)
{
    (void)trig; (void)sine; (void)window;
    if (N != (kMdctN >> shift)) {
        fprintf(stderr, "libnq_celt_b200: processMDCTCudaB1C2: N=%d does not match 1920 >> shift (%d)\n", N, shift);
        abort();
    }
    check_static_mode(nullptr, overlap, shift);
    nq_celt_ctx *ctx = global_ctx();
    int rc = nq_celt_mdct_backward_host(ctx, input, output, 2, shift, stride);
    if (rc != NQ_OK) die(ctx, "processMDCTCudaB1C2", rc);
}

void cleanupCudaBuffers(void)
{
    {
        std::lock_guard<std::mutex> lk(g_mu);
        nq_celt_ctx_destroy(g_ctx);
        g_ctx = nullptr;
    }
    nq_celt_multi_release();
}

void printCudaVersion(void)
{
    int rt = 0, drv = 0;
    cudaRuntimeGetVersion(&rt);
    cudaDriverGetVersion(&drv);
    printf("libnq_celt_b200 (sm_100a): CUDA runtime %d, driver %d\n", rt, drv);
}

}  // extern "C"
