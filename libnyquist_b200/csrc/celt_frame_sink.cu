// Frame sink: the hand-over between phase 1 and phase 2 of the restructured decoder.
//
// Reference (third_party/opus/celt/celt_decoder_clean.c): celt_decode_with_ec decodes one frame's
// bitstream into freq[] (:462-636) and immediately synthesises it (:656 compute_inv_mdcts,
// :658-670 comb_filter, :723 deemphasis).  The restructured decoder stops after :636: phase 1
// (the sequential range decoder / PVQ / denormalisation, unchanged, on the CPU) pushes freq[]
// plus the frame's side information into a sink; phase 2 is a batched GPU call
// (nq_celt_decode_batch_host) that turns pushed frames into float PCM, already routed to the
// output channels of the Opus multistream layout (opus_multistream_decoder.c:237-299).
//
// Two ways to run phase 2:
//   * nq_celt_sink_flush*: one synchronous call over everything pushed so far;
//   * nq_celt_sink_attach + nq_celt_sink_finish: streaming.  Every time a block of 2048 frames is
//     complete a worker thread runs phase 2 on it WHILE phase 1 keeps decoding the next block on
//     the calling thread, and drops the PCM into the caller's buffer (with the positional
//     pre-skip / end-trim window of opusfile.c:2673-2721 applied).  The entropy decoder is the
//     Amdahl bottleneck of a file decode, so hiding phase 2 behind it is
//     what the GPU can contribute end to end.
//
// The sink owns: pinned host blocks for coefficients / flags / side info (so the H2D copies of
// phase 2 run at full PCIe rate) and the decoder state phase 2 needs across blocks -- the raw
// IMDCT tail, the comb-filter history and the de-emphasis memory per decoded channel, i.e. what
// the reference keeps in decode_mem / preemph_memD (celt_decoder_clean.c:90-92).
// Host-only code; no kernel lives here.
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include <atomic>
#include <chrono>

#include "../../include/nq_celt_synth.h"

namespace {

constexpr int kFrame = NQ_CELT_FRAME;
constexpr int kBlockFrames = 2048;   // frames per pinned block (15.7 MB of stereo coefficients)

struct Block {
    float *coef = nullptr;                 // [kBlockFrames][D][960]  pinned
    uint8_t *flags = nullptr;              // [kBlockFrames][streams] pinned
    nq_celt_post_frame *post = nullptr;    // [kBlockFrames][streams] pinned
    long long *dest = nullptr;             // [kBlockFrames] first output sample of the frame, -1: right after the frame before it
    int D = 0, streams = 0;                //                 -2: a side frame (its PCM is kept aside, nq_celt_sink_side_*)
};

// Page-locking host memory costs ~0.3 ms per MB, more than phase 2 itself for a typical file, so
// blocks and output buffers are recycled process-wide instead of being freed with their sink.
struct Pool {
    std::mutex mu;
    std::vector<Block> blocks;
    struct Out { float *p; size_t bytes; };
    std::vector<Out> outs;
    size_t bytes = 0;
    static constexpr size_t kMaxBytes = size_t(2) << 30;
} g_pool;

size_t block_bytes(int D, int streams)
{
    return sizeof(float) * kBlockFrames * D * kFrame + (size_t)kBlockFrames * streams * (1 + sizeof(nq_celt_post_frame));
}

void free_block(Block &b)
{
    nq_celt_host_free(b.coef);
    nq_celt_host_free(b.flags);
    nq_celt_host_free(b.post);
    free(b.dest);
    b = Block();
}

void recycle_block(Block &b)
{
    if (!b.coef) return;
    std::lock_guard<std::mutex> lk(g_pool.mu);
    const size_t n = block_bytes(b.D, b.streams);
    if (g_pool.bytes + n > Pool::kMaxBytes) { free_block(b); return; }
    g_pool.blocks.push_back(b);
    g_pool.bytes += n;
    b = Block();
}

bool take_block(int D, int streams, Block *out)
{
    {
        std::lock_guard<std::mutex> lk(g_pool.mu);
        for (size_t i = 0; i < g_pool.blocks.size(); i++)
            if (g_pool.blocks[i].D == D && g_pool.blocks[i].streams == streams) {
                *out = g_pool.blocks[i];
                g_pool.blocks.erase(g_pool.blocks.begin() + i);
                g_pool.bytes -= block_bytes(D, streams);
                return true;
            }
    }
    Block b;
    b.D = D;
    b.streams = streams;
    b.coef = (float *)nq_celt_host_alloc(sizeof(float) * kBlockFrames * D * kFrame);
    b.flags = (uint8_t *)nq_celt_host_alloc((size_t)kBlockFrames * streams);
    b.post = (nq_celt_post_frame *)nq_celt_host_alloc(sizeof(nq_celt_post_frame) * kBlockFrames * streams);
    b.dest = (long long *)malloc(sizeof(long long) * kBlockFrames);
    if (!b.coef || !b.flags || !b.post || !b.dest) {
        free_block(b);
        return false;
    }
    *out = b;
    return true;
}

float *take_out(size_t bytes, size_t *got)
{
    {
        std::lock_guard<std::mutex> lk(g_pool.mu);
        for (size_t i = 0; i < g_pool.outs.size(); i++)
            if (g_pool.outs[i].bytes >= bytes) {
                float *p = g_pool.outs[i].p;
                *got = g_pool.outs[i].bytes;
                g_pool.bytes -= *got;
                g_pool.outs.erase(g_pool.outs.begin() + i);
                return p;
            }
    }
    *got = bytes;
    return (float *)nq_celt_host_alloc(bytes);
}

void recycle_out(float *p, size_t bytes)
{
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_pool.mu);
    if (g_pool.bytes + bytes > Pool::kMaxBytes) { nq_celt_host_free(p); return; }
    g_pool.outs.push_back({p, bytes});
    g_pool.bytes += bytes;
}

struct Job {
    Block block;
    long long nframes = 0;
};

// Device memory of the many-files path: stream-ordered allocations from the device's default pool,
// which is told to keep what is freed (cudaMalloc / cudaFree of gigabytes per call cost tens of
// milliseconds and synchronise the device).
void keep_pool_memory(int device)
{
    static std::mutex mu;
    static std::vector<char> done;
    std::lock_guard<std::mutex> lk(mu);
    if ((int)done.size() <= device) done.resize(device + 1, 0);
    if (done[device]) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    done[device] = 1;
}

double wall_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

struct nq_celt_sink {
    int channels = 0, streams = 0, coupled = 0, D = 0;
    unsigned char mapping[256] = {};
    std::deque<Block> blocks;        // blocks[k] holds frames [(first_block + k) * 2048, ...)
    long long first_block = 0;       // blocks already handed to the worker (streaming mode)
    std::mutex push_mu;              // pushes of DIFFERENT streams may come from different threads (phase 1
                                     // decodes the streams of a multistream packet in parallel)
    std::vector<long long> pushed;   // frames pushed per stream since the last flush / attach
    std::vector<char> reset_next;    // the stream's next frame follows a decoder reset (flag bit 3)
    // decoder state between phase-2 calls (have_state == false: reset decoder)
    std::vector<float> tail, hist, mem;
    bool have_state = false;
    float *out = nullptr;            // pinned output of nq_celt_sink_flush_pinned / staging of the worker
    size_t out_bytes = 0;
    char err[256] = {0};
    // streaming mode
    nq_celt_ctx *ctx = nullptr;
    float *dst = nullptr;
    long long skip = 0, dst_samples = 0;
    long long produced = 0;          // end of the last frame placed in the output timeline (worker)
    bool placed = false;             // some frame carried an explicit destination or was a side frame
    struct Side { long long tag; int nsamples; std::vector<float> pcm; };
    std::vector<Side> sides;         // side frames in push order (streaming mode)
    std::thread worker;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<Job> queue;
    bool stop = false, busy = false;
    int worker_rc = NQ_OK;
    // upload mode (nq_celt_sink_begin_upload, the many-files loader): the worker only moves every
    // complete block to a device buffer of the sink's own while phase 1 goes on, and phase 2 comes
    // later, for many sinks at once (nq_celt_sink_flush_many)
    bool upload = false;
    cudaStream_t up_stream = nullptr;
    float *d_coef = nullptr;         // [up_cap][D][960]
    uint8_t *d_flags = nullptr;      // [up_cap][streams]
    long long up_cap = 0, up_frames = 0, up_expect = 0;
    int up_device = 0;
    std::vector<nq_celt_post_frame> up_post;   // side info and flags of the uploaded frames stay on the host
    std::vector<uint8_t> up_flags;
};

namespace {

int sink_fail(nq_celt_sink *s, int code, const char *msg)
{
    if (s) snprintf(s->err, sizeof s->err, "%s", msg);
    return code;
}

bool ensure_block(nq_celt_sink *s, size_t k)
{
    while (s->blocks.size() <= k) {
        Block b;
        if (!take_block(s->D, s->streams, &b)) return false;
        s->blocks.push_back(b);
    }
    return true;
}

long long block_samples(const nq_celt_sink *s, const Block &b, long long nframes)
{
    long long total = 0;
    for (long long f = 0; f < nframes; f++) total += b.post[f * s->streams].N;
    return total;
}

// Phase 2 over the first `nframes` frames of a block; PCM to `pcm` ([samples][channels]).
int decode_block(nq_celt_sink *s, nq_celt_ctx *ctx, const Block &b, long long nframes, float *pcm)
{
    for (long long f = 0; f < nframes; f++)
        for (int st = 1; st < s->streams; st++)
            if (b.post[f * s->streams + st].N != b.post[f * s->streams].N)
                return sink_fail(s, NQ_BAD_ARG, "streams of one multistream packet must share the frame size");
    const bool hs = s->have_state;
    const int rc = nq_celt_decode_batch_host(ctx, b.coef, b.flags, b.post, hs ? s->tail.data() : nullptr,
                                             hs ? s->hist.data() : nullptr, hs ? s->mem.data() : nullptr, pcm,
                                             s->tail.data(), s->hist.data(), s->mem.data(), nframes, s->channels,
                                             s->streams, s->coupled, s->mapping);
    if (rc != NQ_OK) {
        snprintf(s->err, sizeof s->err, "phase 2 failed: %s", nq_celt_last_error(ctx));
        return rc;
    }
    s->have_state = true;
    return NQ_OK;
}

// Upload mode: frames [up_frames, up_frames + nframes) of the sink -> its device buffer (grown
// geometrically; the copy returns before the block goes back to the pool).
int upload_block(nq_celt_sink *s, const Block &b, long long nframes)
{
    if (cudaSetDevice(nq_celt_ctx_device(s->ctx)) != cudaSuccess) return sink_fail(s, NQ_INTERNAL_ERROR, "cudaSetDevice");
    const size_t row = sizeof(float) * (size_t)s->D * kFrame;
    if (s->up_frames + nframes > s->up_cap) {
        long long cap = s->up_cap > 0 ? 2 * s->up_cap : (s->up_expect > 0 ? s->up_expect : 4 * kBlockFrames);
        while (cap < s->up_frames + nframes) cap *= 2;
        float *nc = nullptr;
        uint8_t *nf = nullptr;
        if (cudaMallocAsync(&nc, row * cap, s->up_stream) != cudaSuccess || cudaMallocAsync(&nf, (size_t)cap * s->streams, s->up_stream) != cudaSuccess) {
            if (nc) cudaFreeAsync(nc, s->up_stream);
            return sink_fail(s, NQ_ALLOC_FAIL, "device memory for the uploaded frames");
        }
        if (s->up_frames > 0) {
            cudaMemcpyAsync(nc, s->d_coef, row * s->up_frames, cudaMemcpyDeviceToDevice, s->up_stream);
            cudaMemcpyAsync(nf, s->d_flags, (size_t)s->up_frames * s->streams, cudaMemcpyDeviceToDevice, s->up_stream);
            cudaStreamSynchronize(s->up_stream);
        }
        if (s->d_coef) cudaFreeAsync(s->d_coef, s->up_stream);
        if (s->d_flags) cudaFreeAsync(s->d_flags, s->up_stream);
        s->d_coef = nc;
        s->d_flags = nf;
        s->up_cap = cap;
    }
    if (cudaMemcpyAsync(s->d_coef + (size_t)s->up_frames * s->D * kFrame, b.coef, row * nframes, cudaMemcpyHostToDevice, s->up_stream) != cudaSuccess ||
        cudaMemcpyAsync(s->d_flags + (size_t)s->up_frames * s->streams, b.flags, (size_t)nframes * s->streams, cudaMemcpyHostToDevice, s->up_stream) != cudaSuccess)
        return sink_fail(s, NQ_INTERNAL_ERROR, "host to device copy of a block");
    s->up_post.insert(s->up_post.end(), b.post, b.post + (size_t)nframes * s->streams);
    s->up_flags.insert(s->up_flags.end(), b.flags, b.flags + (size_t)nframes * s->streams);
    if (cudaStreamSynchronize(s->up_stream) != cudaSuccess) return sink_fail(s, NQ_INTERNAL_ERROR, "host to device copy of a block");
    s->up_frames += nframes;
    return NQ_OK;
}

void worker_main(nq_celt_sink *s)
{
    for (;;) {
        Job job;
        {
            std::unique_lock<std::mutex> lk(s->mu);
            s->cv.wait(lk, [&] { return s->stop || !s->queue.empty(); });
            if (s->queue.empty()) return;   // stop requested and nothing left
            job = s->queue.front();
            s->queue.pop_front();
            s->busy = true;
        }
        int rc = s->worker_rc;
        if (rc == NQ_OK && s->upload) {
            rc = upload_block(s, job.block, job.nframes);
        } else if (rc == NQ_OK) {
            const long long n = block_samples(s, job.block, job.nframes);
            const size_t need = sizeof(float) * (size_t)n * s->channels;
            if (need > s->out_bytes) {
                recycle_out(s->out, s->out_bytes);
                s->out = take_out(need > (size_t(16) << 20) ? need : (size_t(16) << 20), &s->out_bytes);
                if (!s->out) {
                    s->out_bytes = 0;
                    rc = sink_fail(s, NQ_ALLOC_FAIL, "pinned host memory");
                }
            }
            if (rc == NQ_OK) rc = decode_block(s, s->ctx, job.block, job.nframes, s->out);
            if (rc == NQ_OK) {
                // decoded sample p (per channel) lands at dst[(p - skip) * channels] if inside the window;
                // a frame sits right after the one before it unless it names its own place in the
                // output timeline (files that switch coding modes: the CELT decoder is not called
                // for every stretch of the output, and some of its frames are side frames)
                float *dst = nullptr;
                auto place = [&](long long a, const float *src, long long n) {   // samples [a, a + n) of the timeline
                    long long lo = a > s->skip ? a : s->skip;
                    long long hi = a + n < s->skip + s->dst_samples ? a + n : s->skip + s->dst_samples;
                    if (hi <= lo) return;
                    if (!dst) {   // the destination may be announced after the attach (nq_celt_sink_set_destination)
                        std::unique_lock<std::mutex> lk(s->mu);
                        s->cv.wait(lk, [&] { return s->dst != nullptr || s->stop; });
                        dst = s->dst;
                        if (!dst) {
                            rc = sink_fail(s, NQ_INVALID_STATE, "finished without a destination");
                            return;
                        }
                    }
                    memcpy(dst + (size_t)(lo - s->skip) * s->channels, src + (size_t)(lo - a) * s->channels,
                           sizeof(float) * (size_t)(hi - lo) * s->channels);
                };
                bool plain = true;
                for (long long f = 0; f < job.nframes && plain; f++) plain = job.block.dest[f] == -1;
                if (plain) {
                    place(s->produced, s->out, n);
                    s->produced += n;
                } else {
                    long long off = 0;
                    for (long long f = 0; f < job.nframes && rc == NQ_OK; f++) {
                        const long long N = job.block.post[f * s->streams].N, d = job.block.dest[f];
                        const float *src = s->out + (size_t)off * s->channels;
                        if (d <= -2) {
                            nq_celt_sink::Side sd;
                            sd.tag = -2 - d;
                            sd.nsamples = (int)N;
                            sd.pcm.assign(src, src + (size_t)N * s->channels);
                            s->sides.push_back(std::move(sd));
                        } else {
                            const long long a = d >= 0 ? d : s->produced;
                            place(a, src, N);
                            s->produced = a + N;
                        }
                        off += N;
                    }
                }
            }
        }
        recycle_block(job.block);
        {
            std::lock_guard<std::mutex> lk(s->mu);
            s->worker_rc = rc;
            s->busy = false;
        }
        s->cv.notify_all();
    }
}

void stop_worker(nq_celt_sink *s)
{
    if (!s->worker.joinable()) return;
    {
        std::lock_guard<std::mutex> lk(s->mu);
        s->stop = true;
    }
    s->cv.notify_all();
    s->worker.join();
}

long long min_pushed(const nq_celt_sink *s)
{
    long long n = s->pushed[0];
    for (long long v : s->pushed) n = v < n ? v : n;
    return n;
}

}  // namespace

extern "C" {

int nq_celt_sink_create(nq_celt_sink **out, int channels, int streams, int coupled_streams, const unsigned char *mapping)
{
    if (!out) return NQ_BAD_ARG;
    *out = nullptr;
    // validate_layout, opus_multistream.c:40-55
    if (channels < 1 || channels > 255 || streams < 1 || coupled_streams < 0 || coupled_streams > streams ||
        streams + coupled_streams > 255 || !mapping)
        return NQ_BAD_ARG;
    for (int c = 0; c < channels; c++)
        if (mapping[c] != 255 && mapping[c] >= streams + coupled_streams) return NQ_BAD_ARG;
    nq_celt_sink *s = new (std::nothrow) nq_celt_sink();
    if (!s) return NQ_ALLOC_FAIL;
    s->channels = channels;
    s->streams = streams;
    s->coupled = coupled_streams;
    s->D = streams + coupled_streams;
    memcpy(s->mapping, mapping, channels);
    s->pushed.assign(streams, 0);
    s->reset_next.assign(streams, 0);
    s->tail.assign((size_t)s->D * NQ_CELT_HALF_OVERLAP, 0.f);
    s->hist.assign((size_t)s->D * NQ_CELT_POST_HISTORY, 0.f);
    s->mem.assign(s->D, 0.f);
    *out = s;
    return NQ_OK;
}

void nq_celt_sink_destroy(nq_celt_sink *s)
{
    if (!s) return;
    stop_worker(s);
    for (Job &j : s->queue) recycle_block(j.block);
    for (Block &b : s->blocks) recycle_block(b);
    recycle_out(s->out, s->out_bytes);
    if (s->d_coef || s->d_flags || s->up_stream) {
        cudaSetDevice(s->up_device);
        if (s->up_stream) {
            if (s->d_coef) cudaFreeAsync(s->d_coef, s->up_stream);
            if (s->d_flags) cudaFreeAsync(s->d_flags, s->up_stream);
            cudaStreamSynchronize(s->up_stream);
            cudaStreamDestroy(s->up_stream);
        }
    }
    delete s;
}

void nq_celt_sink_trim_pool(void)
{
    std::lock_guard<std::mutex> lk(g_pool.mu);
    for (Block &b : g_pool.blocks) free_block(b);
    for (auto &o : g_pool.outs) nq_celt_host_free(o.p);
    g_pool.blocks.clear();
    g_pool.outs.clear();
    g_pool.bytes = 0;
}

const char *nq_celt_sink_last_error(const nq_celt_sink *s) { return s ? s->err : ""; }

int nq_celt_sink_push(nq_celt_sink *s, int stream, const float *freq, int CC, int N, int shortBlocks,
                      const nq_celt_post_frame *post)
{
    return nq_celt_sink_push_at(s, stream, freq, CC, N, shortBlocks, post, -1);
}

int nq_celt_sink_push_at(nq_celt_sink *s, int stream, const float *freq, int CC, int N, int shortBlocks,
                         const nq_celt_post_frame *post, int64_t dest)
{
    if (!s || !freq || !post) return NQ_BAD_ARG;
    if (dest != -1) {
        if (s->streams != 1) return sink_fail(s, NQ_UNIMPLEMENTED, "frames with a place of their own: single-stream files only");
        if (!s->ctx || s->upload) return sink_fail(s, NQ_INVALID_STATE, "frames with a place of their own need streaming mode (nq_celt_sink_attach)");
    }
    if (stream < 0 || stream >= s->streams) return sink_fail(s, NQ_BAD_ARG, "stream index out of range");
    const int nch = stream < s->coupled ? 2 : 1;
    if (CC != nch) return sink_fail(s, NQ_BAD_ARG, "channel count of the frame does not match the stream (coupled = 2, mono = 1)");
    int LM = -1;
    for (int k = 0; k < 4; k++)
        if (N == (120 << k)) LM = k;
    if (LM < 0 || post->N != N) return sink_fail(s, NQ_BAD_ARG, "frame size must be 120 << LM and equal post->N");
    if (shortBlocks != 0 && shortBlocks != (1 << LM)) return sink_fail(s, NQ_BAD_ARG, "shortBlocks must be 0 or 1 << LM");
    // The block list is shared between the streams; a frame's rows inside a block are the
    // stream's own.  So: find the block under the lock, copy outside it, publish under the lock.
    // (The block cannot leave the list meanwhile: it is handed to the worker only when EVERY
    // stream has published its last frame in it.)
    long long f;
    size_t fi;
    Block b;
    bool after_reset;
    {
        std::lock_guard<std::mutex> lk(s->push_mu);
        after_reset = s->reset_next[stream] != 0;
        s->reset_next[stream] = 0;
        f = s->pushed[stream];
        const size_t k = (size_t)(f / kBlockFrames - s->first_block);
        fi = (size_t)(f % kBlockFrames);
        if (!ensure_block(s, k)) return sink_fail(s, NQ_ALLOC_FAIL, "pinned host memory");
        b = s->blocks[k];
    }
    const int row = stream < s->coupled ? 2 * stream : stream + s->coupled;
    for (int c = 0; c < nch; c++)   // rows keep the 960-float stride whatever the frame size
        memcpy(b.coef + (fi * s->D + row + c) * kFrame, freq + (size_t)c * N, sizeof(float) * N);
    // a frame with one short block IS a long block of the same size (celt_decoder_clean.c:273-284)
    b.flags[fi * s->streams + stream] = (uint8_t)((shortBlocks > 1 ? 1 : 0) | ((3 - LM) << 1) | (after_reset ? 8 : 0));
    b.post[fi * s->streams + stream] = *post;
    if (stream == 0) b.dest[fi] = dest;
    std::lock_guard<std::mutex> lk(s->push_mu);
    s->pushed[stream] = f + 1;
    // streaming: hand every complete block to the worker
    if (s->ctx && min_pushed(s) >= (s->first_block + 1) * kBlockFrames) {
        Job job;
        job.block = s->blocks.front();
        job.nframes = kBlockFrames;
        s->blocks.pop_front();
        s->first_block++;
        {
            std::lock_guard<std::mutex> qlk(s->mu);
            s->queue.push_back(job);
        }
        s->cv.notify_all();
    }
    return NQ_OK;
}

int64_t nq_celt_sink_pending_frames(const nq_celt_sink *s)
{
    if (!s) return 0;
    return min_pushed(s) - s->first_block * kBlockFrames;
}

int64_t nq_celt_sink_pending_samples(const nq_celt_sink *s)
{
    if (!s) return 0;
    long long n = nq_celt_sink_pending_frames(s), total = 0;
    for (size_t k = 0; n > 0; k++) {
        const long long m = n < kBlockFrames ? n : kBlockFrames;
        total += block_samples(s, s->blocks[k], m);
        n -= m;
    }
    return total;
}

void nq_celt_sink_reset(nq_celt_sink *s)
{
    if (!s) return;
    // OPUS_RESET_STATE, celt_decoder_clean.c:846-859: decode_mem, preemph_memD cleared.  Frames
    // already pushed keep their state; the NEXT frame of every stream carries the reset flag, so a
    // reset may fall anywhere between two pushes (phase 2 zeroes tail / history / memory there).
    std::lock_guard<std::mutex> lk(s->push_mu);
    std::fill(s->reset_next.begin(), s->reset_next.end(), 1);
}

void nq_celt_sink_reset_stream(nq_celt_sink *s, int stream)
{
    if (!s || stream < 0 || stream >= s->streams) return;
    std::lock_guard<std::mutex> lk(s->push_mu);
    s->reset_next[stream] = 1;
}

int nq_celt_sink_flush(nq_celt_sink *s, nq_celt_ctx *ctx, float *pcm_out, int64_t capacity_samples, int64_t *nsamples)
{
    if (!s || !ctx || !nsamples) return NQ_BAD_ARG;
    *nsamples = 0;
    if (s->ctx) return sink_fail(s, NQ_INVALID_STATE, "sink is in streaming mode: use nq_celt_sink_finish");
    const long long n = min_pushed(s);
    for (long long v : s->pushed)
        if (v != n) return sink_fail(s, NQ_INVALID_STATE, "streams have pushed different numbers of frames (flush on packet boundaries)");
    if (n == 0) return NQ_OK;
    if (nq_celt_sink_pending_samples(s) > capacity_samples || !pcm_out) return sink_fail(s, NQ_BAD_ARG, "pcm_out too small");
    long long done = 0, out_pos = 0;
    for (size_t k = 0; done < n; k++) {
        const long long m = n - done < kBlockFrames ? n - done : kBlockFrames;
        const int rc = decode_block(s, ctx, s->blocks[k], m, pcm_out + out_pos * s->channels);
        if (rc != NQ_OK) return rc;
        out_pos += block_samples(s, s->blocks[k], m);
        done += m;
    }
    *nsamples = out_pos;
    std::fill(s->pushed.begin(), s->pushed.end(), 0);
    return NQ_OK;
}

int nq_celt_sink_flush_pinned(nq_celt_sink *s, nq_celt_ctx *ctx, const float **pcm, int64_t *nsamples)
{
    if (!s || !pcm || !nsamples) return NQ_BAD_ARG;
    *pcm = nullptr;
    if (s->ctx) return sink_fail(s, NQ_INVALID_STATE, "sink is in streaming mode: use nq_celt_sink_finish");
    const size_t need = sizeof(float) * (size_t)nq_celt_sink_pending_samples(s) * s->channels;
    if (need > s->out_bytes) {
        recycle_out(s->out, s->out_bytes);
        s->out = take_out(need ? need : 16, &s->out_bytes);
        if (!s->out) {
            s->out_bytes = 0;
            return sink_fail(s, NQ_ALLOC_FAIL, "pinned host memory");
        }
    }
    const int rc = nq_celt_sink_flush(s, ctx, s->out, (int64_t)(s->out_bytes / sizeof(float) / s->channels), nsamples);
    if (rc == NQ_OK) *pcm = s->out;
    return rc;
}

int nq_celt_sink_begin_upload(nq_celt_sink *s, nq_celt_ctx *ctx, int64_t expected_frames)
{
    if (!s || !ctx) return NQ_BAD_ARG;
    if (s->ctx) return sink_fail(s, NQ_INVALID_STATE, "sink already attached");
    if (min_pushed(s) != 0 || s->have_state) return sink_fail(s, NQ_INVALID_STATE, "begin_upload: before the first push of a whole file");
    s->up_device = nq_celt_ctx_device(ctx);
    if (cudaSetDevice(s->up_device) != cudaSuccess) return sink_fail(s, NQ_INTERNAL_ERROR, "cudaSetDevice");
    if (!s->up_stream && cudaStreamCreateWithFlags(&s->up_stream, cudaStreamNonBlocking) != cudaSuccess)
        return sink_fail(s, NQ_INTERNAL_ERROR, "cudaStreamCreate");
    keep_pool_memory(s->up_device);
    s->ctx = ctx;
    s->upload = true;
    s->up_expect = expected_frames > 0 ? expected_frames : 0;
    s->up_frames = 0;
    s->up_post.clear();
    s->up_flags.clear();
    s->dst = nullptr;
    s->first_block = 0;
    s->stop = false;
    s->worker_rc = NQ_OK;
    s->worker = std::thread(worker_main, s);
    return NQ_OK;
}

namespace {

// Upload mode, end of phase 1: the last partial block goes up, the worker stops.
int finish_upload(nq_celt_sink *s)
{
    const long long n = min_pushed(s);
    int rc = NQ_OK;
    for (long long v : s->pushed)
        if (v != n) rc = sink_fail(s, NQ_INVALID_STATE, "streams have pushed different numbers of frames");
    const long long rest = n - s->first_block * kBlockFrames;
    if (rc == NQ_OK && rest > 0) {
        Job job;
        job.block = s->blocks.front();
        job.nframes = rest;
        s->blocks.pop_front();
        s->first_block++;
        {
            std::lock_guard<std::mutex> lk(s->mu);
            s->queue.push_back(job);
        }
        s->cv.notify_all();
    }
    {
        std::unique_lock<std::mutex> lk(s->mu);
        s->cv.wait(lk, [&] { return s->queue.empty() && !s->busy; });
        if (rc == NQ_OK) rc = s->worker_rc;
    }
    stop_worker(s);
    s->ctx = nullptr;
    if (rc == NQ_OK && s->up_frames != n) rc = sink_fail(s, NQ_INTERNAL_ERROR, "upload lost frames");
    return rc;
}

}  // namespace

int nq_celt_sink_flush_many(nq_celt_sink *const *sinks, int nsinks, nq_celt_ctx *ctx, float *const *pcm_out,
                            const int64_t *skip, const int64_t *count, int64_t *decoded)
{
    if (!sinks || nsinks < 1 || !ctx || !pcm_out || !skip || !count) return NQ_BAD_ARG;
    nq_celt_sink *s0 = sinks[0];
    if (!s0) return NQ_BAD_ARG;
    int rc = NQ_OK;
    // ---- what the batch holds ----
    std::vector<long long> nfr(nsinks), first_frame(nsinks + 1, 0), first_sample(nsinks + 1, 0);
    bool any_short = false;
    for (int k = 0; k < nsinks; k++) {
        nq_celt_sink *s = sinks[k];
        if (!s) return NQ_BAD_ARG;
        if (s->channels != s0->channels || s->streams != s0->streams || s->coupled != s0->coupled ||
            memcmp(s->mapping, s0->mapping, s0->channels) != 0)
            return sink_fail(s0, NQ_BAD_ARG, "flush_many: the sinks of one call must share the channel layout");
        if (s->upload) {
            if (s->ctx && finish_upload(s) != NQ_OK) {
                snprintf(s0->err, sizeof s0->err, "flush_many: %s", s->err);
                rc = NQ_INTERNAL_ERROR;
            }
        } else if (s->ctx) {
            return sink_fail(s0, NQ_INVALID_STATE, "flush_many: a sink is in streaming mode");
        }
        if (s->have_state) return sink_fail(s0, NQ_INVALID_STATE, "flush_many: a sink carries state from an earlier flush (whole files only)");
    }
    for (int k = 0; k < nsinks && rc == NQ_OK; k++) {
        nq_celt_sink *s = sinks[k];
        const long long n = s->upload ? s->up_frames : min_pushed(s);
        if (!s->upload)
            for (long long v : s->pushed)
                if (v != n) return sink_fail(s0, NQ_INVALID_STATE, "flush_many: streams have pushed different numbers of frames");
        nfr[k] = n;
        first_frame[k + 1] = first_frame[k] + n;
        long long samples = 0;
        auto scan = [&](const nq_celt_post_frame *post, const uint8_t *flags, long long m, bool head) -> int {
            for (long long f = 0; f < m; f++) {
                const int N = post[f * s->streams].N;
                any_short = any_short || N != kFrame;
                for (int st = 1; st < s->streams; st++)
                    if (post[f * s->streams + st].N != N)
                        return sink_fail(s0, NQ_BAD_ARG, "streams of one multistream packet must share the frame size");
                for (int st = 0; st < s->streams; st++)
                    if ((f > 0 || !head) && (flags[f * s->streams + st] & 8) && s->streams > 1)
                        return sink_fail(s0, NQ_UNIMPLEMENTED, "flush_many: decoder resets inside a multistream file");
                samples += N;
            }
            return NQ_OK;
        };
        if (s->upload) {
            rc = scan(s->up_post.data(), s->up_flags.data(), n, true);
        } else {
            long long done = 0;
            for (size_t b = 0; done < n && rc == NQ_OK; b++) {
                const long long m = n - done < kBlockFrames ? n - done : kBlockFrames;
                rc = scan(s->blocks[b].post, s->blocks[b].flags, m, b == 0);
                done += m;
            }
        }
        first_sample[k + 1] = first_sample[k] + samples;
        if (decoded) decoded[k] = samples;
        if (rc == NQ_OK && (skip[k] < 0 || count[k] < 0 || skip[k] + count[k] > samples || (count[k] > 0 && !pcm_out[k])))
            rc = sink_fail(s0, NQ_BAD_ARG, "flush_many: output window outside the decoded samples");
    }
    const long long T = first_frame[nsinks], S = first_sample[nsinks];
    const int D = s0->D, C = s0->channels, ST = s0->streams;
    float *d_coef = nullptr, *d_pcm = nullptr;
    uint8_t *d_flags = nullptr;
    long long *d_offs = nullptr;
    std::vector<nq_celt_post_frame> post;
    std::vector<int64_t> seg, offs;
    std::vector<uint8_t> head(ST);
    cudaStream_t st = (cudaStream_t)nq_celt_ctx_stream(ctx);
    const int device = nq_celt_ctx_device(ctx);
    const bool timing = getenv("NQ_SINK_TIMING") != nullptr;
    const double tm0 = wall_s();
    double tm_gather = 0, tm_kernels = 0;
    if (rc == NQ_OK && T > 0) {
        post.resize((size_t)T * ST);
        if (cudaSetDevice(device) != cudaSuccess) rc = sink_fail(s0, NQ_INTERNAL_ERROR, "cudaSetDevice");
        keep_pool_memory(device);
        if (rc == NQ_OK &&
            (cudaMallocAsync(&d_coef, sizeof(float) * (size_t)T * D * kFrame, st) != cudaSuccess ||
             cudaMallocAsync(&d_pcm, sizeof(float) * (size_t)S * C + 16, st) != cudaSuccess || cudaMallocAsync(&d_flags, (size_t)T * ST, st) != cudaSuccess ||
             (any_short && cudaMallocAsync(&d_offs, sizeof(long long) * (size_t)(T + 1), st) != cudaSuccess)))
            rc = sink_fail(s0, NQ_ALLOC_FAIL, "flush_many: device memory for the batch");
        // ---- gather: device to device from the sinks that uploaded their blocks during phase 1,
        // host to device from the pinned blocks of the others; every file (and every reset inside a
        // file) opens a segment of the post stage ----
        if (any_short) offs.reserve(T + 1);
        long long pos = 0;
        auto note = [&](const nq_celt_post_frame *p, const uint8_t *fl, long long f0, long long m, bool first_of_file) {
            for (long long f = 0; f < m; f++) {
                if ((fl[f * ST] & 8) || (first_of_file && f == 0)) seg.push_back(f0 + f);
                if (any_short) offs.push_back(pos);
                pos += p[f * ST].N;
            }
            memcpy(post.data() + (size_t)f0 * ST, p, sizeof(nq_celt_post_frame) * (size_t)m * ST);
        };
        bool copy_failed = false;
        for (int k = 0; k < nsinks && rc == NQ_OK; k++) {
            nq_celt_sink *s = sinks[k];
            if (nfr[k] == 0) continue;
            const long long f0 = first_frame[k];
            if (s->upload) {
                note(s->up_post.data(), s->up_flags.data(), f0, nfr[k], true);
                copy_failed |= cudaMemcpyAsync(d_coef + (size_t)f0 * D * kFrame, s->d_coef, sizeof(float) * (size_t)nfr[k] * D * kFrame, cudaMemcpyDeviceToDevice, st) != cudaSuccess;
                copy_failed |= cudaMemcpyAsync(d_flags + (size_t)f0 * ST, s->d_flags, (size_t)nfr[k] * ST, cudaMemcpyDeviceToDevice, st) != cudaSuccess;
            } else {
                long long done = 0;
                for (size_t b = 0; done < nfr[k]; b++) {
                    const long long m = nfr[k] - done < kBlockFrames ? nfr[k] - done : kBlockFrames;
                    const Block &blk = s->blocks[b];
                    note(blk.post, blk.flags, f0 + done, m, b == 0);
                    copy_failed |= cudaMemcpyAsync(d_coef + (size_t)(f0 + done) * D * kFrame, blk.coef, sizeof(float) * (size_t)m * D * kFrame, cudaMemcpyHostToDevice, st) != cudaSuccess;
                    copy_failed |= cudaMemcpyAsync(d_flags + (size_t)(f0 + done) * ST, blk.flags, (size_t)m * ST, cudaMemcpyHostToDevice, st) != cudaSuccess;
                    done += m;
                }
            }
            // the file starts from a reset decoder: flag bit 3 on its first frame, every stream
            const uint8_t *fl0 = s->upload ? s->up_flags.data() : s->blocks[0].flags;
            for (int x = 0; x < ST; x++) head[x] = fl0[x] | 8;
            copy_failed |= cudaMemcpyAsync(d_flags + (size_t)f0 * ST, head.data(), ST, cudaMemcpyHostToDevice, st) != cudaSuccess;
            copy_failed |= cudaStreamSynchronize(st) != cudaSuccess;   // (`head` is reused; the blocks are pinned, the copies short)
        }
        if (copy_failed && rc == NQ_OK) rc = sink_fail(s0, NQ_INTERNAL_ERROR, "flush_many: gathering the batch on the device");
        tm_gather = wall_s();
        seg.push_back(T);
        if (rc == NQ_OK && any_short) {
            offs.push_back(pos);
            if (cudaMemcpyAsync(d_offs, offs.data(), sizeof(long long) * offs.size(), cudaMemcpyHostToDevice, st) != cudaSuccess)
                rc = sink_fail(s0, NQ_INTERNAL_ERROR, "flush_many: host to device copy");
        }
        // ---- ONE synthesis launch, ONE post launch ----
        if (rc == NQ_OK) {
            rc = nq_celt_synth_batch_device_ms(ctx, d_coef, d_flags, nullptr, nullptr, nullptr, d_pcm, nullptr,
                                               any_short ? reinterpret_cast<const int64_t *>(d_offs) : nullptr, T, C, ST, s0->coupled,
                                               s0->mapping, st);
            if (rc == NQ_OK)
                rc = nq_celt_post_segments_device(ctx, d_pcm, post.data(), seg.data(), (int)seg.size() - 1, T, C, ST, s0->coupled,
                                                  s0->mapping, st);
            if (rc != NQ_OK) snprintf(s0->err, sizeof s0->err, "flush_many: phase 2 failed: %s", nq_celt_last_error(ctx));
        }
        // ---- PCM back.  The destinations are ordinary (pageable) memory, where a plain copy is paced by
        // the driver's host-side staging on one thread; so the samples come back in 16 MB pieces through a
        // ring of pinned buffers at the PCIe rate, and a few host threads move the pieces on to their
        // destinations while the next ones are in flight. ----
        if (rc == NQ_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = sink_fail(s0, NQ_INTERNAL_ERROR, "flush_many: CUDA error");
        tm_kernels = wall_s();
        if (rc == NQ_OK) {
            struct Piece { int k; size_t off, n; };   // floats [off, off + n) of file k's window
            std::vector<Piece> pieces;
            const size_t piece_floats = (size_t(16) << 20) / sizeof(float);
            for (int k = 0; k < nsinks; k++)
                for (size_t o = 0, n = (size_t)count[k] * C; o < n; o += piece_floats)
                    pieces.push_back({k, o, n - o < piece_floats ? n - o : piece_floats});
            constexpr int kRing = 12;
            const int nth = 6;
            float *ring[kRing] = {};
            size_t ring_bytes[kRing] = {};
            cudaEvent_t ev[kRing] = {};
            bool ok = true;
            for (int r = 0; r < kRing && ok; r++) {
                ring[r] = take_out(piece_floats * sizeof(float), &ring_bytes[r]);
                ok = ring[r] != nullptr && cudaEventCreateWithFlags(&ev[r], cudaEventDisableTiming) == cudaSuccess;
            }
            if (!ok) rc = sink_fail(s0, NQ_ALLOC_FAIL, "flush_many: pinned staging");
            // piece i uses slot i % kRing; issued[i] / copied[i] hand the slots back and forth
            std::mutex mu;
            std::condition_variable cv;
            size_t issued = 0, next_copy = 0;
            std::vector<char> copied(pieces.size(), 0);
            bool failed = false;
            std::vector<std::thread> th;
            if (rc == NQ_OK)
                for (int t = 0; t < nth; t++)
                    th.emplace_back([&] {
                        cudaSetDevice(device);
                        for (;;) {
                            size_t i;
                            {
                                std::unique_lock<std::mutex> lk(mu);
                                cv.wait(lk, [&] { return failed || next_copy >= pieces.size() || next_copy < issued; });
                                if (failed || next_copy >= pieces.size()) return;
                                i = next_copy++;
                            }
                            const Piece &pc = pieces[i];
                            const bool good = cudaEventSynchronize(ev[i % kRing]) == cudaSuccess;
                            if (good) memcpy(pcm_out[pc.k] + pc.off, ring[i % kRing], pc.n * sizeof(float));
                            {
                                std::lock_guard<std::mutex> lk(mu);
                                copied[i] = 1;
                                if (!good) failed = true;
                            }
                            cv.notify_all();
                        }
                    });
            for (size_t i = 0; i < pieces.size() && rc == NQ_OK; i++) {
                if (i >= kRing) {   // the slot's previous piece must have left it
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return failed || copied[i - kRing]; });
                    if (failed) break;
                }
                const Piece &pc = pieces[i];
                const bool good = cudaMemcpyAsync(ring[i % kRing], d_pcm + (size_t)(first_sample[pc.k] + skip[pc.k]) * C + pc.off, pc.n * sizeof(float),
                                                  cudaMemcpyDeviceToHost, st) == cudaSuccess &&
                                  cudaEventRecord(ev[i % kRing], st) == cudaSuccess;
                {
                    std::lock_guard<std::mutex> lk(mu);
                    if (good) issued = i + 1;
                    else failed = true;
                }
                cv.notify_all();
            }
            {
                std::lock_guard<std::mutex> lk(mu);
                if (issued < pieces.size()) failed = true;   // (nothing more will be issued: let the threads go)
            }
            cv.notify_all();
            for (std::thread &x : th) x.join();
            cudaStreamSynchronize(st);
            if (failed && rc == NQ_OK && !pieces.empty()) {
                bool all = true;
                for (char c : copied) all = all && c;
                if (!all) rc = sink_fail(s0, NQ_INTERNAL_ERROR, "flush_many: device to host copy");
            }
            for (int r = 0; r < kRing; r++) {
                if (ev[r]) cudaEventDestroy(ev[r]);
                recycle_out(ring[r], ring_bytes[r]);
            }
        }
    }
    if (d_coef) cudaFreeAsync(d_coef, st);
    if (d_pcm) cudaFreeAsync(d_pcm, st);
    if (d_flags) cudaFreeAsync(d_flags, st);
    if (d_offs) cudaFreeAsync(d_offs, st);
    if (timing)
        fprintf(stderr, "nq_celt_sink_flush_many: %d files, %lld frames: scan+alloc+gather %.1f ms, kernels %.1f ms, copy back %.1f ms\n",
                nsinks, T, (tm_gather - tm0) * 1e3, (tm_kernels - tm_gather) * 1e3, (wall_s() - tm_kernels) * 1e3);
    for (int k = 0; k < nsinks; k++) {   // empty and reset, whatever happened
        nq_celt_sink *s = sinks[k];
        std::fill(s->pushed.begin(), s->pushed.end(), 0);
        std::fill(s->reset_next.begin(), s->reset_next.end(), 0);
        s->have_state = false;
        s->first_block = 0;
        if (s->upload) {
            s->upload = false;
            s->up_frames = 0;
            s->up_post.clear();
            s->up_flags.clear();
        }
    }
    return rc;
}

int nq_celt_sink_set_destination(nq_celt_sink *s, float *dst)
{
    if (!s || !dst) return NQ_BAD_ARG;
    if (!s->ctx) return sink_fail(s, NQ_INVALID_STATE, "sink is not in streaming mode");
    {
        std::lock_guard<std::mutex> lk(s->mu);
        s->dst = dst;
    }
    s->cv.notify_all();
    return NQ_OK;
}

int nq_celt_sink_attach(nq_celt_sink *s, nq_celt_ctx *ctx, float *dst, int64_t skip_samples, int64_t dst_samples)
{
    // (dst == NULL with dst_samples > 0: the destination follows with nq_celt_sink_set_destination; the
    // worker decodes meanwhile and waits for it before it copies the first block out)
    if (!s || !ctx || skip_samples < 0 || dst_samples < 0) return NQ_BAD_ARG;
    if (s->ctx) return sink_fail(s, NQ_INVALID_STATE, "sink already attached");
    if (min_pushed(s) != 0) return sink_fail(s, NQ_INVALID_STATE, "attach before the first push (or after a flush)");
    s->ctx = ctx;
    s->dst = dst;
    s->skip = skip_samples;
    s->dst_samples = dst_samples;
    s->produced = 0;
    s->sides.clear();
    s->first_block = 0;
    s->stop = false;
    s->worker_rc = NQ_OK;
    s->worker = std::thread(worker_main, s);
    return NQ_OK;
}

int nq_celt_sink_side_count(const nq_celt_sink *s) { return s ? (int)s->sides.size() : 0; }

int nq_celt_sink_side_get(const nq_celt_sink *s, int i, int64_t *tag, int *nsamples, const float **pcm)
{
    if (!s || i < 0 || i >= (int)s->sides.size()) return NQ_BAD_ARG;
    if (tag) *tag = s->sides[i].tag;
    if (nsamples) *nsamples = s->sides[i].nsamples;
    if (pcm) *pcm = s->sides[i].pcm.data();
    return NQ_OK;
}

int nq_celt_sink_finish(nq_celt_sink *s, int64_t *decoded_samples)
{
    if (!s) return NQ_BAD_ARG;
    if (!s->ctx) return sink_fail(s, NQ_INVALID_STATE, "sink is not in streaming mode: use nq_celt_sink_flush");
    const long long n = min_pushed(s);
    int rc = NQ_OK;
    for (long long v : s->pushed)
        if (v != n) rc = sink_fail(s, NQ_INVALID_STATE, "streams have pushed different numbers of frames");
    const long long rest = n - s->first_block * kBlockFrames;
    if (rc == NQ_OK && rest > 0) {   // the last, partial block
        Job job;
        job.block = s->blocks.front();
        job.nframes = rest;
        s->blocks.pop_front();
        s->first_block++;
        {
            std::lock_guard<std::mutex> lk(s->mu);
            s->queue.push_back(job);
        }
        s->cv.notify_all();
    }
    {
        std::unique_lock<std::mutex> lk(s->mu);
        s->cv.wait(lk, [&] { return s->queue.empty() && !s->busy; });
        if (rc == NQ_OK) rc = s->worker_rc;
    }
    stop_worker(s);
    if (decoded_samples) *decoded_samples = s->produced;
    s->ctx = nullptr;
    s->dst = nullptr;
    s->first_block = 0;
    std::fill(s->pushed.begin(), s->pushed.end(), 0);
    return rc;
}

}  // extern "C"
