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

#include "../../include/nq_celt_synth.h"

namespace {

constexpr int kFrame = NQ_CELT_FRAME;
constexpr int kBlockFrames = 2048;   // frames per pinned block (15.7 MB of stereo coefficients)

struct Block {
    float *coef = nullptr;                 // [kBlockFrames][D][960]  pinned
    uint8_t *flags = nullptr;              // [kBlockFrames][streams] pinned
    nq_celt_post_frame *post = nullptr;    // [kBlockFrames][streams] pinned
    int D = 0, streams = 0;
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
    if (!b.coef || !b.flags || !b.post) {
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
    long long produced = 0;          // decoded samples per channel so far (worker)
    std::thread worker;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<Job> queue;
    bool stop = false, busy = false;
    int worker_rc = NQ_OK;
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
        if (rc == NQ_OK) {
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
                // decoded sample p (per channel) lands at dst[(p - skip) * channels] if inside the window
                long long a = s->produced, b = s->produced + n;
                long long lo = a > s->skip ? a : s->skip;
                long long hi = b < s->skip + s->dst_samples ? b : s->skip + s->dst_samples;
                if (hi > lo) {
                    float *dst;
                    {   // the destination may be announced after the attach (nq_celt_sink_set_destination)
                        std::unique_lock<std::mutex> lk(s->mu);
                        s->cv.wait(lk, [&] { return s->dst != nullptr || s->stop; });
                        dst = s->dst;
                    }
                    if (dst)
                        memcpy(dst + (size_t)(lo - s->skip) * s->channels, s->out + (size_t)(lo - a) * s->channels,
                               sizeof(float) * (size_t)(hi - lo) * s->channels);
                    else
                        rc = sink_fail(s, NQ_INVALID_STATE, "finished without a destination");
                }
                s->produced = b;
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
    if (!s || !freq || !post) return NQ_BAD_ARG;
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
    s->first_block = 0;
    s->stop = false;
    s->worker_rc = NQ_OK;
    s->worker = std::thread(worker_main, s);
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
