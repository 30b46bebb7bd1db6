// Frame sink: the hand-over between phase 1 and phase 2 of the restructured decoder.
//
// Reference (third_party/opus/celt/celt_decoder_clean.c): celt_decode_with_ec decodes one frame's
// bitstream into freq[] (:462-636) and immediately synthesises it (:656 compute_inv_mdcts,
// :658-670 comb_filter, :723 deemphasis).  The restructured decoder stops after :636: phase 1
// (the sequential range decoder / PVQ / denormalisation, unchanged, on the CPU) pushes freq[]
// plus the frame's side information into a sink; phase 2 is ONE batched GPU call per flush
// (nq_celt_decode_batch_host) that turns everything pushed so far into float PCM, already
// routed to the output channels of the Opus multistream layout
// (opus_multistream_decoder.c:237-299).
//
// The sink owns: pinned host blocks for coefficients / flags / side info (so the H2D copies of
// phase 2 run at full PCIe rate) and the decoder state phase 2 needs across flushes -- the raw
// IMDCT tail, the comb-filter history and the de-emphasis memory per decoded channel, i.e. what
// the reference keeps in decode_mem / preemph_memD (celt_decoder_clean.c:90-92).
// Host-only code; no kernel lives here.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
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
}

void recycle_block(Block &b)
{
    std::lock_guard<std::mutex> lk(g_pool.mu);
    const size_t n = block_bytes(b.D, b.streams);
    if (g_pool.bytes + n > Pool::kMaxBytes) { free_block(b); return; }
    g_pool.blocks.push_back(b);
    g_pool.bytes += n;
}

bool take_block(int D, int streams, Block *out)
{
    std::lock_guard<std::mutex> lk(g_pool.mu);
    for (size_t i = 0; i < g_pool.blocks.size(); i++)
        if (g_pool.blocks[i].D == D && g_pool.blocks[i].streams == streams) {
            *out = g_pool.blocks[i];
            g_pool.blocks.erase(g_pool.blocks.begin() + i);
            g_pool.bytes -= block_bytes(D, streams);
            return true;
        }
    return false;
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

}  // namespace

struct nq_celt_sink {
    int channels = 0, streams = 0, coupled = 0, D = 0;
    unsigned char mapping[256] = {};
    std::vector<Block> blocks;
    std::vector<long long> pushed;   // frames pushed per stream since the last flush
    // decoder state between flushes (NULL-equivalent = reset decoder)
    std::vector<float> tail, hist, mem;
    bool have_state = false;
    float *out = nullptr;            // pinned output of nq_celt_sink_flush_pinned
    size_t out_bytes = 0;
    char err[256] = {0};
};

namespace {

int sink_fail(nq_celt_sink *s, int code, const char *msg)
{
    if (s) snprintf(s->err, sizeof s->err, "%s", msg);
    return code;
}

bool ensure_block(nq_celt_sink *s, size_t bi)
{
    while (s->blocks.size() <= bi) {
        Block b;
        if (take_block(s->D, s->streams, &b)) {
            s->blocks.push_back(b);
            continue;
        }
        b.D = s->D;
        b.streams = s->streams;
        b.coef = (float *)nq_celt_host_alloc(sizeof(float) * kBlockFrames * s->D * kFrame);
        b.flags = (uint8_t *)nq_celt_host_alloc((size_t)kBlockFrames * s->streams);
        b.post = (nq_celt_post_frame *)nq_celt_host_alloc(sizeof(nq_celt_post_frame) * kBlockFrames * s->streams);
        if (!b.coef || !b.flags || !b.post) {
            free_block(b);
            return false;
        }
        s->blocks.push_back(b);
    }
    return true;
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
    s->tail.assign((size_t)s->D * NQ_CELT_HALF_OVERLAP, 0.f);
    s->hist.assign((size_t)s->D * NQ_CELT_POST_HISTORY, 0.f);
    s->mem.assign(s->D, 0.f);
    *out = s;
    return NQ_OK;
}

void nq_celt_sink_destroy(nq_celt_sink *s)
{
    if (!s) return;
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
    const long long f = s->pushed[stream];
    const size_t bi = (size_t)(f / kBlockFrames), fi = (size_t)(f % kBlockFrames);
    if (!ensure_block(s, bi)) return sink_fail(s, NQ_ALLOC_FAIL, "pinned host memory");
    Block &b = s->blocks[bi];
    const int row = stream < s->coupled ? 2 * stream : stream + s->coupled;
    for (int c = 0; c < nch; c++)   // rows keep the 960-float stride whatever the frame size
        memcpy(b.coef + (fi * s->D + row + c) * kFrame, freq + (size_t)c * N, sizeof(float) * N);
    // a frame with one short block IS a long block of the same size (celt_decoder_clean.c:273-284)
    b.flags[fi * s->streams + stream] = (uint8_t)((shortBlocks > 1 ? 1 : 0) | ((3 - LM) << 1));
    b.post[fi * s->streams + stream] = *post;
    s->pushed[stream] = f + 1;
    return NQ_OK;
}

int64_t nq_celt_sink_pending_frames(const nq_celt_sink *s)
{
    if (!s) return 0;
    long long n = s->pushed[0];
    for (long long v : s->pushed) n = v < n ? v : n;
    return n;
}

int64_t nq_celt_sink_pending_samples(const nq_celt_sink *s)
{
    if (!s) return 0;
    const long long n = nq_celt_sink_pending_frames(s);
    long long total = 0;
    for (long long f = 0; f < n; f++) total += s->blocks[f / kBlockFrames].post[(f % kBlockFrames) * s->streams].N;
    return total;
}

void nq_celt_sink_reset(nq_celt_sink *s)
{
    if (!s) return;
    // OPUS_RESET_STATE, celt_decoder_clean.c:846-859: decode_mem, preemph_memD cleared
    std::fill(s->tail.begin(), s->tail.end(), 0.f);
    std::fill(s->hist.begin(), s->hist.end(), 0.f);
    std::fill(s->mem.begin(), s->mem.end(), 0.f);
    s->have_state = false;
}

int nq_celt_sink_flush(nq_celt_sink *s, nq_celt_ctx *ctx, float *pcm_out, int64_t capacity_samples, int64_t *nsamples)
{
    if (!s || !ctx || !nsamples) return NQ_BAD_ARG;
    *nsamples = 0;
    const long long n = nq_celt_sink_pending_frames(s);
    for (long long v : s->pushed)
        if (v != n) return sink_fail(s, NQ_INVALID_STATE, "streams have pushed different numbers of frames (flush on packet boundaries)");
    if (n == 0) return NQ_OK;
    if (nq_celt_sink_pending_samples(s) > capacity_samples || !pcm_out) return sink_fail(s, NQ_BAD_ARG, "pcm_out too small");
    long long done = 0, out_pos = 0;
    for (size_t bi = 0; done < n; bi++) {
        const long long m = n - done < kBlockFrames ? n - done : kBlockFrames;
        Block &b = s->blocks[bi];
        for (long long f = 0; f < m; f++)
            for (int st = 1; st < s->streams; st++)
                if (b.post[f * s->streams + st].N != b.post[f * s->streams].N)
                    return sink_fail(s, NQ_BAD_ARG, "streams of one multistream packet must share the frame size");
        const bool hs = s->have_state;
        int rc = nq_celt_decode_batch_host(ctx, b.coef, b.flags, b.post, hs ? s->tail.data() : nullptr,
                                           hs ? s->hist.data() : nullptr, hs ? s->mem.data() : nullptr,
                                           pcm_out + out_pos * s->channels, s->tail.data(), s->hist.data(), s->mem.data(), m,
                                           s->channels, s->streams, s->coupled, s->mapping);
        if (rc != NQ_OK) {
            snprintf(s->err, sizeof s->err, "phase 2 failed: %s", nq_celt_last_error(ctx));
            return rc;
        }
        s->have_state = true;
        for (long long f = 0; f < m; f++) out_pos += b.post[f * s->streams].N;
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

}  // extern "C"
