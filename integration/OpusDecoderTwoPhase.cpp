// Two-phase replacement of the reference's src/OpusDecoder.cpp (same public interface:
// nqr::OpusDecoder::LoadFromPath / LoadFromBuffer / GetSupportedFileExtensions, Decoders.h;
// same AudioData contents, Common.h:350-358), compiled INSTEAD of that file.
//
// The reference decodes packet by packet: op_read_float -> ... -> celt_decode_with_ec, which
// entropy-decodes a frame and synthesises it on the spot.  Here:
//
//   phase 1  the same op_read_float loop over the whole file, with the CELT synthesis switched
//            off by integration/overlay: the range decoder / PVQ / denormalisation run unchanged on
//            the CPU and every frame's coefficients + side info land in an nq_celt_sink (pinned
//            host memory).  The PCM opusfile hands back is a placeholder; only its COUNT is used.
//            The streams of a multistream packet are independent decoders
//            (opus_multistream_decoder.c:237-251): with more than one stream, opusfile's decode
//            callback (op_set_decode_callback) splits the packet like opus_multistream_decode_native
//            does and decodes the streams on helper threads at the same time.
//   phase 2  batched inverse MDCT + overlap-add + multistream channel routing + post-filter +
//            de-emphasis on the B200, one call per block of 2048 frames on the sink's worker thread
//            WHILE phase 1 decodes the next block (nq_celt_sink_attach / _finish), float PCM
//            straight into AudioData.samples.
//   then     what the layers above celt_decode_with_ec do to the samples, which is positional:
//            opusfile drops OpusHead.pre_skip samples at the head and trims the end to the final
//            granule position (opusfile.c:2673-2721); opus_decode_frame applies the header gain
//            (opus_decoder_clean.c:578-588).
//
// Scope: single-link files that stay in ONE coding mode:
//   CELT-only   (sb-reverie*.opus, short.opus, the 8-channel file): as above;
//   SILK-only   (test_data/ad_hoc/detodos.opus): such a packet never reaches celt_decode_with_ec
//               (opus_decoder_clean.c:499-513) -- there is no CELT synthesis in it, phase 1, the
//               reference's own SILK decoder on the CPU, IS its whole decode and the PCM opusfile
//               hands back is final;
//   hybrid      the output is a plain sum (opus_decoder_clean.c:553-560): CELT layer above band 17
//               + SILK layer / 32768.  With the CELT synthesis switched off, what opusfile hands
//               back in phase 1 is the SILK layer alone, and the CELT layer comes out of phase 2
//               like any other frame: AudioData.samples = phase-2 PCM + phase-1 PCM.
// A file that SWITCHES modes decodes extra CELT frames into side buffers and cross-fades them in
// (redundancy frames, the fade-out frame, concealment-based transitions: :478-487, :499-513,
// :570-600), which phase 2 does not reproduce: refused with an exception, like chained links and
// lost packets -- there is no CPU CELT synthesis in this build to fall back to.
#include "Decoders.h"
#include "opus/opusfile/include/opusfile.h"

#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iostream>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "nq_celt_synth.h"
#include "nq_phase1_session.h"
#include "OpusBatchLoader.h"

// ---- phase 1 across the streams of a multistream packet ------------------------------------
// Internal entry points of the reference's libopus (opus_private.h:109-111, :140-145; compiled
// into this library with the rest of src/OpusDependencies.c): what opus_multistream_decode_native
// itself uses to split a packet into its self-delimited sub-packets and decode one of them.
extern "C" {
int opus_decode_native(OpusDecoder *st, const unsigned char *data, opus_int32 len, float *pcm, int frame_size,
                       int decode_fec, int self_delimited, opus_int32 *packet_offset, int soft_clip);
int opus_packet_parse_impl(const unsigned char *data, opus_int32 len, int self_delimited, unsigned char *out_toc,
                           const unsigned char *frames[48], opus_int16 size[48], int *payload_offset,
                           opus_int32 *packet_offset);
}

// (before `using namespace nqr`: the ctl macro names ::OpusDecoder unqualified, and nqr has a class of that name)
static OpusDecoder *nq_stream_decoder(OpusMSDecoder *msd, int stream)
{
    OpusDecoder *dec = nullptr;
    if (opus_multistream_decoder_ctl(msd, OPUS_MULTISTREAM_GET_DECODER_STATE(stream, &dec)) != OPUS_OK) return nullptr;
    return dec;
}

using namespace nqr;

static const int OPUS_SAMPLE_RATE = 48000;

namespace {

// Time budget of this thread's last Load, seconds: [0] phase 1 (the op_read_float loop: entropy decode
// on the CPU), [1] phase 2 tail (what is left of the GPU work when phase 1 ends), [2] gain / SILK sum,
// [3] op_test_memory + op_test_open + header, [4] waiting for the zero-filled AudioData.samples (the
// resize runs on a helper thread next to phase 1), [5] context lease + sink + attach, [6] whole
// constructor, [7] the resize itself (overlapped)
thread_local double g_last_timing[8] = {0, 0, 0, 0, 0, 0, 0, 0};

struct OpusFileCloser {
    void operator()(OggOpusFile *f) const { if (f) op_free(f); }
};

// A context is not re-entrant (nq_celt_synth.h), but nqr::NyquistIO::Load may be called from
// several threads at once, as the reference's can: every Load leases a context of its own from a
// process-wide pool (created on demand, kept for the next Load).
class ContextLease
{
public:
    ContextLease()
    {
        {
            std::lock_guard<std::mutex> lk(mu());
            if (!idle().empty()) {
                ctx_ = idle().back();
                idle().pop_back();
                return;
            }
        }
        const int rc = nq_celt_ctx_create(0, &ctx_);
        if (rc != NQ_OK)
            throw std::runtime_error(std::string("two-phase Opus decoder: no usable B200 (") + nq_celt_strerror(rc) +
                                     "); this build has no CPU synthesis");
    }
    ~ContextLease()
    {
        std::lock_guard<std::mutex> lk(mu());
        idle().push_back(ctx_);
    }
    ContextLease(const ContextLease &) = delete;
    ContextLease &operator=(const ContextLease &) = delete;
    nq_celt_ctx *get() const { return ctx_; }

private:
    static std::mutex &mu() { static std::mutex m; return m; }
    static std::vector<nq_celt_ctx *> &idle() { static std::vector<nq_celt_ctx *> v; return v; }
    nq_celt_ctx *ctx_ = nullptr;
};

double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct SinkHolder {
    nq_celt_sink *s = nullptr;
    SinkHolder() = default;
    SinkHolder(const SinkHolder &) = delete;
    SinkHolder &operator=(const SinkHolder &) = delete;
    void reset()
    {
        nq_celt_sink_destroy(s);
        s = nullptr;
    }
    ~SinkHolder() { reset(); }
};

// Helper threads that live for one file: the work items of a packet are tiny (tens of
// microseconds), so the helpers spin on a generation counter instead of sleeping on a condition
// variable, and the loader's thread takes its share of the items.
class StreamPool
{
public:
    explicit StreamPool(int helpers)
    {
        for (int i = 0; i < helpers; i++) threads.emplace_back([this] { helper(); });
    }
    ~StreamPool()
    {
        quit.store(true);
        for (std::thread &t : threads) t.join();
    }
    // fn(item) for item in [0, n), in parallel; returns when all are done.
    template <class F> void run(int n, F &&fn)
    {
        // (no helper is inside work() here: the previous run() waited for `active` to drop to zero)
        job = [&](int i) { fn(i); };
        total.store(n);
        pending.store(n);
        next.store(0);
        generation.fetch_add(1);            // publishes job / total / pending / next to the helpers
        work();
        while (pending.load() > 0 || active.load() > 0) cpu_relax();
    }

private:
    static void cpu_relax()
    {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#else
        std::this_thread::yield();
#endif
    }
    void work()
    {
        for (;;) {
            const int i = next.fetch_add(1);
            if (i >= total.load()) return;
            job(i);
            pending.fetch_sub(1);
        }
    }
    void helper()
    {
        unsigned long long seen = 0;
        int idle = 0;
        while (!quit.load()) {
            const unsigned long long g = generation.load();
            if (g == seen) {
                if (++idle > 20000) std::this_thread::yield(); else cpu_relax();
                continue;
            }
            seen = g;
            idle = 0;
            active.fetch_add(1);
            work();
            active.fetch_sub(1);
        }
        nq_phase1_bind(nullptr, -1);
    }
    std::vector<std::thread> threads;
    std::function<void(int)> job;
    std::atomic<int> total{0}, next{0}, pending{0}, active{0};
    std::atomic<unsigned long long> generation{0};
    std::atomic<bool> quit{false};
};

struct ParallelPhase1 {
    nq_phase1_session *session = nullptr;
    int streams = 0, coupled = 0;
    const unsigned char *mapping = nullptr;   // OpusHead.mapping
    std::unique_ptr<StreamPool> pool;
    std::vector<std::vector<float>> pcm;   // per stream: the placeholder PCM opus_decode_native writes
    struct Item {
        ::OpusDecoder *dec;
        const unsigned char *data;
        opus_int32 len;
        int ret;
    };
    std::vector<Item> items;
};

// op_decode_cb_func (opusfile.h): decode one packet of the link.  Mirrors
// opus_multistream_decode_native (opus_multistream_decoder.c:183-300), channel routing included
// (for CELT packets the PCM of phase 1 is a placeholder, for SILK-only ones it is the output),
// with the per-stream opus_decode_native calls running at the same time.  Anything unusual (a lost packet, a packet
// that does not parse) is left to the reference's own sequential path.
int parallel_decode_cb(void *ctx, OpusMSDecoder *msd, void *pcm, const ogg_packet *op, int nsamples, int nchannels,
                       int format, int /*li*/)
{
    ParallelPhase1 &pp = *static_cast<ParallelPhase1 *>(ctx);
    if (format != OP_DEC_FORMAT_FLOAT || op->bytes < 2 * pp.streams - 1 || nsamples <= 0 || nsamples > 5760)
        return OP_DEC_USE_DEFAULT;
    const unsigned char *data = op->packet;
    opus_int32 len = (opus_int32)op->bytes;
    for (int s = 0; s < pp.streams; s++) {
        unsigned char toc;
        opus_int16 size[48];
        opus_int32 packet_offset = 0;
        if (len <= 0) return OP_DEC_USE_DEFAULT;
        if (opus_packet_parse_impl(data, len, s != pp.streams - 1, &toc, nullptr, size, nullptr, &packet_offset) < 0)
            return OP_DEC_USE_DEFAULT;
        if (opus_packet_get_nb_samples(data, packet_offset, 48000) != nsamples) return OP_DEC_USE_DEFAULT;
        ::OpusDecoder *dec = nq_stream_decoder(msd, s);
        if (!dec) return OP_DEC_USE_DEFAULT;
        pp.items[s] = {dec, data, len, 0};
        data += packet_offset;
        len -= packet_offset;
    }
    pp.pool->run(pp.streams, [&](int s) {
        ParallelPhase1::Item &it = pp.items[s];
        nq_phase1_bind(pp.session, s);
        opus_int32 off = 0;
        it.ret = opus_decode_native(it.dec, it.data, it.len, pp.pcm[s].data(), nsamples, 0, s != pp.streams - 1, &off, 0);
    });
    nq_phase1_bind(pp.session, -1);   // the loader's thread: back to first-seen order for the sequential path
    for (int s = 0; s < pp.streams; s++)
        if (pp.items[s].ret != nsamples) return pp.items[s].ret < 0 ? pp.items[s].ret : OPUS_INVALID_PACKET;
    // channel routing, opus_multistream_decoder.c:260-299 (get_left/right/mono_channel, opus_multistream.c:57-91)
    float *out = static_cast<float *>(pcm);
    for (int c = 0; c < nchannels; c++) {
        const int d = pp.mapping[c];
        if (d == 255) {
            for (int i = 0; i < nsamples; i++) out[(size_t)i * nchannels + c] = 0.f;
            continue;
        }
        const bool is_coupled = d < 2 * pp.coupled;
        const int s = is_coupled ? d / 2 : d - pp.coupled, stride = is_coupled ? 2 : 1;
        const float *src = pp.pcm[s].data() + (is_coupled ? (d & 1) : 0);
        for (int i = 0; i < nsamples; i++) out[(size_t)i * nchannels + c] = src[(size_t)i * stride];
    }
    return 0;
}

int phase1_threads(int streams)
{
    int want = streams;
    if (const char *e = getenv("NQ_PHASE1_THREADS")) want = atoi(e);   // 1: the sequential reference path
    const int hw = (int)std::thread::hardware_concurrency();
    if (hw > 0 && want > hw) want = hw;
    if (want > streams) want = streams;
    return want < 1 ? 1 : want;
}

}  // namespace

extern "C" void nq_twophase_last_timing(double out[8]) { memcpy(out, g_last_timing, sizeof g_last_timing); }

class OpusDecoderInternal
{
public:
    OpusDecoderInternal(AudioData *d, const std::vector<uint8_t> &fileData) : d(d)
    {
        const double tc0 = now_s();
        int err;
        // (owned from here on: a refused or failed file must not leak the OggOpusFile)
        fileHandle.reset(op_test_memory(fileData.data(), fileData.size(), &err));
        if (!fileHandle) throw std::runtime_error("File is not a valid ogg vorbis file");
        if (op_test_open(fileHandle.get()) != 0) {
            fileHandle.release();   // op_test_open frees it on failure
            throw std::runtime_error("Could not open file");
        }
        const OpusHead *header = op_head(fileHandle.get(), 0);

        d->sampleRate = OPUS_SAMPLE_RATE;
        d->channelCount = (uint32_t)header->channel_count;
        d->sourceFormat = MakeFormatForBits(32, true, false);
        const int64_t totalSamples = op_pcm_total(fileHandle.get(), -1);   // samples in a single channel
        d->lengthSeconds = double(uint64_t(totalSamples / OPUS_SAMPLE_RATE));
        d->frameSize = (uint32_t)header->channel_count * GetFormatBitsPerSample(d->sourceFormat);
        if (op_link_count(fileHandle.get()) != 1)
            throw std::runtime_error("two-phase Opus decoder: chained Ogg Opus streams are not supported");
        // samples.resize() zero-fills (86 MB for the bundled 224 s stereo file: tens of milliseconds of
        // page faults); phase 1 does not need the buffer until its first block of 2048 frames is
        // through phase 2, so the resize runs on a helper thread meanwhile
        double resizeSeconds = 0;
        std::thread resizer([&] {
            const double r0 = now_s();
            d->samples.resize(size_t(totalSamples) * d->channelCount);
            resizeSeconds = now_s() - r0;
            resizerDone.store(true);
        });
        struct Joiner {
            std::thread &t;
            ~Joiner() { if (t.joinable()) t.join(); }
        } joinResizer{resizer};
        g_last_timing[3] = now_s() - tc0;
        const bool ok = decodeTwoPhase(header, totalSamples, resizer);
        g_last_timing[6] = now_s() - tc0;
        g_last_timing[7] = resizeSeconds;
        if (!ok) throw std::runtime_error("could not read any data");
    }

private:
    bool decodeTwoPhase(const OpusHead *header, int64_t totalSamples, std::thread &resizer)
    {
        const double ts0 = now_s();
        const int ch = d->channelCount;
        ContextLease ctx;   // (declared before the sink: the sink's worker thread uses it until the sink is destroyed)
        SinkHolder sink;
        if (nq_celt_sink_create(&sink.s, ch, header->stream_count, header->coupled_count, header->mapping) != NQ_OK)
            throw std::runtime_error("two-phase Opus decoder: unsupported channel layout");

        // ---- phase 1: entropy decode of the whole file, frames -> sink; phase 2 follows block by
        // block on the sink's worker thread and writes straight into d->samples through the
        // positional window opusfile applies: pre_skip samples dropped at the head, the end trimmed
        // to the final granule position (opusfile.c:2673-2721) ----
        const int64_t preSkip = header->pre_skip;
        // (destination announced below, when the zero-filled samples exist)
        if (nq_celt_sink_attach(sink.s, ctx.get(), nullptr, preSkip, totalSamples) != NQ_OK)
            throw std::runtime_error(std::string("two-phase Opus decoder: ") + nq_celt_sink_last_error(sink.s));
        const double t0 = now_s();
        g_last_timing[5] = t0 - ts0;
        bool haveOut = false;
        auto announceOut = [&] {
            if (haveOut) return;
            const double w0 = now_s();
            resizer.join();
            g_last_timing[4] = now_s() - w0;
            if (totalSamples > 0) nq_celt_sink_set_destination(sink.s, d->samples.data());
            haveOut = true;
        };
        std::vector<float> placeholder(size_t(5760) * ch);   // 120 ms, the largest Opus packet
        // phase-1 PCM from sample cpuStart on, kept while no CELT frame has turned up (a SILK-only file:
        // this IS the output) and from the first SILK layer on (hybrid files, files that switch modes:
        // the SILK share of the output, fades included)
        std::vector<float> cpuPcm;
        int64_t cpuStart = 0;
        int64_t framesRead = 0;
        bool readError = false;
        nq_phase1_begin(sink.s, header->stream_count);
        ParallelPhase1 pp;
        const int nthreads = phase1_threads(header->stream_count);
        if (nthreads > 1) {
            pp.session = nq_phase1_current();
            pp.streams = header->stream_count;
            pp.coupled = header->coupled_count;
            pp.mapping = header->mapping;
            pp.pcm.assign(pp.streams, std::vector<float>(size_t(5760) * 2));
            pp.items.resize(pp.streams);
            pp.pool.reset(new StreamPool(nthreads - 1));
            op_set_decode_callback(fileHandle.get(), parallel_decode_cb, &pp);
        }
        for (;;) {
            const int n = op_read_float(fileHandle.get(), placeholder.data(), (int)placeholder.size(), nullptr);
            if (!haveOut && resizerDone.load()) announceOut();
            if (n == 0) break;   // EOF
            if (n < 0) {
                std::cerr << "Opus decode error: " << n << std::endl;
                readError = true;
                break;
            }
            if (nq_phase1_frames_so_far() == 0 || nq_phase1_saw_silk_so_far()) {
                if (cpuPcm.empty()) cpuStart = framesRead;
                cpuPcm.insert(cpuPcm.end(), placeholder.begin(), placeholder.begin() + size_t(n) * ch);
            } else if (!cpuPcm.empty()) {
                std::vector<float>().swap(cpuPcm);   // a CELT-only file after all: its placeholder PCM is silence
            }
            framesRead += n;
        }
        op_set_decode_callback(fileHandle.get(), nullptr, nullptr);
        pp.pool.reset();   // helpers joined before the session goes away
        const nq_phase1_stats st = nq_phase1_end();
        const double t1 = now_s();
        announceOut();
        float *out = d->samples.data();

        // ---- phase 2, the rest: last partial block + wait for the worker ----
        int64_t decoded = 0;
        const int rc = nq_celt_sink_finish(sink.s, &decoded);
        const double t2 = now_s();
        if (readError) return false;
        if (st.saw_silk && st.frames == 0 && rc == NQ_OK && !st.error) {
            // SILK-only file: no CELT frame anywhere, nothing for phase 2; opusfile has already
            // applied pre-skip, end trim and the header gain to what it returned
            if (framesRead != totalSamples || cpuPcm.size() != size_t(totalSamples) * ch)
                throw std::runtime_error("two-phase Opus decoder: sample accounting does not match opusfile's");
            memcpy(out, cpuPcm.data(), sizeof(float) * cpuPcm.size());
            g_last_timing[0] = t1 - t0;
            g_last_timing[1] = t2 - t1;
            g_last_timing[2] = 0;
            return totalSamples > 0;
        }
        if (st.irregular_celt)
            throw std::runtime_error("two-phase Opus decoder: lost packets and mode switches without redundancy frames need loss concealment, which the bundled decoder does not define");
        if (st.saw_silk && st.mode_switch && header->stream_count != 1)
            throw std::runtime_error("two-phase Opus decoder: multistream files that switch between the SILK, hybrid and CELT coding modes are not supported");
        if (st.error) throw std::runtime_error(std::string("two-phase Opus decoder: ") + nq_celt_sink_last_error(sink.s));
        if (st.frames && st.streams_seen != header->stream_count)
            throw std::runtime_error("two-phase Opus decoder: stream count mismatch");
        if (rc != NQ_OK)
            throw std::runtime_error(std::string("two-phase Opus decoder: phase 2 failed: ") + nq_celt_sink_last_error(sink.s));
        // CELT-only: phase 2 produced every sample; with a SILK layer the CELT decoder skips stretches
        // of the output, and the packet frames' own accounting must cover the file
        if (framesRead != totalSamples || (st.saw_silk ? st.samples : decoded) < preSkip + totalSamples)
            throw std::runtime_error("two-phase Opus decoder: sample accounting does not match opusfile's");

        // ---- what the layers above the CELT decoder do to its output; everything here is linear in it ----
        // the fades of a mode switch, opus_decoder_clean.c:530-575 (window = the CELT mode's, squared)
        if (!st.fixups.empty()) {
            float window[120];
            nq_celt_debug_tables(nullptr, nullptr, window, nullptr);
            auto at = [&](int64_t p) -> float * {   // sample p of the output timeline, or null outside the file's window
                const int64_t i = p - preSkip;
                return (i >= 0 && i < totalSamples) ? out + size_t(i) * ch : nullptr;
            };
            for (const nq_phase1_fixup &f : st.fixups) {
                const float *side = nullptr;
                if (f.side >= 0) {
                    int64_t tag = -1;
                    int ns = 0;
                    if (nq_celt_sink_side_get(sink.s, f.side, &tag, &ns, &side) != NQ_OK || tag != f.side || ns < 2 * f.n)
                        throw std::runtime_error("two-phase Opus decoder: redundancy frame accounting");
                }
                if (f.n > 120) throw std::runtime_error("two-phase Opus decoder: unexpected fade length");
                for (int i = 0; i < f.n; i++) {
                    const float w = window[i] * window[i];
                    switch (f.kind) {
                    case nq_phase1_fixup::CeltToSilk:
                        if (float *o = at(f.pos + i))
                            for (int c = 0; c < ch; c++) o[c] = side[size_t(i) * ch + c];
                        if (float *o = at(f.pos + f.n + i))
                            for (int c = 0; c < ch; c++) o[c] = w * o[c] + (1.f - w) * side[size_t(f.n + i) * ch + c];
                        break;
                    case nq_phase1_fixup::SilkToCelt:
                        if (float *o = at(f.pos + f.size - f.n + i))
                            for (int c = 0; c < ch; c++) o[c] = w * side[size_t(f.n + i) * ch + c] + (1.f - w) * o[c];
                        break;
                    case nq_phase1_fixup::Transition:
                        if (float *o = at(f.pos + i))
                            for (int c = 0; c < ch; c++) o[c] = 0.f;
                        if (float *o = at(f.pos + f.n + i))
                            for (int c = 0; c < ch; c++) o[c] = w * o[c];
                        break;
                    case nq_phase1_fixup::TransitionShort:
                        if (float *o = at(f.pos + i))
                            for (int c = 0; c < ch; c++) o[c] = w * o[c];
                        break;
                    }
                }
            }
        }
        // header gain: opusfile programs OPUS_SET_GAIN with OpusHead.output_gain (Q8 dB, default
        // OP_HEADER_GAIN), opus_decode_frame scales by celt_exp2(6.48814081e-4f * gain)
        int gainQ8 = header->output_gain;
        if (gainQ8 < -32768) gainQ8 = -32768;
        if (gainQ8 > 32767) gainQ8 = 32767;
        if (gainQ8 != 0) {
            const float gain = (float)std::exp(0.6931471805599453094 * (6.48814081e-4f * gainQ8));
            for (size_t i = 0; i < size_t(totalSamples) * ch; i++) out[i] = out[i] * gain;
        }
        if (st.saw_silk) {
            // + the SILK share, which phase 1 decoded, faded, trimmed and scaled by the header gain
            // already (opus_decoder_clean.c:553-560: pcm = celt + silk / 32768, then the fades, then the gain)
            if (cpuStart * ch + int64_t(cpuPcm.size()) != totalSamples * ch)
                throw std::runtime_error("two-phase Opus decoder: sample accounting does not match opusfile's");
            float *o = out + size_t(cpuStart) * ch;
            for (size_t i = 0; i < cpuPcm.size(); i++) o[i] = o[i] + cpuPcm[i];
        }
        g_last_timing[0] = t1 - t0;
        g_last_timing[1] = t2 - t1;
        g_last_timing[2] = now_s() - t2;
        return totalSamples > 0;
    }

    NO_MOVE(OpusDecoderInternal);
    std::unique_ptr<OggOpusFile, OpusFileCloser> fileHandle;
    std::atomic<bool> resizerDone{false};
    AudioData *d;
};

//////////////////////
// Public Interface //
//////////////////////

void nqr::OpusDecoder::LoadFromPath(AudioData *data, const std::string &path)
{
    auto fileBuffer = nqr::ReadFile(path);
    OpusDecoderInternal decoder(data, fileBuffer.buffer);
}

void nqr::OpusDecoder::LoadFromBuffer(AudioData *data, const std::vector<uint8_t> &memory)
{
    OpusDecoderInternal decoder(data, memory);
}

std::vector<std::string> nqr::OpusDecoder::GetSupportedFileExtensions()
{
    return {"opus"};
}

////////////////////////
// Many files at once //
////////////////////////

namespace {

struct BatchFile {
    std::string path;
    AudioData *d = nullptr;
    SinkHolder sink;
    int64_t preSkip = 0, totalSamples = 0;
    int gainQ8 = 0;
    std::string layoutKey;     // files with equal keys share one phase-2 batch
    bool batched = false;      // false: goes through the ordinary loader
    long long frames = 0;      // CELT frames (per stream)
    std::string error;
};

// Phase 1 of one file into a sink of its own; phase 2 comes later, for all files at once -- only the
// upload of the coefficients to the device runs along (nq_celt_sink_begin_upload).
void batch_phase1(BatchFile &bf, nq_celt_ctx *ctx)
{
    NyquistFileBuffer file = nqr::ReadFile(bf.path);
    int err;
    std::unique_ptr<OggOpusFile, OpusFileCloser> fh(op_test_memory(file.buffer.data(), file.buffer.size(), &err));
    if (!fh) throw std::runtime_error("File is not a valid ogg vorbis file");
    if (op_test_open(fh.get()) != 0) {
        fh.release();
        throw std::runtime_error("Could not open file");
    }
    const OpusHead *header = op_head(fh.get(), 0);
    AudioData *d = bf.d;
    d->sampleRate = OPUS_SAMPLE_RATE;
    d->channelCount = (uint32_t)header->channel_count;
    d->sourceFormat = MakeFormatForBits(32, true, false);
    bf.totalSamples = op_pcm_total(fh.get(), -1);
    d->lengthSeconds = double(uint64_t(bf.totalSamples / OPUS_SAMPLE_RATE));
    d->frameSize = (uint32_t)header->channel_count * GetFormatBitsPerSample(d->sourceFormat);
    if (op_link_count(fh.get()) != 1) return;   // not batched: the ordinary loader reports it
    bf.preSkip = header->pre_skip;
    bf.gainQ8 = header->output_gain;
    const int ch = header->channel_count;
    if (nq_celt_sink_create(&bf.sink.s, ch, header->stream_count, header->coupled_count, header->mapping) != NQ_OK) return;
    bf.layoutKey = std::to_string(ch) + "/" + std::to_string(header->stream_count) + "/" + std::to_string(header->coupled_count) + "/" +
                   std::string(reinterpret_cast<const char *>(header->mapping), (size_t)ch);
    if (nq_celt_sink_begin_upload(bf.sink.s, ctx, (bf.totalSamples + bf.preSkip) / 960 + 8) != NQ_OK)
        throw std::runtime_error(std::string("two-phase Opus decoder: ") + nq_celt_sink_last_error(bf.sink.s));
    std::vector<float> placeholder(size_t(5760) * ch);
    int64_t framesRead = 0;
    nq_phase1_begin(bf.sink.s);
    for (;;) {
        const int n = op_read_float(fh.get(), placeholder.data(), (int)placeholder.size(), nullptr);
        if (n <= 0) {
            if (n < 0) framesRead = -1;
            break;
        }
        framesRead += n;
    }
    const nq_phase1_stats st = nq_phase1_end();
    // the batched phase 2 covers what a CELT-only file needs; SILK layers, mode switches and
    // anything irregular stay with the ordinary loader
    bf.batched = framesRead == bf.totalSamples && !st.saw_silk && !st.mode_switch && !st.irregular_celt && !st.error &&
                 st.frames > 0 && st.streams_seen == header->stream_count;
    bf.frames = st.frames / header->stream_count;
    if (bf.batched) d->samples.resize(size_t(bf.totalSamples) * ch);   // (zero-fill on this file's own thread)
}

}   // namespace

nqr::OpusBatchStats nqr::LoadOpusBatch(const std::vector<std::string> &paths, std::vector<std::shared_ptr<AudioData>> &out, int threads)
{
    OpusBatchStats stats;
    const double t0 = now_s();
    const int K = (int)paths.size();
    out.clear();
    std::vector<BatchFile> files(K);
    for (int i = 0; i < K; i++) {
        out.push_back(std::make_shared<AudioData>());
        files[i].path = paths[i];
        files[i].d = out[i].get();
    }
    ContextLease ctx;
    // ---- phase 1 of all files, `threads` at a time ----
    int nthreads = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    if (nthreads < 1) nthreads = 1;
    if (nthreads > K) nthreads = K;
    std::atomic<int> next{0};
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; t++)
        pool.emplace_back([&] {
            for (;;) {
                const int i = next.fetch_add(1);
                if (i >= K) return;
                try {
                    batch_phase1(files[i], ctx.get());
                } catch (const std::exception &e) {
                    files[i].error = e.what();
                }
            }
        });
    for (std::thread &t : pool) t.join();
    for (const BatchFile &bf : files)
        if (!bf.error.empty()) throw std::runtime_error(bf.path + ": " + bf.error);
    const double t1 = now_s();
    stats.phase1Seconds = t1 - t0;
    // ---- phase 2: one batch per channel layout ----
    const long long launches0 = nq_celt_launch_count(ctx.get());
    std::vector<char> done(K, 0);
    for (int i = 0; i < K; i++) {
        if (done[i] || !files[i].batched) continue;
        std::vector<int> group;
        for (int j = i; j < K; j++)
            if (!done[j] && files[j].batched && files[j].layoutKey == files[i].layoutKey) group.push_back(j);
        std::vector<nq_celt_sink *> sinks;
        std::vector<float *> dst;
        std::vector<int64_t> skip, count, decoded(group.size(), 0);
        for (int j : group) {
            sinks.push_back(files[j].sink.s);
            dst.push_back(files[j].d->samples.data());
            skip.push_back(files[j].preSkip);
            count.push_back(files[j].totalSamples);
            stats.frames += files[j].frames;
            done[j] = 1;
        }
        const int rc = nq_celt_sink_flush_many(sinks.data(), (int)sinks.size(), ctx.get(), dst.data(), skip.data(), count.data(), decoded.data());
        if (rc != NQ_OK)
            throw std::runtime_error(std::string("two-phase Opus decoder: batched phase 2 failed: ") + nq_celt_sink_last_error(sinks[0]));
        for (int j : group) {
            BatchFile &bf = files[j];
            int gainQ8 = bf.gainQ8 < -32768 ? -32768 : (bf.gainQ8 > 32767 ? 32767 : bf.gainQ8);
            if (gainQ8 != 0) {   // opus_decoder_clean.c:578-588
                const float gain = (float)std::exp(0.6931471805599453094 * (6.48814081e-4f * gainQ8));
                for (float &v : bf.d->samples) v *= gain;
            }
            stats.filesBatched++;
        }
    }
    stats.launches = (int)(nq_celt_launch_count(ctx.get()) - launches0);
    const double t2 = now_s();
    stats.phase2Seconds = t2 - t1;
    // ---- the files the batch does not cover: the ordinary two-phase loader, one by one ----
    for (int i = 0; i < K; i++) {
        if (files[i].batched) continue;
        files[i].sink.reset();
        *files[i].d = AudioData();
        nqr::OpusDecoder().LoadFromPath(files[i].d, files[i].path);
        stats.filesSingle++;
    }
    stats.totalSeconds = now_s() - t0;
    return stats;
}
