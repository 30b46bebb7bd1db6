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
//   phase 2  batched inverse MDCT + overlap-add + multistream channel routing + post-filter +
//            de-emphasis on the B200, one call per block of 2048 frames on the sink's worker thread
//            WHILE phase 1 decodes the next block (nq_celt_sink_attach / _finish), float PCM
//            straight into AudioData.samples.
//   then     what the layers above celt_decode_with_ec do to the samples, which is positional:
//            opusfile drops OpusHead.pre_skip samples at the head and trims the end to the final
//            granule position (opusfile.c:2673-2721); opus_decode_frame applies the header gain
//            (opus_decoder_clean.c:578-588).
//
// Scope: single-link, CELT-only streams (every bundled test file).  Chained links and SILK /
// hybrid packets are refused with an exception -- there is no CPU synthesis in this build to fall
// back to.
#include "Decoders.h"
#include "opus/opusfile/include/opusfile.h"

#include <chrono>
#include <cmath>
#include <cstring>
#include <iostream>
#include <mutex>

#include "nq_celt_synth.h"
#include "nq_phase1_session.h"

using namespace nqr;

static const int OPUS_SAMPLE_RATE = 48000;

namespace {

double g_last_timing[3] = {0, 0, 0};   // phase 1, phase 2, trim/gain (seconds) of the last Load

nq_celt_ctx *device_context()
{
    static std::mutex mu;
    static nq_celt_ctx *ctx = nullptr;
    std::lock_guard<std::mutex> lk(mu);
    if (!ctx) {
        int rc = nq_celt_ctx_create(0, &ctx);
        if (rc != NQ_OK)
            throw std::runtime_error(std::string("two-phase Opus decoder: no usable B200 (") + nq_celt_strerror(rc) +
                                     "); this build has no CPU synthesis");
    }
    return ctx;
}

double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct SinkHolder {
    nq_celt_sink *s = nullptr;
    ~SinkHolder() { nq_celt_sink_destroy(s); }
};

}  // namespace

extern "C" void nq_twophase_last_timing(double out[3]) { memcpy(out, g_last_timing, sizeof g_last_timing); }

class OpusDecoderInternal
{
public:
    OpusDecoderInternal(AudioData *d, const std::vector<uint8_t> &fileData) : d(d)
    {
        int err;
        fileHandle = op_test_memory(fileData.data(), fileData.size(), &err);
        if (!fileHandle) throw std::runtime_error("File is not a valid ogg vorbis file");
        if (op_test_open(fileHandle) != 0) {
            fileHandle = nullptr;   // op_test_open frees it on failure
            throw std::runtime_error("Could not open file");
        }
        const OpusHead *header = op_head(fileHandle, 0);

        d->sampleRate = OPUS_SAMPLE_RATE;
        d->channelCount = (uint32_t)header->channel_count;
        d->sourceFormat = MakeFormatForBits(32, true, false);
        const int64_t totalSamples = op_pcm_total(fileHandle, -1);   // samples in a single channel
        d->lengthSeconds = double(uint64_t(totalSamples / OPUS_SAMPLE_RATE));
        d->frameSize = (uint32_t)header->channel_count * GetFormatBitsPerSample(d->sourceFormat);
        d->samples.resize(size_t(totalSamples) * d->channelCount);

        if (op_link_count(fileHandle) != 1)
            throw std::runtime_error("two-phase Opus decoder: chained Ogg Opus streams are not supported");
        if (!decodeTwoPhase(header, totalSamples)) throw std::runtime_error("could not read any data");
    }

    ~OpusDecoderInternal()
    {
        if (fileHandle) op_free(fileHandle);
    }

private:
    bool decodeTwoPhase(const OpusHead *header, int64_t totalSamples)
    {
        const int ch = d->channelCount;
        SinkHolder sink;
        if (nq_celt_sink_create(&sink.s, ch, header->stream_count, header->coupled_count, header->mapping) != NQ_OK)
            throw std::runtime_error("two-phase Opus decoder: unsupported channel layout");

        // ---- phase 1: entropy decode of the whole file, frames -> sink; phase 2 follows block by
        // block on the sink's worker thread and writes straight into d->samples through the
        // positional window opusfile applies: pre_skip samples dropped at the head, the end trimmed
        // to the final granule position (opusfile.c:2673-2721) ----
        const int64_t preSkip = header->pre_skip;
        float *out = d->samples.data();
        if (nq_celt_sink_attach(sink.s, device_context(), out, preSkip, totalSamples) != NQ_OK)
            throw std::runtime_error(std::string("two-phase Opus decoder: ") + nq_celt_sink_last_error(sink.s));
        const double t0 = now_s();
        std::vector<float> placeholder(size_t(5760) * ch);   // 120 ms, the largest Opus packet
        int64_t framesRead = 0;
        bool readError = false;
        nq_phase1_begin(sink.s);
        for (;;) {
            const int n = op_read_float(fileHandle, placeholder.data(), (int)placeholder.size(), nullptr);
            if (n == 0) break;   // EOF
            if (n < 0) {
                std::cerr << "Opus decode error: " << n << std::endl;
                readError = true;
                break;
            }
            framesRead += n;
        }
        const nq_phase1_stats st = nq_phase1_end();
        const double t1 = now_s();

        // ---- phase 2, the rest: last partial block + wait for the worker ----
        int64_t decoded = 0;
        const int rc = nq_celt_sink_finish(sink.s, &decoded);
        const double t2 = now_s();
        if (readError) return false;
        if (st.saw_silk)
            throw std::runtime_error("two-phase Opus decoder: SILK / hybrid packets are not supported (CELT-only streams)");
        if (st.error) throw std::runtime_error(std::string("two-phase Opus decoder: ") + nq_celt_sink_last_error(sink.s));
        if (st.frames && st.streams_seen != header->stream_count)
            throw std::runtime_error("two-phase Opus decoder: stream count mismatch");
        if (rc != NQ_OK)
            throw std::runtime_error(std::string("two-phase Opus decoder: phase 2 failed: ") + nq_celt_sink_last_error(sink.s));
        if (framesRead != totalSamples || decoded < preSkip + totalSamples)
            throw std::runtime_error("two-phase Opus decoder: sample accounting does not match opusfile's");

        // ---- positional post-processing of the layers above the CELT decoder ----
        // header gain: opusfile programs OPUS_SET_GAIN with OpusHead.output_gain (Q8 dB, default
        // OP_HEADER_GAIN), opus_decode_frame scales by celt_exp2(6.48814081e-4f * gain)
        int gainQ8 = header->output_gain;
        if (gainQ8 < -32768) gainQ8 = -32768;
        if (gainQ8 > 32767) gainQ8 = 32767;
        if (gainQ8 != 0) {
            const float gain = (float)std::exp(0.6931471805599453094 * (6.48814081e-4f * gainQ8));
            for (size_t i = 0; i < size_t(totalSamples) * ch; i++) out[i] = out[i] * gain;
        }
        g_last_timing[0] = t1 - t0;
        g_last_timing[1] = t2 - t1;
        g_last_timing[2] = now_s() - t2;
        return totalSamples > 0;
    }

    NO_MOVE(OpusDecoderInternal);
    OggOpusFile *fileHandle = nullptr;
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
