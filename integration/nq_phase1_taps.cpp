// Phase-1 taps -> frame sink.  See nq_phase1_taps.h and overlay/opus/celt/celt_decoder_clean.c.
//
// One decode session at a time per loader thread: the loader brackets its op_read_float loop with
// nq_phase1_begin(sink) / nq_phase1_end().  Inside a session the streams of a multistream packet
// may be decoded by helper threads (OpusDecoderTwoPhase.cpp); a helper binds itself to
// (session, stream) with nq_phase1_bind before it calls into the reference decoder.
#include "nq_phase1_taps.h"
#include "nq_phase1_session.h"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <utility>
#include <vector>

struct nq_phase1_session {
    nq_celt_sink *sink = nullptr;
    std::vector<const void *> decoders;   // unbound calls: CELT decoder states in first-seen order == stream order
    std::atomic<int> max_bound{-1};       // highest stream index a thread was bound to
    std::atomic<bool> silk{false}, mode_switch{false}, lost_celt{false}, short_celt{false};
    std::atomic<int> error{0};
    std::atomic<long long> frames{0};
    std::atomic<int> resets{0};
    bool single_stream = false;           // frames may carry their own place in the output
    std::atomic<bool> positions{false};   // ... and do from now on (a SILK layer or a side frame has turned up)
    std::mutex mu;                        // guards the three below
    std::vector<std::pair<const void *, long long>> opus_pos;   // per OpusDecoder: samples of output so far
    std::vector<nq_phase1_fixup> fixups;
    int nsides = 0;
};

namespace {

struct ThreadState {
    nq_phase1_session *session = nullptr;
    int stream = -1;                      // -1: look the decoder up in session->decoders
    nq_celt_post_frame cur{};             // side info of the frame being assembled (channel 0's calls)
    int calls_c0 = 0;
    long long frame_pos = 0;              // the packet frame opus_decode_frame is working on
    int frame_size = 0;
    int next_kind = 0;                    // the next CELT frame: 0 the frame itself, 1 redundancy (side), 2 fade-out frame
    int last_side = -1;
};

thread_local ThreadState g_t;

}  // namespace

void nq_phase1_begin(nq_celt_sink *sink, int streams)
{
    delete g_t.session;
    g_t = ThreadState();
    g_t.session = new nq_phase1_session();
    g_t.session->sink = sink;
    g_t.session->single_stream = streams == 1;
}

nq_phase1_stats nq_phase1_end(void)
{
    nq_phase1_stats st{};
    if (nq_phase1_session *s = g_t.session) {
        st.frames = s->frames.load();
        const int seen = (int)s->decoders.size();
        const int bound = s->max_bound.load() + 1;
        st.streams_seen = seen > bound ? seen : bound;
        st.saw_silk = s->silk.load() ? 1 : 0;
        st.mode_switch = s->mode_switch.load() ? 1 : 0;
        // celt_decode_with_ec without packet data: loss concealment (also what a mode switch WITHOUT redundancy
        // frames falls back to, opus_decoder_clean.c:318-323, :470-474), which the bundled decoder has no
        // defined behaviour for (celt_decoder_clean.c has no celt_decode_lost)
        st.irregular_celt = s->lost_celt.load() ? 1 : 0;
        st.error = s->error.load();
        st.resets = s->resets.load();
        st.fixups = s->fixups;
        st.samples = s->opus_pos.empty() ? 0 : s->opus_pos[0].second;
        delete s;
    }
    g_t = ThreadState();
    return st;
}

nq_phase1_session *nq_phase1_current(void) { return g_t.session; }

long long nq_phase1_frames_so_far(void) { return g_t.session ? g_t.session->frames.load() : 0; }

int nq_phase1_saw_silk_so_far(void) { return g_t.session && g_t.session->silk.load() ? 1 : 0; }

void nq_phase1_bind(nq_phase1_session *session, int stream)
{
    g_t.session = session;
    g_t.stream = stream;
    g_t.calls_c0 = 0;
    if (session) {
        int seen = session->max_bound.load();
        while (stream > seen && !session->max_bound.compare_exchange_weak(seen, stream)) {}
    }
}

// OPUS_RESET_STATE -> nq_celt_sink_reset_stream.  A decoder the session has not seen a frame from
// yet (its initialisation, or a reset before its first frame) needs nothing: a stream's first
// frame starts from a cleared state anyway.
extern "C" void nq_phase1_reset_tap(const void *dec)
{
    ThreadState &t = g_t;
    nq_phase1_session *s = t.session;
    if (!s || !s->sink) return;
    int stream = t.stream;
    if (stream < 0)
        for (size_t i = 0; i < s->decoders.size(); i++)
            if (s->decoders[i] == dec) stream = (int)i;
    if (stream < 0) return;
    nq_celt_sink_reset_stream(s->sink, stream);
    s->resets.fetch_add(1);
}

extern "C" void nq_phase1_frame_begin(const void *opus_decoder, int audiosize, int mode, int prev_mode)
{
    ThreadState &t = g_t;
    nq_phase1_session *s = t.session;
    if (!s) return;
    if (prev_mode > 0 && prev_mode != mode) s->mode_switch.store(true);
    std::lock_guard<std::mutex> lk(s->mu);
    std::pair<const void *, long long> *e = nullptr;
    for (auto &p : s->opus_pos)
        if (p.first == opus_decoder) e = &p;
    if (!e) {
        s->opus_pos.emplace_back(opus_decoder, 0);
        e = &s->opus_pos.back();
    }
    t.frame_pos = e->second;
    t.frame_size = audiosize;
    t.next_kind = 0;
    e->second += audiosize;
}

extern "C" void nq_phase1_note_silk(int mode, int prev_mode)
{
    nq_phase1_session *s = g_t.session;
    if (!s) return;
    s->silk.store(true);
    if (s->single_stream) s->positions.store(true);   // the CELT decoder no longer covers every stretch of the output
    if (prev_mode > 0 && prev_mode != mode) s->mode_switch.store(true);
}

extern "C" void nq_phase1_note_celt_call(int has_data, int frame_size, int mode, int prev_mode, const char *pcm_arg,
                                         const char *data_arg, int celt_to_silk)
{
    (void)celt_to_silk;
    ThreadState &t = g_t;
    nq_phase1_session *s = t.session;
    if (!s) return;
    if (!has_data) s->lost_celt.store(true);
    if (frame_size < 480) s->short_celt.store(true);
    if (prev_mode > 0 && prev_mode != mode) s->mode_switch.store(true);
    t.next_kind = strncmp(pcm_arg, "redundant_audio", 15) == 0 ? 1 : (strcmp(data_arg, "silence") == 0 ? 2 : 0);
    if (t.next_kind != 0 && s->single_stream) s->positions.store(true);
}

extern "C" void nq_phase1_fade_tap(const char *in1_arg, int audiosize, int overlap)
{
    ThreadState &t = g_t;
    nq_phase1_session *s = t.session;
    if (!s) return;
    nq_phase1_fixup f;
    if (strncmp(in1_arg, "redundant_audio", 15) == 0) f.kind = nq_phase1_fixup::CeltToSilk;
    else if (strncmp(in1_arg, "pcm_transition +", 16) == 0) f.kind = nq_phase1_fixup::Transition;
    else if (strcmp(in1_arg, "pcm_transition") == 0) f.kind = nq_phase1_fixup::TransitionShort;
    else f.kind = nq_phase1_fixup::SilkToCelt;
    f.pos = t.frame_pos;
    f.size = audiosize;
    f.n = overlap;
    f.side = (f.kind == nq_phase1_fixup::CeltToSilk || f.kind == nq_phase1_fixup::SilkToCelt) ? t.last_side : -1;
    std::lock_guard<std::mutex> lk(s->mu);
    s->fixups.push_back(f);
}

extern "C" void nq_phase1_frame_tap(const void *dec, const float *freq, int CC, int N, int LM, int shortBlocks, int c,
                                    int T0, int T1, float g0, float g1, int tapset0, int tapset1)
{
    ThreadState &t = g_t;
    nq_phase1_session *s = t.session;
    if (!s || !s->sink) {
        fprintf(stderr, "nq two-phase decoder: CELT frame decoded outside nq_phase1_begin/end; there is no CPU synthesis in this build\n");
        abort();
    }
    if (c != 0 || s->error.load()) return;   // the calls of the other channel repeat channel 0's arguments
    nq_celt_post_frame &p = t.cur;
    if (t.calls_c0 == 0) {           // celt_decoder_clean.c:663: old -> current over [0, 120)
        p.N = N;
        p.pitch[0] = T0; p.gain[0] = g0; p.tapset[0] = tapset0;
        p.pitch[1] = T1; p.gain[1] = g1; p.tapset[1] = tapset1;
        p.pitch[2] = T1; p.gain[2] = g1; p.tapset[2] = tapset1;
        t.calls_c0 = 1;
        if (LM != 0) return;         // :666 the second call follows
    } else {                          // :666-669: current -> new over [120, N)
        p.pitch[2] = T1; p.gain[2] = g1; p.tapset[2] = tapset1;
    }
    t.calls_c0 = 0;
    int stream = t.stream;
    if (stream < 0) {                 // sequential decode on the session's own thread
        for (size_t i = 0; i < s->decoders.size(); i++)
            if (s->decoders[i] == dec) stream = (int)i;
        if (stream < 0) {
            stream = (int)s->decoders.size();
            s->decoders.push_back(dec);
        }
    }
    // where the frame goes: right after the one before it (CELT-only files, multistream files), or --
    // once a SILK layer or a side frame has turned up in a single-stream file -- to its packet
    // frame's own place / aside (redundancy frames: nq_celt_sink_side_get, cross-faded in by the loader)
    long long dest = -1;
    if (s->positions.load()) {
        if (t.next_kind == 1) {
            std::lock_guard<std::mutex> lk(s->mu);
            t.last_side = s->nsides++;
            dest = -2 - (long long)t.last_side;
        } else {
            dest = t.frame_pos;
        }
    }
    t.next_kind = 0;
    const int rc = nq_celt_sink_push_at(s->sink, stream, freq, CC, N, shortBlocks, &p, dest);
    if (rc != NQ_OK) {
        fprintf(stderr, "nq two-phase decoder: nq_celt_sink_push: %s\n", nq_celt_sink_last_error(s->sink));
        int expected = 0;
        s->error.compare_exchange_strong(expected, rc);
        return;
    }
    s->frames.fetch_add(1);
}
