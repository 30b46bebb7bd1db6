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
#include <vector>

struct nq_phase1_session {
    nq_celt_sink *sink = nullptr;
    std::vector<const void *> decoders;   // unbound calls: CELT decoder states in first-seen order == stream order
    std::atomic<int> max_bound{-1};       // highest stream index a thread was bound to
    std::atomic<bool> silk{false}, mode_switch{false}, lost_celt{false}, short_celt{false};
    std::atomic<int> error{0};
    std::atomic<long long> frames{0};
    std::atomic<int> resets{0};
};

namespace {

struct ThreadState {
    nq_phase1_session *session = nullptr;
    int stream = -1;                      // -1: look the decoder up in session->decoders
    nq_celt_post_frame cur{};             // side info of the frame being assembled (channel 0's calls)
    int calls_c0 = 0;
};

thread_local ThreadState g_t;

}  // namespace

void nq_phase1_begin(nq_celt_sink *sink)
{
    delete g_t.session;
    g_t = ThreadState();
    g_t.session = new nq_phase1_session();
    g_t.session->sink = sink;
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
        // (a CELT-only file may hold 2.5 / 5 ms frames of its own; next to SILK such calls are the mode-switch frames)
        st.irregular_celt = (s->lost_celt.load() || (s->silk.load() && s->short_celt.load())) ? 1 : 0;
        st.error = s->error.load();
        st.resets = s->resets.load();
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

extern "C" void nq_phase1_note_silk(int mode, int prev_mode)
{
    nq_phase1_session *s = g_t.session;
    if (!s) return;
    s->silk.store(true);
    if (prev_mode > 0 && prev_mode != mode) s->mode_switch.store(true);
}

extern "C" void nq_phase1_note_celt_call(int has_data, int frame_size, int mode, int prev_mode)
{
    nq_phase1_session *s = g_t.session;
    if (!s) return;
    if (!has_data) s->lost_celt.store(true);
    if (frame_size < 480) s->short_celt.store(true);
    if (prev_mode > 0 && prev_mode != mode) s->mode_switch.store(true);
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
    const int rc = nq_celt_sink_push(s->sink, stream, freq, CC, N, shortBlocks, &p);
    if (rc != NQ_OK) {
        fprintf(stderr, "nq two-phase decoder: nq_celt_sink_push: %s\n", nq_celt_sink_last_error(s->sink));
        int expected = 0;
        s->error.compare_exchange_strong(expected, rc);
        return;
    }
    s->frames.fetch_add(1);
}
