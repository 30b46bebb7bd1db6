// Phase-1 taps -> frame sink.  See nq_phase1_taps.h and overlay/opus/celt/celt_decoder_clean.c.
//
// One decode session at a time per thread (the reference decoder is single-threaded,
// SURVEY.md section 1): the loader brackets its op_read_float loop with
// nq_phase1_begin(sink) / nq_phase1_end().
#include "nq_phase1_taps.h"
#include "nq_phase1_session.h"

#include <cstdio>
#include <cstdlib>
#include <vector>

namespace {

struct Session {
    nq_celt_sink *sink = nullptr;
    std::vector<const void *> decoders;   // CELT decoder states in first-seen order == multistream stream order
    nq_celt_post_frame cur{};             // side info of the frame being assembled (channel 0's calls)
    int calls_c0 = 0;
    bool silk = false;
    int error = 0;
    long long frames = 0;
};

thread_local Session g_s;

}  // namespace

void nq_phase1_begin(nq_celt_sink *sink)
{
    g_s = Session();
    g_s.sink = sink;
}

nq_phase1_stats nq_phase1_end(void)
{
    nq_phase1_stats st;
    st.frames = g_s.frames;
    st.streams_seen = (int)g_s.decoders.size();
    st.saw_silk = g_s.silk ? 1 : 0;
    st.error = g_s.error;
    g_s = Session();
    return st;
}

extern "C" void nq_phase1_note_silk(void) { g_s.silk = true; }

extern "C" void nq_phase1_frame_tap(const void *dec, const float *freq, int CC, int N, int LM, int shortBlocks, int c,
                                    int T0, int T1, float g0, float g1, int tapset0, int tapset1)
{
    Session &s = g_s;
    if (!s.sink) {
        fprintf(stderr, "nq two-phase decoder: CELT frame decoded outside nq_phase1_begin/end; there is no CPU synthesis in this build\n");
        abort();
    }
    if (c != 0 || s.error) return;   // the calls of the other channel repeat channel 0's arguments
    nq_celt_post_frame &p = s.cur;
    if (s.calls_c0 == 0) {           // celt_decoder_clean.c:663: old -> current over [0, 120)
        p.N = N;
        p.pitch[0] = T0; p.gain[0] = g0; p.tapset[0] = tapset0;
        p.pitch[1] = T1; p.gain[1] = g1; p.tapset[1] = tapset1;
        p.pitch[2] = T1; p.gain[2] = g1; p.tapset[2] = tapset1;
        s.calls_c0 = 1;
        if (LM != 0) return;         // :666 the second call follows
    } else {                          // :666-669: current -> new over [120, N)
        p.pitch[2] = T1; p.gain[2] = g1; p.tapset[2] = tapset1;
    }
    s.calls_c0 = 0;
    int stream = -1;
    for (size_t i = 0; i < s.decoders.size(); i++)
        if (s.decoders[i] == dec) stream = (int)i;
    if (stream < 0) {
        stream = (int)s.decoders.size();
        s.decoders.push_back(dec);
    }
    const int rc = nq_celt_sink_push(s.sink, stream, freq, CC, N, shortBlocks, &p);
    if (rc != NQ_OK) {
        fprintf(stderr, "nq two-phase decoder: nq_celt_sink_push: %s\n", nq_celt_sink_last_error(s.sink));
        s.error = rc;
        return;
    }
    s.frames++;
}
