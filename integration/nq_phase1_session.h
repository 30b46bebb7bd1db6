// C++ side of the phase-1 taps: session control used by the two-phase loader.
#pragma once
#include "nq_celt_synth.h"

struct nq_phase1_stats {
    long long frames;     // CELT frames pushed (all streams)
    int streams_seen;     // distinct CELT decoder states, i.e. multistream streams
    int saw_silk;         // silk_Decode ran: SILK-only or hybrid packets
    int mode_switch;      // consecutive packets of different coding modes (redundancy frames, cross-fades, resets)
    int irregular_celt;   // celt_decode_with_ec without packet data (concealment) or for less than 10 ms next to SILK
    int error;            // first nq_celt_sink_push error, or 0
    int resets;           // OPUS_RESET_STATE of a CELT decoder that had decoded frames before
};

// One decode session per calling thread: the loader brackets its op_read_float loop with these.
void nq_phase1_begin(nq_celt_sink *sink);
nq_phase1_stats nq_phase1_end(void);
long long nq_phase1_frames_so_far(void);   // CELT frames pushed by the calling thread's session so far
int nq_phase1_saw_silk_so_far(void);

// Phase 1 over the streams of a multistream packet in parallel (SURVEY.md section 8(f) row 2:
// every multistream sub-decoder is independent, opus_multistream_decoder.c:237-251).
// A session's helper threads decode streams 1.. while the session's own thread decodes stream 0;
// each of them must say which stream the frames it is about to decode belong to.
struct nq_phase1_session;
nq_phase1_session *nq_phase1_current(void);                       // the calling thread's session (or null)
void nq_phase1_bind(nq_phase1_session *session, int stream);      // this thread now decodes `stream` of `session`; -1: first-seen order
