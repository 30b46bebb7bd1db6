// C++ side of the phase-1 taps: session control used by the two-phase loader.
#pragma once
#include <vector>

#include "nq_celt_synth.h"

// What a fade of opus_decode_frame (opus_decoder_clean.c:530-575) owes the CELT share of the output;
// applied by the loader to phase 2's PCM (positions: samples of the output before the pre-skip).
struct nq_phase1_fixup {
    enum Kind {
        CeltToSilk,        // :545-554  out[pos + i] = side[i]; out[pos + n + i] = w_i out[..] + (1 - w_i) side[n + i]
        SilkToCelt,        // :536-544  out[pos + size - n + i] = w_i side[n + i] + (1 - w_i) out[..]
        Transition,        // :557-563  out[pos + i] = 0; out[pos + n + i] *= w_i      (the other input is not CELT's)
        TransitionShort    // :565-574  out[pos + i] *= w_i
    } kind;
    long long pos;         // first sample of the packet frame
    int size;              // its length
    int n;                 // overlap of the fade (2.5 ms = 120); w_i = window[i]^2
    int side;              // index of the side frame (nq_celt_sink_side_get), -1: none
};

struct nq_phase1_stats {
    long long frames;     // CELT frames pushed (all streams)
    int streams_seen;     // distinct CELT decoder states, i.e. multistream streams
    int saw_silk;         // silk_Decode ran: SILK-only or hybrid packets
    int mode_switch;      // consecutive packets of different coding modes (redundancy frames, cross-fades, resets)
    int irregular_celt;   // celt_decode_with_ec without packet data (loss concealment)
    int error;            // first nq_celt_sink_push error, or 0
    int resets;           // OPUS_RESET_STATE of a CELT decoder that had decoded frames before
    long long samples;    // output samples per channel the packet frames account for (stream 0)
    std::vector<nq_phase1_fixup> fixups;
};

// One decode session per calling thread: the loader brackets its op_read_float loop with these.
void nq_phase1_begin(nq_celt_sink *sink, int streams = 0);   // streams == 1: frames may carry their place in the output (mode switches)
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
