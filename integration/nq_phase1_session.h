// C++ side of the phase-1 taps: session control used by the two-phase loader.
#pragma once
#include "nq_celt_synth.h"

struct nq_phase1_stats {
    long long frames;     // CELT frames pushed (all streams)
    int streams_seen;     // distinct CELT decoder states, i.e. multistream streams
    int saw_silk;         // a SILK / hybrid frame was decoded: not covered by phase 2
    int error;            // first nq_celt_sink_push error, or 0
};

void nq_phase1_begin(nq_celt_sink *sink);
nq_phase1_stats nq_phase1_end(void);
