/* PHASE-1 OVERLAY of third_party/opus/libopus/src/opus_decoder_clean.c (see the CELT overlay
 * next door for the mechanism).  The two-phase decoder covers CELT-only streams: a SILK or hybrid
 * frame mixes a second decoder's output into the PCM on the CPU (opus_decoder_clean.c:388, :553),
 * which phase 2 does not reproduce.  The tap below only RECORDS that silk_Decode ran, so the
 * loader can refuse such a file loudly instead of returning wrong audio; it then calls through.
 */
#include "nq_phase1_taps.h"

#define silk_Decode(...) (nq_phase1_note_silk(), silk_Decode(__VA_ARGS__))
#include_next "opus/libopus/src/opus_decoder_clean.c"
#undef silk_Decode
