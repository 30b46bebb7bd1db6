/* PHASE-1 OVERLAY of third_party/opus/libopus/src/opus_decoder_clean.c (see the CELT overlay
 * next door for the mechanism).  opus_decode_frame mixes a second decoder's output into the PCM on
 * the CPU whenever a packet is not CELT-only (opus_decoder_clean.c:340-410 silk_Decode, :553-560 the
 * sum), and around mode switches it decodes extra CELT frames into side buffers and cross-fades
 * them in (:478-487, :499-513, :570-600), which phase 2 does not reproduce.  The taps below only
 * RECORD what happened -- that silk_Decode ran, in which mode, and what kind of CELT call was made
 * (both macros expand inside opus_decode_frame, where `mode` and `st` are in scope) -- so that the
 * loader can tell the files it covers (CELT-only; SILK-only; hybrid without mode switches) from
 * the ones it must refuse loudly; they then call through.
 */
#include "nq_phase1_taps.h"

#define silk_Decode(...) (nq_phase1_note_silk(mode, st->prev_mode), silk_Decode(__VA_ARGS__))
#define celt_decode_with_ec(dec, data, len, pcm, frame_size, ecdec) \
    (nq_phase1_note_celt_call((data) != NULL, (frame_size), mode, st->prev_mode), celt_decode_with_ec(dec, data, len, pcm, frame_size, ecdec))
#include_next "opus/libopus/src/opus_decoder_clean.c"
#undef silk_Decode
#undef celt_decode_with_ec
