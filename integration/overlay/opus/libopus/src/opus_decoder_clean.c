/* PHASE-1 OVERLAY of third_party/opus/libopus/src/opus_decoder_clean.c (see the CELT overlay
 * next door for the mechanism).  opus_decode_frame mixes a second decoder's output into the PCM on
 * the CPU whenever a packet is not CELT-only (opus_decoder_clean.c:340-410 silk_Decode, :553-560 the
 * sum), and around mode switches it decodes extra CELT frames into side buffers and cross-fades
 * them in (:478-487, :499-513, :530-575).  Every one of these steps is LINEAR in the CELT decoder's
 * output, so with the CELT synthesis switched off phase 1 still computes the SILK share of the
 * final PCM exactly as the reference does, and the CELT share follows from phase 2's output by
 * applying the same fades to it.  The taps below only RECORD what happened, where, and call through
 * (all macros expand inside opus_decode_frame, where st, mode, audiosize, celt_to_silk are in scope):
 *   ec_dec_init          (:265, once per decoded packet frame)  -> where this frame sits in the output
 *   silk_Decode          (:388)                                 -> a SILK layer exists (SILK-only / hybrid)
 *   celt_decode_with_ec  (:485, :540 redundancy into redundant_audio; :500 the frame itself; :513 the
 *                         2.5 ms fade-out frame of a hybrid -> SILK switch)  -> what the next pushed frame is
 *   smooth_fade          (:542 SILK -> CELT redundancy, :552 CELT -> SILK redundancy, :561 / :572 the
 *                         transition of a switch without redundancy)          -> the fade phase 2's PCM owes
 */
#include "nq_phase1_taps.h"

#define ec_dec_init(d, buf, len) \
    (nq_phase1_frame_begin((const void *)st, st->frame_size, st->mode, st->prev_mode), ec_dec_init(d, buf, len))
#define silk_Decode(...) (nq_phase1_note_silk(mode, st->prev_mode), silk_Decode(__VA_ARGS__))
#define celt_decode_with_ec(dec, data, len, pcm, frame_size, ecdec) \
    (nq_phase1_note_celt_call((data) != NULL, (frame_size), mode, st->prev_mode, #pcm, #data, celt_to_silk), \
     celt_decode_with_ec(dec, data, len, pcm, frame_size, ecdec))
/* smooth_fade is defined (:174, fourth parameter `int overlap`) and called (fourth argument F2_5) in the
 * reference file: the definition is left alone under another name, the calls report and call it */
#define NQ_PP_CAT_(a, b) a##b
#define NQ_PP_CAT(a, b) NQ_PP_CAT_(a, b)
#define NQ_PP_SECOND_(a, b, ...) b
#define NQ_PP_SECOND(...) NQ_PP_SECOND_(__VA_ARGS__, 0, )
#define NQ_PP_PROBE_int ~, 1,
#define NQ_PP_IS_DECL(arg) NQ_PP_SECOND(NQ_PP_PROBE_##arg)
#define NQ_FADE_1(in1, in2, out, overlap, ...) nq_ref_smooth_fade(in1, in2, out, overlap, __VA_ARGS__)
#define NQ_FADE_0(in1, in2, out, overlap, ...) \
    (nq_phase1_fade_tap(#in1, audiosize, overlap), nq_ref_smooth_fade(in1, in2, out, overlap, __VA_ARGS__))
#define smooth_fade(in1, in2, out, overlap, ...) NQ_PP_CAT(NQ_FADE_, NQ_PP_IS_DECL(overlap))(in1, in2, out, overlap, __VA_ARGS__)
#include_next "opus/libopus/src/opus_decoder_clean.c"
#undef smooth_fade
#undef silk_Decode
#undef celt_decode_with_ec
#undef ec_dec_init
#undef NQ_FADE_0
#undef NQ_FADE_1
#undef NQ_PP_IS_DECL
#undef NQ_PP_PROBE_int
#undef NQ_PP_SECOND
#undef NQ_PP_SECOND_
#undef NQ_PP_CAT
#undef NQ_PP_CAT_
