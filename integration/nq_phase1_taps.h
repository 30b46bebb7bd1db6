/* Phase-1 taps: what the overlaid reference decoder calls instead of synthesising.
 * C linkage: called from the reference's C translation unit, defined in nq_phase1_taps.cpp. */
#ifndef NQ_PHASE1_TAPS_H
#define NQ_PHASE1_TAPS_H
#ifdef __cplusplus
extern "C" {
#endif

/* Replaces one comb_filter call (celt_decoder_clean.c:663-669).  Called, per frame, once (LM == 0)
 * or twice (LM > 0) per channel c; the calls of channel 0 carry the frame: the first one the
 * (old -> current) filter pair, the second one (current -> new).  The frame is pushed to the
 * active sink when its last channel-0 call arrives. */
void nq_phase1_frame_tap(const void *celt_decoder, const float *freq, int CC, int N, int LM, int shortBlocks, int c,
                         int T0, int T1, float g0, float g1, int tapset0, int tapset1);
/* The CELT decoder state `celt_decoder` is being cleared: OPUS_RESET_STATE (celt_decoder_clean.c:846-859)
 * or its initialisation (:137).  The stream's next frame starts from a reset decoder. */
void nq_phase1_reset_tap(const void *celt_decoder);
/* opus_decode_frame starts on a packet frame (opus_decoder_clean.c:262-266): `audiosize` samples of
 * this OpusDecoder's output, coding mode `mode` (MODE_SILK_ONLY 1000 / MODE_HYBRID 1001 /
 * MODE_CELT_ONLY 1002), the previous frame's having been `prev_mode` (0: none yet). */
void nq_phase1_frame_begin(const void *opus_decoder, int audiosize, int mode, int prev_mode);
/* silk_Decode is about to run for a packet of coding mode `mode`. */
void nq_phase1_note_silk(int mode, int prev_mode);
/* opus_decode_frame is about to call celt_decode_with_ec: with or without packet data (NULL: loss
 * concealment, which the bundled decoder has no defined behaviour for), for `frame_size` samples.
 * pcm_arg / data_arg: the call's own argument texts, which tell the three call sites apart --
 * "redundant_audio": a 5 ms redundancy frame (:485 CELT -> SILK, :540 SILK -> CELT, see celt_to_silk);
 * data "silence": the 2.5 ms fade-out frame of a hybrid -> SILK switch (:513); else the frame itself. */
void nq_phase1_note_celt_call(int has_data, int frame_size, int mode, int prev_mode, const char *pcm_arg,
                              const char *data_arg, int celt_to_silk);
/* smooth_fade is about to run over `overlap` samples of the current frame (`audiosize` samples long).
 * in1_arg, the text of its first argument, names the call site: "pcm + ..." :542 the frame's last
 * 2.5 ms fade into the second half of the redundancy frame (SILK -> CELT); "redundant_audio + ..."
 * :552 (CELT -> SILK: the first 2.5 ms ARE the redundancy frame's, the next 2.5 ms fade from it);
 * "pcm_transition + ..." :561 / "pcm_transition" :572 a switch without redundancy. */
void nq_phase1_fade_tap(const char *in1_arg, int audiosize, int overlap);

#ifdef __cplusplus
}
#endif
#endif
