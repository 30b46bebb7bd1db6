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
void nq_phase1_note_silk(void);

#ifdef __cplusplus
}
#endif
#endif
