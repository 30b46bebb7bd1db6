/* PHASE-1 OVERLAY of third_party/opus/celt/celt_decoder_clean.c -- the reference-side patch of
 * the two-phase decoder, expressed without touching or copying
 * the reference file: src/OpusDependencies.c:96 does
 *     #include "opus/celt/celt_decoder_clean.c"
 * and integration/Makefile puts integration/overlay first on the -I list, so that include lands
 * here; this file redirects three call sites and then continues with the reference's own file
 * through #include_next.
 *
 * What changes inside celt_decode_with_ec / compute_inv_mdcts:
 *   - clt_mdct_backward_B1_C2 / clt_mdct_backward (celt_decoder_clean.c:290,298,309): no work.
 *     The inverse MDCT + overlap-add is phase 2 (GPU).
 *   - comb_filter (:663-669): no filtering.  The call site has everything phase 2 needs in
 *     scope -- st, freq (the frame's denormalised coefficients, intact since :620-636), CC, N,
 *     LM, shortBlocks, the channel index c and the filter arguments -- so this is where the frame
 *     is pushed into the sink (integration/nq_phase1_taps.cpp -> nq_celt_sink_push).
 *   - deemphasis (:723) still runs, on the silent out_syn, and its output is ignored: the PCM the
 *     reference returns in phase 1 is a placeholder that only carries sample COUNTS through
 *     opus_decode_native / opus_multistream / opusfile (pre-skip, end trim are positional).
 * (celt.h and mdct.h were already included by the unity build before this point, so the
 * function-like macros below never meet a prototype.)
 */
#include "nq_phase1_taps.h"

#define clt_mdct_backward_B1_C2(l, in, out, window, overlap, shift, stride) ((void)(in), (void)(out))
#define clt_mdct_backward(l, in, out, window, overlap, shift, stride) ((void)0)
#define comb_filter(y, x, T0, T1, n, g0, g1, tapset0, tapset1, window, overlap) \
    nq_phase1_frame_tap((const void *)st, freq, CC, N, LM, shortBlocks, c, T0, T1, g0, g1, tapset0, tapset1)
#include_next "opus/celt/celt_decoder_clean.c"
#undef clt_mdct_backward_B1_C2
#undef clt_mdct_backward
#undef comb_filter
