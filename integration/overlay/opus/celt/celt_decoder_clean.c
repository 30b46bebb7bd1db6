/* PHASE-1 OVERLAY of third_party/opus/celt/celt_decoder_clean.c -- the reference-side patch of
 * the two-phase decoder, expressed without touching or copying
 * the reference file: src/OpusDependencies.c:96 does
 *     #include "opus/celt/celt_decoder_clean.c"
 * and integration/Makefile puts integration/overlay first on the -I list, so that include lands
 * here; this file redirects a few call sites and then continues with the reference's own file
 * through #include_next.
 *
 * What changes inside celt_decode_with_ec / compute_inv_mdcts / opus_custom_decoder_ctl:
 *   - clt_mdct_backward_B1_C2 / clt_mdct_backward (celt_decoder_clean.c:290,298,309): no work.
 *     The inverse MDCT + overlap-add is phase 2 (GPU).
 *   - comb_filter (:663-669): no filtering.  The call site has everything phase 2 needs in
 *     scope -- st, freq (the frame's denormalised coefficients, intact since :620-636), CC, N,
 *     LM, shortBlocks, the channel index c and the filter arguments -- so this is where the frame
 *     is pushed into the sink (integration/nq_phase1_taps.cpp -> nq_celt_sink_push).
 *   - the history shift OPUS_MOVE(decode_mem[c], decode_mem[c] + N, ...) (:622-626; the file's only
 *     OPUS_MOVE): no work -- nothing reads decode_mem in this build, the tail lives on the GPU.
 *   - deemphasis (:723): the CALL clears the frame's PCM instead of filtering a silent out_syn
 *     (7.7 % of the CELT decode): the PCM the reference returns in phase 1
 *     is a placeholder that carries sample COUNTS through opus_decode_native / opus_multistream /
 *     opusfile (pre-skip, end trim are positional) and, for hybrid packets, receives the SILK layer.
 *     `deemphasis` is both defined and called in the reference file, with eight arguments each;
 *     the macro below tells the two apart by the third argument (`int N` in the definition, `N` in
 *     the call) and leaves the definition alone under another name.
 *   - OPUS_RESET_STATE (:846-859; OPUS_CLEAR of the decoder state, also at :137 in the init): the
 *     sink is told, so that phase 2 clears its tail / history / memory before the stream's next frame.
 * (celt.h, mdct.h and os_support.h were already included by the unity build before this point, so
 * the function-like macros below never meet a prototype.)
 */
#include "nq_phase1_taps.h"

#define clt_mdct_backward_B1_C2(l, in, out, window, overlap, shift, stride) ((void)(in), (void)(out))
#define clt_mdct_backward(l, in, out, window, overlap, shift, stride) ((void)0)
#define comb_filter(y, x, T0, T1, n, g0, g1, tapset0, tapset1, window, overlap) \
    nq_phase1_frame_tap((const void *)st, freq, CC, N, LM, shortBlocks, c, T0, T1, g0, g1, tapset0, tapset1)

#pragma push_macro("OPUS_MOVE")
#pragma push_macro("OPUS_CLEAR")
#undef OPUS_MOVE
#undef OPUS_CLEAR
#define OPUS_MOVE(dst, src, n) ((void)0)
#define OPUS_CLEAR(dst, n) (nq_phase1_reset_tap((const void *)st), memset((dst), 0, (n) * sizeof(*(dst))))

/* definition `deemphasis(celt_sig *in[], opus_val16 *pcm, int N, ...)` vs call `deemphasis(out_syn, pcm, N, ...)` */
#define NQ_PP_CAT_(a, b) a##b
#define NQ_PP_CAT(a, b) NQ_PP_CAT_(a, b)
#define NQ_PP_SECOND_(a, b, ...) b
#define NQ_PP_SECOND(...) NQ_PP_SECOND_(__VA_ARGS__, 0, )
#define NQ_PP_PROBE_int ~, 1,
#define NQ_PP_IS_DECL(arg) NQ_PP_SECOND(NQ_PP_PROBE_##arg)
#define NQ_DEEMPH_1(in, pcm, N, ...) nq_ref_deemphasis_unused(in, pcm, N, __VA_ARGS__)
#define NQ_DEEMPH_0(in, pcm, N, C, downsample, coef, mem, scratch) \
    memset((pcm), 0, sizeof(*(pcm)) * (size_t)((N) / (downsample)) * (size_t)(C))
#define deemphasis(in, pcm, N, ...) NQ_PP_CAT(NQ_DEEMPH_, NQ_PP_IS_DECL(N))(in, pcm, N, __VA_ARGS__)

#include_next "opus/celt/celt_decoder_clean.c"

#undef deemphasis
#undef NQ_DEEMPH_0
#undef NQ_DEEMPH_1
#undef NQ_PP_IS_DECL
#undef NQ_PP_PROBE_int
#undef NQ_PP_SECOND
#undef NQ_PP_SECOND_
#undef NQ_PP_CAT
#undef NQ_PP_CAT_
#pragma pop_macro("OPUS_CLEAR")
#pragma pop_macro("OPUS_MOVE")
#undef clt_mdct_backward_B1_C2
#undef clt_mdct_backward
#undef comb_filter
