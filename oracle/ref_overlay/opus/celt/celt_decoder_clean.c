/* TEST INFRASTRUCTURE (oracle/_ref build only) -- not part of the product.
 *
 * Include-path overlay: the reference's unity file does
 *   #include "opus/celt/celt_decoder_clean.c"      (src/OpusDependencies.c:96)
 * and oracle/Makefile puts oracle/ref_overlay first on the -I list, so that
 * include lands here.  This file contains NO reference code: it redirects the
 * call sites of comb_filter (:663-669) and the three call sites of the inverse MDCT inside compute_inv_mdcts
 * (celt_decoder_clean.c:290,298,309) to taps defined in oracle/ref_harness.c
 * and then continues with the reference's own, unmodified file through
 * #include_next.  The taps call straight through to the reference functions
 * (mdct.c:258,267), so the decode result is unchanged.
 */
#include "opus/celt/mdct.h"

void nqref_tap_mdct_b1c2(const mdct_lookup *l, float *in[2], float *out[2],
                         const float *window, int overlap, int shift, int stride);
void nqref_tap_mdct(const mdct_lookup *l, float *in, float *out,
                    const float *window, int overlap, int shift, int stride);

/* comb_filter (celt.c:114) is called twice per channel per frame at
 * celt_decoder_clean.c:663-669; the tap logs the arguments and calls through. */
void nqref_tap_comb_filter(const void *decoder, float *y, float *x, int T0, int T1, int N, float g0, float g1,
                           int tapset0, int tapset1, const float *window, int overlap);

#define clt_mdct_backward_B1_C2 nqref_tap_mdct_b1c2
#define clt_mdct_backward nqref_tap_mdct
#define comb_filter(...) nqref_tap_comb_filter((const void *)st, __VA_ARGS__)   /* st: the CELTDecoder in scope at :663-669 */
#include_next "opus/celt/celt_decoder_clean.c"
#undef clt_mdct_backward_B1_C2
#undef clt_mdct_backward
#undef comb_filter
