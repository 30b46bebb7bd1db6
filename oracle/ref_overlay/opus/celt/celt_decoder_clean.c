/* TEST INFRASTRUCTURE (oracle/_ref build only) -- not part of the product.
 *
 * Include-path overlay: the reference's unity file does
 *   #include "opus/celt/celt_decoder_clean.c"      (src/OpusDependencies.c:96)
 * and oracle/Makefile puts oracle/ref_overlay first on the -I list, so that
 * include lands here.  This file contains NO reference code: it redirects the
 * three call sites of the inverse MDCT inside compute_inv_mdcts
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

#define clt_mdct_backward_B1_C2 nqref_tap_mdct_b1c2
#define clt_mdct_backward nqref_tap_mdct
#include_next "opus/celt/celt_decoder_clean.c"
#undef clt_mdct_backward_B1_C2
#undef clt_mdct_backward
