// In-register inverse-DFT codelets for the CELT synthesis kernels (sm_100a).
//
// The reference computes its N/4-point inverse FFT with kiss_fft's mixed-radix
// passes over memory (third_party/opus/celt/kiss_fft.c:696-747, radices
// 5,3,2,4,4 for 480 and 5,3,4 for 60).  Here a transform is split as
//     480 = 30 x 16   (long block)        60 = 30 x 2   (short block)
// and each factor is done entirely in registers by one thread:
//   * 30 points = 2 x 3 x 5, pairwise coprime -> Good-Thomas prime-factor
//     mapping, i.e. a 2x3x5 three-dimensional DFT with NO internal twiddles.
//     Index map p = (15a + 10b + 6c) mod 30 is its own CRT inverse, so the
//     transform is in place: slot p holds x[p] before and X[p] after.
//   * 16 points = 4 x 4 with one layer of W16 twiddles.
// All loops are fully unrolled so every array index is a compile-time
// constant and the arrays live in registers; the rotation constants come from
// celt_consts.cuh as float literals (immediate operands).
//
// Sign convention: INVERSE transform, X[k] = sum_n x[n] exp(+j 2 pi n k / N),
// unnormalised, like opus_ifft.
#pragma once
#include <cuda_runtime.h>
#include "celt_consts.cuh"

// __host__ too, so tests/host_codelet_check.cu can run the very same codelets
// on the CPU against a naive DFT (test only; the product has no CPU path).
#define NQ_HD __host__ __device__ __forceinline__

namespace nq {

NQ_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
NQ_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
NQ_HD float2 cmul(float2 a, float2 b)
{
    return make_float2(fmaf(a.x, b.x, -(a.y * b.y)), fmaf(a.x, b.y, a.y * b.x));
}
// a * (wr + j wi) with literal wr, wi
NQ_HD float2 cmulc(float2 a, float wr, float wi)
{
    return make_float2(fmaf(a.x, wr, -(a.y * wi)), fmaf(a.x, wi, a.y * wr));
}
// a * (+j) and a * (-j)
NQ_HD float2 mul_pj(float2 a) { return make_float2(-a.y, a.x); }
NQ_HD float2 mul_mj(float2 a) { return make_float2(a.y, -a.x); }

NQ_HD void idft2(float2 &x0, float2 &x1)
{
    float2 t = x0;
    x0 = cadd(t, x1);
    x1 = csub(t, x1);
}

// w = exp(+j 2pi/3) = -1/2 + j sin(2pi/3)
NQ_HD void idft3(float2 &x0, float2 &x1, float2 &x2)
{
    const float2 t = cadd(x1, x2), d = csub(x1, x2);
    const float2 m = make_float2(fmaf(-0.5f, t.x, x0.x), fmaf(-0.5f, t.y, x0.y));
    x0 = cadd(x0, t);
    x1 = make_float2(fmaf(-NQ_SIN_2PI_3, d.y, m.x), fmaf(NQ_SIN_2PI_3, d.x, m.y));   // m + j s d
    x2 = make_float2(fmaf(NQ_SIN_2PI_3, d.y, m.x), fmaf(-NQ_SIN_2PI_3, d.x, m.y));   // m - j s d
}

// w = exp(+j 2pi/5)
NQ_HD void idft5(float2 &x0, float2 &x1, float2 &x2, float2 &x3, float2 &x4)
{
    const float2 t1 = cadd(x1, x4), t3 = csub(x1, x4);
    const float2 t2 = cadd(x2, x3), t4 = csub(x2, x3);
    const float2 m1 = make_float2(fmaf(NQ_COS_4PI_5, t2.x, fmaf(NQ_COS_2PI_5, t1.x, x0.x)),
                                  fmaf(NQ_COS_4PI_5, t2.y, fmaf(NQ_COS_2PI_5, t1.y, x0.y)));
    const float2 m2 = make_float2(fmaf(NQ_COS_2PI_5, t2.x, fmaf(NQ_COS_4PI_5, t1.x, x0.x)),
                                  fmaf(NQ_COS_2PI_5, t2.y, fmaf(NQ_COS_4PI_5, t1.y, x0.y)));
    // s1 = sin(2pi/5) t3 + sin(4pi/5) t4 ; s2 = sin(4pi/5) t3 - sin(2pi/5) t4
    const float2 s1 = make_float2(fmaf(NQ_SIN_4PI_5, t4.x, NQ_SIN_2PI_5 * t3.x),
                                  fmaf(NQ_SIN_4PI_5, t4.y, NQ_SIN_2PI_5 * t3.y));
    const float2 s2 = make_float2(fmaf(-NQ_SIN_2PI_5, t4.x, NQ_SIN_4PI_5 * t3.x),
                                  fmaf(-NQ_SIN_2PI_5, t4.y, NQ_SIN_4PI_5 * t3.y));
    x0 = cadd(x0, cadd(t1, t2));
    x1 = make_float2(m1.x - s1.y, m1.y + s1.x);   // m1 + j s1
    x4 = make_float2(m1.x + s1.y, m1.y - s1.x);   // m1 - j s1
    x2 = make_float2(m2.x - s2.y, m2.y + s2.x);   // m2 + j s2
    x3 = make_float2(m2.x + s2.y, m2.y - s2.x);   // m2 - j s2
}

// inverse 4-point: X[q] = sum_n x[n] (+j)^(n q)
NQ_HD void idft4(float2 &x0, float2 &x1, float2 &x2, float2 &x3)
{
    const float2 a = cadd(x0, x2), b = csub(x0, x2);
    const float2 c = cadd(x1, x3), d = mul_pj(csub(x1, x3));
    x0 = cadd(a, c);
    x2 = csub(a, c);
    x1 = cadd(b, d);
    x3 = csub(b, d);
}

// 30-point inverse DFT, in place, natural order in and out (prime-factor map).
NQ_HD void idft30(float2 (&g)[30])
{
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
        for (int b = 0; b < 3; b++) {
            const int base = 15 * a + 10 * b;
            idft5(g[base % 30], g[(base + 6) % 30], g[(base + 12) % 30], g[(base + 18) % 30], g[(base + 24) % 30]);
        }
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
        for (int c = 0; c < 5; c++) {
            const int base = 15 * a + 6 * c;
            idft3(g[base % 30], g[(base + 10) % 30], g[(base + 20) % 30]);
        }
#pragma unroll
    for (int b = 0; b < 3; b++)
#pragma unroll
        for (int c = 0; c < 5; c++) {
            const int base = 10 * b + 6 * c;
            idft2(g[base % 30], g[(base + 15) % 30]);
        }
}

// 16-point inverse DFT.  Input z[n] natural order; output returned through
// the accessor convention out(k) = z[slot16(k)]: the transform is done in
// place as 4x4 and the result for frequency k = c + 4d sits in slot 4c + d.
NQ_HD constexpr int slot16(int k) { return 4 * (k & 3) + (k >> 2); }

NQ_HD void idft16(float2 (&z)[16])
{
    constexpr float wr[16] = {NQ_W16_RE};
    constexpr float wi[16] = {NQ_W16_IM};
    // step 1: for each b, 4-point DFT over a of z[4a + b]  -> u[b][c] stored at z[4c + b]
#pragma unroll
    for (int b = 0; b < 4; b++) idft4(z[b], z[4 + b], z[8 + b], z[12 + b]);
    // twiddle u[b][c] *= W16^(b c)
#pragma unroll
    for (int b = 1; b < 4; b++)
#pragma unroll
        for (int c = 1; c < 4; c++) {
            const int q = b * c;   // 1,2,3,2,4,6,3,6,9
            if (q == 4) z[4 * c + b] = mul_pj(z[4 * c + b]);
            else z[4 * c + b] = cmulc(z[4 * c + b], wr[q], wi[q]);
        }
    // step 2: for each c, 4-point DFT over b of z[4c + b] -> Z[c + 4d] stored at z[4c + d]
#pragma unroll
    for (int c = 0; c < 4; c++) idft4(z[4 * c], z[4 * c + 1], z[4 * c + 2], z[4 * c + 3]);
}

}  // namespace nq
