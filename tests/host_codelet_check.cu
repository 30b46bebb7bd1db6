// Test-only: runs the product's in-register DFT codelets on the HOST against
// a naive double-precision inverse DFT.  Built and run by tests/test_host_logic.py.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include "../libnyquist_b200/csrc/celt_fft_codelets.cuh"

template <int N> static double check(void (*fn)(float2 (&)[N]), int (*slot)(int))
{
    float2 x[N], y[N];
    double worst = 0;
    for (int trial = 0; trial < 20; trial++) {
        for (int n = 0; n < N; n++) { x[n].x = (float)rand() / RAND_MAX - .5f; x[n].y = (float)rand() / RAND_MAX - .5f; y[n] = x[n]; }
        fn(y);
        for (int k = 0; k < N; k++) {
            double re = 0, im = 0;
            for (int n = 0; n < N; n++) {
                double ph = 2 * M_PI * ((n * k) % N) / N;
                re += x[n].x * cos(ph) - x[n].y * sin(ph);
                im += x[n].x * sin(ph) + x[n].y * cos(ph);
            }
            int s = slot(k);
            worst = fmax(worst, fmax(fabs(re - y[s].x), fabs(im - y[s].y)));
        }
    }
    return worst;
}
static int ident(int k) { return k; }
static int s16(int k) { return nq::slot16(k); }
static void f30(float2 (&g)[30]) { nq::idft30(g); }
static void f16(float2 (&g)[16]) { nq::idft16(g); }
int main()
{
    double e30 = check<30>(f30, ident), e16 = check<16>(f16, s16);
    printf("idft30 %.3e idft16 %.3e\n", e30, e16);
    return (e30 < 5e-6 && e16 < 5e-6) ? 0 : 1;
}
