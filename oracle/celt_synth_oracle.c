/* TEST INFRASTRUCTURE -- CPU oracle for the CELT synthesis hot path.
 *
 * A plain-C restatement of the reference algorithm (dafx/libnyquist, bundled
 * Opus 1.1 float build), written to evaluate every float expression in the
 * same order as the reference so that results can be compared BIT FOR BIT
 * with the compiled reference (oracle/_ref/libnq_ref.so) and with the
 * reference's golden vectors test_data/ifft_{input,output}_N{480,60}.bin.
 * Parity status: PINNED (tests/test_oracle.py: ifft goldens bit-exact; whole
 * path bit-exact against the compiled reference on synthetic batches and on
 * frames recorded from the bundled .opus files).
 *
 * Reference file:line followed by each function is cited at the function.
 * (paths relative to /root/reference/third_party/opus/celt/)
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, load or call this file.  The product
 * (libnyquist_b200/) has no CPU path and never links it.
 *
 * Build: make -C oracle oracle   (gcc -O3 -ffp-contract=off, no -march)
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define NQO_API __attribute__((visibility("default")))

enum { MDCT_N = 1920, NFFT0 = 480, OVERLAP = 120, HALF_OVL = 60, FRAME = 960, MAXSTAGES = 8 };

typedef struct { float r, i; } cpx;

/* One inverse-FFT plan per shift (mdct.h:49 kfft[4]; kiss_fft.h:78). */
typedef struct {
    int nfft, shift, nstages;
    int radix[MAXSTAGES], m[MAXSTAGES];  /* kf_factor output: radix[s]*m[s] = m[s-1] */
    int16_t bitrev[NFFT0];
} ifft_plan;

static float g_window[OVERLAP];   /* static_modes_float.h:9   window120            */
static float g_trig[NFFT0 + 1];   /* static_modes_float.h:477 mdct_twiddles960     */
static cpx g_tw[NFFT0];           /* static_modes_float.h:99  fft_twiddles48000_960 */
static ifft_plan g_plan[4];
static int g_ready;

/* kiss_fft.c:489-525 kf_factor: pull out 4s, then 2s, then 3s, then 5s. */
static void factorise(ifft_plan *pl)
{
    int n = pl->nfft, p = 4, s = 0;
    do {
        while (n % p) {
            if (p == 4) p = 2;
            else if (p == 2) p = 3;
            else p += 2;
            if (p * p > n) p = n;
        }
        n /= p;
        pl->radix[s] = p;
        pl->m[s] = n;
        s++;
    } while (n > 1);
    pl->nstages = s;
}

/* kiss_fft.c:459-487 compute_bitrev_table: output slot of every input index. */
static void fill_bitrev(int fout, int16_t *f, int fstride, const int *radix, const int *m)
{
    int j;
    if (*m == 1) {
        for (j = 0; j < *radix; j++) { *f = (int16_t)(fout + j); f += fstride; }
    } else {
        for (j = 0; j < *radix; j++) {
            fill_bitrev(fout, f, fstride * *radix, radix + 1, m + 1);
            f += fstride;
            fout += *m;
        }
    }
}

static void build_plans(void)
{
    int s;
    for (s = 0; s < 4; s++) {
        ifft_plan *pl = &g_plan[s];
        pl->nfft = NFFT0 >> s;
        pl->shift = s;   /* static states carry shift -1/1/2/3; opus_ifft clamps -1 to 0 (kiss_fft.c:706) */
        factorise(pl);
        fill_bitrev(0, pl->bitrev, 1, pl->radix, pl->m);
    }
}

/* Self-generated tables (modes.c:374, mdct.c:99, kiss_fft.c:527-545).  The
 * reference ships them as 8-significant-digit literals, so a few entries can
 * differ by one ulp from these; nqo_set_tables() installs the reference's
 * exact values (tests/golden/ref_tables.npz) for bit-exact work. */
NQO_API void nqo_init_default(void)
{
    int i;
    const double pi = 3.14159265358979323846264338327;
    for (i = 0; i < OVERLAP; i++) {
        double s = sin(.5 * pi * (i + .5) / OVERLAP);
        g_window[i] = (float)sin(.5 * pi * s * s);
    }
    /* mdct.c:99 with PI = 3.141592653f (mathops.h:83): the argument is float */
    for (i = 0; i <= NFFT0; i++) g_trig[i] = (float)cos(2 * 3.141592653f * i / MDCT_N);
    for (i = 0; i < NFFT0; i++) {
        double ph = (-2 * pi / NFFT0) * i;
        g_tw[i].r = (float)cos(ph);
        g_tw[i].i = (float)sin(ph);
    }
    build_plans();
    g_ready = 1;
}

NQO_API void nqo_set_tables(const float *window120, const float *trig481, const float *twiddles480_ri)
{
    build_plans();
    memcpy(g_window, window120, sizeof g_window);
    memcpy(g_trig, trig481, sizeof g_trig);
    memcpy(g_tw, twiddles480_ri, sizeof g_tw);
    g_ready = 1;
}

NQO_API void nqo_get_tables(float *window120, float *trig481, float *twiddles480_ri,
                            int16_t *bitrev480, int16_t *bitrev240, int16_t *bitrev120, int16_t *bitrev60)
{
    if (!g_ready) nqo_init_default();
    memcpy(window120, g_window, sizeof g_window);
    memcpy(trig481, g_trig, sizeof g_trig);
    memcpy(twiddles480_ri, g_tw, sizeof g_tw);
    memcpy(bitrev480, g_plan[0].bitrev, 480 * 2);
    memcpy(bitrev240, g_plan[1].bitrev, 240 * 2);
    memcpy(bitrev120, g_plan[2].bitrev, 120 * 2);
    memcpy(bitrev60, g_plan[3].bitrev, 60 * 2);
}

/* a * conj(b): _kiss_fft_guts.h:111-113 C_MULC */
static inline cpx mulc(cpx a, cpx b)
{
    cpx o;
    o.r = a.r * b.r + a.i * b.i;
    o.i = a.i * b.r - a.r * b.i;
    return o;
}

/* kiss_fft.c:82-110 ki_bfly2 */
static void ibfly2(cpx *F, int ts, int m, int groups, int pitch)
{
    int g, j;
    for (g = 0; g < groups; g++) {
        cpx *a = F + g * pitch, *b = a + m;
        const cpx *w = g_tw;
        for (j = 0; j < m; j++, a++, b++, w += ts) {
            cpx t = mulc(*b, *w);
            b->r = a->r - t.r; b->i = a->i - t.i;
            a->r += t.r; a->i += t.i;
        }
    }
}

/* kiss_fft.c:158-200 ki_bfly4 */
static void ibfly4(cpx *F, int ts, int m, int groups, int pitch)
{
    int g, j;
    for (g = 0; g < groups; g++) {
        cpx *p = F + g * pitch;
        const cpx *w1 = g_tw, *w2 = g_tw, *w3 = g_tw;
        for (j = 0; j < m; j++, p++, w1 += ts, w2 += 2 * ts, w3 += 3 * ts) {
            cpx s0 = mulc(p[m], *w1), s1 = mulc(p[2 * m], *w2), s2 = mulc(p[3 * m], *w3);
            cpx s3, s4, s5;
            s5.r = p->r - s1.r; s5.i = p->i - s1.i;
            p->r += s1.r; p->i += s1.i;
            s3.r = s0.r + s2.r; s3.i = s0.i + s2.i;
            s4.r = s0.r - s2.r; s4.i = s0.i - s2.i;
            p[2 * m].r = p->r - s3.r; p[2 * m].i = p->i - s3.i;
            p->r += s3.r; p->i += s3.i;
            p[m].r = s5.r - s4.i; p[m].i = s5.i + s4.r;
            p[3 * m].r = s5.r + s4.i; p[3 * m].i = s5.i - s4.r;
        }
    }
}

/* kiss_fft.c:258-306 ki_bfly3 */
static void ibfly3(cpx *F, int ts, int m, int groups, int pitch)
{
    int g, j;
    const cpx epi3 = g_tw[ts * m];
    for (g = 0; g < groups; g++) {
        cpx *p = F + g * pitch;
        const cpx *w1 = g_tw, *w2 = g_tw;
        for (j = 0; j < m; j++, p++, w1 += ts, w2 += 2 * ts) {
            cpx s1 = mulc(p[m], *w1), s2 = mulc(p[2 * m], *w2), s3, s0;
            s3.r = s1.r + s2.r; s3.i = s1.i + s2.i;
            s0.r = s1.r - s2.r; s0.i = s1.i - s2.i;
            p[m].r = p->r - s3.r * .5f;
            p[m].i = p->i - s3.i * .5f;
            s0.r *= -epi3.i; s0.i *= -epi3.i;
            p->r += s3.r; p->i += s3.i;
            p[2 * m].r = p[m].r + s0.i;
            p[2 * m].i = p[m].i - s0.r;
            p[m].r -= s0.i;
            p[m].i += s0.r;
        }
    }
}

/* kiss_fft.c:385-455 ki_bfly5 */
static void ibfly5(cpx *F, int ts, int m, int groups, int pitch)
{
    int g, u;
    const cpx ya = g_tw[ts * m], yb = g_tw[ts * 2 * m];
    for (g = 0; g < groups; g++) {
        cpx *f0 = F + g * pitch, *f1 = f0 + m, *f2 = f0 + 2 * m, *f3 = f0 + 3 * m, *f4 = f0 + 4 * m;
        for (u = 0; u < m; u++, f0++, f1++, f2++, f3++, f4++) {
            cpx s0 = *f0;
            cpx s1 = mulc(*f1, g_tw[u * ts]), s2 = mulc(*f2, g_tw[2 * u * ts]);
            cpx s3 = mulc(*f3, g_tw[3 * u * ts]), s4 = mulc(*f4, g_tw[4 * u * ts]);
            cpx s5, s6, s7, s8, s9, s10, s11, s12;
            s7.r = s1.r + s4.r; s7.i = s1.i + s4.i;
            s10.r = s1.r - s4.r; s10.i = s1.i - s4.i;
            s8.r = s2.r + s3.r; s8.i = s2.i + s3.i;
            s9.r = s2.r - s3.r; s9.i = s2.i - s3.i;
            f0->r += s7.r + s8.r;
            f0->i += s7.i + s8.i;
            s5.r = s0.r + s7.r * ya.r + s8.r * yb.r;
            s5.i = s0.i + s7.i * ya.r + s8.i * yb.r;
            s6.r = -(s10.i * ya.i) - s9.i * yb.i;
            s6.i = s10.r * ya.i + s9.r * yb.i;
            f1->r = s5.r - s6.r; f1->i = s5.i - s6.i;
            f4->r = s5.r + s6.r; f4->i = s5.i + s6.i;
            s11.r = s0.r + s7.r * yb.r + s8.r * ya.r;
            s11.i = s0.i + s7.i * yb.r + s8.i * ya.r;
            s12.r = s10.i * yb.i - s9.i * ya.i;
            s12.i = -(s10.r * yb.i) + s9.r * ya.i;
            f2->r = s11.r + s12.r; f2->i = s11.i + s12.i;
            f3->r = s11.r - s12.r; f3->i = s11.i - s12.i;
        }
    }
}

/* kiss_fft.c:696-747 opus_ifft: unnormalised inverse DFT, out of place.
 * Stage s runs butterflies of radix[s] on sub-transforms of length
 * radix[s]*m[s]; groups = product of the radices before s (fstride[s]). */
static void ifft(const ifft_plan *pl, const cpx *in, cpx *out)
{
    int s, i, groups[MAXSTAGES];
    for (i = 0; i < pl->nfft; i++) out[pl->bitrev[i]] = in[i];
    groups[0] = 1;
    for (s = 1; s < pl->nstages; s++) groups[s] = groups[s - 1] * pl->radix[s - 1];
    for (s = pl->nstages - 1; s >= 0; s--) {
        int m = pl->m[s], pitch = s ? pl->m[s - 1] : 1, ts = groups[s] << pl->shift;
        switch (pl->radix[s]) {
        case 2: ibfly2(out, ts, m, groups[s], pitch); break;
        case 4: ibfly4(out, ts, m, groups[s], pitch); break;
        case 3: ibfly3(out, ts, m, groups[s], pitch); break;
        case 5: ibfly5(out, ts, m, groups[s], pitch); break;
        }
    }
}

NQO_API void nqo_ifft(int shift, const float *in_ri, float *out_ri)
{
    if (!g_ready) nqo_init_default();
    ifft(&g_plan[shift], (const cpx *)in_ri, (cpx *)out_ri);
}

/* mdct.c:267-379 clt_mdct_backward (float build).  in: N2 coefficients at
 * `stride`; out: read [0,60) (previous raw tail), written [0,N2+60). */
static void mdct_backward(const float *in, float *out, int shift, int stride)
{
    const int N = MDCT_N >> shift, N2 = N >> 1, N4 = N >> 2;
    const float sine = (float)2 * 3.141592653f * (.125f) / N;   /* mdct.c:292, PI from mathops.h:83 */
    float f2[FRAME];
    const float *t = g_trig;
    int i;
    /* pre-rotate, mdct.c:295-313 */
    {
        const float *xp1 = in, *xp2 = in + stride * (N2 - 1);
        float *yp = f2;
        for (i = 0; i < N4; i++) {
            float yr = -(*xp2 * t[i << shift]) + *xp1 * t[(N4 - i) << shift];
            float yi = -(*xp2 * t[(N4 - i) << shift]) - *xp1 * t[i << shift];
            *yp++ = yr - yi * sine;
            *yp++ = yi + yr * sine;
            xp1 += 2 * stride;
            xp2 -= 2 * stride;
        }
    }
    /* mdct.c:316 */
    ifft(&g_plan[shift], (const cpx *)f2, (cpx *)(out + HALF_OVL));
    /* post-rotate + de-shuffle from both ends, mdct.c:320-359 */
    {
        float *yp0 = out + HALF_OVL, *yp1 = out + HALF_OVL + N2 - 2;
        for (i = 0; i < (N4 + 1) >> 1; i++) {
            float re = yp0[0], im = yp0[1];
            float t0 = t[i << shift], t1 = t[(N4 - i) << shift];
            float yr = re * t0 - im * t1;
            float yi = im * t0 + re * t1;
            re = yp1[0];
            im = yp1[1];
            yp0[0] = -(yr - yi * sine);
            yp1[1] = yi + yr * sine;
            t0 = t[(N4 - i - 1) << shift];
            t1 = t[(i + 1) << shift];
            yr = re * t0 - im * t1;
            yi = im * t0 + re * t1;
            yp1[0] = -(yr - yi * sine);
            yp0[1] = yi + yr * sine;
            yp0 += 2;
            yp1 -= 2;
        }
    }
    /* TDAC mirror, mdct.c:361-377 */
    {
        float *xp1 = out + OVERLAP - 1, *yp1 = out;
        const float *wp1 = g_window, *wp2 = g_window + OVERLAP - 1;
        for (i = 0; i < OVERLAP / 2; i++) {
            float x1 = *xp1, x2 = *yp1;
            *yp1++ = *wp2 * x2 - *wp1 * x1;
            *xp1-- = *wp1 * x2 + *wp2 * x1;
            wp1++;
            wp2--;
        }
    }
}

NQO_API void nqo_mdct_backward(const float *in, float *out, int shift, int stride)
{
    if (!g_ready) nqo_init_default();
    mdct_backward(in, out, shift, stride);
}

/* celt_decoder_clean.c:264-312 compute_inv_mdcts.  All three branches of the
 * reference perform the same sequence of clt_mdct_backward calls (the _B1_C2
 * wrapper is two plain calls, mdct.c:258-265), so one loop restates them. */
static void inv_mdcts(int shortBlocks, const float *X, float **out_mem, int C, int LM)
{
    int B, N, shift, b, c;
    if (shortBlocks) { B = shortBlocks; N = 120; shift = 3; }
    else { B = 1; N = 120 << LM; shift = 3 - LM; }
    for (c = 0; c < C; c++)
        for (b = 0; b < B; b++)
            mdct_backward(X + b + c * N * B, out_mem[c] + N * b, shift, B);
}

NQO_API void nqo_compute_inv_mdcts(int shortBlocks, const float *X, float **out_mem, int C, int LM)
{
    if (!g_ready) nqo_init_default();
    inv_mdcts(shortBlocks, X, out_mem, C, LM);
}

/* Frame loop of celt_decode_with_ec reduced to synthesis: tail hand-over
 * (celt_decoder_clean.c:622-626), out_syn (:638-642), call (:656). */
static void synth_range(const float *coef, const uint8_t *transient, const float *tail_in,
                        float *pcm_out, float *tail_out, long f0, long f1, int C, int warm)
{
    float *mem = (float *)calloc((size_t)C * (FRAME + HALF_OVL), sizeof(float));
    float **out_syn = (float **)malloc((size_t)C * sizeof(float *));
    long f;
    int c, i;
    for (c = 0; c < C; c++) {
        out_syn[c] = mem + (size_t)c * (FRAME + HALF_OVL);
        if (tail_in) memcpy(out_syn[c] + FRAME, tail_in + c * HALF_OVL, HALF_OVL * sizeof(float));
    }
    for (f = warm ? f0 - 1 : f0; f < f1; f++) {
        for (c = 0; c < C; c++) memmove(out_syn[c], out_syn[c] + FRAME, HALF_OVL * sizeof(float));
        inv_mdcts(transient[f] ? 8 : 0, coef + (size_t)f * C * FRAME, out_syn, C, 3);
        if (f < f0 || !pcm_out) continue;
        for (c = 0; c < C; c++)
            for (i = 0; i < FRAME; i++)
                pcm_out[((size_t)f * FRAME + i) * C + c] = out_syn[c][i];
    }
    if (tail_out)
        for (c = 0; c < C; c++) memcpy(tail_out + c * HALF_OVL, out_syn[c] + FRAME, HALF_OVL * sizeof(float));
    free(out_syn);
    free(mem);
}

typedef struct {
    const float *coef; const uint8_t *transient; const float *tail_in;
    float *pcm_out, *tail_out; long f0, f1; int C, warm;
} job_t;

static void *job_main(void *p)
{
    job_t *j = (job_t *)p;
    synth_range(j->coef, j->transient, j->tail_in, j->pcm_out, j->tail_out, j->f0, j->f1, j->C, j->warm);
    return NULL;
}

/* coef [nframes][C][960], transient [nframes], tail_in/out [C][60] (NULL ok),
 * pcm_out [nframes*960][C] (NULL => compute only).  Contiguous frame ranges
 * on nthreads threads; returns seconds spent in the threaded region. */
NQO_API double nqo_synth_batch(const float *coef, const uint8_t *transient, const float *tail_in,
                               float *pcm_out, float *tail_out, long nframes, int C, int nthreads)
{
    pthread_t *th;
    job_t *jobs;
    struct timespec t0, t1;
    int t;
    if (!g_ready) nqo_init_default();
    if (nthreads < 1) nthreads = 1;
    if (nthreads > nframes) nthreads = nframes > 0 ? (int)nframes : 1;
    th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    jobs = (job_t *)malloc(sizeof(job_t) * nthreads);
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (t = 0; t < nthreads; t++) {
        job_t *j = &jobs[t];
        j->coef = coef; j->transient = transient; j->pcm_out = pcm_out; j->C = C;
        j->f0 = nframes * t / nthreads;
        j->f1 = nframes * (t + 1) / nthreads;
        j->warm = j->f0 > 0;
        j->tail_in = j->f0 == 0 ? tail_in : NULL;
        j->tail_out = t == nthreads - 1 ? tail_out : NULL;
        pthread_create(&th[t], NULL, job_main, j);
    }
    for (t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(jobs);
    free(th);
    return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

/* ======================================================================
 * pitch post-filter + de-emphasis + PCM scale
 * ====================================================================== */

/* celt.c:114-172 comb_filter, in place or not, float build.  The constant
 * part follows the SSE override the reference actually compiles on x86-64
 * (celt/x86/pitch_sse.h:104-151: partial sums, four samples per step, loads
 * before stores -- equivalent to the per-sample order below because T >= 15). */
static void comb_filter(float *y, const float *x, int T0, int T1, int N, float g0, float g1,
                        int tapset0, int tapset1)
{
    static const float gains[3][3] = {
        {0.3066406250f, 0.2170410156f, 0.1296386719f},
        {0.4638671875f, 0.2680664062f, 0.f},
        {0.7998046875f, 0.1000976562f, 0.f}};
    float g00, g01, g02, g10, g11, g12, x0, x1, x2, x3, x4;
    int i;
    if (g0 == 0 && g1 == 0) {
        if (x != y) memmove(y, x, N * sizeof(float));
        return;
    }
    g00 = g0 * gains[tapset0][0]; g01 = g0 * gains[tapset0][1]; g02 = g0 * gains[tapset0][2];
    g10 = g1 * gains[tapset1][0]; g11 = g1 * gains[tapset1][1]; g12 = g1 * gains[tapset1][2];
    x1 = x[-T1 + 1]; x2 = x[-T1]; x3 = x[-T1 - 1]; x4 = x[-T1 - 2];
    for (i = 0; i < OVERLAP; i++) {
        float f;
        x0 = x[i - T1 + 2];
        f = g_window[i] * g_window[i];
        y[i] = x[i]
             + ((1.0f - f) * g00) * x[i - T0]
             + ((1.0f - f) * g01) * (x[i - T0 + 1] + x[i - T0 - 1])
             + ((1.0f - f) * g02) * (x[i - T0 + 2] + x[i - T0 - 2])
             + (f * g10) * x2
             + (f * g11) * (x1 + x3)
             + (f * g12) * (x0 + x4);
        x4 = x3; x3 = x2; x2 = x1; x1 = x0;
    }
    if (g1 == 0) {
        if (x != y) memmove(y + OVERLAP, x + OVERLAP, (N - OVERLAP) * sizeof(float));
        return;
    }
    for (; i < N; i++)
        y[i] = (x[i] + g10 * x[i - T1])
             + (g11 * (x[i - T1 + 1] + x[i - T1 - 1]) + g12 * (x[i - T1 + 2] + x[i - T1 - 2]));
}

NQO_API void nqo_comb_filter(float *y, float *x, int T0, int T1, int N, float g0, float g1, int tapset0, int tapset1)
{
    if (!g_ready) nqo_init_default();
    comb_filter(y, x, T0, T1, N, g0, g1, tapset0, tapset1);
}

/* Per-frame side information phase 1 hands to the post stage: the arguments of the two
 * comb_filter calls at celt_decoder_clean.c:663-669 (old -> cur over [0,120), cur -> new
 * over [120,N)); layout shared with include/nq_celt_synth.h nq_celt_post_frame. */
typedef struct {
    int32_t N;          /* samples per channel of this frame: 120 << LM */
    int32_t pitch[3];   /* postfilter_period_old, postfilter_period, postfilter_pitch */
    float gain[3];
    int32_t tapset[3];
} post_frame;

enum { HIST = 1026 };   /* COMBFILTER_MAXPERIOD + 2 samples of filtered history, celt.h:187 */

/* comb_filter x2 (celt_decoder_clean.c:658-670) then deemphasis (:723, :192-256) over a run of
 * frames.  sig [nsamples][C] interleaved celt_sig (the synthesis output), pcm [nsamples][C]
 * float in [-1,1].  hist_in/out [C][1026]: last filtered samples; mem_in/out [C]: preemph_memD. */
NQO_API void nqo_post_batch(const float *sig, const post_frame *frames, long nframes, int C,
                            const float *hist_in, const float *mem_in, float *pcm,
                            float *hist_out, float *mem_out)
{
    const float coef0 = 0.85000610f;   /* mode->preemph[0], static_modes_float.h:583 */
    long nsamples = 0, f, n0;
    int c, j;
    float *buf;
    if (!g_ready) nqo_init_default();
    for (f = 0; f < nframes; f++) nsamples += frames[f].N;
    buf = (float *)calloc((size_t)HIST + nsamples, sizeof(float));
    for (c = 0; c < C; c++) {
        float m = mem_in ? mem_in[c] : 0.f;
        float *x = buf + HIST;
        if (hist_in) memcpy(buf, hist_in + (size_t)c * HIST, HIST * sizeof(float));
        else memset(buf, 0, HIST * sizeof(float));
        for (n0 = 0; n0 < nsamples; n0++) x[n0] = sig[n0 * C + c];
        n0 = 0;
        for (f = 0; f < nframes; f++) {
            const post_frame *p = &frames[f];
            float *o = x + n0;
            comb_filter(o, o, p->pitch[0], p->pitch[1], 120, p->gain[0], p->gain[1], p->tapset[0], p->tapset[1]);
            if (p->N > 120)
                comb_filter(o + 120, o + 120, p->pitch[1], p->pitch[2], p->N - 120, p->gain[1], p->gain[2],
                            p->tapset[1], p->tapset[2]);
            for (j = 0; j < p->N; j++) {
                float tmp = o[j] + m + 1e-30f;
                m = coef0 * tmp;
                pcm[(n0 + j) * C + c] = tmp * (1 / 32768.f);
            }
            n0 += p->N;
        }
        if (hist_out) memcpy(hist_out + (size_t)c * HIST, x + nsamples - HIST, HIST * sizeof(float));
        if (mem_out) mem_out[c] = m;
    }
    free(buf);
}
