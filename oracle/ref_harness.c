/* TEST INFRASTRUCTURE -- builds oracle/_ref/libnq_ref.so, the UNMODIFIED
 * reference compiled in place from /root/reference (see oracle/Makefile).
 *
 * Nothing of the reference is copied: this file unity-includes the
 * reference's own translation unit (src/OpusDependencies.c, which itself
 * #includes every CELT/SILK/opusfile/ogg .c file, lines 78-270) and adds thin
 * exported wrappers so Python tests / bench.py's cpu_baseline and reference
 * arm can call the reference functions on arbitrary buffers:
 *
 *   opus_ifft            third_party/opus/celt/kiss_fft.c:696
 *   clt_mdct_backward    third_party/opus/celt/mdct.c:267
 *   compute_inv_mdcts    third_party/opus/celt/celt_decoder_clean.c:264 (static)
 *   whole-file decode    opusfile op_open_memory / op_read_float
 *
 * Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline leg and
 * --impl reference) may load the resulting library.  The product
 * (libnyquist_b200/) never does.
 */
#include "OpusDependencies.c"

#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define NQREF_API __attribute__((visibility("default")))

#define FRAME 960   /* samples per channel per 20 ms frame (LM = 3)        */
#define OVL 120     /* mode->overlap                                        */
#define HALF_OVL 60 /* raw tail carried between consecutive (sub-)blocks   */

static const CELTMode *the_mode(void)
{
    return opus_custom_mode_create(48000, 960, NULL);
}

/* ---- static tables (static_modes_float.h:9,99,343..417,477) ------------- */
NQREF_API void nqref_tables(float *window120, float *trig481, float *twiddles480_ri,
                            int16_t *bitrev480, int16_t *bitrev240,
                            int16_t *bitrev120, int16_t *bitrev60)
{
    const CELTMode *m = the_mode();
    int i;
    for (i = 0; i < 120; i++) window120[i] = m->window[i];
    for (i = 0; i <= 480; i++) trig481[i] = m->mdct.trig[i];
    for (i = 0; i < 480; i++) {
        twiddles480_ri[2 * i] = m->mdct.kfft[0]->twiddles[i].r;
        twiddles480_ri[2 * i + 1] = m->mdct.kfft[0]->twiddles[i].i;
    }
    for (i = 0; i < 480; i++) bitrev480[i] = m->mdct.kfft[0]->bitrev[i];
    for (i = 0; i < 240; i++) bitrev240[i] = m->mdct.kfft[1]->bitrev[i];
    for (i = 0; i < 120; i++) bitrev120[i] = m->mdct.kfft[2]->bitrev[i];
    for (i = 0; i < 60; i++) bitrev60[i] = m->mdct.kfft[3]->bitrev[i];
}

/* ---- a3.2: opus_ifft with the static state kfft[shift] ------------------ */
NQREF_API void nqref_opus_ifft(int shift, const float *in_ri, float *out_ri)
{
    const CELTMode *m = the_mode();
    opus_ifft(m->mdct.kfft[shift], (const kiss_fft_cpx *)in_ri, (kiss_fft_cpx *)out_ri);
}

/* ---- a3: one clt_mdct_backward call (out is read-modify-write) ---------- */
NQREF_API void nqref_clt_mdct_backward(float *in, float *out, int shift, int stride)
{
    const CELTMode *m = the_mode();
    clt_mdct_backward(&m->mdct, in, out, m->window, OVL, shift, stride);
}

/* ---- a1: compute_inv_mdcts on caller buffers ---------------------------- */
NQREF_API void nqref_compute_inv_mdcts(int shortBlocks, float *X, float **out_mem,
                                       int C, int LM)
{
    compute_inv_mdcts(the_mode(), shortBlocks, X, out_mem, C, LM);
}

/* ---- batch driver: the frame loop of celt_decode_with_ec reduced to the
 *      synthesis stage (history shift :622-626, out_syn :638-642, call :656).
 *      coef      [nframes][C][960]
 *      transient [nframes]           (non-zero => shortBlocks = 8)
 *      tail_in   [C][60] or NULL     (raw tail of the frame before the batch)
 *      pcm_out   [nframes*960][C]    interleaved celt_sig (pre post-filter)
 *      tail_out  [C][60] or NULL
 */
static void synth_range(const float *coef, const uint8_t *transient,
                        const float *tail_in, float *pcm_out, float *tail_out,
                        long f0, long f1, int C, int warm)
{
    const CELTMode *m = the_mode();
    float *mem = (float *)calloc((size_t)C * (FRAME + HALF_OVL), sizeof(float));
    float *X = (float *)malloc((size_t)C * FRAME * sizeof(float));
    float **out_syn = (float **)malloc((size_t)C * sizeof(float *));
    long f;
    int c, i;
    for (c = 0; c < C; c++) {
        out_syn[c] = mem + (size_t)c * (FRAME + HALF_OVL);
        if (tail_in)
            memcpy(out_syn[c] + FRAME, tail_in + c * HALF_OVL, HALF_OVL * sizeof(float));
    }
    /* warm: recompute frame f0-1 only to obtain its raw tail (shard halo) */
    for (f = warm ? f0 - 1 : f0; f < f1; f++) {
        for (c = 0; c < C; c++) /* what OPUS_MOVE at :625 does to the tail */
            memmove(out_syn[c], out_syn[c] + FRAME, HALF_OVL * sizeof(float));
        memcpy(X, coef + (size_t)f * C * FRAME, (size_t)C * FRAME * sizeof(float));
        compute_inv_mdcts(m, transient[f] ? 8 : 0, X, out_syn, C, 3);
        if (f < f0) continue;
        if (pcm_out)
            for (c = 0; c < C; c++)
                for (i = 0; i < FRAME; i++)
                    pcm_out[((size_t)f * FRAME + i) * C + c] = out_syn[c][i];
    }
    if (tail_out)
        for (c = 0; c < C; c++)
            memcpy(tail_out + c * HALF_OVL, out_syn[c] + FRAME, HALF_OVL * sizeof(float));
    free(out_syn);
    free(X);
    free(mem);
}

NQREF_API void nqref_synth_batch(const float *coef, const uint8_t *transient,
                                 const float *tail_in, float *pcm_out, float *tail_out,
                                 long nframes, int C)
{
    synth_range(coef, transient, tail_in, pcm_out, tail_out, 0, nframes, C, 0);
}

typedef struct {
    const float *coef; const uint8_t *transient; const float *tail_in;
    float *pcm_out; float *tail_out; long f0, f1; int C; int warm;
} range_job;

static void *range_thread(void *p)
{
    range_job *j = (range_job *)p;
    synth_range(j->coef, j->transient, j->tail_in, j->pcm_out, j->tail_out,
                j->f0, j->f1, j->C, j->warm);
    return NULL;
}

/* Same result as nqref_synth_batch, contiguous frame ranges on nthreads host
 * threads; a range that does not start at 0 recomputes the previous frame
 * for its tail.  Returns the wall time in seconds of the threaded region. */
NQREF_API double nqref_synth_batch_mt(const float *coef, const uint8_t *transient,
                                      const float *tail_in, float *pcm_out, float *tail_out,
                                      long nframes, int C, int nthreads)
{
    pthread_t *th;
    range_job *jobs;
    struct timespec t0, t1;
    int t;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > nframes) nthreads = nframes > 0 ? (int)nframes : 1;
    th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    jobs = (range_job *)malloc(sizeof(range_job) * nthreads);
    the_mode();
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (t = 0; t < nthreads; t++) {
        range_job *j = &jobs[t];
        j->coef = coef; j->transient = transient; j->pcm_out = pcm_out; j->C = C;
        j->f0 = nframes * t / nthreads;
        j->f1 = nframes * (t + 1) / nthreads;
        j->warm = j->f0 > 0;
        j->tail_in = j->f0 == 0 ? tail_in : NULL;
        j->tail_out = t == nthreads - 1 ? tail_out : NULL;
        pthread_create(&th[t], NULL, range_thread, j);
    }
    for (t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(jobs);
    free(th);
    return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

/* ---- whole-file decode with taps on the inverse-MDCT call sites ----------
 * The overlay oracle/ref_overlay/opus/celt/celt_decoder_clean.c routes the
 * calls at celt_decoder_clean.c:290,298,309 through these taps.  When a
 * recording is active each tap appends one record:
 *   header  int32[4] = {nch_in_call, shift, stride(B), N2*B}
 *   coef    nch * N2*B floats  (the frame's freq[] for these channels, as
 *                               handed to compute_inv_mdcts)
 *   out     nch * N2*B floats  (out_syn after the last sub-block, i.e. the
 *                               synthesis output before comb_filter)
 */
typedef struct {
    int active;
    float *buf; size_t len, cap;   /* in floats */
    float *cbuf; size_t clen, ccap; /* comb_filter call log, 8 floats per call */
    long nrecords;
    int b;                         /* sub-block counter inside a stride-B group */
    float *xbase[2], *obase[2];
} recorder;
static recorder g_rec;

static void rec_reserve(size_t extra)
{
    if (g_rec.len + extra > g_rec.cap) {
        size_t ncap = g_rec.cap ? g_rec.cap * 2 : (1u << 20);
        while (ncap < g_rec.len + extra) ncap *= 2;
        g_rec.buf = (float *)realloc(g_rec.buf, ncap * sizeof(float));
        g_rec.cap = ncap;
    }
}

static void rec_group(int nch, int shift, int B, float **in, float **out)
{
    int N2 = (1920 >> shift) >> 1, c;
    if (g_rec.b == 0)
        for (c = 0; c < nch; c++) { g_rec.xbase[c] = in[c]; g_rec.obase[c] = out[c]; }
    if (++g_rec.b < B) return;
    g_rec.b = 0;
    {
        int32_t hdr[4] = {nch, shift, B, N2 * B};
        size_t n = (size_t)N2 * B;
        rec_reserve(4 + 2 * nch * n);
        memcpy(g_rec.buf + g_rec.len, hdr, sizeof hdr); g_rec.len += 4;
        for (c = 0; c < nch; c++) { memcpy(g_rec.buf + g_rec.len, g_rec.xbase[c], n * 4); g_rec.len += n; }
        for (c = 0; c < nch; c++) { memcpy(g_rec.buf + g_rec.len, g_rec.obase[c], n * 4); g_rec.len += n; }
        g_rec.nrecords++;
    }
}

void nqref_tap_mdct_b1c2(const mdct_lookup *l, float *in[2], float *out[2],
                         const float *window, int overlap, int shift, int stride)
{
    clt_mdct_backward_B1_C2(l, in, out, window, overlap, shift, stride);
    if (g_rec.active) rec_group(2, shift, stride, in, out);
}

void nqref_tap_mdct(const mdct_lookup *l, float *in, float *out,
                    const float *window, int overlap, int shift, int stride)
{
    clt_mdct_backward(l, in, out, window, overlap, shift, stride);
    if (g_rec.active) rec_group(1, shift, stride, &in, &out);
}

/* comb_filter tap: one 8-float record per call {N, T0, T1, g0, g1, tapset0, tapset1, stream},
 * stream = index of the calling CELT decoder in first-seen order, which is the order
 * opus_multistream_decode_native walks the streams in (opus_multistream_decoder.c:237). */
static const void *g_decoders[256];
static int g_ndecoders;

void nqref_tap_comb_filter(const void *decoder, float *y, float *x, int T0, int T1, int N, float g0, float g1,
                           int tapset0, int tapset1, const float *window, int overlap)
{
    comb_filter(y, x, T0, T1, N, g0, g1, tapset0, tapset1, window, overlap);
    if (g_rec.active) {
        float *r;
        int sidx = 0;
        while (sidx < g_ndecoders && g_decoders[sidx] != decoder) sidx++;
        if (sidx == g_ndecoders && g_ndecoders < 256) g_decoders[g_ndecoders++] = decoder;
        if (g_rec.clen + 8 > g_rec.ccap) {
            g_rec.ccap = g_rec.ccap ? g_rec.ccap * 2 : (1u << 16);
            g_rec.cbuf = (float *)realloc(g_rec.cbuf, g_rec.ccap * sizeof(float));
        }
        r = g_rec.cbuf + g_rec.clen;
        r[0] = (float)N; r[1] = (float)T0; r[2] = (float)T1; r[3] = g0; r[4] = g1;
        r[5] = (float)tapset0; r[6] = (float)tapset1; r[7] = (float)sidx;
        g_rec.clen += 8;
    }
}

static int g_last_pre_skip, g_last_output_gain, g_last_streams, g_last_coupled, g_last_channels;
static unsigned char g_last_mapping[256];
NQREF_API int nqref_last_pre_skip(void) { return g_last_pre_skip; }
NQREF_API int nqref_last_output_gain(void) { return g_last_output_gain; }
/* OpusHead layout of the file decoded last: returns channels, fills streams / coupled / mapping[channels] */
NQREF_API int nqref_last_layout(int *streams, int *coupled, unsigned char *mapping)
{
    *streams = g_last_streams; *coupled = g_last_coupled;
    memcpy(mapping, g_last_mapping, (size_t)g_last_channels);
    return g_last_channels;
}
NQREF_API size_t nqref_comb_floats(void) { return g_rec.clen; }
NQREF_API void nqref_comb_copy(float *dst) { memcpy(dst, g_rec.cbuf, g_rec.clen * sizeof(float)); }

/* comb_filter (celt.c:114) and deemphasis (celt_decoder_clean.c:192) on caller buffers, for
 * pinning the oracle's restatement.  x points INSIDE a buffer with >= T+2 samples of history. */
NQREF_API void nqref_comb_filter(float *y, float *x, int T0, int T1, int N, float g0, float g1,
                                 int tapset0, int tapset1)
{
    comb_filter(y, x, T0, T1, N, g0, g1, tapset0, tapset1, the_mode()->window, OVL);
}

NQREF_API void nqref_deemphasis(float **in, float *pcm, int N, int C, float *mem)
{
    float scratch[FRAME];
    deemphasis(in, pcm, N, C, 1, the_mode()->preemph, mem, scratch);
}

/* Decode an in-memory Ogg Opus file the way src/OpusDecoder.cpp:57-119 does
 * (op_test_memory/op_test_open/op_read_float loop).  Returns samples per
 * channel decoded (<0 on error).  pcm may be NULL (count only).  If record
 * != 0 the taps above log every synthesis call; fetch with nqref_record_*. */
NQREF_API long nqref_decode_memory(const unsigned char *data, size_t nbytes,
                                   float *pcm, long pcm_capacity_floats,
                                   int *channels_out, int record)
{
    int err = 0, ch;
    long total = 0;
    float scratch[5760 * 8];
    OggOpusFile *of = op_test_memory(data, nbytes, &err);
    if (!of) return -1;
    if (op_test_open(of) != 0) return -2;   /* frees of on failure */
    ch = op_head(of, 0)->channel_count;
    g_last_pre_skip = (int)op_head(of, 0)->pre_skip;
    g_last_output_gain = op_head(of, 0)->output_gain;
    g_last_streams = op_head(of, 0)->stream_count;
    g_last_coupled = op_head(of, 0)->coupled_count;
    g_last_channels = ch;
    memcpy(g_last_mapping, op_head(of, 0)->mapping, (size_t)(ch < 8 ? ch : 8));
    g_ndecoders = 0;
    if (channels_out) *channels_out = ch;
    g_rec.active = record; g_rec.len = 0; g_rec.clen = 0; g_rec.nrecords = 0; g_rec.b = 0;
    for (;;) {
        float *dst = scratch;
        int room = (int)(sizeof scratch / sizeof scratch[0]);
        int got;
        if (pcm) {
            long left = pcm_capacity_floats - total * ch;
            if (left <= 0) break;
            dst = pcm + total * ch;
            room = left > (1 << 30) ? (1 << 30) : (int)left;
        }
        got = op_read_float(of, dst, room, NULL);
        if (got == 0) break;
        if (got < 0) { total = got; break; }
        total += got;
    }
    g_rec.active = 0;
    op_free(of);
    return total;
}

/* ---- an Ogg Opus multistream file made with the reference's own ENCODER --------------------
 * The reference mount lacks its 8-channel test file (test_data/Rachel8ch.opus is listed in
 * .MISSING_LARGE_BLOBS), so BASELINE config 4 is exercised on a file produced here: the bundled
 * libopus surround encoder (opus_multistream_encoder.c:557, mapping family 1, CELT-only through
 * OPUS_APPLICATION_RESTRICTED_LOWDELAY) + the bundled libogg for the container (RFC 7845:
 * OpusHead, OpusTags, audio pages with 48 kHz granule positions).
 * pcm [nsamples][channels] float, 20 ms frames.  Returns the number of bytes written to out
 * (<0 on error). */
static int put_page(ogg_stream_state *os, int flush, unsigned char *out, long cap, long *pos)
{
    ogg_page og;
    while (flush ? ogg_stream_flush(os, &og) : ogg_stream_pageout(os, &og)) {
        if (*pos + og.header_len + og.body_len > cap) return -1;
        memcpy(out + *pos, og.header, (size_t)og.header_len); *pos += og.header_len;
        memcpy(out + *pos, og.body, (size_t)og.body_len); *pos += og.body_len;
    }
    return 0;
}

/* force_mode: 0 = the encoder decides (RESTRICTED_LOWDELAY: CELT only), else MODE_SILK_ONLY (1000) /
 * MODE_HYBRID (1001) / MODE_CELT_ONLY (1002) through the private OPUS_SET_FORCE_MODE ctl
 * (opus_private.h:90-96), with the VOIP application -- test files for the SILK / hybrid paths of
 * opus_decode_frame (opus_decoder_clean.c:340-600). */
static int g_switch_frame = -1, g_switch_mode = 0;   /* nqref_encode_mode_switch: from this frame on, that mode */
static const int *g_sched_frames = NULL, *g_sched_modes = NULL;   /* nqref_encode_mode_schedule: several switches */
static int g_sched_n = 0;

static long encode_ogg(const float *pcm, long nsamples, int channels, int bitrate, int force_mode,
                       unsigned char *out, long cap)
{
    int err = 0, streams = 0, coupled = 0, lookahead = 0, i;
    unsigned char mapping[255], head[32 + 255], pkt[8 * 1500];
    const char *vendor = "nq-oracle";
    unsigned char tags[64];
    long pos = 0, f, nframes = nsamples / FRAME;
    ogg_stream_state os;
    ogg_packet op;
    OpusMSEncoder *enc = opus_multistream_surround_encoder_create(48000, channels, channels > 2 ? 1 : 0, &streams, &coupled,
                                                                  mapping, force_mode ? OPUS_APPLICATION_VOIP : OPUS_APPLICATION_RESTRICTED_LOWDELAY, &err);
    if (!enc || err != OPUS_OK) return -1;
    opus_multistream_encoder_ctl(enc, OPUS_SET_BITRATE(bitrate));
    if (force_mode) {
        opus_multistream_encoder_ctl(enc, OPUS_SET_FORCE_MODE(force_mode));
        opus_multistream_encoder_ctl(enc, OPUS_SET_BANDWIDTH(force_mode == MODE_SILK_ONLY ? OPUS_BANDWIDTH_WIDEBAND : OPUS_BANDWIDTH_FULLBAND));
    }
    opus_multistream_encoder_ctl(enc, OPUS_GET_LOOKAHEAD(&lookahead));
    if (ogg_stream_init(&os, 0x0B200) != 0) return -2;
    /* OpusHead, RFC 7845 section 5.1 */
    memcpy(head, "OpusHead", 8);
    head[8] = 1; head[9] = (unsigned char)channels;
    head[10] = (unsigned char)(lookahead & 255); head[11] = (unsigned char)(lookahead >> 8);
    head[12] = 0x80; head[13] = 0xBB; head[14] = 0; head[15] = 0;   /* 48000 */
    head[16] = 0; head[17] = 0;                                     /* output gain 0 */
    head[18] = (unsigned char)(channels > 2 ? 1 : 0);
    i = 19;
    if (channels > 2) {
        head[i++] = (unsigned char)streams; head[i++] = (unsigned char)coupled;
        memcpy(head + i, mapping, (size_t)channels); i += channels;
    }
    memset(&op, 0, sizeof op);
    op.packet = head; op.bytes = i; op.b_o_s = 1; op.packetno = 0;
    ogg_stream_packetin(&os, &op);
    if (put_page(&os, 1, out, cap, &pos)) return -3;
    /* OpusTags, section 5.2 */
    memcpy(tags, "OpusTags", 8);
    i = (int)strlen(vendor);
    tags[8] = (unsigned char)i; tags[9] = tags[10] = tags[11] = 0;
    memcpy(tags + 12, vendor, (size_t)i);
    memset(tags + 12 + i, 0, 4);
    memset(&op, 0, sizeof op);
    op.packet = tags; op.bytes = 16 + i; op.packetno = 1;
    ogg_stream_packetin(&os, &op);
    if (put_page(&os, 1, out, cap, &pos)) return -3;
    for (f = 0; f < nframes; f++) {
        int n;
        if (f == g_switch_frame) opus_multistream_encoder_ctl(enc, OPUS_SET_FORCE_MODE(g_switch_mode));
        for (i = 0; i < g_sched_n; i++)
            if (f == g_sched_frames[i]) {   /* (SILK-only needs a SILK bandwidth, or the encoder makes it hybrid) */
                opus_multistream_encoder_ctl(enc, OPUS_SET_FORCE_MODE(g_sched_modes[i]));
                opus_multistream_encoder_ctl(enc, OPUS_SET_BANDWIDTH(g_sched_modes[i] == MODE_SILK_ONLY ? OPUS_BANDWIDTH_WIDEBAND : OPUS_BANDWIDTH_FULLBAND));
            }
        n = opus_multistream_encode_float(enc, pcm + f * FRAME * channels, FRAME, pkt, (opus_int32)sizeof pkt);
        if (n < 0) return -4;
        memset(&op, 0, sizeof op);
        op.packet = pkt; op.bytes = n; op.packetno = 2 + f;
        op.e_o_s = f == nframes - 1;
        /* granule position = samples up to and including this packet; the last one is trimmed so the
         * file plays nsamples - lookahead... keep every decoded sample: total = all frames - pre-skip */
        op.granulepos = (f + 1) * FRAME;
        ogg_stream_packetin(&os, &op);
        if (put_page(&os, 0, out, cap, &pos)) return -3;
    }
    if (put_page(&os, 1, out, cap, &pos)) return -3;
    ogg_stream_clear(&os);
    opus_multistream_encoder_destroy(enc);
    return pos;
}

NQREF_API long nqref_encode_surround(const float *pcm, long nsamples, int channels, int bitrate,
                                     unsigned char *out, long cap)
{
    return encode_ogg(pcm, nsamples, channels, bitrate, 0, out, cap);
}

NQREF_API long nqref_encode_forced_mode(const float *pcm, long nsamples, int channels, int bitrate, int force_mode,
                                        unsigned char *out, long cap)
{
    return encode_ogg(pcm, nsamples, channels, bitrate, force_mode, out, cap);
}

/* A file that switches coding mode at `switch_frame` (the encoder inserts the redundancy frames and
 * the decoder cross-fades, opus_decoder_clean.c:570-600): what the two-phase loader must refuse. */
NQREF_API long nqref_encode_mode_switch(const float *pcm, long nsamples, int channels, int bitrate, int mode_a,
                                        int mode_b, int switch_frame, unsigned char *out, long cap)
{
    long n;
    g_switch_frame = switch_frame;
    g_switch_mode = mode_b;
    n = encode_ogg(pcm, nsamples, channels, bitrate, mode_a, out, cap);
    g_switch_frame = -1;
    return n;
}

/* A file that walks through several coding modes: from frame frames[i] on, mode modes[i]. */
NQREF_API long nqref_encode_mode_schedule(const float *pcm, long nsamples, int channels, int bitrate, int first_mode,
                                          const int *frames, const int *modes, int nswitch, unsigned char *out, long cap)
{
    long n;
    g_sched_frames = frames;
    g_sched_modes = modes;
    g_sched_n = nswitch;
    n = encode_ogg(pcm, nsamples, channels, bitrate, first_mode, out, cap);
    g_sched_n = 0;
    return n;
}

/* Rewrites the output-gain field (Q7.8 dB, RFC 7845 section 5.1) of an in-memory Ogg Opus file's
 * OpusHead and fixes the page checksum, so the header-gain path (opusfile OP_HEADER_GAIN ->
 * OPUS_SET_GAIN -> opus_decoder_clean.c:578-588) can be exercised on the bundled files. */
NQREF_API int nqref_patch_output_gain(unsigned char *data, long nbytes, int gain_q8)
{
    ogg_page og;
    long i, body_len = 0;
    int nsegs;
    if (nbytes < 47 || memcmp(data, "OggS", 4) != 0) return -1;
    nsegs = data[26];
    for (i = 0; i < nsegs; i++) body_len += data[27 + i];
    if (27 + nsegs + body_len > nbytes || body_len < 19 || memcmp(data + 27 + nsegs, "OpusHead", 8) != 0) return -2;
    data[27 + nsegs + 16] = (unsigned char)(gain_q8 & 255);
    data[27 + nsegs + 17] = (unsigned char)((gain_q8 >> 8) & 255);
    og.header = data; og.header_len = 27 + nsegs;
    og.body = data + 27 + nsegs; og.body_len = body_len;
    ogg_page_checksum_set(&og);
    return 0;
}

NQREF_API long nqref_record_count(void) { return g_rec.nrecords; }
NQREF_API size_t nqref_record_floats(void) { return g_rec.len; }
NQREF_API void nqref_record_copy(float *dst) { memcpy(dst, g_rec.buf, g_rec.len * sizeof(float)); }
NQREF_API void nqref_record_free(void)
{
    free(g_rec.buf); g_rec.buf = NULL; g_rec.len = g_rec.cap = 0; g_rec.nrecords = 0;
    free(g_rec.cbuf); g_rec.cbuf = NULL; g_rec.clen = g_rec.ccap = 0;
}
