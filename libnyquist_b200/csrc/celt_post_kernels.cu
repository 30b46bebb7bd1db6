// CELT post stage for sm_100a: pitch post-filter (comb filter) + de-emphasis + PCM scaling.
//
// What the reference does per channel per frame (third_party/opus/celt/):
//   comb_filter x2      celt_decoder_clean.c:658-670 -> celt.c:114-172 (constant part: x86/pitch_sse.h:104)
//                       IN PLACE on out_syn, i.e. recursive: y[i] reads y[i-T-2 .. i-T+2], T >= 15
//   parameter hand-over celt_decoder_clean.c:672-683 (old <- cur <- new)
//   deemphasis          celt_decoder_clean.c:192-256, called at :723: tmp = x + m; m = 0.85000610 tmp;
//                       pcm = tmp / 32768
//
// Both are recurrences along time, so -- unlike the synthesis -- a stream cannot be cut into
// independent runs; the parallelism is
//   * across streams (one 128-thread CTA per stream of 1 or 2 channels: a batch of many files /
//     multistream sub-decoders keeps the GPU busy, one file keeps one CTA busy);
//   * inside a frame: the comb filter's nearest tap is T-2 samples back, so blocks of T-2 >= 13
//     samples are independent (typical pitch periods give 100-1000 samples per step), and both
//     channels of a coupled stream share one instruction stream;
//   * the de-emphasis IIR is evaluated as 120 thread-private segments plus a scan of the segment
//     carries (affine maps m -> m_seg + a^len * m), not sample by sample.
// The last 1024+2 filtered samples per channel live in shared memory in front of the current
// frame (the reference keeps them in decode_mem, celt_decoder_clean.c:92); a frame is read from
// HBM once (coalesced, one frame ahead), filtered in place, de-emphasised into a staging buffer and
// written once (coalesced).  The kernel works in place on the interleaved [nsamples][C] buffer the
// synthesis kernel wrote.
#include "celt_synth_kernels.cuh"

namespace nq {

constexpr int kPostThreads = 128;            // one CTA = one stream (1 or 2 channels)
constexpr int kPostSegs = 120;               // de-emphasis segments: N / 120 = 8, 4, 2, 1 samples each
constexpr unsigned kFullMask = 0xffffffffu;

constexpr int kOff = 1028;                   // the current frame starts here: 16-byte aligned, history in [2, 1028)
constexpr int kStagePad = 4;                 // floats of padding after every 8 samples of the staging buffer

struct __align__(16) PostSmem {
    float buf[2][kOff + kFrame + 4];         // per channel: filtered history, then the current frame
    // de-emphasised frame, 8 samples x nch channels then 4 floats of padding: a thread that owns
    // 8 consecutive samples reads / writes them as 128-bit words without bank conflicts
    float stage[(kFrame / 8) * (16 + kStagePad)];
    float win2[kOverlap];                   // window[i]^2, celt.c:147
    float wtot[2][4];                       // de-emphasis: per-warp carries
    float mem[2];                           // de-emphasis state (preemph_memD)
};

size_t post_kernel_smem_bytes() { return sizeof(PostSmem); }

// celt.c:121-124
__constant__ float c_tap_gains[3][3] = {{0.3066406250f, 0.2170410156f, 0.1296386719f},
                                        {0.4638671875f, 0.2680664062f, 0.f},
                                        {0.7998046875f, 0.1000976562f, 0.f}};

struct Taps {
    int T;
    float g0, g1, g2;   // g * gains[tapset][0..2]
    bool on;            // g != 0
};

__device__ __forceinline__ Taps make_taps(int T, float g, int tapset)
{
    Taps t;
    t.T = T;
    t.g0 = g * c_tap_gains[tapset][0];
    t.g1 = g * c_tap_gains[tapset][1];
    t.g2 = g * c_tap_gains[tapset][2];
    t.on = g != 0.f;
    return t;
}

// One region [a, b) of a frame, in place (frame sample i lives at buf[ch][kOff + i]).
//   xfade: celt.c:142-166, the filter fades from `t0` to `t1` with window^2 over the region
//   else : celt.c:176 / pitch_sse.h:104, constant filter `t1`
// Samples inside a block of min(T)-2 are independent of each other; blocks run in order, one
// CTA barrier apart.  The branch structure is uniform over the CTA; which of the two filters are
// live is a template parameter, so that the sample loop carries no flags.
template <int NCH, bool kXfade, bool kUse0, bool kUse1>
__device__ __forceinline__ void comb_region_t(PostSmem &sm, int a, int b, const Taps &t0, const Taps &t1, int tid)
{
    int B = b - a;
    if (kUse0 && t0.T - 2 < B) B = t0.T - 2;
    if (kUse1 && t1.T - 2 < B) B = t1.T - 2;
    for (int i0 = a; i0 < b; i0 += B) {
        const int iend = i0 + B < b ? i0 + B : b;
        for (int i = i0 + tid; i < iend; i += kPostThreads) {
            float f = 1.f;
            if (kXfade) f = sm.win2[i - a];
            const float e = 1.f - f;
#pragma unroll
            for (int ch = 0; ch < NCH; ch++) {
                float *x = sm.buf[ch] + kOff + i;
                float acc = x[0];
                if (kUse0) {
                    const float *q = x - t0.T;
                    acc += (e * t0.g0) * q[0];
                    acc += (e * t0.g1) * (q[1] + q[-1]);
                    acc += (e * t0.g2) * (q[2] + q[-2]);
                }
                if (kUse1) {
                    const float *q = x - t1.T;
                    if (kXfade) {
                        acc += (f * t1.g0) * q[0];
                        acc += (f * t1.g1) * (q[1] + q[-1]);
                        acc += (f * t1.g2) * (q[2] + q[-2]);
                    } else {   // partial sums as pitch_sse.h:136-139
                        acc += t1.g0 * q[0];
                        acc += t1.g1 * (q[1] + q[-1]) + t1.g2 * (q[2] + q[-2]);
                    }
                }
                x[0] = acc;
            }
        }
        __syncthreads();
    }
}

template <int NCH>
__device__ __forceinline__ void comb_region(PostSmem &sm, int a, int b, const Taps &t0, const Taps &t1, bool xfade, int tid)
{
    const bool use0 = xfade && t0.on, use1 = t1.on;
    if (!use0 && !use1) return;   // celt.c:126-132 and :167-173: the filter is the identity here
    if (!xfade) comb_region_t<NCH, false, false, true>(sm, a, b, t0, t1, tid);
    else if (use0 && use1) comb_region_t<NCH, true, true, true>(sm, a, b, t0, t1, tid);
    else if (use0) comb_region_t<NCH, true, true, false>(sm, a, b, t0, t1, tid);
    else comb_region_t<NCH, true, false, true>(sm, a, b, t0, t1, tid);
}

// float offset of sample n (channel 0) in the staging buffer
template <int NCH>
__device__ __forceinline__ int stage_at(int n) { return (n >> 3) * (8 * NCH + kStagePad) + (n & 7) * NCH; }

// deemphasis, celt_decoder_clean.c:232-241, over the N filtered samples of the frame: thread
// t < 120 owns samples [t*seg, (t+1)*seg), seg = N/120; the segment carries (affine maps
// m -> m_seg + a^seg m) are combined by a warp scan plus a 4-entry hand-over between the warps.
template <int NCH>
__device__ __forceinline__ void deemphasis_frame(PostSmem &sm, int N, int tid, float A8, float Al8)
{
    constexpr float a = 0.85000610f;   // mode->preemph[0], static_modes_float.h:583
    const int seg = N / kPostSegs;     // 8, 4, 2, 1
    const int lane = tid & 31, warp = tid >> 5;
    const bool act = tid < kPostSegs;
    float A = A8;                      // a^seg (20 ms frames: precomputed by the caller)
    if (seg != 8) {
        A = a;
        for (int k = 1; k < seg; k <<= 1) A *= A;
    }
    float m[NCH];
    float v[NCH][8];                   // the thread's segment (20 ms frames: 8 samples per channel, in registers)
    const bool fast = seg == 8;
#pragma unroll
    for (int ch = 0; ch < NCH; ch++) {
        m[ch] = tid == 0 ? sm.mem[ch] : 0.f;   // thread 0 starts from the incoming state, the others from 0
        if (act) {
            const float *x = sm.buf[ch] + kOff + tid * seg;
            if (fast) {
                const float4 lo = *reinterpret_cast<const float4 *>(x), hi = *reinterpret_cast<const float4 *>(x + 4);
                v[ch][0] = lo.x; v[ch][1] = lo.y; v[ch][2] = lo.z; v[ch][3] = lo.w;
                v[ch][4] = hi.x; v[ch][5] = hi.y; v[ch][6] = hi.z; v[ch][7] = hi.w;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const float tmp = v[ch][j] + m[ch] + 1e-30f;   // VERY_SMALL, arch.h:195
                    m[ch] = a * tmp;
                    v[ch][j] = tmp;
                }
            } else {
                for (int j = 0; j < seg; j++) {
                    const float tmp = x[j] + m[ch] + 1e-30f;
                    m[ch] = a * tmp;
                    sm.stage[stage_at<NCH>(tid * seg + j) + ch] = tmp;
                }
            }
        }
    }
    // inclusive scan over the lanes of each warp (Kogge-Stone); Ad = A^d
    float Ad = A;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
        for (int ch = 0; ch < NCH; ch++) {
            const float mp = __shfl_up_sync(kFullMask, m[ch], d);
            if (lane >= d) m[ch] = fmaf(Ad, mp, m[ch]);
        }
        Ad *= Ad;
    }
    const float A32 = Ad;   // A^32: a whole warp of segments
    if (lane == 31)
#pragma unroll
        for (int ch = 0; ch < NCH; ch++) sm.wtot[ch][warp] = m[ch];
    __syncthreads();
    float Al = Al8;         // A^lane; precomputed by the caller for 20 ms frames
    if (!fast) {
        Al = 1.f;
        for (int k = 0; k < lane; k++) Al *= A;
    }
#pragma unroll
    for (int ch = 0; ch < NCH; ch++) {
        // state entering this warp's first segment: Horner over the carries of the warps before it
        const float w0 = sm.wtot[ch][0], w1 = sm.wtot[ch][1], w2 = sm.wtot[ch][2];
        const float c1 = w0, c2 = fmaf(A32, c1, w1), c3 = fmaf(A32, c2, w2);
        const float cw = warp == 0 ? 0.f : (warp == 1 ? c1 : (warp == 2 ? c2 : c3));
        const float prev = __shfl_up_sync(kFullMask, m[ch], 1);
        const float cin = lane == 0 ? cw : fmaf(Al, cw, prev);   // state entering this thread's segment
        if (tid == kPostSegs - 1) sm.mem[ch] = fmaf(Al * A, cw, m[ch]);
        if (act && fast) {
            if (tid > 0) {
                float pw = 1.f;   // a^j
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    v[ch][j] = fmaf(pw, cin, v[ch][j]);
                    pw *= a;
                }
            }
        } else if (act && tid > 0) {
            float pw = 1.f;
            for (int j = 0; j < seg; j++) {
                sm.stage[stage_at<NCH>(tid * seg + j) + ch] += pw * cin;
                pw *= a;
            }
        }
    }
    if (act && fast) {   // 8 samples x NCH channels, interleaved, as 128-bit words (conflict-free thanks to the padding)
        float4 *dst = reinterpret_cast<float4 *>(sm.stage + tid * (8 * NCH + kStagePad));
        if (NCH == 2) {
#pragma unroll
            for (int j = 0; j < 8; j += 2) dst[j >> 1] = make_float4(v[0][j], v[NCH - 1][j], v[0][j + 1], v[NCH - 1][j + 1]);
        } else {
            dst[0] = make_float4(v[0][0], v[0][1], v[0][2], v[0][3]);
            dst[1] = make_float4(v[0][4], v[0][5], v[0][6], v[0][7]);
        }
    }
    __syncthreads();
}

template <int NCH>
__device__ __forceinline__ void post_job(const PostParams &p, const PostJob &job, PostSmem &sm, int tid)
{
    const int C = p.C;
    // incoming state
#pragma unroll
    for (int ch = 0; ch < NCH; ch++) {
        const int row = job.state_row + ch;
        const float *h = (p.hist_in && !job.reset) ? p.hist_in + (size_t)row * kPostHist : nullptr;
        for (int i = tid; i < kPostHist; i += kPostThreads) sm.buf[ch][kOff - kPostHist + i] = h ? h[i] : 0.f;
        if (tid == 0) sm.mem[ch] = (p.mem_in && !job.reset) ? p.mem_in[row] : 0.f;
    }
    __syncthreads();
    long long s0 = job.sample0;
    const PostFrame *fr = p.frames + (size_t)job.frame0 * p.frame_stride + job.stream_col;
    // (0.85000610^8)^lane: how far the de-emphasis state entering this thread's warp decays before its
    // segment of a 20 ms frame (deemphasis_frame); a chain of up to 31 dependent multiplies, done once
    float Al8 = 1.f, A8 = 0.85000610f;
    A8 *= A8; A8 *= A8; A8 *= A8;
    for (int k = 0; k < (tid & 31); k++) Al8 *= A8;
    // Plain stereo, 20 ms frames (the common case): the NEXT frame's 7680 bytes are fetched into
    // registers before the current frame is filtered, so the HBM latency hides behind the
    // recurrences instead of adding to every frame.  480 float4 = 120 threads x 4.
    const bool stereo2 = NCH == 2 && C == 2;
    float4 nxt[4];
    bool have_nxt = false;
    PostFrame pf_next;
    if (job.nframes > 0) pf_next = *fr;
    if (stereo2 && job.nframes > 0 && pf_next.N == kFrame) {
        const float4 *g4 = reinterpret_cast<const float4 *>(p.pcm + s0 * 2);
        if (tid < 120)
#pragma unroll
            for (int k = 0; k < 4; k++) nxt[k] = __ldcs(g4 + tid + 120 * k);
        have_nxt = true;
    }
    for (int f = 0; f < job.nframes; f++, fr += p.frame_stride) {
        const PostFrame pf = pf_next;
        if (f + 1 < job.nframes) pf_next = fr[p.frame_stride];   // side info one frame ahead, like the samples
        const int N = pf.N;
        float *g = p.pcm + s0 * C + job.ch0;
        // frame -> buffer (raw synthesis output)
        if (NCH == 2 && have_nxt) {
            if (tid < 120)
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int i = 2 * (tid + 120 * k);
                    sm.buf[0][kOff + i] = nxt[k].x;
                    sm.buf[1 % NCH][kOff + i] = nxt[k].y;
                    sm.buf[0][kOff + i + 1] = nxt[k].z;
                    sm.buf[1 % NCH][kOff + i + 1] = nxt[k].w;
                }
            have_nxt = false;
        } else if (NCH == 2 && C == 2) {
            const float4 *g4 = reinterpret_cast<const float4 *>(g);
            for (int i = tid; i < N / 2; i += kPostThreads) {
                const float4 v = __ldcs(g4 + i);
                sm.buf[0][kOff + 2 * i] = v.x;
                sm.buf[1 % NCH][kOff + 2 * i] = v.y;
                sm.buf[0][kOff + 2 * i + 1] = v.z;
                sm.buf[1 % NCH][kOff + 2 * i + 1] = v.w;
            }
        } else if (NCH == 2 && ((C | job.ch0) & 1) == 0) {   // 8-byte aligned {ch0, ch0+1} pairs
#pragma unroll 4
            for (int i = tid; i < N; i += kPostThreads) {
                const float2 v = __ldcs(reinterpret_cast<const float2 *>(g + (size_t)i * C));
                sm.buf[0][kOff + i] = v.x;
                sm.buf[1 % NCH][kOff + i] = v.y;
            }
        } else {
#pragma unroll 4
            for (int i = tid; i < N; i += kPostThreads)
#pragma unroll
                for (int ch = 0; ch < NCH; ch++) sm.buf[ch][kOff + i] = __ldcs(g + (size_t)i * C + ch);
        }
        if (stereo2 && f + 1 < job.nframes && pf_next.N == kFrame) {
            const float4 *g4 = reinterpret_cast<const float4 *>(p.pcm + (s0 + N) * 2);
            if (tid < 120)
#pragma unroll
                for (int k = 0; k < 4; k++) nxt[k] = __ldcs(g4 + tid + 120 * k);
            have_nxt = true;
        }
        __syncthreads();
        // celt_decoder_clean.c:660-669: [0,120) fades old -> cur; [120,240) fades cur -> new; [240,N) new
        {
            const Taps told = make_taps(pf.pitch[0], pf.gain[0], pf.tapset[0]);
            const Taps tcur = make_taps(pf.pitch[1], pf.gain[1], pf.tapset[1]);
            comb_region<NCH>(sm, 0, kOverlap, told, tcur, true, tid);
            if (N > kOverlap) {
                const Taps tnew = make_taps(pf.pitch[2], pf.gain[2], pf.tapset[2]);
                const int mid = N < 2 * kOverlap ? N : 2 * kOverlap;
                comb_region<NCH>(sm, kOverlap, mid, tcur, tnew, true, tid);
                if (N > mid) comb_region<NCH>(sm, mid, N, tcur, tnew, false, tid);
            }
        }
        deemphasis_frame<NCH>(sm, N, tid, A8, Al8);
        // staging -> HBM, scaled to [-1, 1] (SCALEOUT, arch.h:202)
        constexpr float kScale = 1.f / 32768.f;
        if (NCH == 2 && C == 2) {
            const float4 *s4 = reinterpret_cast<const float4 *>(sm.stage);
            float4 *g4 = reinterpret_cast<float4 *>(g);
            for (int i = tid; i < N / 2; i += kPostThreads) {
                float4 v = s4[(i >> 2) * ((16 + kStagePad) / 4) + (i & 3)];   // samples 2i, 2i+1 of both channels
                v.x *= kScale; v.y *= kScale; v.z *= kScale; v.w *= kScale;
                __stcs(g4 + i, v);
            }
        } else if (NCH == 2 && ((C | job.ch0) & 1) == 0) {
            for (int i = tid; i < N; i += kPostThreads) {
                float2 v = *reinterpret_cast<const float2 *>(sm.stage + stage_at<NCH>(i));
                v.x *= kScale; v.y *= kScale;
                __stcs(reinterpret_cast<float2 *>(g + (size_t)i * C), v);
            }
        } else {
            for (int i = tid; i < N; i += kPostThreads)
#pragma unroll
                for (int ch = 0; ch < NCH; ch++) __stcs(g + (size_t)i * C + ch, sm.stage[stage_at<NCH>(i) + ch] * kScale);
        }
        // slide the history: the last 1026 filtered samples move to the front (celt_decoder_clean.c:622-626)
        {
            constexpr int kPer = (kPostHist + kPostThreads - 1) / kPostThreads;   // 9
            float keep[NCH][kPer];
#pragma unroll
            for (int ch = 0; ch < NCH; ch++)
#pragma unroll
                for (int k = 0; k < kPer; k++) {
                    const int i = tid + kPostThreads * k;
                    keep[ch][k] = i < kPostHist ? sm.buf[ch][kOff - kPostHist + N + i] : 0.f;
                }
            __syncthreads();
#pragma unroll
            for (int ch = 0; ch < NCH; ch++)
#pragma unroll
                for (int k = 0; k < kPer; k++) {
                    const int i = tid + kPostThreads * k;
                    if (i < kPostHist) sm.buf[ch][kOff - kPostHist + i] = keep[ch][k];
                }
        }
        __syncthreads();
        s0 += N;
    }
    // outgoing state: the last kPostHist filtered samples and the de-emphasis memory
#pragma unroll
    for (int ch = 0; ch < NCH; ch++) {
        const int row = job.state_row + ch;
        if (p.hist_out && job.write_state)
            for (int i = tid; i < kPostHist; i += kPostThreads) p.hist_out[(size_t)row * kPostHist + i] = sm.buf[ch][kOff - kPostHist + i];
        if (p.mem_out && job.write_state && tid == 0) p.mem_out[row] = sm.mem[ch];
    }
}

#ifndef NQ_POST_MIN_BLOCKS
#define NQ_POST_MIN_BLOCKS 5
#endif
__global__ void __launch_bounds__(kPostThreads, NQ_POST_MIN_BLOCKS) celt_post_kernel(const __grid_constant__ PostParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PostSmem &sm = *reinterpret_cast<PostSmem *>(smem_raw);
    const int tid = threadIdx.x;
    for (int i = tid; i < kOverlap; i += kPostThreads) {
        const float w = p.window[i];
        sm.win2[i] = w * w;
    }
    const PostJob job = p.jobs[blockIdx.x];
    if (job.nch == 2) post_job<2>(p, job, sm, tid);
    else post_job<1>(p, job, sm, tid);
}

cudaError_t prepare_post_kernel()
{
    return cudaFuncSetAttribute(celt_post_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)post_kernel_smem_bytes());
}

cudaError_t launch_post(const PostParams &p, int njobs, cudaStream_t stream)
{
    if (njobs <= 0) return cudaSuccess;
    celt_post_kernel<<<njobs, kPostThreads, post_kernel_smem_bytes(), stream>>>(p);
    return cudaGetLastError();
}

}  // namespace nq
