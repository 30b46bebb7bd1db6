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
//   * across streams (one WARP, in its own one-warp CTA, per stream of 1 or 2 channels: a batch
//     of many files / multistream sub-decoders keeps the GPU busy, one file keeps one warp busy);
//   * inside a frame: the comb filter's nearest tap is T-2 samples back, so blocks of T-2 >= 13
//     samples are independent (typical pitch periods give 100-1000 samples per step), and both
//     channels of a coupled stream share one instruction stream;
//   * the de-emphasis IIR is evaluated as 32 lane-private segments plus a warp-level scan of the
//     segment carries (affine maps m -> m_seg + a^len * m), not sample by sample.
// The last 1024+2 filtered samples per channel live in a shared-memory ring (the reference keeps
// them in decode_mem, celt_decoder_clean.c:92); a frame is read from HBM once (coalesced), filtered
// in the ring, de-emphasised into a staging buffer and written once (coalesced).  The kernel works
// in place on the interleaved [nsamples][C] buffer the synthesis kernel wrote.
#include "celt_synth_kernels.cuh"

namespace nq {

constexpr int kRing = 2048;                  // >= kPostHist + 960, power of two
constexpr int kRingMask = kRing - 1;
constexpr unsigned kFullMask = 0xffffffffu;

struct __align__(16) PostSmem {
    float ring[2][kRing];      // filtered samples (comb output), per channel
    float stage[2 * kFrame];   // de-emphasised frame [n][nch]
    float win2[kOverlap];      // window[i]^2, celt.c:147
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

// One region [a, b) of a frame, in place in the ring (base = ring index of frame sample 0).
//   xfade: celt.c:142-166, the filter fades from `t0` to `t1` with window^2 over the region
//   else : celt.c:176 / pitch_sse.h:104, constant filter `t1`
// Samples inside a block of min(T)-2 are independent of each other; blocks run in order.
template <int NCH>
__device__ __forceinline__ void comb_region(PostSmem &sm, int base, int a, int b, const Taps &t0, const Taps &t1,
                                            bool xfade, int lane)
{
    const bool use0 = xfade && t0.on, use1 = t1.on;
    if (!use0 && !use1) return;   // celt.c:126-132 and :167-173: the filter is the identity here
    int B = b - a;
    if (use0 && t0.T - 2 < B) B = t0.T - 2;
    if (use1 && t1.T - 2 < B) B = t1.T - 2;
    for (int i0 = a; i0 < b; i0 += B) {
        const int iend = i0 + B < b ? i0 + B : b;
        for (int i = i0 + lane; i < iend; i += 32) {
            float f = 1.f;
            if (xfade) f = sm.win2[i - a];
#pragma unroll
            for (int ch = 0; ch < NCH; ch++) {
                float *r = sm.ring[ch];
                const int p = base + i;
                float acc = r[p & kRingMask];
                if (use0) {
                    const int q = p - t0.T;
                    const float e = 1.f - f;
                    acc += (e * t0.g0) * r[q & kRingMask];
                    acc += (e * t0.g1) * (r[(q + 1) & kRingMask] + r[(q - 1) & kRingMask]);
                    acc += (e * t0.g2) * (r[(q + 2) & kRingMask] + r[(q - 2) & kRingMask]);
                }
                if (use1) {
                    const int q = p - t1.T;
                    if (xfade) {
                        acc += (f * t1.g0) * r[q & kRingMask];
                        acc += (f * t1.g1) * (r[(q + 1) & kRingMask] + r[(q - 1) & kRingMask]);
                        acc += (f * t1.g2) * (r[(q + 2) & kRingMask] + r[(q - 2) & kRingMask]);
                    } else {   // partial sums as pitch_sse.h:136-139
                        acc += t1.g0 * r[q & kRingMask];
                        acc += t1.g1 * (r[(q + 1) & kRingMask] + r[(q - 1) & kRingMask]) +
                               t1.g2 * (r[(q + 2) & kRingMask] + r[(q - 2) & kRingMask]);
                    }
                }
                r[p & kRingMask] = acc;
            }
        }
        __syncwarp();
    }
}

// deemphasis, celt_decoder_clean.c:232-241, over the N filtered samples at ring[base ..):
// lane l owns samples [l*seg, (l+1)*seg); the carries are combined with a warp scan.
template <int NCH>
__device__ __forceinline__ void deemphasis_frame(PostSmem &sm, int base, int N, float (&mem)[NCH], int lane)
{
    constexpr float a = 0.85000610f;   // mode->preemph[0], static_modes_float.h:583
    const int nl = (N & 31) == 0 ? 32 : 30;   // N = 120 << LM: 960, 480 -> 32 lanes; 240, 120 -> 30 lanes
    const int seg = N / nl;                   // 30, 15, 8, 4
    const bool act = lane < nl;
    float A = 1.f;                            // a^seg
    for (int j = 0; j < seg; j++) A *= a;
#pragma unroll
    for (int ch = 0; ch < NCH; ch++) {
        const float *r = sm.ring[ch];
        const int p0 = base + lane * seg;
        // local pass from m = 0 (lane 0: from the incoming state), result kept in the staging buffer
        float m = lane == 0 ? mem[ch] : 0.f;
        if (act) {
            for (int j = 0; j < seg; j++) {
                const float tmp = r[(p0 + j) & kRingMask] + m + 1e-30f;   // VERY_SMALL, arch.h:195
                m = a * tmp;
                sm.stage[(lane * seg + j) * NCH + ch] = tmp;
            }
        } else {
            m = 0.f;
        }
        // inclusive scan of the affine maps c -> m + A*c over the lanes (Kogge-Stone)
        float Ad = A;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const float mp = __shfl_up_sync(kFullMask, m, d);
            if (lane >= d) m = fmaf(Ad, mp, m);
            Ad *= Ad;
        }
        // m = state after this lane's segment; the carry INTO the segment is the previous lane's
        const float cin = __shfl_up_sync(kFullMask, m, 1);
        mem[ch] = __shfl_sync(kFullMask, m, nl - 1);
        if (act && lane > 0) {
            float pw = 1.f;   // a^j
            for (int j = 0; j < seg; j++) {
                sm.stage[(lane * seg + j) * NCH + ch] += pw * cin;
                pw *= a;
            }
        }
    }
    __syncwarp();
}

template <int NCH>
__device__ __forceinline__ void post_job(const PostParams &p, const PostJob &job, PostSmem &sm, int lane)
{
    const int C = p.C;
    // incoming state
    float mem[NCH];
#pragma unroll
    for (int ch = 0; ch < NCH; ch++) {
        const int row = job.state_row + ch;
        const float *h = (p.hist_in && !job.reset) ? p.hist_in + (size_t)row * kPostHist : nullptr;
        for (int i = lane; i < kPostHist; i += 32) sm.ring[ch][(kRing - kPostHist + i) & kRingMask] = h ? h[i] : 0.f;
        mem[ch] = (p.mem_in && !job.reset) ? p.mem_in[row] : 0.f;
    }
    __syncwarp();
    int base = 0;   // ring index of the current frame's sample 0 (history sits just below it)
    long long s0 = job.sample0;
    const PostFrame *fr = p.frames + (size_t)job.frame0 * p.frame_stride + job.stream_col;
    // Plain stereo, 20 ms frames (the common case): the NEXT frame's 7680 bytes are fetched into
    // registers before the current frame is filtered, so the HBM latency hides behind the
    // recurrences instead of adding to every frame.
    constexpr bool kCanPrefetch = NCH == 2;
    const bool stereo2 = NCH == 2 && C == 2;
    float4 nxt[15];
    bool have_nxt = false;
    if (kCanPrefetch && stereo2 && job.nframes > 0 && fr->N == kFrame) {
        const float4 *g4 = reinterpret_cast<const float4 *>(p.pcm + s0 * 2);
#pragma unroll
        for (int k = 0; k < 15; k++) nxt[k] = __ldcs(g4 + lane + 32 * k);
        have_nxt = true;
    }
    PostFrame pf_next;
    if (job.nframes > 0) pf_next = *fr;
    for (int f = 0; f < job.nframes; f++, fr += p.frame_stride) {
        const PostFrame pf = pf_next;
        if (f + 1 < job.nframes) pf_next = fr[p.frame_stride];   // side info one frame ahead, like the samples
        const int N = pf.N;
        float *g = p.pcm + s0 * C + job.ch0;
        // frame -> ring (raw synthesis output)
        if (kCanPrefetch && have_nxt) {
#pragma unroll
            for (int k = 0; k < 15; k++) {
                const int i = lane + 32 * k;
                sm.ring[0][(base + 2 * i) & kRingMask] = nxt[k].x;
                sm.ring[1][(base + 2 * i) & kRingMask] = nxt[k].y;
                sm.ring[0][(base + 2 * i + 1) & kRingMask] = nxt[k].z;
                sm.ring[1][(base + 2 * i + 1) & kRingMask] = nxt[k].w;
            }
            have_nxt = false;
        } else if (NCH == 2 && C == 2) {
            const float4 *g4 = reinterpret_cast<const float4 *>(g);
#pragma unroll 5
            for (int i = lane; i < N / 2; i += 32) {
                const float4 v = __ldcs(g4 + i);
                sm.ring[0][(base + 2 * i) & kRingMask] = v.x;
                sm.ring[1][(base + 2 * i) & kRingMask] = v.y;
                sm.ring[0][(base + 2 * i + 1) & kRingMask] = v.z;
                sm.ring[1][(base + 2 * i + 1) & kRingMask] = v.w;
            }
        } else if (NCH == 2 && ((C | job.ch0) & 1) == 0) {   // 8-byte aligned {ch0, ch0+1} pairs
#pragma unroll 6
            for (int i = lane; i < N; i += 32) {
                const float2 v = __ldcs(reinterpret_cast<const float2 *>(g + (size_t)i * C));
                sm.ring[0][(base + i) & kRingMask] = v.x;
                sm.ring[1][(base + i) & kRingMask] = v.y;
            }
        } else {
#pragma unroll 6
            for (int i = lane; i < N; i += 32)
#pragma unroll
                for (int ch = 0; ch < NCH; ch++) sm.ring[ch][(base + i) & kRingMask] = __ldcs(g + (size_t)i * C + ch);
        }
        if (kCanPrefetch && stereo2 && f + 1 < job.nframes && pf_next.N == kFrame) {
            const float4 *g4 = reinterpret_cast<const float4 *>(p.pcm + (s0 + N) * 2);
#pragma unroll
            for (int k = 0; k < 15; k++) nxt[k] = __ldcs(g4 + lane + 32 * k);
            have_nxt = true;
        }
        __syncwarp();
        // celt_decoder_clean.c:660-669: [0,120) fades old -> cur; [120,240) fades cur -> new; [240,N) new
        {
            const Taps told = make_taps(pf.pitch[0], pf.gain[0], pf.tapset[0]);
            const Taps tcur = make_taps(pf.pitch[1], pf.gain[1], pf.tapset[1]);
            comb_region<NCH>(sm, base, 0, kOverlap, told, tcur, true, lane);
            if (N > kOverlap) {
                const Taps tnew = make_taps(pf.pitch[2], pf.gain[2], pf.tapset[2]);
                const int mid = N < 2 * kOverlap ? N : 2 * kOverlap;
                // celt.c:126: comb_filter as a whole is the identity when both gains are zero
                comb_region<NCH>(sm, base, kOverlap, mid, tcur, tnew, true, lane);
                if (N > mid) comb_region<NCH>(sm, base, mid, N, tcur, tnew, false, lane);
            }
        }
        deemphasis_frame<NCH>(sm, base, N, mem, lane);
        // staging -> HBM, scaled to [-1, 1] (SCALEOUT, arch.h:202)
        constexpr float kScale = 1.f / 32768.f;
        if (NCH == 2 && C == 2) {
            const float4 *s4 = reinterpret_cast<const float4 *>(sm.stage);
            float4 *g4 = reinterpret_cast<float4 *>(g);
            for (int i = lane; i < N / 2; i += 32) {
                float4 v = s4[i];
                v.x *= kScale; v.y *= kScale; v.z *= kScale; v.w *= kScale;
                __stcs(g4 + i, v);
            }
        } else if (NCH == 2 && ((C | job.ch0) & 1) == 0) {   // 8-byte aligned {ch0, ch0+1} pairs
            const float2 *s2 = reinterpret_cast<const float2 *>(sm.stage);
            for (int i = lane; i < N; i += 32) {
                float2 v = s2[i];
                v.x *= kScale; v.y *= kScale;
                __stcs(reinterpret_cast<float2 *>(g + (size_t)i * C), v);
            }
        } else {
            for (int i = lane; i < N; i += 32)
#pragma unroll
                for (int ch = 0; ch < NCH; ch++) __stcs(g + (size_t)i * C + ch, sm.stage[i * NCH + ch] * kScale);
        }
        __syncwarp();
        base = (base + N) & kRingMask;
        s0 += N;
    }
    // outgoing state: the last kPostHist filtered samples and the de-emphasis memory
#pragma unroll
    for (int ch = 0; ch < NCH; ch++) {
        const int row = job.state_row + ch;
        if (p.hist_out)
            for (int i = lane; i < kPostHist; i += 32)
                p.hist_out[(size_t)row * kPostHist + i] = sm.ring[ch][(base - kPostHist + i) & kRingMask];
        if (p.mem_out && lane == 0) p.mem_out[row] = mem[ch];
    }
}

__global__ void __launch_bounds__(32) celt_post_kernel(const __grid_constant__ PostParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PostSmem &sm = *reinterpret_cast<PostSmem *>(smem_raw);
    const int lane = threadIdx.x;
    for (int i = lane; i < kOverlap; i += 32) {
        const float w = p.window[i];
        sm.win2[i] = w * w;
    }
    __syncwarp();
    const PostJob job = p.jobs[blockIdx.x];
    if (job.nch == 2) post_job<2>(p, job, sm, lane);
    else post_job<1>(p, job, sm, lane);
}

cudaError_t prepare_post_kernel()
{
    return cudaFuncSetAttribute(celt_post_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)post_kernel_smem_bytes());
}

cudaError_t launch_post(const PostParams &p, int njobs, cudaStream_t stream)
{
    if (njobs <= 0) return cudaSuccess;
    celt_post_kernel<<<njobs, 32, post_kernel_smem_bytes(), stream>>>(p);
    return cudaGetLastError();
}

}  // namespace nq
