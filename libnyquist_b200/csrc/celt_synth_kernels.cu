// CELT synthesis (inverse MDCT + TDAC windowed overlap-add) for sm_100a.
//
// What the reference does per frame (third_party/opus/celt/):
//   compute_inv_mdcts     celt_decoder_clean.c:264-312  long: 1 x N2=960, transient: 8 x N2=120 at stride 8
//   clt_mdct_backward     mdct.c:267-379   pre-rotate :295-313, opus_ifft :316, post-rotate :320-359, mirror :361-377
//   opus_ifft             kiss_fft.c:696-747
//   tail hand-over        celt_decoder_clean.c:622-626 (the 60-sample raw tail lives in decode_mem)
//
// B200 design: the path is HBM-bound (15 360 B per stereo
// frame, ~4 flop/B), so the kernel is organised around moving every sample
// across HBM exactly once:
//   * one WARP owns a run of consecutive frames of one channel pair and keeps
//     the 60-sample raw tail on chip between frames (a run that does not
//     start the batch re-computes the frame before it instead of reading a
//     tail from memory); runs are short (64 frames) and, beyond the first
//     wave, claimed with an atomic so the SMs finish together;
//   * the frame's two 3840-byte coefficient rows arrive by TMA bulk copy
//     (cp.async.bulk + mbarrier) into a warp-private shared-memory buffer,
//     prefetched one frame ahead while the current frame is being computed;
//   * the 480-point inverse FFT is 30 x 16, the 60-point one 30 x 2: stage 1
//     = 32 lanes x (30-point prime-factor DFT in registers) is ONE code path
//     for long and transient frames (lane-dependent coefficient offsets and
//     twiddle rows), so the hot code fits the instruction cache whatever
//     the mix; one padded, conflict-free shared-memory transpose; long stage
//     2 = 30 lanes x 2 channels x (16-point DFT in registers).  The MDCT
//     pre-/post-rotations (incl. the reference's sin(x)~x factor (1+js)^2)
//     are folded into the single inter-stage twiddle table plus literal
//     per-slot rotations, so there is one table load per complex point;
//   * even/odd output pairing is done with one warp shuffle per point (bins
//     k and N4-1-k live in mirrored lanes);
//   * window, overlap-add and the stereo channel interleave are fused into
//     the epilogue: each lane stores float4 = {L[n], R[n], L[n+1], R[n+1]},
//     30 lanes writing 480 contiguous bytes per instruction;
//   * transient frames (8 short blocks): lanes = (channel, sub-block, half);
//     the radix-2 step reads the partner row of the transpose buffer, the
//     sub-block to sub-block tail is a warp shuffle, the loop over the 30
//     bins is rolled; output is parked in the consumed coefficient rows so
//     global stores stay 128-bit and coalesced;
//   * a mono stream pairs two consecutive frames as the warp's two channels;
//   * channel layouts other than plain stereo / mono (3, 8 channels, Opus
//     multistream with a channel mapping) are warp-specialised: a GROUP of
//     synthesis warps, one per stream, works on the same run; each leaves its
//     frame as a [960][2] plane in its (idle) transpose buffer and arrives on
//     the group's `full` mbarrier; a STORE warp waits for it, writes the
//     interleaved [960][C] frame with contiguous float4 stores -- the
//     multistream channel mapping is a gather in this pass -- and arrives on
//     `empty`, which a synthesis warp only waits for when it is about to
//     overwrite its plane, well into the next frame.
// No tensor cores: this is an FFT, not a dense contraction.
#include <cstdlib>
#include "celt_synth_kernels.cuh"
#include "celt_fft_codelets.cuh"

namespace nq {

// ------------------------------------------------------------------ PTX ----
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_test_wait(unsigned long long *bar, uint32_t parity)   // non-blocking
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {}
}
// For waits that are expected to last microseconds (a store warp waiting for the next frame): sleep
// between tries, so that the waiting warp leaves the issue slots to the others.
__device__ __forceinline__ void mbar_wait_long(unsigned long long *bar, uint32_t parity, unsigned ns)
{
    while (!mbar_test_wait(bar, parity)) __nanosleep(ns);
}
// TMA 1-D bulk copy global -> shared, completion counted in bytes on `bar`.
__device__ __forceinline__ void tma_load_row(void *dst_smem, const void *src_gmem, uint32_t bytes, unsigned long long *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

constexpr unsigned kFull = 0xffffffffu;

struct __align__(16) WarpSmem {
    float in[2 * kInRowFloats];       // two coefficient rows (TMA destination)
    float2 x[2 * kXChanF2];           // inter-stage transpose buffer / short-frame output staging
    float tail[2 * kHalfOvl];         // raw tail y[N2-60..N2) of the previous (sub-)block, per channel
    unsigned long long bar;           // mbarrier for the TMA prefetch
    // Group mode, used in the WarpSmem of the group's first warp only: the hand-over between the
    // group's synthesis warps and its store warp(s), and the ring through which the group's first
    // warp publishes the runs it claims.
    unsigned long long full;          // mbarrier: every synthesis warp's plane of the frame is complete
    unsigned long long empty;         // (low 32 bits) a COUNTER: frames x store warps done reading the planes.  Not an mbarrier: the
                                      // warp of a lone mono stream hands over two frames at a time, and a parity wait cannot skip a phase
    unsigned long long pad_;
    unsigned long long claim[4];      // (run << 8 | sequence number), see celt_synth_kernel
};
static_assert(sizeof(WarpSmem) % 16 == 0, "warp smem slice must keep 16-byte alignment");
static_assert((kInRowFloats * 4) % 16 == 0, "TMA destination rows must be 16-byte aligned");

constexpr int kPlaneOffFloats = 2 * kInRowFloats;   // offsetof(WarpSmem, x) / 4

// Short-block frames park their finished samples in the channel's own (consumed) coefficient
// row.  Sample n of channel c sits at ws.in[park_index(c, n)]: rows of 120 samples per sub-block,
// the second half of the frame shifted by 4 floats and channel 1 by 2 -- the 32 lanes (channel,
// sub-block, half) of a parking store then hit 32 different banks (plain rows: 120 b = 24 b mod 32
// puts sub-blocks b and b + 4, and both channels, on the same banks: 4-way conflicts, 240 of the
// short path's 900 shared-memory wavefronts per frame).  The 8 floats of padding per row hold it.
// kModeMono keeps channel 1 (the next frame of the stream) 16-byte aligned instead: its frames
// leave as float4 copies of the rows, which is worth more there than the last factor of two.
template <int kMode>
__device__ __forceinline__ int park_index(int c, int n)
{
    return c * (kInRowFloats + (kMode == kModeMono ? 0 : 2)) + n + (n >= kFrame / 2 ? 4 : 0);
}


// This file is compiled four times (-DNQ_PART=0..3, libnyquist_b200/build.py) so the kernel
// instantiations build in parallel: part 0 = stereo + mono variants, the generic kernel and the
// host-side dispatch; part 1 = group variants; part 2 = group variants with paired mono streams;
// part 3 = direct variants.
#ifndef NQ_PART
#define NQ_PART 0
#endif

#if NQ_PART == 0
size_t fast_kernel_smem_bytes() { return sizeof(FastTables) + kWarpsPerCta * sizeof(WarpSmem); }
#endif

// Group mode is warp-specialised: the group's SYNTHESIS warps (one per stream) each leave their
// frame as a [960][2] plane in shared memory and arrive on the group's `full` mbarrier; the group's
// STORE warp(s) wait for it, write the interleaved [960][C] output frame and arrive on `empty`,
// which a synthesis warp waits for only when it is about to overwrite its plane (after stage 1's
// loads and 30-point transforms of the next frame).  The synthesis warps therefore never wait for
// each other, carry none of the store pass's state or code, and run one frame ahead of the store.
//
// Store pass: the output frame is 240*C float4; thread tg of the first T2 store threads owns float4
// q = tg + T2*i.  T2 is chosen on the host so that 4*T2 is a multiple of C: the four output channels
// a thread serves are then the same in every iteration, and only the sample index advances (by
// 4*T2/C).
struct StoreCtx {
    uint32_t src[4];       // shared-memory byte address of this thread's four elements at iteration 0
    uint32_t step;         // byte advance per iteration: (4*T2/C) samples * 8 bytes (planes are [960][2])
    int q0, T2, niter;     // first float4, stride, iterations for a 20 ms frame (0 for the threads beyond T2)
    int T;                 // store threads of the group
    uint32_t planes;       // shared-memory byte address of the group's first plane
    unsigned mute;         // bit j: element j belongs to a silent output channel
    unsigned lone;         // bit j: element j comes from the group's lone mono stream (its plane holds frame pairs)
    int shape;             // store-loop shape, see group_store_frame (uniform over the group)
};

// A synthesis warp's view of its group.
struct GroupLink {
    unsigned long long *full;
    volatile unsigned *stored;   // the group's counter of finished store passes (frames x store warps)
    uint32_t nstored;      // frames handed to the store warps so far
    uint32_t store_warps;  // store warps of the group: each counts a frame once
    bool pending;          // the store pass of the last one may still be reading this warp's plane
};

// Called before anything is written to ws.x (the plane of the previous stored frame(s)): every frame
// this warp has handed over must have been stored.
__device__ __forceinline__ void group_wait_plane_free(GroupLink &g)
{
    if (g.pending) {
        const unsigned need = g.nstored * g.store_warps;
#ifndef NQ_PLANE_WAIT_NS
#define NQ_PLANE_WAIT_NS 50
#endif
        while ((int)(*g.stored - need) < 0) {
            if (NQ_PLANE_WAIT_NS > 0) __nanosleep(NQ_PLANE_WAIT_NS);   // leave the issue slots to the warps that have work
        }
        g.pending = false;
    }
}

__device__ __forceinline__ float lds_f32(uint32_t addr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ float2 lds_f32x2(uint32_t addr)
{
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}

// Interleaved output frame [960][C] from the group's planes: output channel c is decoded
// channel mapping[c] (a gather, opus_multistream_decoder.c:260-299) or silence.  Three loop
// shapes, chosen once per kernel: both element pairs of a thread are the two channels of one
// coupled stream in order (plain layouts: 2 x LDS.64), the general gather (4 x LDS.32), and the
// general gather with silent channels.
// General form (shape 3), for channel counts no store-thread count divides evenly (4*T2 % C != 0 for
// every usable T2: many duplicated output channels on few streams): every float4 of the output frame
// is gathered element by element, channel and sample re-derived per element.  Slow, correct, rare.
static __device__ __noinline__ void group_store_frame_general(const SynthParams &p, uint32_t planes, int tg, int T, float *frame_out, int nsamples,
                                                              int lone_slot, int col)
{
    const int C = p.C, nq4 = (nsamples / 4) * C;
    float4 *dst = reinterpret_cast<float4 *>(frame_out);
    for (int q = tg; q < nq4; q += T) {
        int n = (4 * q) / C, c = (4 * q) % C;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const unsigned sc = p.chan_src[c];
            const unsigned sub = (int)(sc >> 1) == lone_slot ? (unsigned)col : (sc & 1);   // a lone mono stream's plane: column = frame of the pair
            v[j] = sc == 0xffffu ? 0.f : lds_f32(planes + (sc >> 1) * (uint32_t)sizeof(WarpSmem) + sub * 4 + n * 8);
            if (++c == C) { c = 0; n++; }
        }
        __stcs(dst + q, make_float4(v[0], v[1], v[2], v[3]));
    }
}

// col: which column of a lone mono stream's plane this frame is (its warp synthesises frame pairs)
__device__ __forceinline__ void group_store_frame(const StoreCtx &g, uint32_t base, float *frame_out, int niter, int col)
{
    float4 *dst = reinterpret_cast<float4 *>(frame_out) + g.q0;
    const unsigned lm = col ? g.lone : 0u;
    uint32_t a0 = base + g.src[0] + ((lm & 1) ? 4u : 0u), a1 = base + g.src[1] + ((lm & 2) ? 4u : 0u),
             a2 = base + g.src[2] + ((lm & 4) ? 4u : 0u), a3 = base + g.src[3] + ((lm & 8) ? 4u : 0u);
    const uint32_t step = g.step;
    const int T2 = g.T2;
    if (g.shape == 0) {
#pragma unroll 5
        for (int i = 0; i < niter; i++) {
            const float2 lo = lds_f32x2(a0), hi = lds_f32x2(a2);
            __stcs(dst, make_float4(lo.x, lo.y, hi.x, hi.y));
            dst += T2;
            a0 += step; a2 += step;
        }
    } else if (g.shape == 1) {
#pragma unroll 5
        for (int i = 0; i < niter; i++) {
            __stcs(dst, make_float4(lds_f32(a0), lds_f32(a1), lds_f32(a2), lds_f32(a3)));
            dst += T2;
            a0 += step; a1 += step; a2 += step; a3 += step;
        }
    } else {
        for (int i = 0; i < niter; i++) {
            float4 v = make_float4(lds_f32(a0), lds_f32(a1), lds_f32(a2), lds_f32(a3));
            if (g.mute & 1) v.x = 0.f;
            if (g.mute & 2) v.y = 0.f;
            if (g.mute & 4) v.z = 0.f;
            if (g.mute & 8) v.w = 0.f;
            __stcs(dst, v);
            dst += T2;
            a0 += step; a1 += step; a2 += step; a3 += step;
        }
    }
}

// ------------------------------------------------- prefetch of the next item ---
// TMA bulk copies of the next item's coefficient rows into ws.in (lane 0 issues, completion on
// the warp's mbarrier).  Only once every lane is done with ws.in.
// after_generic_writes: ws.in was WRITTEN with ordinary stores since the last copy landed (a short
// frame parks its samples there): a proxy fence orders those stores before the copy engine's writes
// (the pattern of a TMA store after shared-memory writes: fence in every writer, barrier, one issuer).
// row2: where the second row sits relative to the first, in floats: kFrame (the other channel of the
// frame) or D * kFrame (the same channel of the NEXT frame: a mono stream's frame pair).
__device__ __forceinline__ void prefetch_rows(const SynthParams &p, WarpSmem &ws, int lane, long long fnext, int cb, int rows,
                                              bool after_generic_writes, long long row2 = kFrame)
{
    if (after_generic_writes) {   // every lane fences its own stores, then the warp meets, then lane 0 issues
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
    }
    if (lane == 0) {
        const float *src = (fnext < 0 ? p.halo_coef : p.coef + fnext * p.D * kFrame) + cb * kFrame;
        mbar_expect_tx(&ws.bar, rows * kFrame * 4);
        tma_load_row(ws.in, src, kFrame * 4, &ws.bar);
        if (rows == 2) tma_load_row(ws.in + kInRowFloats, src + row2, kFrame * 4, &ws.bar);
    }
}

// ---------------------------------------------------------------- stage 1 ---
// Common to long and transient frames -- ONE copy of the 30-point transform in the instruction
// cache, which the two paths do not fit together otherwise (20 % transient frames: 37 % i-cache
// misses with separate copies).  Lane = (channel c, r):
//   long block : r = residue n2, bins i = 16 n1 + n2;  x[2i] = X[32 n1 + 2 n2], x[959-2i] = X[959 - 2 n2 - 32 n1]
//   short block: r = 2b + h (sub-block b, half h), bins i = 2 n1 + h of the 60-point transform of sub-block b;
//                x[2i] = X[b + 8 (4 n1 + 2h)], x[119-2i] = X[b + 8 (119 - 4 n1 - 2h)]   (celt_decoder_clean.c:296)
// i.e. both read xa = row[oa + 32 n1], xb = row[ob - 32 n1] with lane-dependent (oa, ob), rotate by the
// literal exp(j 2pi n1/120) [mdct.c:303-312 without the lane-constant part], run the 30-point DFT and
// multiply by a lane-dependent row of a twiddle table (rotations + inter-stage twiddle folded, DESIGN.md
// section 3); the products go to the transpose buffer ws.x[c][r][k1].
template <int kModeT>
__device__ __forceinline__ void stage1_common(const FastTables &tb, WarpSmem &ws, int lane, bool is_short, GroupLink &grp)
{
    constexpr int kMode = kModeT == kModeGroupPaired ? kModeGroup : kModeT;
    constexpr float pre_re[30] = {NQ_PRE30_RE};
    constexpr float pre_im[30] = {NQ_PRE30_IM};
    const int c = lane >> 4, r = lane & 15, b = r >> 1, h = r & 1;
    const int oa = is_short ? 16 * h + b : 2 * r;
    const int ob = is_short ? 952 + b - 16 * h : 959 - 2 * r;
    const float *row = ws.in + c * kInRowFloats;
    float2 g[30];
#pragma unroll
    for (int n1 = 0; n1 < 30; n1++) {
        const float xa = row[oa + 32 * n1], xb = row[ob - 32 * n1];
        g[n1] = make_float2(fmaf(xb, pre_re[n1], -(xa * pre_im[n1])), fmaf(xb, pre_im[n1], xa * pre_re[n1]));
    }
    idft30(g);
    if (kMode == kModeGroup) group_wait_plane_free(grp);   // the store pass of the previous frame reads ws.x
    const float2 *tw = is_short ? tb.t_short + h * 30 : tb.t_long + r * kXRowF2;
    float2 *dst = ws.x + c * kXChanF2 + r * kXRowF2;
#pragma unroll
    for (int k1 = 0; k1 < 30; k1++) dst[k1] = cmul(g[k1], tw[k1]);
    __syncwarp();   // ws.in is consumed, ws.x is complete
}

// ------------------------------------------------------ long block, stage 2 ---
// kModeMono: the two "channels" are two CONSECUTIVE FRAMES of one mono stream (rows f and f+1 of
// the coefficient array), so a mono stream fills the warp like a stereo one; channel 1 then
// overlap-adds against channel 0's fresh tail instead of a stored one.
// vmask (paired group mode): bit ch set = channel ch of this warp is a long block in this frame.
// Two mono streams that share a warp switch blocks independently; when they disagree the frame
// goes through both paths and each keeps only its own channel.
template <int kModeT>
// lone (group mode): this warp carries a LONE mono stream and treats it like kModeMono does -- `lone`
// consecutive frames (1 or 2) as its channels, tail in slot 0 -- while still computing two channels
// (the second one of a single frame is scratch); 0: an ordinary warp of the group.
__device__ __forceinline__ void long_stage2(const SynthParams &p, WarpSmem &ws, int lane, long long off, int cb, int nch,
                                            bool store, const float (&w4)[4], int vmask, int lone = 0)
{
    constexpr bool kPaired = kModeT == kModeGroupPaired;
    constexpr int kMode = kPaired ? kModeGroup : kModeT;
    constexpr bool kStereo = kMode == kModeStereo;
    if (!kPaired) vmask = 3;
    const bool mono = kMode == kModeMono || (kMode == kModeGroup && lone > 0);   // (compile time except in group mode)
    const int mframes = kMode == kModeMono ? nch : lone;                          // frames of a mono item
    constexpr float post_re[16] = {NQ_POST16_RE};
    constexpr float post_im[16] = {NQ_POST16_IM};

    // ---- stage 2: lane = k1 (30 active); bins k = k1 + 30*k2, both channels ----
    const bool active = lane < 30;
    const int k1 = active ? lane : 29;
    float E[2][16], O[2][16], H0[2], H1[2];
    float2 told[2];
#pragma unroll
    for (int ch = 0; ch < 2; ch++) {
        if (ch < nch) {
            told[ch] = *reinterpret_cast<const float2 *>(ws.tail + (mono ? 0 : ch * kHalfOvl) + 58 - 2 * k1);
            float2 z[16];
            const float2 *src = ws.x + ch * kXChanF2 + k1;
#pragma unroll
            for (int q = 0; q < 16; q++) z[q] = src[q * kXRowF2];
            idft16(z);
            float im[16];
#pragma unroll
            for (int k2 = 0; k2 < 16; k2++) {
                // Y = Z * exp(j 2pi k2/64); y[2k] = -Re Y, y[959-2k] = Im Y   [mdct.c:330-358]
                const float2 zz = z[slot16(k2)];
                E[ch][k2] = fmaf(zz.y, post_im[k2], -(zz.x * post_re[k2]));
                im[k2] = fmaf(zz.x, post_im[k2], zz.y * post_re[k2]);
            }
            // y[2k+1] = y[959-2k'] with k' = 479-k = (29-k1) + 30*(15-k2)
#pragma unroll
            for (int k2 = 0; k2 < 16; k2++) O[ch][k2] = __shfl_sync(kFull, im[15 - k2], (29 - lane) & 31);
        }
    }
    __syncwarp();   // all lanes are done with ws.x and with the old tail
    if (mono && mframes == 2) {
        // frame f+1 mirrors against frame f's raw tail: tail[58-2k1], tail[59-2k1] sit in lane 29-k1
        told[1].x = __shfl_sync(kFull, E[0][15], (29 - lane) & 31);
        told[1].y = __shfl_sync(kFull, O[0][15], (29 - lane) & 31);
    }
#pragma unroll
    for (int ch = 0; ch < 2; ch++) {
        if (ch < nch) {
            if (mono) {   // one stream: only the LAST frame's tail is kept, in slot 0
                if (active && ch == mframes - 1) *reinterpret_cast<float2 *>(ws.tail + 2 * k1) = make_float2(E[ch][15], O[ch][15]);
            } else if (active && (!kPaired || ((vmask >> ch) & 1)))   // raw tail for the next frame: y[900+2k1], y[901+2k1]
                *reinterpret_cast<float2 *>(ws.tail + ch * kHalfOvl + 2 * k1) = make_float2(E[ch][15], O[ch][15]);
            // TDAC mirror with the previous raw tail [mdct.c:361-377]; m = 2k1 and 2k1+1
            const float t0 = told[ch].y, t1 = told[ch].x;   // tail[59-2k1], tail[58-2k1]
            const float y0 = E[ch][0], y1 = O[ch][0];
            E[ch][0] = fmaf(w4[0], t0, w4[1] * y0);      // out[60+2k1] = w[59-2k1] t0 + w[60+2k1] y0
            O[ch][0] = fmaf(w4[2], t1, w4[3] * y1);      // out[61+2k1] = w[58-2k1] t1 + w[61+2k1] y1
            H0[ch] = fmaf(w4[3], t1, -(w4[2] * y1));     // out[58-2k1] = w[61+2k1] t1 - w[58-2k1] y1
            H1[ch] = fmaf(w4[1], t0, -(w4[0] * y0));     // out[59-2k1] = w[60+2k1] t0 - w[59-2k1] y0
        }
    }
    if (store && active) {
        if (kStereo) {
            float4 *dst = reinterpret_cast<float4 *>(p.pcm + off * 2);
            __stcs(dst + 29 - k1, make_float4(H0[0], H0[1], H1[0], H1[1]));
#pragma unroll
            for (int k2 = 0; k2 < 15; k2++)
                __stcs(dst + 30 + k1 + 30 * k2, make_float4(E[0][k2], E[1][k2], O[0][k2], O[1][k2]));
        } else if (kMode == kModeMono) {
            // frame f (+ frame f+1): each lane owns y[2k], y[2k+1] -> 8-byte stores, 240 contiguous bytes per instruction
            float *dst = p.pcm + off;
#pragma unroll
            for (int ch = 0; ch < 2; ch++) {
                if (ch < nch) {
                    __stcs(reinterpret_cast<float2 *>(dst + ch * kFrame + 58 - 2 * k1), make_float2(H0[ch], H1[ch]));
#pragma unroll
                    for (int k2 = 0; k2 < 15; k2++)
                        __stcs(reinterpret_cast<float2 *>(dst + ch * kFrame + 60 + 2 * (k1 + 30 * k2)), make_float2(E[ch][k2], O[ch][k2]));
                }
            }
        } else if (kMode == kModeGroup) {
            // the frame as a [960][2] plane in the (now idle) transpose buffer; column 1 of a mono stream is unused
            float4 *pl = reinterpret_cast<float4 *>(ws.x);
            if (nch == 2 && vmask == 3) {
                pl[29 - k1] = make_float4(H0[0], H0[1], H1[0], H1[1]);
#pragma unroll
                for (int k2 = 0; k2 < 15; k2++) pl[30 + k1 + 30 * k2] = make_float4(E[0][k2], E[1][k2], O[0][k2], O[1][k2]);
            } else if (nch == 2) {   // only one column is this pass's to write
                const bool second = vmask == 2;
                float *col = reinterpret_cast<float *>(ws.x) + (second ? 1 : 0);
                col[2 * (58 - 2 * k1)] = second ? H0[1] : H0[0];
                col[2 * (59 - 2 * k1)] = second ? H1[1] : H1[0];
#pragma unroll
                for (int k2 = 0; k2 < 15; k2++) {
                    const int n = 60 + 2 * (k1 + 30 * k2);
                    col[2 * n] = second ? E[1][k2] : E[0][k2];
                    col[2 * n + 2] = second ? O[1][k2] : O[0][k2];
                }
            } else {
                pl[29 - k1] = make_float4(H0[0], 0.f, H1[0], 0.f);
#pragma unroll
                for (int k2 = 0; k2 < 15; k2++) pl[30 + k1 + 30 * k2] = make_float4(E[0][k2], 0.f, O[0][k2], 0.f);
            }
        } else {
            // generic channel count: channel pair (cb, cb+1) of an interleaved [n][C] frame
            float *dst = p.pcm + off * p.C + cb;
            const int C = p.C;
            if (nch == 2 && (C & 1) == 0) {          // 8-byte aligned {ch, ch+1} pairs
                __stcs(reinterpret_cast<float2 *>(dst + (58 - 2 * k1) * C), make_float2(H0[0], H0[1]));
                __stcs(reinterpret_cast<float2 *>(dst + (59 - 2 * k1) * C), make_float2(H1[0], H1[1]));
#pragma unroll
                for (int k2 = 0; k2 < 15; k2++) {
                    const int n = 60 + 2 * (k1 + 30 * k2);
                    __stcs(reinterpret_cast<float2 *>(dst + n * C), make_float2(E[0][k2], E[1][k2]));
                    __stcs(reinterpret_cast<float2 *>(dst + (n + 1) * C), make_float2(O[0][k2], O[1][k2]));
                }
            } else if (C == 1) {                     // mono: {y[n], y[n+1]} is contiguous
                __stcs(reinterpret_cast<float2 *>(dst + 58 - 2 * k1), make_float2(H0[0], H1[0]));
#pragma unroll
                for (int k2 = 0; k2 < 15; k2++)
                    __stcs(reinterpret_cast<float2 *>(dst + 60 + 2 * (k1 + 30 * k2)), make_float2(E[0][k2], O[0][k2]));
            } else {
#pragma unroll
                for (int ch = 0; ch < 2; ch++) {
                    if (ch < nch) {
                        dst[(58 - 2 * k1) * C + ch] = H0[ch];
                        dst[(59 - 2 * k1) * C + ch] = H1[ch];
#pragma unroll
                        for (int k2 = 0; k2 < 15; k2++) {
                            const int n = 60 + 2 * (k1 + 30 * k2);
                            dst[n * C + ch] = E[ch][k2];
                            dst[(n + 1) * C + ch] = O[ch][k2];
                        }
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------- short blocks, stage 2 ---
// 8 short blocks per channel (N = 240, N2 = 120, N4 = 60).  Lane = (c, b, h).  The radix-2 step that
// completes the 60-point transform pairs the rows r and r^1 of the transpose buffer; window +
// overlap-add against the previous sub-block's raw tail (lane - 2, a shuffle) follow.  A ROLLED loop
// of 30 trips (the inputs sit in shared memory, not in registers): a few hundred bytes of code.
//
// Trip i handles bin k1 = i in the h = 0 lanes and k1 = 29 - i in the h = 1 lanes, so that for
// everybody the head sample is y[m] with m = 2i + h: every address of the trip (tail, window pair,
// output) is then a lane-constant base plus a multiple of i with the same sign in all lanes, and
// the compiler folds it into the instruction.  The transpose buffer is read "even row of the pair,
// then odd row" (first, second) instead of "own row, partner's row": the h = 0 lanes read column i
// while the h = 1 lanes read column 29 - i, and that way the two halves of the warp land on banks
// of opposite parity (own-row reads would collide two ways).
//   Y = Z * exp(j 2pi 30 h / 240) with bins k = k1 + 30 h; y[2k] = -Re Y, y[119-2k] = Im Y:
//   h = 0: Z = first + second, head = y[2k1]    = -Re Z,        tl = y[119-2k1] = Im Z
//   h = 1: Z = first - second, head = y[59-2k1] = Im(Z e^{jpi/4}), tl = y[60+2k1] = -Re(Z e^{jpi/4})
// both written head = c1 Zx + c2 Zy, tl = c3 Zx + c4 Zy with lane constants.
// Output: kModeStereo stores straight to global memory -- one shuffle with lane ^ 16 turns the
// two samples of a trip into one {L, R} pair per lane (c = 0 keeps sample 60 + m, c = 1 sample
// 59 - m), 16 contiguous bytes per lane pair (h = 0, 1) -- so the consumed coefficient rows are free
// for the next frame's TMA copy as early as in a long frame.  The other modes park the finished
// samples in the channel's own -- consumed -- coefficient row, ws.in[c][0..960), for short_output.
// The raw tail of sub-block 7 replaces the old one in ws.tail; the old entries a group of 15
// trips needs are read before the group's first write (one warp barrier per group; groups of 5:
// all-transient 0.825 of the HBM peak, 15: 0.835, all 30 unrolled: the code outgrows the
// instruction cache, 0.89 instead of 0.97 on the 2.8 % mix).
template <int kModeT>
__device__ __forceinline__ void short_stage2(const SynthParams &p, const FastTables &tb, WarpSmem &ws, int lane, long long off, int nch,
                                             bool store, int vmask, int lone = 0)
{
    constexpr bool kPaired = kModeT == kModeGroupPaired;
    constexpr int kMode = kPaired ? kModeGroup : kModeT;
    const bool mono = kMode == kModeMono || (kMode == kModeGroup && lone > 0);   // see long_stage2
    const int mframes = kMode == kModeMono ? nch : lone;
#ifndef NQ_SHORT_ST
#define NQ_SHORT_ST 2
#endif
    constexpr bool kDirect = kMode == kModeStereo && NQ_SHORT_ST != 2;   // samples leave from registers
    const int c = lane >> 4, r = lane & 15, b = r >> 1, h = r & 1;
    const bool mine = !kPaired || ((vmask >> c) & 1);   // bit c = this channel is transient in this frame
    // column of trip i: i (h = 0) or 29 - i (h = 1); rows: the even and the odd row of the pair.
    // Shared-memory byte addresses that advance by a lane-constant step per trip (one add each,
    // where indexing by the trip number costs the compiler four instructions per access).
    uint32_t xa = smem_u32(ws.x + c * kXChanF2 + (r & ~1) * kXRowF2 + (h ? 29 : 0));
    const uint32_t xstep = h ? (uint32_t)-8 : 8u;
    float *st_lo = ws.in + park_index<kMode>(c, 120 * b) + 59 - h, *st_hi = st_lo + 1 + 2 * h;   // trip i: st_lo[-2i], st_hi[2i]
    // kModeMono: c = 1 is the NEXT frame of the same stream; its block 0 follows block 7 of c = 0 (lane - 2)
    float *tail = ws.tail + (mono ? 0 : c * kHalfOvl) + 59 - h;   // trip i: tail[-2i]
    const bool tail_from_smem = b == 0 && (!mono || c == 0);
    const bool tail_writer = b == 7 && (mono ? c == mframes - 1 : mine);
    const float sg = h ? -1.0f : 1.0f;
    const float c1 = h ? NQ_SQRT1_2 : -1.0f, c2 = h ? NQ_SQRT1_2 : 0.0f, c3 = h ? -NQ_SQRT1_2 : 0.0f, c4 = h ? NQ_SQRT1_2 : 1.0f;
    const float2 *wp = tb.wpair + h * 30;
    // direct stores: c = 0 writes {L, R}[60 + m] (ascending), c = 1 writes {L, R}[59 - m] (descending)
    float2 *gdst = nullptr;
    int gstep = 0;
    if (kDirect) {
        gdst = reinterpret_cast<float2 *>(p.pcm + off * 2) + 120 * b + (c ? 59 - h : 60 + h);
        gstep = c ? -2 : 2;
    }
#ifndef NQ_SHORT_GROUP
#define NQ_SHORT_GROUP 15
#endif
    constexpr int kGroup = NQ_SHORT_GROUP;   // trips per warp barrier (divides 30)
#pragma unroll 1
    for (int i0 = 0; i0 < 30; i0 += kGroup) {
        float told[kGroup];
#pragma unroll
        for (int k = 0; k < kGroup; k++) told[k] = tail_from_smem ? tail[-2 * k] : 0.f;
        __syncwarp();   // the old tail entries of this group are in registers before sub-block 7 replaces them
#pragma unroll
        for (int k = 0; k < kGroup; k++) {
            const float2 fa = lds_f32x2(xa), fb = lds_f32x2(xa + kXRowF2 * 8);
            xa += xstep;
            const float zx = fmaf(sg, fb.x, fa.x), zy = fmaf(sg, fb.y, fa.y);
            const float head = fmaf(c2, zy, c1 * zx), tl = fmaf(c4, zy, c3 * zx);
            float tp = __shfl_up_sync(kFull, tl, 2);     // same (c, h), sub-block b-1
            if (tail_from_smem) tp = told[k];
            const float2 w = wp[k];                      // (window[59-m], window[60+m])
            const float lo = fmaf(w.y, tp, -(w.x * head));   // out[59-m]
            const float hi = fmaf(w.x, tp, w.y * head);      // out[60+m]
            if (kDirect) {
                const float other = __shfl_xor_sync(kFull, c ? hi : lo, 16);
#if NQ_SHORT_ST == 1
                if (store) gdst[gstep * k] = c ? make_float2(other, lo) : make_float2(hi, other);
#else
                if (store) __stcs(gdst + gstep * k, c ? make_float2(other, lo) : make_float2(hi, other));
#endif
            } else if (mine) {
                st_lo[-2 * k] = lo;
                st_hi[2 * k] = hi;
            }
            if (tail_writer) tail[-2 * k] = tl;          // y_7[60 + (59-m)]
        }
        wp += kGroup;
        tail -= 2 * kGroup;
        st_lo -= 2 * kGroup;
        st_hi += 2 * kGroup;
        if (kDirect) gdst += gstep * kGroup;
    }
    __syncwarp();
}

// Finished short-block frame: planar ws.in[c][0..960) -> where the mode wants it.
template <int kModeT>
__device__ __forceinline__ void short_output(const SynthParams &p, WarpSmem &ws, int lane, long long off, int cb, int nch, bool store)
{
    constexpr int kMode = kModeT == kModeGroupPaired ? kModeGroup : kModeT;
    if (kMode == kModeGroup) {   // the stream's [960][2] plane for the group's store pass
        float4 *pl = reinterpret_cast<float4 *>(ws.x);
#pragma unroll 5
        for (int j = 0; j < 15; j++) {
            const int n = 2 * (lane + 32 * j);   // (n even: the pair n, n + 1 never straddles the shifted half)
            const float2 l = *reinterpret_cast<const float2 *>(ws.in + park_index<kMode>(0, n));
            const float2 rr = *reinterpret_cast<const float2 *>(ws.in + park_index<kMode>(1, n));
            pl[n >> 1] = make_float4(l.x, rr.x, l.y, rr.y);
        }
    } else if (!store) {
    } else if (kMode == kModeStereo) {
        float4 *dst = reinterpret_cast<float4 *>(p.pcm + off * 2);
#pragma unroll 5
        for (int j = 0; j < 15; j++) {
            const int n = 2 * (lane + 32 * j);
            const float2 l = *reinterpret_cast<const float2 *>(ws.in + park_index<kMode>(0, n));
            const float2 rr = *reinterpret_cast<const float2 *>(ws.in + park_index<kMode>(1, n));
            __stcs(dst + (n >> 1), make_float4(l.x, rr.x, l.y, rr.y));
        }
    } else if (kMode == kModeMono) {
        for (int ch = 0; ch < nch; ch++) {
            float4 *dst = reinterpret_cast<float4 *>(p.pcm + off + ch * kFrame);
            for (int j = lane; j < kFrame / 4; j += 32)
                __stcs(dst + j, *reinterpret_cast<const float4 *>(ws.in + park_index<kMode>(ch, 4 * j)));
        }
    } else {   // kModeDirect
        float *dst = p.pcm + off * p.C + cb;
        for (int idx = lane; idx < 2 * kFrame; idx += 32) {
            const int n = idx >> 1, ch = idx & 1;
            if (ch < nch) dst[n * p.C + ch] = ws.in[park_index<kMode>(ch, n)];
        }
    }
    __syncwarp();   // ws.in may be refilled
}

// ------------------------------------------- frames shorter than 20 ms ----
// LM < 3 (2.5 / 5 / 10 ms frames: N2 = 120 / 240 / 480, or 1 / 2 / 4 short blocks), SURVEY.md
// section 8(f) row 4.  Rare in practice (the bundled files hold one such frame), so this is a
// compact, loop-based warp routine that follows the reference formulas literally
// (mdct.c:295-313, :320-359, :361-377) with the reference's trig table, the N4-point inverse DFT
// as 30 x R (30-point codelet, then an R-term direct sum) -- the same arithmetic as
// mdct_backward_generic_kernel below -- and leaves the frame as a [N][2] plane at the start of
// ws.x.  What matters is that such frames can sit anywhere inside a batch and hand their tail on.
static __device__ __noinline__ void small_frame_planes(const GenericTables *gt, const float *win, float *in_rows, float2 *xbuf,
                                                float *tails, int lane, int nch, int sh, int trmask)
{
    const int Nf = kFrame >> sh;
    float *plane = reinterpret_cast<float *>(xbuf);   // [Nf][2], at most 480 x 2 floats
    float2 *a_buf = xbuf + 480, *b_buf = a_buf + 240;
    float *y = reinterpret_cast<float *>(a_buf);      // y[N2] takes a_buf's place once stage 1 has read it
    const float *trig = gt->trig;
    for (int ch = 0; ch < nch; ch++) {
        const int is_tr = (trmask >> ch) & 1;         // bit ch: channel ch uses short blocks in this frame
        const int nb = is_tr ? (8 >> sh) : 1;
        const int N2 = is_tr ? 120 : Nf, N4 = N2 >> 1, R = N4 / 30;
        const int shift = is_tr ? 3 : sh;
        const float sine = (float)2 * 3.141592653f * (.125f) / (float)(kMdctN >> shift);   // mdct.c:292
        const float *row = in_rows + ch * kInRowFloats;
        float *tail = tails + ch * kHalfOvl;
        for (int b = 0; b < nb; b++) {
            for (int i = lane; i < N4; i += 32) {
                const float x1 = row[b + 2 * i * nb], x2 = row[b + (N2 - 1 - 2 * i) * nb];
                const float t0 = trig[i << shift], t1 = trig[(N4 - i) << shift];
                const float yr = -(x2 * t0) + x1 * t1, yi = -(x2 * t1) - x1 * t0;
                a_buf[i] = make_float2(yr - yi * sine, yi + yr * sine);
            }
            __syncwarp();
            if (lane < R) {
                float2 g[30];
#pragma unroll
                for (int n1 = 0; n1 < 30; n1++) g[n1] = a_buf[R * n1 + lane];
                idft30(g);
#pragma unroll
                for (int k1 = 0; k1 < 30; k1++) {
                    float sn, cs;
                    sincospif(2.0f * (float)(lane * k1) / (float)N4, &sn, &cs);
                    b_buf[lane * 30 + k1] = cmulc(g[k1], cs, sn);
                }
            }
            __syncwarp();
            for (int k = lane; k < N4; k += 32) {
                const int k1 = k % 30, k2 = k / 30;
                float2 acc = make_float2(0.f, 0.f);
                for (int n2 = 0; n2 < R; n2++) {
                    float sn, cs;
                    sincospif(2.0f * (float)((n2 * k2) % R) / (float)R, &sn, &cs);
                    acc = cadd(acc, cmulc(b_buf[n2 * 30 + k1], cs, sn));
                }
                const float t0 = trig[k << shift], t1 = trig[(N4 - k) << shift];
                const float yr = acc.x * t0 - acc.y * t1, yi = acc.y * t0 + acc.x * t1;
                y[2 * k] = -(yr - yi * sine);
                y[N2 - 1 - 2 * k] = yi + yr * sine;
            }
            __syncwarp();
            for (int n = lane; n < N2; n += 32) {
                float o;
                if (n < kHalfOvl) o = win[kOverlap - 1 - n] * tail[n] - win[n] * y[kHalfOvl - 1 - n];
                else if (n < kOverlap) o = win[kOverlap - 1 - n] * tail[kOverlap - 1 - n] + win[n] * y[n - kHalfOvl];
                else o = y[n - kHalfOvl];
                plane[(b * N2 + n) * 2 + ch] = o;
            }
            __syncwarp();
            for (int i = lane; i < kHalfOvl; i += 32) tail[i] = y[N2 - kHalfOvl + i];
            __syncwarp();
        }
    }
}

// ------------------------------------------------ group mode: store warps ---
// Next run of a group.  Static: item + stride.  Dynamic: the group's first synthesis warp claims the
// run after next with an atomic when it STARTS a run and publishes it, tagged with the run's
// sequence number, in a ring of four words in shared memory; everybody else in the group (the other
// synthesis warps, the store warps) polls that word when it FINISHES the run.  (The warps of a group
// are never more than two frames apart, so a ring of four cannot be overrun; a sequence tag rather
// than an mbarrier because the claimer may be a phase ahead of a reader.)
__device__ __forceinline__ long long group_next_item(const volatile unsigned long long *claim, unsigned seq)
{
    unsigned long long w;
    do w = claim[seq & 3]; while ((unsigned)(w & 0xff) != (seq & 0xff));
    return (long long)(w >> 8);
}

// One store warp.  A group's frame is stored by SW store warps ("units" (group, part)); a CTA has
// NS store warps, and store warp s serves the units s, s + NS, ... (one unit each except in the
// layouts with many narrow groups): it polls their `full` barriers, writes the unit's part of the
// interleaved frame and releases the planes.  Every unit walks the same runs and frames as the
// synthesis warps of its group.
#ifndef NQ_LONE_PAIRS
#define NQ_LONE_PAIRS 1   // group mode: the warp of a lone mono stream takes two consecutive frames as its two channels
#endif
#ifndef NQ_STORE_SLEEP_NS
#define NQ_STORE_SLEEP_NS 256
#endif
constexpr unsigned kStoreSleepNs = NQ_STORE_SLEEP_NS;   // a store warp's nap between polls of its `full` barriers
constexpr int kMaxStoreUnits = 3;   // 12 groups of one synthesis warp over 4 store warps (more per store warp measured slower: 7 x 2 + 2 warps 0.79, 6 x 2 + 4 warps 0.90)

template <bool kAnySize>
__device__ __forceinline__ void group_store_role(const SynthParams &p, WarpSmem *wsmem, int s, int lane)
{
    const int C = p.C, W = p.nstreams, G = p.groups_per_cta, SW = p.store_warps, NS = p.store_warps_cta;
    const int tg = (s % SW) * 32 + lane;   // (NS is a multiple of SW: every unit of this warp is the same part)
    StoreCtx st;
    st.T = SW * 32;
    st.T2 = p.store_threads > 0 ? p.store_threads : 1;   // store_threads == 0: general store loop (shape 3)
    st.q0 = tg;
    st.step = (uint32_t)(4 * st.T2 / C) * 8u;
    st.niter = tg < st.T2 ? ((kFrame / 4) * C - tg + st.T2 - 1) / st.T2 : 0;
    st.mute = 0;
    st.lone = 0;
    st.planes = kPlaneOffFloats * 4;   // relative to the group's first slice
    st.shape = p.store_shape;
    // the group's lone mono stream, if any: its warp synthesises frame PAIRS (celt_synth_kernel), so its
    // plane holds frame f in column 0 and frame f + 1 in column 1; `second` below follows the same
    // greedy pairing rule as the synthesis warp
    int lone_slot = -1, lone_col = 0;
    if (NQ_LONE_PAIRS && p.lone_pairs)
        for (int w = 0; w < W; w++)
            if (p.streams[w].nch == 1) {
                lone_slot = w;
                lone_col = p.streams[w].flag_col;
            }
    {
        int n = (4 * tg) / C, c = (4 * tg) % C;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const unsigned sc = p.chan_src[c];
            if (sc == 0xffffu) st.mute |= 1u << j;
            else if ((int)(sc >> 1) == lone_slot) st.lone |= 1u << j;
            st.src[j] = sc == 0xffffu ? st.planes : st.planes + (sc >> 1) * (uint32_t)sizeof(WarpSmem) + (sc & 1) * 4 + n * 8;
            if (++c == C) { c = 0; n++; }
        }
    }
    const bool dynamic = p.work_counter != nullptr;
    const long long item_stride = (long long)gridDim.x * G;
    long long item[kMaxStoreUnits] = {}, f[kMaxStoreUnits] = {}, f1[kMaxStoreUnits] = {};
    uint32_t nstored[kMaxStoreUnits] = {}, seq[kMaxStoreUnits] = {};
    bool second[kMaxStoreUnits] = {};   // the unit's next frame is the second one of a lone mono stream's pair
    WarpSmem *ugws[kMaxStoreUnits] = {};
    int nunits = 0, alive = 0;
    auto start_run = [&](int k) {
        second[k] = false;
        if (item[k] < p.nruns) {
            seq[k]++;
            run_range(p, item[k], &f[k], &f1[k]);
        } else {
            f[k] = f1[k] = 0;
            alive--;
        }
    };
    for (int u = s; u < G * SW && nunits < kMaxStoreUnits; u += NS, nunits++) {
        item[nunits] = (long long)blockIdx.x * G + u / SW;
        ugws[nunits] = wsmem + (u / SW) * W;
        alive++;
        start_run(nunits);
    }
    while (alive > 0) {
        bool any = false;
        for (int k = 0; k < nunits; k++) {
            if (f[k] >= f1[k]) continue;
            WarpSmem *gws = ugws[k];
            if (nunits == 1) mbar_wait_long(&gws->full, nstored[k] & 1, kStoreSleepNs);
            else if (!mbar_test_wait(&gws->full, nstored[k] & 1)) continue;
            any = true;
            const long long fr = f[k];
            int niter = st.niter, nsamples = kFrame;
            float *out = p.pcm + fr * kFrame * C;
            if (kAnySize) {   // (every stream of a frame carries the same size bits)
                const int sh = (p.transient[fr * p.flag_stride] >> 1) & 3;
                nsamples = kFrame >> sh;
                niter = tg < st.T2 ? ((nsamples / 4) * C - tg + st.T2 - 1) / st.T2 : 0;
                out = p.pcm + p.frame_offset[fr] * C;
            }
            int col = 0;
            if (lone_slot >= 0) {
                if (second[k]) {
                    col = 1;
                    second[k] = false;
                } else if (fr + 1 < f1[k]) {   // (mono_pair of celt_synth_kernel)
                    const int fl = p.transient[fr * p.flag_stride + lone_col];
                    second[k] = (fl >> 1) == 0 && p.transient[(fr + 1) * p.flag_stride + lone_col] == fl;
                }
            }
            const uint32_t base = smem_u32(gws);
            if (st.shape == 3) group_store_frame_general(p, base + st.planes, st.q0, st.T, out, nsamples, lone_slot, col);
            else group_store_frame(st, base, out, niter, col);
            __syncwarp();   // every lane's reads of the planes are done (their values have left in the stores)
            if (lane == 0) atomicAdd(reinterpret_cast<unsigned *>(&gws->empty), 1u);
            nstored[k]++;
            if (++f[k] == f1[k]) {
                item[k] = dynamic ? group_next_item(gws->claim, seq[k]) : item[k] + item_stride;
                start_run(k);
            }
        }
        if (!any) __nanosleep(kStoreSleepNs);   // several units, none ready: leave the issue slots to the synthesis warps
    }
}

// ----------------------------------------------------------- fast kernel ---
// kWarps: stereo / mono / direct: 14 warps per CTA (what shared memory allows).  Group: up to 16
// (G*W synthesis warps followed by the CTA's store warps), 128 registers each like the others.
// kAnySize: the batch may hold frames shorter than 20 ms (p.frame_offset != nullptr).  A separate
// instantiation, so that the common all-20-ms kernel keeps its register allocation.
template <int kModeT, int kWarps, bool kAnySize>
__global__ void __launch_bounds__(kWarps * 32, 1) celt_synth_kernel(const __grid_constant__ SynthParams p)
{
    // kModeGroupPaired = kModeGroup + warps that carry two mono streams with independent block
    // switching (the masked two-pass path); a separate instantiation keeps that code out of the
    // plain group kernel, which measurably pays for it otherwise.
    constexpr bool kPaired = kModeT == kModeGroupPaired;
    constexpr int kMode = kPaired ? kModeGroup : kModeT;
    constexpr bool kStereo = kMode == kModeStereo;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FastTables &tb = *reinterpret_cast<FastTables *>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpSmem *wsmem = reinterpret_cast<WarpSmem *>(smem_raw + sizeof(FastTables));

    // Roles.  Stereo / mono / direct: every warp synthesises.  Group: warps [0, G*W) synthesise
    // (group gi = warp / W, stream slot = warp % W), the next G*SW warps store.
    int gi = 0, slot = 0, store_warp = -1;
    const int W = kMode == kModeGroup ? p.nstreams : 1, G = kMode == kModeGroup ? p.groups_per_cta : 0;
    if (kMode == kModeGroup) {
        if (warp < G * W) {
            gi = warp / W;
            slot = warp - gi * W;
        } else {
            store_warp = warp - G * W;
        }
    }
    const bool synth_warp = store_warp < 0;
    WarpSmem &ws = wsmem[synth_warp ? warp : 0];   // (only synthesis warps have a slice of their own)
    WarpSmem *gws = wsmem + gi * W;                 // group mode: the group's first slice holds its barriers

    {
        const float *src = reinterpret_cast<const float *>(p.tables);
        float *dst = reinterpret_cast<float *>(&tb);
        for (int i = threadIdx.x; i < int(sizeof(FastTables) / 4); i += blockDim.x) dst[i] = src[i];
    }
    if (synth_warp) {
        for (int i = lane; i < 2 * kInRowFloats; i += 32) ws.in[i] = 0.f;
        if (lane == 0) {
            mbar_init(&ws.bar, 1);
            if (kMode == kModeGroup && slot == 0) {
                mbar_init(&ws.full, 32 * W);
                ws.empty = 0;
                for (int i = 0; i < 4; i++) ws.claim[i] = 0;
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    __syncthreads();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");

    if (kMode == kModeGroup && !synth_warp) {   // (no block-wide barrier below this point)
        group_store_role<kAnySize>(p, wsmem, store_warp, lane);
        return;
    }

    // lane-constant window taps of the long-block mirror: k1 = lane (< 30)
    float w4[4];
    {
        const int k1 = lane < 30 ? lane : 29;
        const float *win = p.gen->window;   // (once per kernel: global memory)
        w4[0] = win[59 - 2 * k1];
        w4[1] = win[60 + 2 * k1];
        w4[2] = win[58 - 2 * k1];
        w4[3] = win[61 + 2 * k1];
    }

    // Work items.  Stereo / direct: one warp per (run, channel pair).  Group: a group owns a run,
    // synthesis warp `slot` of the group owns stream `slot`.
    long long item, item_stride, nitems;
    GroupLink grp;
    grp.full = &gws->full;
    grp.stored = reinterpret_cast<volatile unsigned *>(&gws->empty);
    grp.store_warps = (uint32_t)p.store_warps;
    grp.nstored = 0;
    grp.pending = false;
    uint32_t seq = 0;               // group mode: runs started so far
    long long claimed = 0;          // group mode, first warp: the run after this one
    if (kMode == kModeGroup) {
        item = (long long)blockIdx.x * G + gi;
        item_stride = (long long)gridDim.x * G;
        nitems = p.nruns;
    } else {
        item = (long long)blockIdx.x * kWarps + warp;
        item_stride = (long long)gridDim.x * kWarps;
        nitems = p.nruns * p.npairs;
    }
    // Dynamic distribution (stereo / mono with a work counter): the first kWarps * gridDim items are
    // assigned statically, every later one is claimed with an atomic when a warp finishes its run --
    // SMs do not progress at the same pace (HBM channel contention), and short runs claimed on
    // demand keep the tail of the launch short.
    const bool dynamic = p.work_counter != nullptr;
    uint32_t phase = 0;
    for (; item < nitems;) {
        long long run;
        int cb, nch, rows, flag_col, halo_bit, flag_col1 = -1;
        bool lone = false;   // group mode: this warp carries a lone mono stream
        if (kMode == kModeGroup) {
            if (dynamic) {   // see group_next_item
                seq++;
                if (slot == 0) {
                    unsigned long long next = 0;
                    if (lane == 0) {
                        next = atomicAdd(p.work_counter, 1ULL) + (unsigned long long)item_stride;
                        if (next > (unsigned long long)nitems) next = (unsigned long long)nitems;
                        *reinterpret_cast<volatile unsigned long long *>(&gws->claim[seq & 3]) = next << 8 | (seq & 0xff);
                    }
                    claimed = (long long)__shfl_sync(kFull, next, 0);
                }
            }
            const StreamDesc sd = p.streams[slot];
            run = item;
            cb = sd.row;
            // A group's warp always computes two channels (a compile-time width keeps the stage
            // functions inside the register budget): a lone mono stream's second channel has no
            // coefficient row -- its buffer stays zero -- and no reader.
            rows = sd.nch;
            nch = 2;
            lone = NQ_LONE_PAIRS && p.lone_pairs && sd.nch == 1;   // ... unless two consecutive frames can share the warp (below)
            flag_col = sd.flag_col;
            halo_bit = sd.flag_col;
            if (kPaired && sd.flag_col1 != sd.flag_col) flag_col1 = sd.flag_col1;   // two mono streams in one warp
        } else {
            run = kStereo ? item : item / p.npairs;
            const int pair = kStereo ? 0 : int(item - run * p.npairs);
            cb = 2 * pair;
            rows = nch = kStereo ? 2 : (p.D - cb >= 2 ? 2 : 1);
            flag_col = p.flag_per_stream ? pair : 0;
            halo_bit = p.flag_per_stream ? (pair & 31) : 0;
        }
        long long f0, f1;
        run_range(p, run, &f0, &f1);
        // A run that does not open the batch (or a batch with a halo frame)
        // first re-computes the frame before it, only to obtain its raw tail.
        // ... unless the run opens on a frame that follows a decoder reset (flag bit 3: the first
        // frame of another file in a batch of many): its tail is zero by definition.
        bool opens_on_reset = (p.transient[f0 * p.flag_stride + flag_col] & kFlagReset) != 0;
        if (kPaired && flag_col1 >= 0)   // (two streams in this warp: both must have been reset)
            opens_on_reset = opens_on_reset && (p.transient[f0 * p.flag_stride + flag_col1] & kFlagReset) != 0;
        const bool warm = (f0 > 0 || p.halo_coef != nullptr) && !opens_on_reset;
        for (int i = lane; i < 2 * kHalfOvl; i += 32) {
            const int ch = i / kHalfOvl;
            float t = 0.f;
            if (!warm && p.tail_in != nullptr && ch < rows) t = p.tail_in[(cb + ch) * kHalfOvl + (i - ch * kHalfOvl)];
            ws.tail[i] = t;
        }
        __syncwarp();
        long long f = warm ? f0 - 1 : f0;
        const uint8_t *flags = p.transient + flag_col;
        // flag byte of a frame: bit 0 = transient (short blocks), bits 1-2 = 3 - LM (0: 20 ms frame)
        int flag = f < 0 ? (p.halo_lm_shift << 1) | ((p.halo_transient >> halo_bit) & 1) : flags[f * p.flag_stride];
        // transient bit of the second channel when it is a stream of its own
        int tr1 = -1;
        if (kPaired && flag_col1 >= 0)
            tr1 = f < 0 ? (p.halo_transient >> flag_col1) & 1 : p.transient[f * p.flag_stride + flag_col1] & (kFlagTransient | kFlagReset);
        // kModeMono, and the warp of a LONE mono stream in group mode (a layout like 3.0 or 5.0: its
        // second channel would be idle): an item is one frame, or two consecutive 20 ms frames of the
        // same block type as the warp's two channels (never the warm-up frame, whose output is not
        // stored); nfr = frames of the current item.  (Compile time except in group mode.)
        const bool mono_like = kMode == kModeMono || (kMode == kModeGroup && lone);
        const long long row2 = mono_like ? (long long)p.D * kFrame : kFrame;   // second row: next frame / other channel
        auto mono_pair = [&](long long g, int gflag) -> bool {
            return mono_like && g >= f0 && g + 1 < f1 && (gflag >> 1) == 0 && flags[(g + 1) * p.flag_stride] == gflag;
        };
        int nfr = (mono_like && mono_pair(f, flag)) ? 2 : 1;
        const int state_nch = rows;   // channels with a tail of their own (kModeMono: nch is reused as frames per item)
        prefetch_rows(p, ws, lane, f, cb, mono_like ? nfr : rows, true, row2);   // (the previous run may have ended on a short frame)
        while (f < f1) {
            if (kMode == kModeMono) nch = nfr;
            const int lone_frames = (kMode == kModeGroup && lone) ? nfr : 0;
            const bool more = f + nfr < f1;
            const int next_flag = more ? flags[(f + nfr) * p.flag_stride] : 0;
            const int next_nfr = (mono_like && more && mono_pair(f + nfr, next_flag)) ? 2 : 1;
            int next_tr1 = -1;
            if (kPaired && flag_col1 >= 0 && more) next_tr1 = p.transient[(f + 1) * p.flag_stride + flag_col1] & (kFlagTransient | kFlagReset);
            // OPUS_RESET_STATE before this frame (celt_decoder_clean.c:846-859); two mono streams that share a
            // warp are decoders of their own: tr1 carries the second one's reset bit next to its transient bit
            const bool reset0 = (flag & kFlagReset) != 0, reset1 = kPaired && tr1 >= 0 ? (tr1 & kFlagReset) != 0 : reset0;
            if (reset0 || reset1) {
                for (int i = lane; i < 2 * kHalfOvl; i += 32)
                    if (i < kHalfOvl ? reset0 : reset1) ws.tail[i] = 0.f;
                __syncwarp();
            }
            if (kPaired && tr1 >= 0) tr1 &= kFlagTransient;
            while (!mbar_try_wait(&ws.bar, phase)) {}
            phase ^= 1;
            const bool store = f >= f0;
            const long long off = (kAnySize && store) ? p.frame_offset[f] : f * kFrame;
            const int sh = kAnySize ? (flag >> 1) & 3 : 0;   // (bits above 2 are not part of the size)
            const int tr0 = flag & 1;
            const bool split = kPaired && tr1 >= 0 && tr1 != tr0;
            if (sh == 0) {
                const int next_nch = mono_like ? next_nfr : rows;
                // One pass -- or, for two mono streams in one warp that disagree about block switching,
                // two: the short pass goes first and parks its channel's samples in that channel's own
                // coefficient row; the long pass (which needs ws.x as its transpose buffer) leaves its
                // column in the plane; then the parked column joins it.  vmask bit ch = channel ch
                // belongs to the pass.  (A loop, so that each stage has ONE call site.)
                const int cs = tr0 ? 0 : 1;   // split: the transient channel
                for (int ps = 0; ps < (kPaired && split ? 2 : 1); ps++) {
                    const bool is_short = split ? ps == 0 : tr0 != 0;
                    const int vmask = !split ? 3 : (is_short ? 1 << cs : 2 >> cs);
                    stage1_common<kModeT>(tb, ws, lane, is_short, grp);
                    if (!is_short) {
                        // every lane has consumed its part of ws.in: the next item's rows can land while stage 2 runs
                        if (more && !split) prefetch_rows(p, ws, lane, f + nfr, cb, next_nch, false, row2);
                        long_stage2<kModeT>(p, ws, lane, off, cb, nch, store, w4, vmask, lone_frames);
                    } else if (kStereo && NQ_SHORT_ST != 2) {
                        // (samples leave from registers: ws.in is free as early as in a long frame)
                        if (more) prefetch_rows(p, ws, lane, f + nfr, cb, next_nch, false);
                        short_stage2<kModeT>(p, tb, ws, lane, off, nch, store, vmask);
                    } else {
                        short_stage2<kModeT>(p, tb, ws, lane, off, nch, store, vmask, lone_frames);
                        if (!split) {
                            short_output<kModeT>(p, ws, lane, off, cb, nch, store);
                            if (more) prefetch_rows(p, ws, lane, f + nfr, cb, next_nch, true, row2);
                        }
                    }
                }
                if (kPaired && split) {
                    __syncwarp();
                    float *col = reinterpret_cast<float *>(ws.x) + cs;
                    for (int n = lane; n < kFrame; n += 32) col[2 * n] = ws.in[park_index<kMode>(cs, n)];
                    __syncwarp();
                    if (more) prefetch_rows(p, ws, lane, f + nfr, cb, next_nch, true, row2);
                }
            } else {
                if (kMode == kModeGroup) group_wait_plane_free(grp);
                small_frame_planes(p.gen, p.gen->window, ws.in, ws.x, ws.tail, lane, kMode == kModeGroup ? rows : nch, sh, tr0 | ((tr1 >= 0 ? tr1 : tr0) << 1));
                if (more) prefetch_rows(p, ws, lane, f + 1, cb, mono_like ? next_nfr : rows, true, row2);   // ws.in fully consumed
                const int Nf = kFrame >> sh;
                if (kMode != kModeGroup && store) {
                    const float *plane = reinterpret_cast<const float *>(ws.x);
                    if (kStereo) {
                        const float4 *s4 = reinterpret_cast<const float4 *>(plane);
                        float4 *dst = reinterpret_cast<float4 *>(p.pcm + off * 2);
                        for (int i = lane; i < Nf / 2; i += 32) __stcs(dst + i, s4[i]);
                    } else {
                        float *dst = p.pcm + off * p.C + cb;
                        for (int idx = lane; idx < 2 * Nf; idx += 32) {
                            const int n = idx >> 1, ch = idx & 1;
                            if (ch < nch) dst[n * p.C + ch] = plane[idx];
                        }
                    }
                    __syncwarp();   // the plane is overwritten by the next frame's stage 1
                }
            }
            if (kMode == kModeGroup && store) {
                mbar_arrive(grp.full);   // this warp's plane of frame f is complete (release) ...
                if (lone_frames == 2) {
                    // ... and it holds frame f + 1 as well, in its second column: that frame's phase of `full`
                    // opens when frame f's has completed (the other warps of the group have arrived for f)
                    mbar_wait_long(grp.full, grp.nstored & 1, 64);
                    grp.nstored++;
                    mbar_arrive(grp.full);
                }
                grp.nstored++;
                grp.pending = true;      // ... and stays intact until the store warps have arrived on `empty`
            }
            flag = next_flag;
            tr1 = next_tr1;
            f += nfr;
            nfr = next_nfr;
        }
        __syncwarp();
        if (f1 == p.nframes && p.tail_out != nullptr) {
            for (int i = lane; i < 2 * kHalfOvl; i += 32) {
                const int ch = i / kHalfOvl;
                if (ch < state_nch) p.tail_out[(cb + ch) * kHalfOvl + (i - ch * kHalfOvl)] = ws.tail[i];
            }
        }
        __syncwarp();
        if (dynamic && kMode == kModeGroup) {
            item = slot == 0 ? claimed : group_next_item(gws->claim, seq);
        } else if (dynamic) {
            unsigned long long next = 0;
            if (lane == 0) next = atomicAdd(p.work_counter, 1ULL) + (unsigned long long)item_stride;
            item = (long long)__shfl_sync(kFull, next, 0);
        } else {
            item += item_stride;
        }
    }
    // (group mode: the store warps finish the last frame on their own; nobody overwrites a plane any more)
}

template <int kMode, int kWarps>
static cudaError_t prepare_variant(int smem)
{
    cudaError_t e = cudaFuncSetAttribute(celt_synth_kernel<kMode, kWarps, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(celt_synth_kernel<kMode, kWarps, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
}

template <int kMode, int kWarps>
static void launch_variant(const SynthParams &p, unsigned grid, int warps, size_t smem, cudaStream_t stream)
{
    if (p.frame_offset) celt_synth_kernel<kMode, kWarps, true><<<grid, warps * 32, smem, stream>>>(p);
    else celt_synth_kernel<kMode, kWarps, false><<<grid, warps * 32, smem, stream>>>(p);
}

// prepare / launch of the variants each part owns (warps = 12 or 14 for the group parts)
cudaError_t prepare_part1(int smem);
cudaError_t prepare_part2(int smem);
cudaError_t prepare_part3(int smem);
void launch_part1(const SynthParams &p, int warps, unsigned grid, size_t smem, cudaStream_t stream);
void launch_part2(const SynthParams &p, int warps, unsigned grid, size_t smem, cudaStream_t stream);
void launch_part3(const SynthParams &p, int warps, unsigned grid, size_t smem, cudaStream_t stream);

#if NQ_PART == 1
cudaError_t prepare_part1(int smem)
{
    return prepare_variant<kModeGroup, 16>(smem);
}
void launch_part1(const SynthParams &p, int warps, unsigned grid, size_t smem, cudaStream_t stream)
{
    launch_variant<kModeGroup, 16>(p, grid, warps, smem, stream);
}
#elif NQ_PART == 2
cudaError_t prepare_part2(int smem)
{
    return prepare_variant<kModeGroupPaired, 16>(smem);
}
void launch_part2(const SynthParams &p, int warps, unsigned grid, size_t smem, cudaStream_t stream)
{
    launch_variant<kModeGroupPaired, 16>(p, grid, warps, smem, stream);
}
#elif NQ_PART == 3
cudaError_t prepare_part3(int smem) { return prepare_variant<kModeDirect, kWarpsPerCta>(smem); }
void launch_part3(const SynthParams &p, int, unsigned grid, size_t smem, cudaStream_t stream)
{
    launch_variant<kModeDirect, kWarpsPerCta>(p, grid, kWarpsPerCta, smem, stream);
}
#endif

#if NQ_PART == 0
// Group mode, warp-specialised: a group is W synthesis warps (one per stream) served by SW store
// warps; a CTA holds G groups and NS store warps, G the largest count that fits shared memory (14
// synthesis warps), 16 warps per CTA and at most kMaxStoreUnits (group, part) units per store warp.
int group_store_warps(int nstreams) { return nstreams >= 8 ? 2 : 1; }

int groups_per_cta(int nstreams)
{
    if (nstreams < 1 || nstreams > kMaxGroupStreams) return 0;
    const int SW = group_store_warps(nstreams);
    int g = kWarpsPerCta / nstreams;
    while (g > 1 && g * nstreams + (g * SW + kMaxStoreUnits - 1) / kMaxStoreUnits > 16) g--;
    if (const char *e = getenv("NQ_GROUPS_PER_CTA")) {   // tuning knob: fewer groups than fit
        const int v = atoi(e);
        if (v >= 1 && v < g) g = v;
    }
    return g;
}

// Store warps of a CTA: one per unit where 16 warps allow it (a multiple of SW, so that every unit
// of a store warp is the same part of its group's frame).
int group_store_warps_cta(int nstreams)
{
    const int G = groups_per_cta(nstreams), SW = group_store_warps(nstreams);
    const int units = G * SW, room = 16 - G * nstreams;
    return units <= room ? units : room / SW * SW;
}

// Threads of a group's store warps that take part in the store pass: the largest T2 <= 32*SW with
// 4*T2 a multiple of C (see StoreCtx); 0 if that leaves less than half of them busy.
int group_store_threads(int C, int nstreams)
{
    int g = C % 4 == 0 ? 4 : (C % 2 == 0 ? 2 : 1);
    const int m = C / g, T = 32 * group_store_warps(nstreams);
    const int T2 = (T / m) * m;
    return 2 * T2 >= T ? T2 : 0;
}

// one_decoder: a single CELT decoder (1 or 2 channels sharing one transient flag).  Two mono
// STREAMS routed to two output channels also have D = C = 2, but each switches blocks on its own.
int synth_mode(int D, int C, int nstreams, bool identity_map, bool one_decoder)
{
    if (D == 2 && C == 2 && nstreams == 1 && identity_map && one_decoder) return kModeStereo;
    if (D == 1 && C == 1 && identity_map) return kModeMono;
    if (nstreams <= kMaxGroupStreams) return kModeGroup;   // (store pass: group_store_threads() or the general loop)
    return kModeDirect;
}

cudaError_t prepare_kernels()
{
    const int smem = (int)fast_kernel_smem_bytes();
    cudaError_t e = prepare_variant<kModeStereo, kWarpsPerCta>(smem);
    if (e == cudaSuccess) e = prepare_variant<kModeMono, kWarpsPerCta>(smem);
    if (e == cudaSuccess) e = prepare_part1(smem);
    if (e == cudaSuccess) e = prepare_part2(smem);
    if (e == cudaSuccess) e = prepare_part3(smem);
    return e;
}

cudaError_t launch_synth(const SynthParams &p, int mode, int num_sms, cudaStream_t stream, int *launched_ctas)
{
    long long per_cta = kWarpsPerCta, nitems = p.nruns * p.npairs;
    int warps = kWarpsPerCta, smem_warps = kWarpsPerCta;
    bool paired = false;
    if (mode == kModeGroup) {
        per_cta = p.groups_per_cta;
        nitems = p.nruns;
        smem_warps = p.groups_per_cta * p.nstreams;   // synthesis warps own a shared-memory slice, store warps do not
        warps = smem_warps + p.store_warps_cta;
        for (int s = 0; s < p.nstreams; s++) paired = paired || p.streams[s].flag_col1 != p.streams[s].flag_col;
    }
    long long ctas = (nitems + per_cta - 1) / per_cta;
    if (ctas > num_sms) ctas = num_sms;   // persistent: one CTA per SM, warps stride over the items
    if (ctas < 1) ctas = 1;
    if (launched_ctas) *launched_ctas = (int)ctas;
    const size_t smem = sizeof(FastTables) + (size_t)smem_warps * sizeof(WarpSmem);
    const unsigned grid = (unsigned)ctas;
    if (mode == kModeStereo) launch_variant<kModeStereo, kWarpsPerCta>(p, grid, warps, smem, stream);
    else if (mode == kModeMono) launch_variant<kModeMono, kWarpsPerCta>(p, grid, warps, smem, stream);
    else if (mode == kModeGroup && paired) launch_part2(p, warps, grid, smem, stream);
    else if (mode == kModeGroup) launch_part1(p, warps, grid, smem, stream);
    else launch_part3(p, warps, grid, smem, stream);
    return cudaGetLastError();
}

// -------------------------------------------------------- generic kernel ---
// One clt_mdct_backward call per CTA for ANY (shift, stride): the drop-in
// entry points (single calls, all four transform sizes) and the LM < 3
// shapes.  Follows the reference formulas literally for the rotations
// (mdct.c:295-313, :320-359, :361-377) with the reference's trig/window
// tables; the N4-point inverse DFT is 30 x R (R = 16, 8, 4, 2): 30-point
// prime-factor DFTs in registers, then an R-term direct sum.  Latency-bound
// by design (a batch of one); throughput work belongs to the fast kernel.
constexpr int kGenericThreads = 128;

__global__ void __launch_bounds__(kGenericThreads) mdct_backward_generic_kernel(const MdctCall *calls, const GenericTables *tabs)
{
    __shared__ float2 a_buf[480];   // pre-rotated input, later the FFT output
    __shared__ float2 b_buf[480];   // stage-1 output, later y[] as 960 floats
    __shared__ float s_trig[481];
    __shared__ float s_win[kOverlap];

    const MdctCall call = calls[blockIdx.x];
    const int shift = call.shift, stride = call.stride;
    const int N = kMdctN >> shift, N2 = N >> 1, N4 = N >> 2, R = N4 / 30;
    const float sine = (float)2 * 3.141592653f * (.125f) / N;   // mdct.c:292
    const int tid = threadIdx.x;

    for (int i = tid; i < 481; i += kGenericThreads) s_trig[i] = tabs->trig[i];
    for (int i = tid; i < kOverlap; i += kGenericThreads) s_win[i] = tabs->window[i];
    __syncthreads();

    // pre-rotate, mdct.c:303-312
    for (int i = tid; i < N4; i += kGenericThreads) {
        if (call.ifft_only) {
            a_buf[i] = reinterpret_cast<const float2 *>(call.in)[i];
            continue;
        }
        const float x1 = call.in[(size_t)2 * i * stride];
        const float x2 = call.in[(size_t)(N2 - 1 - 2 * i) * stride];
        const float t0 = s_trig[i << shift], t1 = s_trig[(N4 - i) << shift];
        const float yr = -(x2 * t0) + x1 * t1;
        const float yi = -(x2 * t1) - x1 * t0;
        a_buf[i] = make_float2(yr - yi * sine, yi + yr * sine);
    }
    __syncthreads();

    // inverse DFT, N4 = 30 * R: i = R*n1 + n2, k = k1 + 30*k2
    if (tid < R) {
        const int n2 = tid;
        float2 g[30];
#pragma unroll
        for (int n1 = 0; n1 < 30; n1++) g[n1] = a_buf[R * n1 + n2];
        idft30(g);
#pragma unroll
        for (int k1 = 0; k1 < 30; k1++) {
            float sn, cs;
            sincospif(2.0f * (float)(n2 * k1) / (float)N4, &sn, &cs);   // exp(+j 2pi n2 k1 / N4)
            b_buf[n2 * 30 + k1] = cmulc(g[k1], cs, sn);
        }
    }
    __syncthreads();
    for (int k = tid; k < N4; k += kGenericThreads) {
        const int k1 = k % 30, k2 = k / 30;
        float2 acc = make_float2(0.f, 0.f);
        for (int n2 = 0; n2 < R; n2++) {
            float sn, cs;
            sincospif(2.0f * (float)((n2 * k2) % R) / (float)R, &sn, &cs);
            acc = cadd(acc, cmulc(b_buf[n2 * 30 + k1], cs, sn));
        }
        a_buf[k] = acc;
        if (call.ifft_only) reinterpret_cast<float2 *>(call.out)[k] = acc;
    }
    if (call.ifft_only) return;
    __syncthreads();

    // post-rotate + de-shuffle, mdct.c:330-358: bin k -> y[2k], y[N2-1-2k]
    float *y = reinterpret_cast<float *>(b_buf);
    for (int k = tid; k < N4; k += kGenericThreads) {
        const float re = a_buf[k].x, im = a_buf[k].y;
        const float t0 = s_trig[k << shift], t1 = s_trig[(N4 - k) << shift];
        const float yr = re * t0 - im * t1;
        const float yi = im * t0 + re * t1;
        y[2 * k] = -(yr - yi * sine);
        y[N2 - 1 - 2 * k] = yi + yr * sine;
    }
    __syncthreads();

    // out[60+m] = y[m]; TDAC mirror on out[0..120) with the previous tail, mdct.c:368-376
    for (int m = kHalfOvl + tid; m < N2; m += kGenericThreads) call.out[kHalfOvl + m] = y[m];
    for (int i = tid; i < kHalfOvl; i += kGenericThreads) {
        const float x1 = y[kHalfOvl - 1 - i];   // out[119 - i]
        const float x2 = call.out[i];
        call.out[i] = s_win[kOverlap - 1 - i] * x2 - s_win[i] * x1;
        call.out[kOverlap - 1 - i] = s_win[i] * x2 + s_win[kOverlap - 1 - i] * x1;
    }
}

cudaError_t launch_mdct_generic(const MdctCall *d_calls, int ncalls, const GenericTables *d_tables, cudaStream_t stream)
{
    if (ncalls <= 0) return cudaSuccess;
    mdct_backward_generic_kernel<<<ncalls, kGenericThreads, 0, stream>>>(d_calls, d_tables);
    return cudaGetLastError();
}

#endif   // NQ_PART == 0

}  // namespace nq
