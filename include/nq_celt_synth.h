/* nq_celt_synth.h -- C ABI of the B200-native CELT synthesis stage
 * (inverse MDCT + TDAC windowed overlap-add + channel interleave).
 *
 * Drop-in boundary for ONE path of dafx/libnyquist's bundled Opus decoder
 * (reference paths relative to /root/reference/):
 *
 *   compute_inv_mdcts         third_party/opus/celt/celt_decoder_clean.c:264-312
 *   clt_mdct_backward         third_party/opus/celt/mdct.c:267-379, proto mdct.h:66-68
 *   clt_mdct_backward_B1_C2   third_party/opus/celt/mdct.c:258-265 (CPU), :223-243 (USE_CUDA hook)
 *   opus_ifft                 third_party/opus/celt/kiss_fft.c:696-747
 *   fork's GPU seam           cuda/mdct_cuda.hpp:79-103 (processMDCTCuda, processMDCTCudaB1C2,
 *                             cleanupCudaBuffers, printCudaVersion)
 *
 * Plain C, plain pointers and sizes; no CUDA or torch types in any signature
 * (streams are passed as void*, i.e. a cudaStream_t).  The library has NO CPU
 * fallback: without a CUDA device every entry fails (NQ_INTERNAL_ERROR) or,
 * for the void reference-shaped entries, aborts loudly.
 *
 * Arithmetic: IEEE float32 throughout (the reference's float build,
 * celt/arch.h:133-138).  Results match the reference within 1e-5 of full
 * scale (32768) -- the FFT factorisation differs from kiss_fft's, so parity is
 * tolerance-based, not bit-exact (observed ~2e-7 of full scale).
 */
#ifndef NQ_CELT_SYNTH_H
#define NQ_CELT_SYNTH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define NQ_API
#else
#define NQ_API __attribute__((visibility("default")))
#endif

/* Error codes mirror Opus (third_party/opus/libopus/include/opus_defines.h:46-60). */
#define NQ_OK 0
#define NQ_BAD_ARG (-1)
#define NQ_INTERNAL_ERROR (-3)   /* a CUDA call failed; see nq_celt_last_error() */
#define NQ_UNIMPLEMENTED (-5)
#define NQ_INVALID_STATE (-6)
#define NQ_ALLOC_FAIL (-7)

#define NQ_CELT_FRAME 960      /* samples per channel per 20 ms frame (LM = 3)             */
#define NQ_CELT_HALF_OVERLAP 60 /* raw tail carried between blocks (overlap/2, mdct.c:361) */

/* Per-device context: device tables, one stream, scratch buffers for the
 * host-pointer entries.  Not re-entrant; different contexts may be used
 * concurrently from different threads (reference threading contract:
 * one decoder state per stream). */
typedef struct nq_celt_ctx nq_celt_ctx;

NQ_API int nq_celt_ctx_create(int device, nq_celt_ctx **out);
NQ_API void nq_celt_ctx_destroy(nq_celt_ctx *ctx);
NQ_API const char *nq_celt_strerror(int code);
NQ_API const char *nq_celt_last_error(const nq_celt_ctx *ctx);
NQ_API int nq_celt_device_count(void);
/* The context's device ordinal and its own stream (a cudaStream_t), for callers that enqueue copies
 * next to the batch entries. */
NQ_API int nq_celt_ctx_device(const nq_celt_ctx *ctx);
NQ_API void *nq_celt_ctx_stream(const nq_celt_ctx *ctx);
/* Number of kernel launches issued through this context so far. */
NQ_API long long nq_celt_launch_count(const nq_celt_ctx *ctx);

/* Pinned host memory for the host-pointer batch entry (optional but needed
 * for full PCIe bandwidth). */
NQ_API void *nq_celt_host_alloc(size_t bytes);
NQ_API void nq_celt_host_free(void *p);

/* ---- batched synthesis: phase 2 of the restructured decoder ---------------
 * Replaces the per-frame call compute_inv_mdcts(mode, shortBlocks, freq,
 * out_syn, C, LM=3) at celt_decoder_clean.c:656 together with the history
 * shift at :622-626, for nframes consecutive 20 ms frames of one stream.
 *
 *   coef       [nframes][C][960]  spectral coefficients exactly as
 *              denormalise_bands leaves them in freq[] (channel-major; a
 *              transient frame keeps its 8 sub-blocks interleaved,
 *              freq[c*960 + j*8 + b], celt_decoder_clean.c:296)
 *   transient  [nframes]          flag byte per frame: 0 = long block, 1 = isTransient
 *              (shortBlocks = 8; celt_decoder_clean.c:586, :656).  Bit 3 = decoder
 *              reset before the frame.  These three entries take 20 ms frames
 *              only: bits 1-2 (3 - LM) must be zero -- the _host / _host_multi
 *              entries return NQ_BAD_ARG otherwise; _device cannot inspect device
 *              memory and would synthesise such a frame as a 20 ms one, so frames
 *              shorter than 20 ms go through nq_celt_synth_batch_device_ms with
 *              frame_offset (or nq_celt_decode_batch_host, whose side info carries N)
 *   tail_in    [C][60] or NULL    raw tail left by the frame before the batch
 *              (= out_syn[c][960..1020) after that frame); NULL => zeros,
 *              i.e. a freshly reset decoder (celt_decoder_clean.c:846-859)
 *   halo_coef  [C][960] or NULL   alternative to tail_in for a shard that
 *              starts mid-stream: coefficients of the frame before the
 *              batch; that frame is re-synthesised only for its tail.
 *              Ignored when tail_in is non-NULL.
 *   pcm_out    [nframes*960][C]   interleaved celt_sig samples (the values of
 *              out_syn[c][0..960) per frame, before comb_filter/deemphasis)
 *   tail_out   [C][60] or NULL    raw tail after the last frame
 *
 * _device: all pointers are device pointers on ctx's device, 16-byte aligned
 * (coef, pcm_out); the work is enqueued on `stream` (NULL => ctx's stream)
 * and the call returns without synchronising.
 * _host: all pointers are host pointers; chunks are pipelined H2D / kernel /
 * D2H on internal streams and the call returns when pcm_out is complete.
 */
NQ_API int nq_celt_synth_batch_device(nq_celt_ctx *ctx, const float *coef, const uint8_t *transient,
                                      const float *tail_in, const float *halo_coef, int halo_transient,
                                      float *pcm_out, float *tail_out, int64_t nframes, int C, void *stream);

NQ_API int nq_celt_synth_batch_host(nq_celt_ctx *ctx, const float *coef, const uint8_t *transient,
                                    const float *tail_in, float *pcm_out, float *tail_out,
                                    int64_t nframes, int C);

/* Opus multistream batch: `streams` independent CELT decoders, the first
 * `coupled_streams` of them stereo, routed to `channels` output channels by
 * `mapping` -- the arguments of opus_multistream_decoder_create
 * (third_party/opus/libopus/src/opus_multistream_decoder.c:110).  Replaces, for
 * nframes frames, the per-stream compute_inv_mdcts calls AND the strided
 * channel copy opus_multistream_decode_native does afterwards (:237-299,
 * get_left/right/mono_channel opus_multistream.c:57-91): the interleave is
 * fused into the synthesis kernel's store pass.
 *
 *   D = streams + coupled_streams decoded channels; decoded channel d is
 *       stream d/2 (left, right) for d < 2*coupled_streams, else the mono
 *       stream d - coupled_streams
 *   coef       [nframes][D][960]   rows in decoded-channel order
 *   transient  [nframes][streams]  every stream has its own block switching
 *   tail_in / tail_out [D][60], halo_coef [D][960] (device), halo_transient
 *              [streams] (HOST pointer; needed with halo_coef)
 *   mapping    [channels] (HOST pointer): output channel c carries decoded
 *              channel mapping[c]; 255 = silent channel; a decoded channel may
 *              feed several outputs or none
 *   pcm_out    [nframes*960][channels]
 *   frame_offset [nframes] (device) or NULL: first output sample of every
 *              frame; needed when the batch holds frames shorter than 20 ms
 *              (SURVEY.md 8(f) row 4).  A flag byte then carries the frame
 *              size too: bit 0 = transient, bits 1-2 = 3 - LM (0 = 20 ms, so
 *              plain 0/1 flags keep meaning 20 ms frames).  Bit 3 (any batch,
 *              with or without frame_offset): the decoder was reset before
 *              this frame, i.e. it overlap-adds against a zero tail -- the
 *              first frame of another file in a batch of many.  Such a frame has
 *              N = 120 << LM coefficients at the start of each 960-float row
 *              (transient: 1 << LM short blocks interleaved, as the reference).
 *              NULL: every frame is a 20 ms frame, frame f starts at 960 f.
 * mapping == NULL: one CELT decoder (streams = 1, channels 1 or 2).
 * One warp per coupled stream and one per PAIR of mono streams, at most 14
 * warps per batch layout (NQ_UNIMPLEMENTED beyond; 7.1 surround needs 4).  All
 * pointers except mapping / halo_transient are device pointers; enqueued on
 * `stream`, no synchronisation. */
NQ_API int nq_celt_synth_batch_device_ms(nq_celt_ctx *ctx, const float *coef, const uint8_t *transient,
                                         const float *tail_in, const float *halo_coef,
                                         const uint8_t *halo_transient, float *pcm_out, float *tail_out,
                                         const int64_t *frame_offset, int64_t nframes, int channels, int streams,
                                         int coupled_streams, const unsigned char *mapping, void *stream);

/* ---- post stage: pitch post-filter + de-emphasis ---
 * What celt_decode_with_ec does to out_syn after compute_inv_mdcts: comb_filter
 * twice (celt_decoder_clean.c:658-670, celt/celt.c:114-172), the parameter
 * hand-over (:672-683) and deemphasis (:723, :192-256), which also scales to
 * [-1, 1] and interleaves.  Side information per frame per stream, produced by
 * phase 1 (the entropy decoder): */
typedef struct nq_celt_post_frame {
    int32_t N;          /* samples per channel of the frame: 120 << LM                       */
    int32_t pitch[3];   /* postfilter_period_old, postfilter_period, postfilter_pitch (:436) */
    float gain[3];      /* ..._gain_old, ..._gain, postfilter_gain                           */
    int32_t tapset[3];  /* ..._tapset_old, ..._tapset, postfilter_tapset                     */
} nq_celt_post_frame;   /* i.e. [0] -> [1] fades over samples [0,120), [1] -> [2] over [120,240), [2] from there on */
#define NQ_CELT_POST_HISTORY 1026   /* COMBFILTER_MAXPERIOD + 2 (celt.h:187) filtered samples per channel */

/* In place on the buffer the synthesis wrote: pcm [nsamples][channels] holds
 * celt_sig on entry and float PCM in [-1, 1] on return (nsamples = sum of N).
 *   frames   [nframes][streams]  HOST pointer (40 bytes per stream-frame)
 *   hist_*   [D][1026]           last filtered samples per decoded channel
 *                                (the part of decode_mem the comb filter reads,
 *                                celt_decoder_clean.c:92); NULL in = reset decoder
 *   mem_*    [D]                 preemph_memD (celt_decoder_clean.c:90)
 * pcm / hist / mem are device pointers.  mapping == NULL: one CELT decoder,
 * channels = 1 or 2, streams = 1.  One warp per stream walks the frames in
 * order (the filters are recurrences): parallel over streams, not over time. */
NQ_API int nq_celt_post_batch_device(nq_celt_ctx *ctx, float *pcm, const nq_celt_post_frame *frames,
                                     const float *hist_in, const float *mem_in, float *hist_out, float *mem_out,
                                     int64_t nframes, int channels, int streams, int coupled_streams,
                                     const unsigned char *mapping, void *stream);

/* A batch of MANY independent streams (files): segment k = frames
 * [seg_start[k], seg_start[k+1]) (HOST array of nseg+1 entries, 0 .. nframes),
 * every segment filtered from a reset decoder by its own CTA(s) -- the way the
 * post stage, whose filters are recurrences along time, fills the GPU.  The
 * synthesis side of such a batch is the ordinary nq_celt_synth_batch_device*
 * call with flag bit 3 (value 8, "decoder reset before this frame",
 * celt_decoder_clean.c:846-859) set on the first frame of every segment. */
NQ_API int nq_celt_post_segments_device(nq_celt_ctx *ctx, float *pcm, const nq_celt_post_frame *frames,
                                        const int64_t *seg_start, int nseg, int64_t nframes, int channels,
                                        int streams, int coupled_streams, const unsigned char *mapping,
                                        void *stream);

/* Whole phase 2 on host buffers: coefficients + side info in, float PCM out
 * (synthesis, channel mapping, post-filter, de-emphasis), chunks pipelined
 * H2D / kernels / D2H.  Decoder state (tail [D][60], hist [D][1026], mem [D];
 * NULL in = reset decoder, NULL out = discard) lets a stream be fed in pieces.
 * Frames shorter than 20 ms are allowed anywhere: frames[].N = 120 << LM and
 * the flag bytes (see frame_offset above) must agree (NQ_BAD_ARG otherwise);
 * pcm_out then holds sum(N) samples per channel. */
NQ_API int nq_celt_decode_batch_host(nq_celt_ctx *ctx, const float *coef, const uint8_t *transient,
                                     const nq_celt_post_frame *frames, const float *tail_in, const float *hist_in,
                                     const float *mem_in, float *pcm_out, float *tail_out, float *hist_out,
                                     float *mem_out, int64_t nframes, int channels, int streams,
                                     int coupled_streams, const unsigned char *mapping);

/* ---- frame sink: phase 1 -> phase 2 hand-over ------
 * The restructured celt_decode_with_ec stops after denormalise_bands
 * (celt_decoder_clean.c:620-636) and, instead of compute_inv_mdcts / comb_filter
 * / deemphasis (:656-723), pushes the frame here; INTEGRATION.md shows the
 * reference-side patch.  One sink per Ogg Opus link / multistream decoder
 * (arguments of opus_multistream_decoder_create, opus_multistream_decoder.c:110).
 * The sink keeps pushed frames in pinned host memory and the phase-2 decoder
 * state (IMDCT tail, comb-filter history, de-emphasis memory) across flushes. */
typedef struct nq_celt_sink nq_celt_sink;
NQ_API int nq_celt_sink_create(nq_celt_sink **out, int channels, int streams, int coupled_streams,
                               const unsigned char *mapping);
NQ_API void nq_celt_sink_destroy(nq_celt_sink *sink);
NQ_API const char *nq_celt_sink_last_error(const nq_celt_sink *sink);
/* One decoded CELT frame of stream `stream` (opus_multistream order), in decode
 * order (pushes of DIFFERENT streams may come from different threads at the same
 * time -- the streams of a multistream packet are independent decoders,
 * opus_multistream_decoder.c:237-251 -- every other sink call belongs to one
 * thread): freq [CC][N] as it stands at celt_decoder_clean.c:636 (CC = 2 for a
 * coupled stream, 1 for a mono one), N = 120 << LM, shortBlocks = 0 or 1 << LM
 * (:273-284), post = the comb_filter arguments of the frame (:660-669). */
NQ_API int nq_celt_sink_push(nq_celt_sink *sink, int stream, const float *freq, int CC, int N, int shortBlocks,
                             const nq_celt_post_frame *post);
/* Same, for a frame that does not simply follow the one before it in the output (files that switch
 * between the SILK, hybrid and CELT coding modes, opus_decoder_clean.c:478-487, :499-513, :530-555:
 * the CELT decoder is not called for every stretch of the output, and the 5 ms redundancy frames are
 * decoded into a side buffer and cross-faded in by the caller).  Streaming mode, single-stream files.
 *   dest >= 0   first output sample (per channel, before the pre-skip window) of this frame;
 *   dest == -1  right after the previous frame (what nq_celt_sink_push does);
 *   dest <= -2  a side frame with the caller's tag -2 - dest: decoded in sequence like any frame -- it
 *               takes part in the overlap-add, the post-filter and the de-emphasis state -- but its PCM is
 *               kept aside: nq_celt_sink_side_get after nq_celt_sink_finish, in push order. */
NQ_API int nq_celt_sink_push_at(nq_celt_sink *sink, int stream, const float *freq, int CC, int N, int shortBlocks,
                                const nq_celt_post_frame *post, int64_t dest);
NQ_API int nq_celt_sink_side_count(const nq_celt_sink *sink);
NQ_API int nq_celt_sink_side_get(const nq_celt_sink *sink, int i, int64_t *tag, int *nsamples, const float **pcm);
NQ_API int64_t nq_celt_sink_pending_frames(const nq_celt_sink *sink);
NQ_API int64_t nq_celt_sink_pending_samples(const nq_celt_sink *sink);   /* per channel */
/* Phase 2 for everything pushed since the last flush (every stream must have
 * pushed the same number of frames): pcm_out [*nsamples][channels] host buffer,
 * float PCM in [-1, 1] in output-channel order. */
NQ_API int nq_celt_sink_flush(nq_celt_sink *sink, nq_celt_ctx *ctx, float *pcm_out, int64_t capacity_samples,
                              int64_t *nsamples);
/* Same, into a pinned buffer the sink owns (valid until the next flush or
 * nq_celt_sink_destroy): saves the caller a page-locked allocation per file. */
NQ_API int nq_celt_sink_flush_pinned(nq_celt_sink *sink, nq_celt_ctx *ctx, const float **pcm, int64_t *nsamples);
/* Streaming phase 2: after attach, every complete block of 2048 frames is
 * decoded on a worker thread while the caller keeps pushing (phase 1 and phase 2
 * overlap), and decoded sample p of every channel is written to
 * dst[(p - skip_samples) * channels + c] when 0 <= p - skip_samples < dst_samples
 * -- the positional pre-skip / end-trim window opusfile applies
 * (opusfile.c:2673-2721).  finish() decodes the last partial block, waits for
 * the worker and reports how many samples per channel were decoded in all. */
NQ_API int nq_celt_sink_attach(nq_celt_sink *sink, nq_celt_ctx *ctx, float *dst, int64_t skip_samples,
                               int64_t dst_samples);
NQ_API int nq_celt_sink_finish(nq_celt_sink *sink, int64_t *decoded_samples);
/* attach() may be given dst == NULL when the output buffer is still being allocated (zero-filling
 * the samples of a long file takes tens of milliseconds that phase 1 can use): the worker decodes
 * meanwhile and waits for this call before it copies the first block out.  Must come before finish(). */
NQ_API int nq_celt_sink_set_destination(nq_celt_sink *sink, float *dst);
/* Pinned blocks of destroyed sinks are recycled process-wide (page-locking is
 * slow); this releases them. */
NQ_API void nq_celt_sink_trim_pool(void);
/* Phase 2 for MANY files at once -- the producer of the many-streams kernels.  `nsinks` sinks of
 * the same channel layout, each holding one whole file (pushed from a reset decoder, not attached,
 * not flushed before): their frames are gathered into ONE device batch (every file's first frame
 * carries the reset flag), synthesised by ONE kernel launch and post-filtered by ONE launch with
 * one CTA per file and channel pair (the reference runs the loop src/OpusDecoder.cpp:101-119 once
 * per file).  pcm_out[k] receives decoded samples [skip[k], skip[k] + count[k]) of file k
 * (interleaved, `channels` wide; the positional pre-skip / end-trim window of opusfile.c:2673-2721);
 * decoded[k] = samples per channel file k holds in all.  The destinations may be ordinary
 * (pageable) memory: the samples come back through a ring of pinned buffers and a few host threads.
 * The sinks are empty and reset afterwards. */
/* Optional, before the first push of a file that will go through nq_celt_sink_flush_many: a worker
 * thread moves every complete block of 2048 frames to a device buffer of the sink's own while phase
 * 1 keeps pushing (and returns the pinned block to the pool), so that flush_many only gathers
 * device to device instead of uploading whole files after phase 1 has ended.  expected_frames:
 * how many frames the file is expected to hold (sizes the device buffer; 0 = unknown, it grows). */
NQ_API int nq_celt_sink_begin_upload(nq_celt_sink *sink, nq_celt_ctx *ctx, int64_t expected_frames);
NQ_API int nq_celt_sink_flush_many(nq_celt_sink *const *sinks, int nsinks, nq_celt_ctx *ctx, float *const *pcm_out,
                                   const int64_t *skip, const int64_t *count, int64_t *decoded);
/* OPUS_RESET_STATE (celt_decoder_clean.c:846-859): the next frame pushed for every
 * stream starts from a cleared decoder (tail, history, memory), wherever that
 * falls relative to the flushes. */
NQ_API void nq_celt_sink_reset(nq_celt_sink *sink);
/* Same for one stream: each stream of a multistream file is a decoder of its own
 * (opus_multistream_decoder.c:237-251) and is reset on its own. */
NQ_API void nq_celt_sink_reset_stream(nq_celt_sink *sink, int stream);

/* Same as _host, frames sharded contiguously over `ndev` devices (devices[i]
 * = CUDA ordinal; NULL => 0..ndev-1) with one host thread + context per
 * device and NO device-to-device traffic: a shard that starts mid-stream
 * uploads the previous frame's coefficients as its halo.  The per-device
 * contexts are created on first use and kept for the process (calls are
 * serialised); nq_celt_multi_release() -- or cleanupCudaBuffers() -- frees them. */
NQ_API int nq_celt_synth_batch_host_multi(const int *devices, int ndev, const float *coef,
                                          const uint8_t *transient, const float *tail_in,
                                          float *pcm_out, float *tail_out, int64_t nframes, int C);
NQ_API void nq_celt_multi_release(void);

/* ---- single-call entries with the reference's semantics -------------------
 * nq_clt_mdct_backward == clt_mdct_backward (mdct.c:267): HOST pointers,
 * synchronous.  `in`: N2 = (1920 >> shift)/2 coefficients at `stride`
 * (untouched); `out`: read [0, overlap/2), written [0, N2 + overlap/2).
 * Only the reference's static mode is supported (mdct n = 1920, overlap = 120,
 * shift 0..3, stride >= 1); `l` and `window` may be NULL (the library owns
 * bit-compatible tables) and are otherwise only sanity-checked.
 * These are `void` like the reference; on any CUDA failure they print the
 * error and abort() -- there is no CPU fallback by design. */
typedef struct nq_mdct_lookup {   /* layout of mdct_lookup, mdct.h:49-54 */
    int n;
    int maxshift;
    const void *kfft[4];
    const float *trig;
} nq_mdct_lookup;

NQ_API void nq_clt_mdct_backward(const nq_mdct_lookup *l, float *in, float *out, const float *window,
                                 int overlap, int shift, int stride);
NQ_API void nq_clt_mdct_backward_B1_C2(const nq_mdct_lookup *l, float *in[2], float *out[2],
                                       const float *window, int overlap, int shift, int stride);
/* Error-returning form used by the two above; ctx == NULL => process-global
 * context on the current device. */
NQ_API int nq_celt_mdct_backward_host(nq_celt_ctx *ctx, const float *const *in, float *const *out, int ncalls,
                                      int shift, int stride);

/* opus_ifft (kiss_fft.c:696-747) with the static state mode->mdct.kfft[shift]:
 * `count` independent N4 = 480 >> shift point UNNORMALISED inverse DFTs on
 * interleaved (re, im) float32 host buffers, out of place (kiss_fft.c:707).
 * This is the entry the reference's orphan golden vectors
 * test_data/ifft_{input,output}_N{480,60}.bin pin. */
NQ_API int nq_opus_ifft_host(nq_celt_ctx *ctx, int shift, const float *in_ri, float *out_ri, int count);

/* compute_inv_mdcts (celt_decoder_clean.c:264) on host buffers: out_mem[c]
 * points at out_syn[c]; [0,60) holds the previous raw tail on entry and
 * [0, N*B + 60) is written.  LM in 0..3; shortBlocks = 0 or 1 << LM. */
NQ_API int nq_compute_inv_mdcts(nq_celt_ctx *ctx, int shortBlocks, const float *X, float *const *out_mem,
                                int C, int LM);

/* ---- the fork's existing GPU seam (cuda/mdct_cuda.hpp:79-103) -------------
 * Same names and signatures, so the reference built with -DUSE_CUDA links
 * against this library instead of cuda/mdct_cuda.cu + cuFFT unchanged
 * (mdct.c:223-254 calls these).  N, sine, overlap, trig and window are
 * accepted for signature compatibility; N must be 1920 >> shift. */
NQ_API void processMDCTCuda(const float *input, float *output, const float *trig, int N, int shift,
                            int stride, float sine, int overlap, const float *window);
NQ_API void processMDCTCudaB1C2(const float *input[2], float *output[2], const float *trig, int N,
                                int shift, int stride, float sine, int overlap, const float *window);
NQ_API void cleanupCudaBuffers(void);
NQ_API void printCudaVersion(void);

/* ---- introspection for tests ------------------------------------------- */
/* How a batch of `nframes` frames with this channel layout would be launched on
 * a device with `num_sms` SMs (pure host code, no device needed; mapping == NULL:
 * plain `channels`-channel batch with shared flags).  out[0] kernel variant
 * (0 stereo, 1 group, 2 direct, 4 mono), [1] warps per group, [2] groups per CTA,
 * [3] threads in the store pass, [4] store-loop shape, [5] mono streams paired in
 * one warp, [6] frames per run, [7] runs, [8] post-stage CTAs, [9] of which
 * two-channel, [10] decoded channels, [11] identity mapping.  NQ_UNIMPLEMENTED /
 * NQ_BAD_ARG exactly where the batch entries return them. */
NQ_API int nq_celt_debug_plan(int channels, int streams, int coupled_streams, const unsigned char *mapping,
                              int64_t nframes, int num_sms, int64_t out[12]);
/* The runs a batch of `nframes` frames is cut into (pure host code): run_first[r] = first frame of run r
 * for r < *nruns, run_first[*nruns] = nframes (as far as `capacity` entries allow). */
NQ_API int nq_celt_debug_runs(int channels, int64_t nframes, int num_sms, int64_t *run_first, int64_t capacity,
                              int64_t *nruns);
/* Copies the host-built tables: t_long [16*31*2], t_short [2*30*2],
 * window [120], trig [481] (any pointer may be NULL). */
NQ_API void nq_celt_debug_tables(float *t_long, float *t_short, float *window, float *trig);

#ifdef __cplusplus
}
#endif
#endif /* NQ_CELT_SYNTH_H */
