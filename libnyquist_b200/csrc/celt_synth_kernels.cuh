// Shared declarations between the kernels (celt_synth_kernels.cu) and the
// host layer / C-ABI (celt_synth_api.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace nq {

constexpr int kFrame = 960;      // samples per channel per 20 ms frame (N2 of the long MDCT)
constexpr int kHalfOvl = 60;     // overlap/2: raw tail carried between (sub-)blocks
constexpr int kOverlap = 120;    // mode->overlap, static_modes_float.h:579
constexpr int kMdctN = 1920;     // mode->mdct.n,  static_modes_float.h:591

// Fast kernel geometry: one warp owns one run of consecutive frames of one
// channel pair; warps are independent (no block-wide barrier in the loop).
constexpr int kWarpsPerCta = 14;
constexpr int kInRowFloats = 968;        // 960 coefficients + 8 pad (bank offset between the two rows)
constexpr int kXRowF2 = 31;              // padded row of the 16x30 inter-stage buffer (float2 units)
constexpr int kXChanF2 = 16 * kXRowF2;   // 496 float2 per channel

// Tables the fast kernel keeps in shared memory (one copy per CTA).  Built on
// the host in double precision by build_tables() (celt_synth_api.cu).
struct FastTables {
    float2 t_long[kXChanF2];   // [n2][k1] (row stride 31): -(1+js)^2 e^{j2pi(n2+k1)/1920} e^{j2pi n2 k1/480}
    float2 t_short[2 * 30];    // [h][k1]:                  -(1+js)^2 e^{j2pi(h+k1)/240}  e^{j2pi h k1/60}
    // short-block mirror: wpair[h][i] = (window[59 - m], window[60 + m]) with m = 2i + h, the two
    // window taps a lane (.., h) needs in trip i of short_stage2's loop (window120:
    // static_modes_float.h:9, modes.c:374 formula) -- one 64-bit load per trip
    float2 wpair[2 * 30];
};

// Tables of the generic (any shift / stride) kernel: the reference's own
// trig + window tables (mdct.c:99, modes.c:374).
struct GenericTables {
    float trig[481];
    float window[kOverlap];
};

// How the fast kernel is specialised (template parameter):
//   kModeStereo  D = C = 2, one warp per run, float4 {L,R,L,R} stores straight from registers;
//   kModeGroup   any channel layout with <= kMaxGroupStreams streams, warp-specialised: a GROUP of
//                synthesis warps (one per stream = one coupled pair or one mono channel) works on a
//                run; every warp leaves its frame as a [960][2] plane in shared memory, and the
//                group's store warp writes the interleaved [960][C] output frame with contiguous
//                float4 stores, gathering output channel c from decoded channel mapping[c]
//                (opus_multistream_decoder.c:260-299);
//   kModeDirect  more streams than a CTA has warps: channel pairs, scattered stores (slow, rare);
//   kModeMono    D = C = 1: one warp per run like stereo, its two "channels" being two consecutive
//                frames of the stream whenever they have the same block type.
// flag byte of a frame: bit 0 transient, bits 1-2 = 3 - LM, bit 3 = the decoder was reset before this frame
constexpr int kFlagTransient = 1, kFlagReset = 8;
constexpr int kModeStereo = 0, kModeGroup = 1, kModeDirect = 2;
constexpr int kModeMono = 4;
constexpr int kModeGroupPaired = 3;   // kernel-internal: kModeGroup whose warps may carry two independent mono streams
constexpr int kMaxGroupStreams = kWarpsPerCta;
constexpr int kMaxChannels = 255;

// One warp of a group: a coupled stream, a mono stream, or TWO consecutive mono streams (their
// rows are adjacent; each keeps its own transient flag).
struct StreamDesc {
    uint8_t nch;         // channels this warp synthesises (1 or 2)
    uint8_t row;         // first decoded channel (coefficient row inside a frame)
    uint8_t flag_col;    // column of channel 0's flag inside a frame's flag record
    uint8_t flag_col1;   // column of channel 1's flag (== flag_col for a coupled stream)
};

struct SynthParams {
    const float *coef;          // [nframes][D][960]
    const uint8_t *transient;   // [nframes][flag_stride]
    const float *tail_in;       // [D][60] or nullptr
    const float *halo_coef;     // [D][960] coefficients of frame -1, or nullptr
    float *pcm;                 // [nframes*960][C]
    float *tail_out;            // [D][60] or nullptr
    const FastTables *tables;
    const GenericTables *gen;          // reference trig table: frames shorter than 20 ms
    const long long *frame_offset;     // [nframes] first sample of every frame, or nullptr: frame f starts at 960 f
    unsigned long long *work_counter;  // zeroed before the launch: runs beyond the first wave are claimed dynamically; nullptr: static
    long long nframes;
    long long frames_per_run;   // frames of runs [0, big_runs)
    long long nruns;            // all runs
    long long big_runs;         // the runs from here on hold small_run frames each: the launch ends on short runs,
    long long small_run;        //   so the warps finish within a few frames of each other (run_range())
    int D;                      // decoded channels per frame (coefficient rows)
    int C;                      // output channels (pcm row width); == D without a channel mapping
    int npairs;                 // kModeDirect: channel pairs per frame
    int nstreams;               // kModeGroup: synthesis warps per group (coupled streams + pairs of mono streams)
    int groups_per_cta;         // kModeGroup: groups a CTA holds (groups_per_cta())
    int store_warps;            // kModeGroup: store warps per group (group_store_warps())
    int store_warps_cta;        // kModeGroup: store warps per CTA (group_store_warps_cta())
    int store_threads;          // kModeGroup: threads of a group's store warps in the store pass (group_store_threads())
    int store_shape;            // kModeGroup: loop shape of the store pass (0: 2 x LDS.64, 1: 4 x LDS.32, 2: with silent channels)
    int lone_pairs;             // kModeGroup: the warp of a lone mono stream takes two consecutive frames as its two channels
    int halo_lm_shift;          // 3 - LM of the halo frame
    int halo_transient;         // flag(s) of the halo frame: bit s = stream s (bit 0 for everybody if !flag_per_stream)
    int flag_stride;            // bytes between the flag records of consecutive frames
    int flag_per_stream;        // 0: every stream reads column 0; 1: stream / pair s reads column s
    StreamDesc streams[kMaxGroupStreams];
    uint16_t chan_src[kMaxChannels + 1];   // kModeGroup: output channel c <- (stream slot << 1 | sub), 0xffff = silent
};

// One clt_mdct_backward call (mdct.c:267): device pointers.
struct MdctCall {
    const float *in;   // N2 coefficients at `stride`
    float *out;        // N2 + 60 floats, [0,60) read (previous raw tail), all written
    int shift;
    int stride;
    int ifft_only;     // 1: `in`/`out` are N4 interleaved complex values; only opus_ifft (kiss_fft.c:696) is run
};

// ---- post stage (celt_post_kernels.cu): comb filter + de-emphasis ----------------------------
constexpr int kPostHist = 1026;   // COMBFILTER_MAXPERIOD + 2 filtered samples of history per channel (celt.h:187)

// Side information of one frame of one stream: the arguments of the two comb_filter calls at
// celt_decoder_clean.c:660-669.  Same layout as nq_celt_post_frame (include/nq_celt_synth.h).
struct PostFrame {
    int32_t N;           // samples per channel of this frame, 120 << LM
    int32_t pitch[3];    // postfilter_period_old, postfilter_period, postfilter_pitch (new)
    float gain[3];
    int32_t tapset[3];
};

// One warp's work: `nframes` consecutive frames of 1 or 2 adjacent output channels.
struct PostJob {
    long long sample0;   // first sample (per channel) of frame0 inside the pcm buffer
    int frame0, nframes;
    int ch0, nch;        // output channel(s) ch0 .. ch0+nch-1
    int stream_col;      // column of the stream inside a frame's side-info record
    int state_row;       // row of channel ch0 in the state arrays (decoded-channel order)
    int reset;           // 1: start from a reset decoder (zero history and memory), ignore *_in
    int write_state;     // 1: leave history / memory in *_out (the piece that ends the batch)
};

struct PostParams {
    float *pcm;              // [nsamples][C]: celt_sig in, PCM out (in place)
    const PostFrame *frames; // [nframes][frame_stride]
    const PostJob *jobs;
    const float *window;     // window120
    const float *hist_in;    // [rows][kPostHist] or nullptr
    const float *mem_in;     // [rows] or nullptr
    float *hist_out;
    float *mem_out;
    int C;
    int frame_stride;        // streams
};

// Frames [*f0, *f1) of run `run`: big_runs runs of frames_per_run frames, then runs of small_run frames.
__host__ __device__ inline void run_range(const SynthParams &p, long long run, long long *f0, long long *f1)
{
    long long a, len;
    if (run < p.big_runs) {
        a = run * p.frames_per_run;
        len = p.frames_per_run;
    } else {
        a = p.big_runs * p.frames_per_run + (run - p.big_runs) * p.small_run;
        len = p.small_run;
    }
    *f0 = a;
    *f1 = a + len < p.nframes ? a + len : p.nframes;
}

size_t post_kernel_smem_bytes();
cudaError_t prepare_post_kernel();
cudaError_t launch_post(const PostParams &p, int njobs, cudaStream_t stream);

size_t fast_kernel_smem_bytes();
int synth_mode(int D, int C, int nstreams, bool identity_map, bool one_decoder);
int groups_per_cta(int nstreams);
int group_store_warps(int nstreams);
int group_store_warps_cta(int nstreams);
int group_store_threads(int C, int nstreams);
cudaError_t launch_synth(const SynthParams &p, int mode, int num_sms, cudaStream_t stream, int *launched_ctas);
cudaError_t launch_mdct_generic(const MdctCall *d_calls, int ncalls, const GenericTables *d_tables, cudaStream_t stream);
cudaError_t prepare_kernels();

}  // namespace nq
