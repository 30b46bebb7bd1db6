"""libnyquist_b200 -- B200-native CELT synthesis stage (inverse MDCT + TDAC
overlap-add + channel interleave) behind the reference's own interfaces.

This package is a thin host-side mirror (ctypes) of the C ABI declared in
include/nq_celt_synth.h.  The names follow the reference's functions for this
path (paths relative to /root/reference/third_party/opus/celt/):

    compute_inv_mdcts        celt_decoder_clean.c:264-312
    clt_mdct_backward        mdct.c:267-379
    clt_mdct_backward_B1_C2  mdct.c:258-265

plus the batched phase-2 entry (`CeltSynth.synth_batch*`).  There is NO CPU
path: importing works anywhere (so host logic can be tested), but every
compute call needs the CUDA library and a B200 and raises otherwise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .build import LIB_PATH, build_library
from .sharding import shard_range, shard_plan  # noqa: F401

FRAME = 960
HALF_OVERLAP = 60
OVERLAP = 120
MDCT_N = 1920

NQ_OK = 0
POST_HISTORY = 1026

# nq_celt_post_frame (include/nq_celt_synth.h): the arguments of the two comb_filter calls of a
# frame, celt_decoder_clean.c:660-669
POST_FRAME_DTYPE = np.dtype([("N", np.int32), ("pitch", np.int32, 3), ("gain", np.float32, 3),
                             ("tapset", np.int32, 3)])

EXPORTED_SYMBOLS = [
    "nq_celt_ctx_create", "nq_celt_ctx_destroy", "nq_celt_strerror", "nq_celt_last_error",
    "nq_celt_device_count", "nq_celt_launch_count", "nq_celt_host_alloc", "nq_celt_host_free",
    "nq_celt_synth_batch_device", "nq_celt_synth_batch_device_ms", "nq_celt_synth_batch_host",
    "nq_celt_synth_batch_host_multi", "nq_celt_multi_release", "nq_celt_post_batch_device", "nq_celt_post_segments_device", "nq_celt_decode_batch_host",
    "nq_celt_sink_create", "nq_celt_sink_destroy", "nq_celt_sink_last_error", "nq_celt_sink_push", "nq_celt_sink_push_at", "nq_celt_sink_side_count", "nq_celt_sink_side_get",
    "nq_celt_sink_pending_frames", "nq_celt_sink_pending_samples", "nq_celt_sink_flush", "nq_celt_sink_reset", "nq_celt_sink_reset_stream", "nq_celt_sink_set_destination",
    "nq_celt_sink_flush_pinned", "nq_celt_sink_flush_many", "nq_celt_sink_begin_upload", "nq_celt_sink_trim_pool", "nq_celt_ctx_device", "nq_celt_ctx_stream", "nq_celt_sink_attach", "nq_celt_sink_finish",
    "nq_clt_mdct_backward", "nq_clt_mdct_backward_B1_C2", "nq_celt_mdct_backward_host",
    "nq_compute_inv_mdcts", "nq_opus_ifft_host", "processMDCTCuda", "processMDCTCudaB1C2", "cleanupCudaBuffers",
    "printCudaVersion", "nq_celt_debug_tables", "nq_celt_debug_plan", "nq_celt_debug_runs",
]


class NqError(RuntimeError):
    def __init__(self, code: int, detail: str = ""):
        self.code = code
        super().__init__(f"nq_celt error {code}: {detail}")


_lib = None


def load_library():
    """Loads lib/libnq_celt_b200.so.  Raises (never falls back) if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not built: run `python -m libnyquist_b200.build` (needs nvcc). "
            "libnyquist_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    fp, vp = C.POINTER(C.c_float), C.c_void_p
    L.nq_celt_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.nq_celt_ctx_destroy.argtypes = [vp]
    L.nq_celt_strerror.argtypes = [C.c_int]
    L.nq_celt_strerror.restype = C.c_char_p
    L.nq_celt_last_error.argtypes = [vp]
    L.nq_celt_last_error.restype = C.c_char_p
    L.nq_celt_launch_count.argtypes = [vp]
    L.nq_celt_launch_count.restype = C.c_longlong
    L.nq_celt_host_alloc.argtypes = [C.c_size_t]
    L.nq_celt_host_alloc.restype = vp
    L.nq_celt_host_free.argtypes = [vp]
    L.nq_celt_synth_batch_device.argtypes = [vp, vp, vp, vp, vp, C.c_int, vp, vp, C.c_int64, C.c_int, vp]
    L.nq_celt_synth_batch_host.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int64, C.c_int]
    L.nq_celt_synth_batch_device_ms.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, C.c_int64, C.c_int, C.c_int,
                                                C.c_int, vp, vp]
    L.nq_celt_post_batch_device.argtypes = [vp, vp, vp, vp, vp, vp, vp, C.c_int64, C.c_int, C.c_int, C.c_int, vp, vp]
    L.nq_celt_post_segments_device.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int, vp, vp]
    L.nq_celt_decode_batch_host.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, C.c_int64, C.c_int, C.c_int,
                                            C.c_int, vp]
    L.nq_celt_sink_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, vp]
    L.nq_celt_sink_destroy.argtypes = [vp]
    L.nq_celt_sink_last_error.argtypes = [vp]
    L.nq_celt_sink_last_error.restype = C.c_char_p
    L.nq_celt_sink_push.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, vp]
    L.nq_celt_sink_pending_frames.argtypes = [vp]
    L.nq_celt_sink_pending_frames.restype = C.c_int64
    L.nq_celt_sink_pending_samples.argtypes = [vp]
    L.nq_celt_sink_pending_samples.restype = C.c_int64
    L.nq_celt_sink_flush.argtypes = [vp, vp, vp, C.c_int64, C.POINTER(C.c_int64)]
    L.nq_celt_sink_reset.argtypes = [vp]
    L.nq_celt_sink_attach.argtypes = [vp, vp, vp, C.c_int64, C.c_int64]
    L.nq_celt_sink_finish.argtypes = [vp, C.POINTER(C.c_int64)]
    L.nq_celt_synth_batch_host_multi.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, C.c_int64, C.c_int]
    L.nq_celt_multi_release.restype = None
    L.nq_clt_mdct_backward.argtypes = [vp, fp, fp, fp, C.c_int, C.c_int, C.c_int]
    L.nq_clt_mdct_backward.restype = None
    L.nq_clt_mdct_backward_B1_C2.argtypes = [vp, C.POINTER(fp), C.POINTER(fp), fp, C.c_int, C.c_int, C.c_int]
    L.nq_clt_mdct_backward_B1_C2.restype = None
    L.nq_celt_mdct_backward_host.argtypes = [vp, C.POINTER(fp), C.POINTER(fp), C.c_int, C.c_int, C.c_int]
    L.nq_compute_inv_mdcts.argtypes = [vp, C.c_int, fp, C.POINTER(fp), C.c_int, C.c_int]
    L.nq_opus_ifft_host.argtypes = [vp, C.c_int, fp, fp, C.c_int]
    L.processMDCTCuda.argtypes = [fp, fp, fp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, fp]
    L.processMDCTCuda.restype = None
    L.processMDCTCudaB1C2.argtypes = [C.POINTER(fp), C.POINTER(fp), fp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, fp]
    L.processMDCTCudaB1C2.restype = None
    L.cleanupCudaBuffers.restype = None
    L.printCudaVersion.restype = None
    L.nq_celt_debug_tables.argtypes = [fp, fp, fp, fp]
    L.nq_celt_debug_plan.argtypes = [C.c_int, C.c_int, C.c_int, vp, C.c_int64, C.c_int, C.POINTER(C.c_int64)]
    _lib = L
    return L


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _vp(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _f32c(a, what):
    if not isinstance(a, np.ndarray) or a.dtype != np.float32 or not a.flags.c_contiguous:
        raise TypeError(f"{what} must be a C-contiguous float32 numpy array")
    return a


def debug_tables():
    """Host-built tables of the fast kernel (pure host code, no GPU needed)."""
    L = load_library()
    t_long = np.zeros((16, 31, 2), np.float32)
    t_short = np.zeros((2, 30, 2), np.float32)
    window = np.zeros(120, np.float32)
    trig = np.zeros(481, np.float32)
    L.nq_celt_debug_tables(_fp(t_long), _fp(t_short), _fp(window), _fp(trig))
    return dict(t_long=t_long, t_short=t_short, window=window, trig=trig)


PLAN_FIELDS = ("mode", "warps_per_group", "groups_per_cta", "store_threads", "store_shape", "paired_mono",
               "frames_per_run", "runs", "post_ctas", "post_ctas_two_channel", "decoded_channels", "identity")
MODE_STEREO, MODE_GROUP, MODE_DIRECT, MODE_MONO = 0, 1, 2, 4


def debug_plan(channels, streams=0, coupled_streams=0, mapping=None, nframes=1_000_000, num_sms=148):
    """How a batch with this layout would be launched (pure host code, no GPU needed)."""
    L = load_library()
    out = (C.c_int64 * 12)()
    mp = None if mapping is None else np.ascontiguousarray(mapping, np.uint8)
    rc = L.nq_celt_debug_plan(int(channels), int(streams), int(coupled_streams), _vp(mp), int(nframes), int(num_sms), out)
    if rc != NQ_OK:
        raise NqError(rc, "nq_celt_debug_plan")
    return dict(zip(PLAN_FIELDS, [int(v) for v in out]))


def debug_runs(channels, nframes, num_sms=148):
    """First frame of every run a batch is cut into, plus nframes at the end (pure host code)."""
    L = load_library()
    L.nq_celt_debug_runs.argtypes = [C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
    n = C.c_int64(0)
    rc = L.nq_celt_debug_runs(int(channels), int(nframes), int(num_sms), None, 0, C.byref(n))
    if rc != NQ_OK:
        raise NqError(rc, "nq_celt_debug_runs")
    first = np.zeros(n.value + 1, np.int64)
    rc = L.nq_celt_debug_runs(int(channels), int(nframes), int(num_sms), _vp(first), first.size, C.byref(n))
    if rc != NQ_OK:
        raise NqError(rc, "nq_celt_debug_runs")
    return first


# ---- reference-shaped single calls (host buffers, synchronous) ------------
def clt_mdct_backward(inp: np.ndarray, out: np.ndarray, shift: int, stride: int) -> None:
    """mdct.c:267 semantics: `out` [0,60) is the previous raw tail on entry and
    [0, N2+60) is written; `inp` is left untouched."""
    L = load_library()
    _f32c(inp, "inp"); _f32c(out, "out")
    N2 = (MDCT_N >> shift) >> 1
    if inp.size < (N2 - 1) * stride + 1 or out.size < N2 + HALF_OVERLAP:
        raise ValueError("buffer too small for this (shift, stride)")
    L.nq_clt_mdct_backward(None, _fp(inp), _fp(out), None, OVERLAP, shift, stride)


def clt_mdct_backward_B1_C2(inp, out, shift: int, stride: int) -> None:
    """mdct.c:258: the two-channel convenience wrapper (one GPU launch)."""
    L = load_library()
    fp = C.POINTER(C.c_float)
    ins = (fp * 2)(_fp(_f32c(inp[0], "inp[0]")), _fp(_f32c(inp[1], "inp[1]")))
    outs = (fp * 2)(_fp(_f32c(out[0], "out[0]")), _fp(_f32c(out[1], "out[1]")))
    L.nq_clt_mdct_backward_B1_C2(None, ins, outs, None, OVERLAP, shift, stride)


class CeltSynth:
    """One per-device context (nq_celt_ctx)."""

    def __init__(self, device: int = 0):
        self._L = load_library()
        h = C.c_void_p()
        rc = self._L.nq_celt_ctx_create(int(device), C.byref(h))
        if rc != NQ_OK:
            raise NqError(rc, self._L.nq_celt_strerror(rc).decode() +
                          " (no CUDA device / not a B200? there is no CPU fallback)")
        self._h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            self._L.nq_celt_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != NQ_OK:
            raise NqError(rc, self._L.nq_celt_last_error(self._h).decode())

    @property
    def launch_count(self) -> int:
        return int(self._L.nq_celt_launch_count(self._h))

    # -- compute_inv_mdcts, celt_decoder_clean.c:264 ------------------------
    def compute_inv_mdcts(self, shortBlocks: int, X: np.ndarray, out_mem, C_: int, LM: int) -> None:
        """out_mem: list of C float32 arrays (out_syn[c]), each >= N*B+60 long;
        [0,60) = previous raw tail on entry; [0, N*B+60) written."""
        _f32c(X, "X")
        fp = C.POINTER(C.c_float)
        outs = (fp * C_)(*[_fp(_f32c(o, "out_mem[c]")) for o in out_mem])
        self._check(self._L.nq_compute_inv_mdcts(self._h, int(shortBlocks), _fp(X), outs, int(C_), int(LM)))

    # -- opus_ifft, kiss_fft.c:696 -------------------------------------------
    def opus_ifft(self, x_ri: np.ndarray, shift: int) -> np.ndarray:
        """x_ri [count][2*N4] (or [2*N4]) interleaved re/im float32; N4 = 480 >> shift."""
        x = np.ascontiguousarray(x_ri, np.float32)
        n = 2 * (480 >> shift)
        if x.size % n:
            raise ValueError("size must be a multiple of 2*N4")
        y = np.empty_like(x)
        self._check(self._L.nq_opus_ifft_host(self._h, int(shift), _fp(x), _fp(y), x.size // n))
        return y

    # -- batched phase 2, host buffers --------------------------------------
    def synth_batch(self, coef: np.ndarray, transient: np.ndarray, tail_in=None, out=None):
        """coef [nframes][C][960] f32, transient [nframes] u8, tail_in [C][60] or None.
        Returns (pcm [nframes*960][C] f32, tail_out [C][60] f32)."""
        _f32c(coef, "coef")
        if coef.ndim != 3 or coef.shape[2] != FRAME:
            raise ValueError("coef must be [nframes][C][960]")
        nframes, Cn, _ = coef.shape
        tr = np.ascontiguousarray(transient, np.uint8)
        if tr.shape != (nframes,):
            raise ValueError("transient must be [nframes]")
        ti = None if tail_in is None else np.ascontiguousarray(tail_in, np.float32)
        if ti is not None and ti.shape != (Cn, HALF_OVERLAP):
            raise ValueError("tail_in must be [C][60]")
        pcm = np.empty((nframes * FRAME, Cn), np.float32) if out is None else _f32c(out, "out")
        tail = np.zeros((Cn, HALF_OVERLAP), np.float32)
        self._check(self._L.nq_celt_synth_batch_host(self._h, _vp(coef), _vp(tr), _vp(ti), _vp(pcm), _vp(tail),
                                                     nframes, Cn))
        return pcm, tail

    # -- batched phase 2, raw pointers (device or pinned host) ---------------
    def synth_batch_device_ptr(self, coef_ptr, transient_ptr, tail_in_ptr, halo_ptr, halo_transient,
                               pcm_ptr, tail_out_ptr, nframes, Cn, stream=0):
        self._check(self._L.nq_celt_synth_batch_device(
            self._h, C.c_void_p(coef_ptr), C.c_void_p(transient_ptr), C.c_void_p(tail_in_ptr or None),
            C.c_void_p(halo_ptr or None), int(halo_transient), C.c_void_p(pcm_ptr),
            C.c_void_p(tail_out_ptr or None), int(nframes), int(Cn), C.c_void_p(stream or None)))

    def synth_batch_host_ptr(self, coef_ptr, transient_ptr, tail_in_ptr, pcm_ptr, tail_out_ptr, nframes, Cn):
        self._check(self._L.nq_celt_synth_batch_host(
            self._h, C.c_void_p(coef_ptr), C.c_void_p(transient_ptr), C.c_void_p(tail_in_ptr or None),
            C.c_void_p(pcm_ptr), C.c_void_p(tail_out_ptr or None), int(nframes), int(Cn)))

    # -- torch convenience: device tensors on this context's device ----------
    def synth_batch_torch(self, coef, transient, tail_in=None, halo_coef=None, halo_transient=0,
                          out=None, want_tail=True, stream=None):
        """coef: cuda float32 [nframes][C][960]; transient: cuda uint8 [nframes].
        Enqueues on torch's current stream (or `stream`); returns (pcm, tail_out)."""
        import torch
        assert coef.is_cuda and coef.dtype == torch.float32 and coef.is_contiguous()
        assert transient.is_cuda and transient.dtype == torch.uint8 and transient.is_contiguous()
        nframes, Cn, n = coef.shape
        assert n == FRAME and transient.shape == (nframes,)
        pcm = out if out is not None else torch.empty((nframes * FRAME, Cn), dtype=torch.float32, device=coef.device)
        tail = torch.empty((Cn, HALF_OVERLAP), dtype=torch.float32, device=coef.device) if want_tail else None
        st = (stream if stream is not None else torch.cuda.current_stream(coef.device)).cuda_stream
        if st == 0:
            st = 1   # torch's default stream is the legacy default stream: cudaStreamLegacy (NULL would mean ctx's own stream)
        self.synth_batch_device_ptr(coef.data_ptr(), transient.data_ptr(),
                                    0 if tail_in is None else tail_in.data_ptr(),
                                    0 if halo_coef is None else halo_coef.data_ptr(), halo_transient,
                                    pcm.data_ptr(), 0 if tail is None else tail.data_ptr(), nframes, Cn, st)
        return pcm, tail


    def synth_batch_ms_torch(self, coef, transient, streams: int, coupled_streams: int, mapping,
                             tail_in=None, halo_coef=None, halo_transient=None, out=None, want_tail=True,
                             stream=None, frame_offset=None):
        """Opus multistream batch (opus_multistream_decoder.c:110, :237-299): coef cuda f32
        [nframes][streams+coupled][960], transient cuda u8 [nframes][streams], mapping: sequence of
        `channels` decoded-channel indices (255 = silent).  Returns (pcm [nframes*960][channels], tail)."""
        import torch
        D = streams + coupled_streams
        assert coef.is_cuda and coef.dtype == torch.float32 and coef.is_contiguous()
        assert transient.is_cuda and transient.dtype == torch.uint8 and transient.is_contiguous()
        nframes = coef.shape[0]
        assert coef.shape == (nframes, D, FRAME) and transient.shape == (nframes, streams)
        mp = None if mapping is None else np.ascontiguousarray(mapping, np.uint8)
        ch = D if mp is None else mp.size
        nsamples = nframes * FRAME
        if frame_offset is not None:   # cuda int64 [nframes + 1]: frames shorter than 20 ms in the batch
            assert frame_offset.is_cuda and frame_offset.dtype == torch.int64 and frame_offset.shape == (nframes + 1,)
            nsamples = int(frame_offset[-1])
        pcm = out if out is not None else torch.empty((nsamples, ch), dtype=torch.float32, device=coef.device)
        tail = torch.empty((D, HALF_OVERLAP), dtype=torch.float32, device=coef.device) if want_tail else None
        st = (stream if stream is not None else torch.cuda.current_stream(coef.device)).cuda_stream
        if st == 0:
            st = 1
        ht = None if halo_transient is None else np.ascontiguousarray(halo_transient, np.uint8)
        self._check(self._L.nq_celt_synth_batch_device_ms(
            self._h, C.c_void_p(coef.data_ptr()), C.c_void_p(transient.data_ptr()),
            C.c_void_p(0 if tail_in is None else tail_in.data_ptr()),
            C.c_void_p(0 if halo_coef is None else halo_coef.data_ptr()), _vp(ht), C.c_void_p(pcm.data_ptr()),
            C.c_void_p(0 if tail is None else tail.data_ptr()),
            C.c_void_p(0 if frame_offset is None else frame_offset.data_ptr()), nframes, ch, int(streams),
            int(coupled_streams), _vp(mp), C.c_void_p(st)))
        return pcm, tail


    # -- post stage: comb_filter x2 + deemphasis, celt_decoder_clean.c:658-670, :723 ----------
    def post_batch_torch(self, pcm, frames, hist_in=None, mem_in=None, streams=1, coupled_streams=None,
                         mapping=None, want_state=True, stream=None):
        """In place on pcm (cuda f32 [nsamples][channels], celt_sig in, PCM out).  frames: numpy
        POST_FRAME_DTYPE [nframes] or [nframes][streams].  Returns (hist_out [D][1026], mem_out [D])."""
        import torch
        assert pcm.is_cuda and pcm.dtype == torch.float32 and pcm.is_contiguous() and pcm.dim() == 2
        ch = pcm.shape[1]
        fr = np.ascontiguousarray(frames, POST_FRAME_DTYPE).reshape(-1, streams)
        if coupled_streams is None:
            coupled_streams = 1 if (mapping is None and ch == 2) else 0
        D = streams + coupled_streams
        assert int(fr["N"][:, 0].sum()) == pcm.shape[0], "sum of frame sizes must equal the sample count"
        mp = None if mapping is None else np.ascontiguousarray(mapping, np.uint8)
        hist = torch.empty((D, POST_HISTORY), dtype=torch.float32, device=pcm.device) if want_state else None
        mem = torch.empty((D,), dtype=torch.float32, device=pcm.device) if want_state else None
        st = (stream if stream is not None else torch.cuda.current_stream(pcm.device)).cuda_stream
        if st == 0:
            st = 1
        ptr = lambda t: C.c_void_p(0 if t is None else t.data_ptr())
        self._check(self._L.nq_celt_post_batch_device(
            self._h, ptr(pcm), _vp(fr), ptr(hist_in), ptr(mem_in), ptr(hist), ptr(mem), fr.shape[0], ch,
            int(streams), int(coupled_streams), _vp(mp), C.c_void_p(st)))
        return hist, mem

    def post_segments_torch(self, pcm, frames, seg_start, streams=1, coupled_streams=None, mapping=None, stream=None):
        """Post stage over a batch of independent segments (files): seg_start = nseg+1 frame indices;
        every segment starts from a reset decoder and gets its own CTA(s).  In place on pcm."""
        import torch
        assert pcm.is_cuda and pcm.dtype == torch.float32 and pcm.is_contiguous() and pcm.dim() == 2
        ch = pcm.shape[1]
        fr = np.ascontiguousarray(frames, POST_FRAME_DTYPE).reshape(-1, streams)
        if coupled_streams is None:
            coupled_streams = 1 if (mapping is None and ch == 2) else 0
        ss = np.ascontiguousarray(seg_start, np.int64)
        mp = None if mapping is None else np.ascontiguousarray(mapping, np.uint8)
        st = (stream if stream is not None else torch.cuda.current_stream(pcm.device)).cuda_stream
        if st == 0:
            st = 1
        self._check(self._L.nq_celt_post_segments_device(self._h, C.c_void_p(pcm.data_ptr()), _vp(fr), _vp(ss), ss.size - 1,
                                                         fr.shape[0], ch, int(streams), int(coupled_streams), _vp(mp),
                                                         C.c_void_p(st)))

    # -- whole phase 2 on host buffers ----------------------------------------------------------
    def decode_batch(self, coef, transient, frames, state=None, streams=1, coupled_streams=None, mapping=None):
        """coef [nframes][D][960] f32, transient [nframes] or [nframes][streams] u8, frames
        POST_FRAME_DTYPE [nframes] or [nframes][streams]; state = (tail [D][60], hist [D][1026],
        mem [D]) or None for a reset decoder.  Returns (pcm [nframes*960][channels], new state)."""
        _f32c(coef, "coef")
        nframes, D, n = coef.shape
        if coupled_streams is None:
            coupled_streams = 1 if (mapping is None and D == 2) else 0
        assert n == FRAME and D == streams + coupled_streams
        mp = None if mapping is None else np.ascontiguousarray(mapping, np.uint8)
        ch = D if mp is None else mp.size
        tr = np.ascontiguousarray(transient, np.uint8).reshape(nframes, -1)
        assert tr.shape[1] == (streams if mp is not None else 1)
        fr = np.ascontiguousarray(frames, POST_FRAME_DTYPE).reshape(nframes, streams)
        ti, hi, mi = (None, None, None) if state is None else [np.ascontiguousarray(a, np.float32) for a in state]
        pcm = np.empty((int(fr["N"][:, 0].sum()), ch), np.float32)
        to = np.zeros((D, HALF_OVERLAP), np.float32)
        ho = np.zeros((D, POST_HISTORY), np.float32)
        mo = np.zeros((D,), np.float32)
        self._check(self._L.nq_celt_decode_batch_host(self._h, _vp(coef), _vp(tr), _vp(fr), _vp(ti), _vp(hi), _vp(mi),
                                                      _vp(pcm), _vp(to), _vp(ho), _vp(mo), nframes, ch, int(streams),
                                                      int(coupled_streams), _vp(mp)))
        return pcm, (to, ho, mo)


class FrameSink:
    """nq_celt_sink: what the restructured celt_decode_with_ec pushes frames into (phase 1) and
    the one call that turns them into PCM (phase 2)."""

    def __init__(self, channels: int, streams: int, coupled_streams: int, mapping):
        self._L = load_library()
        mp = np.ascontiguousarray(mapping, np.uint8)
        assert mp.size == channels
        h = C.c_void_p()
        rc = self._L.nq_celt_sink_create(C.byref(h), channels, streams, coupled_streams, _vp(mp))
        if rc != NQ_OK:
            raise NqError(rc, "nq_celt_sink_create: bad layout")
        self._h, self.channels = h, channels

    def close(self):
        if getattr(self, "_h", None):
            self._L.nq_celt_sink_destroy(self._h)
            self._h = None

    __del__ = close

    def _check(self, rc):
        if rc != NQ_OK:
            raise NqError(rc, self._L.nq_celt_sink_last_error(self._h).decode())

    def push(self, stream: int, freq: np.ndarray, shortBlocks: int, post) -> None:
        """freq [CC][N] float32; post: one POST_FRAME_DTYPE record."""
        _f32c(freq, "freq")
        CC, N = freq.shape
        pf = np.ascontiguousarray(post, POST_FRAME_DTYPE).reshape(1)
        self._check(self._L.nq_celt_sink_push(self._h, int(stream), _vp(freq), CC, N, int(shortBlocks), _vp(pf)))

    @property
    def pending_frames(self) -> int:
        return int(self._L.nq_celt_sink_pending_frames(self._h))

    def flush(self, synth: "CeltSynth") -> np.ndarray:
        n = int(self._L.nq_celt_sink_pending_samples(self._h))
        pcm = np.empty((n, self.channels), np.float32)
        got = C.c_int64(0)
        self._check(self._L.nq_celt_sink_flush(self._h, synth._h, _vp(pcm), n, C.byref(got)))
        assert got.value == n
        return pcm

    def reset(self) -> None:
        self._L.nq_celt_sink_reset(self._h)

    # streaming phase 2 (worker thread per sink, overlaps with the pushes)
    def attach(self, synth: "CeltSynth", dst: np.ndarray, skip_samples: int = 0) -> None:
        """dst: float32 [nsamples][channels], filled in the background as blocks complete."""
        _f32c(dst, "dst")
        assert dst.ndim == 2 and dst.shape[1] == self.channels
        self._dst = dst   # keep alive
        self._check(self._L.nq_celt_sink_attach(self._h, synth._h, _vp(dst), int(skip_samples), dst.shape[0]))

    def finish(self) -> int:
        got = C.c_int64(0)
        self._check(self._L.nq_celt_sink_finish(self._h, C.byref(got)))
        return got.value


def synth_batch_multi_gpu(coef: np.ndarray, transient: np.ndarray, tail_in=None, devices=None):
    """Frames sharded contiguously over the given devices inside ONE process
    (one host thread + context per device, no collective)."""
    L = load_library()
    _f32c(coef, "coef")
    nframes, Cn, _ = coef.shape
    tr = np.ascontiguousarray(transient, np.uint8)
    ti = None if tail_in is None else np.ascontiguousarray(tail_in, np.float32)
    if devices is None:
        devices = list(range(L.nq_celt_device_count()))
    dev = np.asarray(devices, np.int32)
    pcm = np.empty((nframes * FRAME, Cn), np.float32)
    tail = np.zeros((Cn, HALF_OVERLAP), np.float32)
    rc = L.nq_celt_synth_batch_host_multi(_vp(dev), len(devices), _vp(coef), _vp(tr), _vp(ti), _vp(pcm), _vp(tail),
                                          nframes, Cn)
    if rc != NQ_OK:
        raise NqError(rc, L.nq_celt_strerror(rc).decode())
    return pcm, tail
