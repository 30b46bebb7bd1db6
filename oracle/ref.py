"""TEST INFRASTRUCTURE -- ctypes view of oracle/_ref/libnq_ref.so.

That library is the UNMODIFIED reference (dafx/libnyquist, bundled Opus float
build) compiled in place by oracle/Makefile; see oracle/ref_harness.c for the
reference file:line each entry wraps.  Only tests/, __graft_entry__.smoke()
and bench.py (cpu_baseline leg, --impl reference) may import this module; the
product package libnyquist_b200 never does.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libnq_ref.so")

FRAME = 960
HALF_OVL = 60

_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{LIB_PATH} missing: run `make -C oracle ref` where /root/reference exists")
        L = C.CDLL(LIB_PATH)
        fp = C.POINTER(C.c_float)
        L.nqref_opus_ifft.argtypes = [C.c_int, fp, fp]
        L.nqref_clt_mdct_backward.argtypes = [fp, fp, C.c_int, C.c_int]
        L.nqref_compute_inv_mdcts.argtypes = [C.c_int, fp, C.POINTER(fp), C.c_int, C.c_int]
        L.nqref_synth_batch.argtypes = [fp, C.c_void_p, fp, fp, fp, C.c_long, C.c_int]
        L.nqref_synth_batch_mt.argtypes = [fp, C.c_void_p, fp, fp, fp, C.c_long, C.c_int, C.c_int]
        L.nqref_synth_batch_mt.restype = C.c_double
        L.nqref_decode_memory.argtypes = [C.c_void_p, C.c_size_t, fp, C.c_long, C.POINTER(C.c_int), C.c_int]
        L.nqref_decode_memory.restype = C.c_long
        L.nqref_record_count.restype = C.c_long
        L.nqref_record_floats.restype = C.c_size_t
        L.nqref_record_copy.argtypes = [fp]
        L.nqref_comb_floats.restype = C.c_size_t
        L.nqref_comb_copy.argtypes = [fp]
        L.nqref_last_layout.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p]
        L.nqref_encode_surround.argtypes = [fp, C.c_long, C.c_int, C.c_int, C.c_void_p, C.c_long]
        L.nqref_encode_surround.restype = C.c_long
        L.nqref_patch_output_gain.argtypes = [C.c_void_p, C.c_long, C.c_int]
        L.nqref_comb_filter.argtypes = [fp, fp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int]
        L.nqref_deemphasis.argtypes = [C.POINTER(fp), fp, C.c_int, C.c_int, fp]
        _lib = L
    return _lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None else None


def tables():
    """The reference's static tables (static_modes_float.h:9,99,343-417,477)."""
    L = lib()
    out = dict(
        window120=np.zeros(120, np.float32),
        trig481=np.zeros(481, np.float32),
        twiddles480=np.zeros(960, np.float32),
        bitrev480=np.zeros(480, np.int16),
        bitrev240=np.zeros(240, np.int16),
        bitrev120=np.zeros(120, np.int16),
        bitrev60=np.zeros(60, np.int16),
    )
    L.nqref_tables(*[v.ctypes.data_as(C.c_void_p) for v in out.values()])
    return out


def opus_ifft(x_ri: np.ndarray, shift: int) -> np.ndarray:
    """kiss_fft.c:696 with the static state mode->mdct.kfft[shift]; interleaved re/im."""
    x = np.ascontiguousarray(x_ri, np.float32)
    assert x.size == 2 * (480 >> shift)
    y = np.zeros_like(x)
    lib().nqref_opus_ifft(shift, _fp(x), _fp(y))
    return y


def clt_mdct_backward(inp: np.ndarray, out: np.ndarray, shift: int, stride: int) -> None:
    """mdct.c:267.  `out` (>= N2+60 floats) is read-modify-written in place."""
    assert inp.dtype == np.float32 and out.dtype == np.float32
    assert inp.flags.c_contiguous and out.flags.c_contiguous
    lib().nqref_clt_mdct_backward(_fp(inp), _fp(out), shift, stride)


def synth_batch(coef, transient, tail_in=None, nthreads=1):
    """compute_inv_mdcts (celt_decoder_clean.c:264) over a batch of LM=3 frames.

    coef [nframes][C][960] f32, transient [nframes] u8, tail_in [C][60] or None.
    Returns (pcm [nframes*960][C], tail_out [C][60], seconds).
    """
    coef = np.ascontiguousarray(coef, np.float32)
    nframes, Cn, n = coef.shape
    assert n == FRAME
    tr = np.ascontiguousarray(transient, np.uint8)
    assert tr.shape == (nframes,)
    ti = None if tail_in is None else np.ascontiguousarray(tail_in, np.float32)
    pcm = np.zeros((nframes * FRAME, Cn), np.float32)
    tail = np.zeros((Cn, HALF_OVL), np.float32)
    sec = lib().nqref_synth_batch_mt(_fp(coef), tr.ctypes.data_as(C.c_void_p), _fp(ti),
                                     _fp(pcm), _fp(tail), nframes, Cn, int(nthreads))
    return pcm, tail, sec


def decode_file(path: str, record: bool = False):
    """Whole-file decode through opusfile, as src/OpusDecoder.cpp:57-119 does.

    Returns (pcm [nsamples][channels] f32, records or None).  Each record is
    a dict(nch, shift, B, coef [nch][n], out [nch][n]) captured at the
    inverse-MDCT call sites (celt_decoder_clean.c:290,298,309).
    """
    L = lib()
    data = np.fromfile(path, np.uint8)
    ch = C.c_int(0)
    n = L.nqref_decode_memory(data.ctypes.data_as(C.c_void_p), data.size, None, 0, C.byref(ch), 0)
    if n < 0:
        raise RuntimeError(f"reference could not decode {path}: {n}")
    pcm = np.zeros((n, ch.value), np.float32)
    n2 = L.nqref_decode_memory(data.ctypes.data_as(C.c_void_p), data.size, _fp(pcm), pcm.size,
                               C.byref(ch), 1 if record else 0)
    assert n2 == n
    recs = None
    if record:
        flat = np.zeros(L.nqref_record_floats(), np.float32)
        L.nqref_record_copy(_fp(flat))
        comb = np.zeros(L.nqref_comb_floats(), np.float32)
        L.nqref_comb_copy(_fp(comb))
        comb = comb.reshape(-1, 8)
        L.nqref_record_free()
        recs = []
        cpos = 0
        p = 0
        while p < flat.size:
            nch, shift, B, nn = flat[p:p + 4].view(np.int32)
            p += 4
            coef = flat[p:p + nch * nn].reshape(nch, nn); p += nch * nn
            out = flat[p:p + nch * nn].reshape(nch, nn); p += nch * nn
            # comb_filter calls of this synthesis record: per channel 1 call (LM = 0) or 2 (LM > 0)
            # (celt_decoder_clean.c:663-669); LM follows from the record's frame length.
            ncalls = nch * (1 if nn == 120 else 2)
            calls = comb[cpos:cpos + ncalls]; cpos += ncalls
            recs.append(dict(nch=int(nch), shift=int(shift), B=int(B), coef=coef, out=out, comb=calls,
                             stream=int(calls[0][7])))   # multistream: index of the CELT decoder that produced it
        assert cpos == len(comb), (cpos, len(comb))
    return pcm, recs


def header_info():
    """(pre_skip, output_gain) of the file decoded last (opusfile OpusHead)."""
    L = lib()
    return int(L.nqref_last_pre_skip()), int(L.nqref_last_output_gain())


def layout_info():
    """(channels, streams, coupled_streams, mapping) of the file decoded last (OpusHead)."""
    st, cp = C.c_int(0), C.c_int(0)
    mp = np.zeros(256, np.uint8)
    ch = lib().nqref_last_layout(C.byref(st), C.byref(cp), mp.ctypes.data_as(C.c_void_p))
    return ch, st.value, cp.value, [int(v) for v in mp[:ch]]


def decode_bytes(data: bytes, record: bool = False):
    """decode_file for an in-memory Ogg Opus file."""
    import tempfile
    with tempfile.NamedTemporaryFile(suffix=".opus") as f:
        f.write(data)
        f.flush()
        return decode_file(f.name, record)


def encode_surround(pcm: np.ndarray, bitrate: int = 512000) -> bytes:
    """The reference's own surround ENCODER (opus_multistream_encoder.c:557, CELT-only) + libogg:
    pcm [nsamples][channels] float32 in [-1, 1], nsamples a multiple of 960 -> an Ogg Opus file."""
    pcm = np.ascontiguousarray(pcm, np.float32)
    n, ch = pcm.shape
    assert n % FRAME == 0
    cap = 1 << 24
    out = np.zeros(cap, np.uint8)
    got = lib().nqref_encode_surround(_fp(pcm), n, ch, int(bitrate), out.ctypes.data_as(C.c_void_p), cap)
    if got < 0:
        raise RuntimeError(f"reference encoder failed: {got}")
    return out[:got].tobytes()


MODE_SILK_ONLY, MODE_HYBRID, MODE_CELT_ONLY = 1000, 1001, 1002   # opus_private.h:90-92


def encode_mode_switch(pcm: np.ndarray, mode_a: int, mode_b: int, switch_frame: int, bitrate: int = 40000) -> bytes:
    """Like encode_forced_mode, switching from mode_a to mode_b at frame `switch_frame`."""
    pcm = np.ascontiguousarray(pcm, np.float32)
    n, ch = pcm.shape
    assert n % FRAME == 0 and ch in (1, 2)
    cap = 1 << 22
    out = np.zeros(cap, np.uint8)
    L = lib()
    L.nqref_encode_mode_switch.restype = C.c_long
    L.nqref_encode_mode_switch.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_long]
    got = L.nqref_encode_mode_switch(pcm.ctypes.data_as(C.c_void_p), n, ch, int(bitrate), int(mode_a), int(mode_b),
                                     int(switch_frame), out.ctypes.data_as(C.c_void_p), cap)
    if got < 0:
        raise RuntimeError(f"reference encoder failed: {got}")
    return out[:got].tobytes()


def encode_mode_schedule(pcm: np.ndarray, first_mode: int, schedule, bitrate: int = 40000) -> bytes:
    """A file that walks through several coding modes: schedule = [(frame, mode), ...]."""
    pcm = np.ascontiguousarray(pcm, np.float32)
    n, ch = pcm.shape
    assert n % FRAME == 0 and ch in (1, 2)
    cap = 1 << 22
    out = np.zeros(cap, np.uint8)
    frames = np.asarray([f for f, _ in schedule], np.int32)
    modes = np.asarray([m for _, m in schedule], np.int32)
    L = lib()
    L.nqref_encode_mode_schedule.restype = C.c_long
    L.nqref_encode_mode_schedule.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                             C.c_void_p, C.c_long]
    got = L.nqref_encode_mode_schedule(pcm.ctypes.data_as(C.c_void_p), n, ch, int(bitrate), int(first_mode),
                                       frames.ctypes.data_as(C.c_void_p), modes.ctypes.data_as(C.c_void_p), len(schedule),
                                       out.ctypes.data_as(C.c_void_p), cap)
    if got < 0:
        raise RuntimeError(f"reference encoder failed: {got}")
    return out[:got].tobytes()


def encode_forced_mode(pcm: np.ndarray, mode: int, bitrate: int = 32000) -> bytes:
    """The reference's own encoder (VOIP application) forced into one coding mode with the private
    OPUS_SET_FORCE_MODE ctl: test files for the SILK / hybrid branches of opus_decode_frame
    (opus_decoder_clean.c:340-600).  pcm [nsamples][1 or 2] float32, nsamples a multiple of 960."""
    pcm = np.ascontiguousarray(pcm, np.float32)
    n, ch = pcm.shape
    assert n % FRAME == 0 and ch in (1, 2)
    cap = 1 << 22
    out = np.zeros(cap, np.uint8)
    L = lib()
    L.nqref_encode_forced_mode.restype = C.c_long
    L.nqref_encode_forced_mode.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_long]
    got = L.nqref_encode_forced_mode(pcm.ctypes.data_as(C.c_void_p), n, ch, int(bitrate), int(mode),
                                     out.ctypes.data_as(C.c_void_p), cap)
    if got < 0:
        raise RuntimeError(f"reference encoder failed: {got}")
    return out[:got].tobytes()


def with_output_gain(data: bytes, gain_q8: int) -> bytes:
    """The same Ogg Opus file with OpusHead.output_gain (Q7.8 dB) rewritten and the page CRC fixed."""
    a = np.frombuffer(data, np.uint8).copy()
    rc = lib().nqref_patch_output_gain(a.ctypes.data_as(C.c_void_p), a.size, int(gain_q8))
    if rc != 0:
        raise RuntimeError(f"not an Ogg Opus file ({rc})")
    return a.tobytes()


def comb_filter(buf: np.ndarray, start: int, T0, T1, N, g0, g1, tapset0, tapset1) -> None:
    """celt.c:114 in place on buf[start:start+N] (buf[:start] is the history, >= max(T)+2 long)."""
    assert buf.dtype == np.float32 and buf.flags.c_contiguous and start >= max(T0, T1) + 2
    p = C.cast(buf.ctypes.data + 4 * start, C.POINTER(C.c_float))
    lib().nqref_comb_filter(p, p, int(T0), int(T1), int(N), float(g0), float(g1), int(tapset0), int(tapset1))


def deemphasis(x: np.ndarray, mem: np.ndarray) -> np.ndarray:
    """celt_decoder_clean.c:192: x [C][N] celt_sig -> pcm [N][C]; mem [C] updated in place."""
    Cn, N = x.shape
    x = np.ascontiguousarray(x, np.float32)
    fp = C.POINTER(C.c_float)
    rows = (fp * Cn)(*[C.cast(x.ctypes.data + 4 * N * c, fp) for c in range(Cn)])
    pcm = np.zeros((N, Cn), np.float32)
    assert mem.dtype == np.float32 and mem.shape == (Cn,)
    lib().nqref_deemphasis(rows, _fp(pcm), N, Cn, _fp(mem))
    return pcm


# ---- the reference's own end-to-end file decode: nqr::NyquistIO::Load -------------------------
LOAD_LIB_PATH = os.path.join(HERE, "_ref", "libnyquist_ref.so")
_load_lib = None


def load_available() -> bool:
    return os.path.exists(LOAD_LIB_PATH)


def nyquist_load(path: str):
    """nqr::NyquistIO::Load of the UNMODIFIED reference (Common.cpp:36, OpusDecoder.cpp:39-183).
    Returns (samples [n][channels] f32, sample_rate, seconds spent inside Load incl. the copy out)."""
    import time
    global _load_lib
    if _load_lib is None:
        L = C.CDLL(LOAD_LIB_PATH)
        L.nqref_load.argtypes = [C.c_char_p, C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.c_size_t),
                                 C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.nqref_free.argtypes = [C.POINTER(C.c_float)]
        _load_lib = L
    p = C.POINTER(C.c_float)()
    n, ch, sr = C.c_size_t(0), C.c_int(0), C.c_int(0)
    t0 = time.perf_counter()
    rc = _load_lib.nqref_load(path.encode(), C.byref(p), C.byref(n), C.byref(ch), C.byref(sr))
    dt = time.perf_counter() - t0
    if rc != 0:
        raise RuntimeError(f"reference NyquistIO::Load failed for {path}")
    a = np.ctypeslib.as_array(p, shape=(n.value,)).copy()
    _load_lib.nqref_free(p)
    return a.reshape(-1, ch.value), sr.value, dt
