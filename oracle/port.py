"""TEST INFRASTRUCTURE -- ctypes view of the plain-C oracle (oracle/celt_synth_oracle.c).

The oracle restates, in float arithmetic evaluated in the reference's order,
  opus_ifft           third_party/opus/celt/kiss_fft.c:696-747
  clt_mdct_backward   third_party/opus/celt/mdct.c:267-379
  compute_inv_mdcts   third_party/opus/celt/celt_decoder_clean.c:264-312
  tail hand-over      third_party/opus/celt/celt_decoder_clean.c:622-642
Parity status: pinned (tests/test_oracle.py).

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline leg,
--impl reference) may import this module.  It is the checker, never the
product: libnyquist_b200 has no CPU path.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libnq_oracle.so")
GOLDEN_TABLES = os.path.join(os.path.dirname(HERE), "tests", "golden", "ref_tables.npz")

FRAME = 960
HALF_OVL = 60

_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle (gcc, a second or two).  Also tries `make ref`, which
    is a no-op when /root/reference is absent (GPU box: prebuilt .so travels)."""
    src = os.path.join(HERE, "celt_synth_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        fp = C.POINTER(C.c_float)
        L.nqo_set_tables.argtypes = [fp, fp, fp]
        L.nqo_ifft.argtypes = [C.c_int, fp, fp]
        L.nqo_mdct_backward.argtypes = [fp, fp, C.c_int, C.c_int]
        L.nqo_compute_inv_mdcts.argtypes = [C.c_int, fp, C.POINTER(fp), C.c_int, C.c_int]
        L.nqo_synth_batch.argtypes = [fp, C.c_void_p, fp, fp, fp, C.c_long, C.c_int, C.c_int]
        L.nqo_synth_batch.restype = C.c_double
        L.nqo_comb_filter.argtypes = [fp, fp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int]
        L.nqo_post_batch.argtypes = [fp, C.c_void_p, C.c_long, C.c_int, fp, fp, fp, fp, fp]
        # Install the reference's exact static tables when the fixture exists
        # (one-ulp differences from the formula otherwise; see the C header).
        if os.path.exists(GOLDEN_TABLES):
            t = np.load(GOLDEN_TABLES)
            w = np.ascontiguousarray(t["window120"], np.float32)
            g = np.ascontiguousarray(t["trig481"], np.float32)
            k = np.ascontiguousarray(t["twiddles480"], np.float32)
            L.nqo_set_tables(_fp(w), _fp(g), _fp(k))
        else:
            L.nqo_init_default()
        _lib = L
    return _lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None else None


def use_default_tables():
    lib().nqo_init_default()


def tables():
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
    L.nqo_get_tables(*[v.ctypes.data_as(C.c_void_p) for v in out.values()])
    return out


def opus_ifft(x_ri: np.ndarray, shift: int) -> np.ndarray:
    x = np.ascontiguousarray(x_ri, np.float32)
    assert x.size == 2 * (480 >> shift)
    y = np.zeros_like(x)
    lib().nqo_ifft(shift, _fp(x), _fp(y))
    return y


def clt_mdct_backward(inp: np.ndarray, out: np.ndarray, shift: int, stride: int) -> None:
    assert inp.dtype == np.float32 and out.dtype == np.float32
    assert inp.flags.c_contiguous and out.flags.c_contiguous
    lib().nqo_mdct_backward(_fp(inp), _fp(out), shift, stride)


def synth_batch(coef, transient, tail_in=None, nthreads=1, want_pcm=True):
    """coef [nframes][C][960], transient [nframes] -> (pcm [nframes*960][C], tail [C][60], seconds)."""
    coef = np.ascontiguousarray(coef, np.float32)
    nframes, Cn, n = coef.shape
    assert n == FRAME
    tr = np.ascontiguousarray(transient, np.uint8)
    assert tr.shape == (nframes,)
    ti = None if tail_in is None else np.ascontiguousarray(tail_in, np.float32)
    pcm = np.zeros((nframes * FRAME, Cn), np.float32) if want_pcm else None
    tail = np.zeros((Cn, HALF_OVL), np.float32)
    sec = lib().nqo_synth_batch(_fp(coef), tr.ctypes.data_as(C.c_void_p), _fp(ti),
                                _fp(pcm), _fp(tail), nframes, Cn, int(nthreads))
    return pcm, tail, sec


def compute_inv_mdcts(shortBlocks: int, X: np.ndarray, out_mem, Cn: int, LM: int) -> None:
    """celt_decoder_clean.c:264 on caller buffers: X [C][N*B], out_mem = list of C
    float32 arrays (>= N*B+60), [0,60) = previous raw tail on entry."""
    X = np.ascontiguousarray(X, np.float32)
    fp = C.POINTER(C.c_float)
    outs = (fp * Cn)(*[_fp(o) for o in out_mem])
    lib().nqo_compute_inv_mdcts(int(shortBlocks), _fp(X), outs, int(Cn), int(LM))


# ---- post-filter + de-emphasis ------------------------
POST_FRAME_DTYPE = np.dtype([("N", np.int32), ("pitch", np.int32, 3), ("gain", np.float32, 3),
                             ("tapset", np.int32, 3)])
HIST = 1026


def comb_filter(buf: np.ndarray, start: int, T0, T1, N, g0, g1, tapset0, tapset1) -> None:
    """celt.c:114 in place on buf[start:start+N]; buf[:start] is the history."""
    assert buf.dtype == np.float32 and buf.flags.c_contiguous and start >= max(T0, T1) + 2
    p = C.cast(buf.ctypes.data + 4 * start, C.POINTER(C.c_float))
    lib().nqo_comb_filter(p, p, int(T0), int(T1), int(N), float(g0), float(g1), int(tapset0), int(tapset1))


def post_batch(sig, frames, hist_in=None, mem_in=None):
    """comb_filter x2 + deemphasis over a run of frames (celt_decoder_clean.c:658-670, :723).
    sig [nsamples][C] celt_sig, frames: POST_FRAME_DTYPE [nframes].
    Returns (pcm [nsamples][C], hist_out [C][1026], mem_out [C])."""
    sig = np.ascontiguousarray(sig, np.float32)
    frames = np.ascontiguousarray(frames, POST_FRAME_DTYPE)
    nsamples, Cn = sig.shape
    assert int(frames["N"].sum()) == nsamples
    hi = None if hist_in is None else np.ascontiguousarray(hist_in, np.float32)
    mi = None if mem_in is None else np.ascontiguousarray(mem_in, np.float32)
    pcm = np.zeros_like(sig)
    ho = np.zeros((Cn, HIST), np.float32)
    mo = np.zeros(Cn, np.float32)
    lib().nqo_post_batch(_fp(sig), frames.ctypes.data_as(C.c_void_p), len(frames), Cn, _fp(hi), _fp(mi),
                         _fp(pcm), _fp(ho), _fp(mo))
    return pcm, ho, mo


def post_frames_from_records(recs):
    """Side info for the post stage from oracle.ref.decode_file(..., record=True) records."""
    out = np.zeros(len(recs), POST_FRAME_DTYPE)
    for i, r in enumerate(recs):
        c = r["comb"]
        nn = r["coef"].shape[1]
        out[i]["N"] = nn
        # call 1: (old -> cur); call 2 (LM > 0): (cur -> new)
        out[i]["pitch"][0], out[i]["pitch"][1] = int(c[0][1]), int(c[0][2])
        out[i]["gain"][0], out[i]["gain"][1] = c[0][3], c[0][4]
        out[i]["tapset"][0], out[i]["tapset"][1] = int(c[0][5]), int(c[0][6])
        if nn > 120:
            assert int(c[1][1]) == int(c[0][2]) and c[1][3] == c[0][4]
            out[i]["pitch"][2], out[i]["gain"][2], out[i]["tapset"][2] = int(c[1][2]), c[1][4], int(c[1][6])
    return out
