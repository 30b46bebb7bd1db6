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
