"""Design prototype (float64 numpy) of the factorisation used by the fast
CUDA kernel: 480 = 30 x 16 (long) and 60 = 30 x 2 (short), with the MDCT
pre-/post-rotations folded into one inter-stage twiddle table.  Checked
against the oracle; not used by the product."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import port

PI_F = np.float32(3.141592653)

def sine_of(N):
    return float(np.float32(2) * PI_F * np.float32(.125) / np.float32(N))

def imdct_long(X):
    N = 1920
    s = sine_of(N)
    n1 = np.arange(30)[:, None]; n2 = np.arange(16)[None, :]
    i = 16 * n1 + n2
    a = X[2 * i]; b = X[959 - 2 * i]
    c = np.exp(2j * np.pi * np.arange(30) / 120)[:, None]
    g = (b + 1j * a) * c                                   # [n1][n2]
    k1 = np.arange(30)
    W30 = np.exp(2j * np.pi * np.outer(k1, np.arange(30)) / 30)   # [k1][n1]
    A = W30 @ g                                            # [k1][n2]
    T = -(1 + 1j * s) ** 2 * np.exp(2j * np.pi * n2 / N) * np.exp(2j * np.pi * n2 * k1[:, None] / 480) * np.exp(2j * np.pi * k1[:, None] / N)
    Bv = A * T                                             # [k1][n2]
    k2 = np.arange(16)
    W16 = np.exp(2j * np.pi * np.outer(np.arange(16), k2) / 16)   # [n2][k2]
    Z = Bv @ W16                                           # [k1][k2]
    Y = Z * np.exp(2j * np.pi * k2 / 64)[None, :]
    y = np.zeros(960)
    k = k1[:, None] + 30 * k2[None, :]
    y[2 * k] = -Y.real
    y[959 - 2 * k] = Y.imag
    return y

def imdct_short(x):      # x: 120 coefficients of one sub-block
    N = 240
    s = sine_of(N)
    n1 = np.arange(30)[:, None]; h = np.arange(2)[None, :]
    i = 2 * n1 + h
    a = x[2 * i]; b = x[119 - 2 * i]
    c = np.exp(2j * np.pi * np.arange(30) / 120)[:, None]
    g = (b + 1j * a) * c
    k1 = np.arange(30)
    W30 = np.exp(2j * np.pi * np.outer(k1, np.arange(30)) / 30)
    A = W30 @ g                                            # [k1][h]
    T = -(1 + 1j * s) ** 2 * np.exp(2j * np.pi * h / N) * np.exp(2j * np.pi * h * k1[:, None] / 60) * np.exp(2j * np.pi * k1[:, None] / N)
    Bv = A * T
    Z = np.stack([Bv[:, 0] + Bv[:, 1], Bv[:, 0] - Bv[:, 1]], 1)   # [k1][k2]
    Y = Z * np.exp(1j * np.pi * np.arange(2) / 4)[None, :]
    y = np.zeros(120)
    k = k1[:, None] + 30 * np.arange(2)[None, :]
    y[2 * k] = -Y.real
    y[119 - 2 * k] = Y.imag
    return y

def synth(coef, transient, tail_in, window):
    nframes, C, _ = coef.shape
    pcm = np.zeros((nframes * 960, C))
    tail = np.zeros((C, 60)) if tail_in is None else tail_in.astype(np.float64).copy()
    w = window.astype(np.float64)
    def blend(t, y):    # returns the 120 finished samples from tail t[60] and y[0..60)
        o = np.zeros(120)
        m = np.arange(60)
        o[59 - m] = w[60 + m] * t[59 - m] - w[59 - m] * y[m]
        o[60 + m] = w[59 - m] * t[59 - m] + w[60 + m] * y[m]
        return o
    for f in range(nframes):
        for c in range(C):
            out = np.zeros(960)
            if not transient[f]:
                y = imdct_long(coef[f, c].astype(np.float64))
                out[:120] = blend(tail[c], y)
                out[120:] = y[60:900]
                tail[c] = y[900:]
            else:
                for b in range(8):
                    y = imdct_short(coef[f, c, b::8].astype(np.float64))
                    out[120 * b:120 * b + 120] = blend(tail[c], y)
                    tail[c] = y[60:]
            pcm[f * 960:(f + 1) * 960, c] = out
    return pcm, tail

if __name__ == "__main__":
    rng = np.random.default_rng(1)
    nf, C = 8, 2
    coef = (rng.standard_normal((nf, C, 960)) * 1000).astype(np.float32)
    tr = np.array([0, 1, 0, 0, 1, 1, 0, 0], np.uint8)
    tail_in = (rng.standard_normal((C, 60)) * 300).astype(np.float32)
    want, wtail, _ = port.synth_batch(coef, tr, tail_in)
    got, gtail = synth(coef, tr, tail_in, port.tables()["window120"])
    print("max err", np.abs(got - want).max(), "of max", np.abs(want).max(), "tail err", np.abs(gtail - wtail).max())
