"""Lane-accurate numpy model of the fast CUDA kernel's data movement
(libnyquist_b200/csrc/celt_synth_kernels.cu): which lane holds which bin,
the mirrored-lane shuffles, the padded transpose buffer, the folded twiddle
tables (taken from the product's host code via nq_celt_debug_tables) and the
fused window / overlap-add / interleave epilogue.  The small DFTs themselves
are np.fft here; the CUDA codelets are checked by tests/host_codelet_check.cu.
Test-only: validates the kernel DESIGN against the oracle without a GPU.
"""
import numpy as np


def _idft(x, axis):
    return np.fft.ifft(x, axis=axis) * x.shape[axis]


class WarpModel:
    def __init__(self, tables):
        self.t_long = tables["t_long"][..., 0].astype(np.float64) + 1j * tables["t_long"][..., 1]   # [16][31]
        self.t_short = tables["t_short"][..., 0].astype(np.float64) + 1j * tables["t_short"][..., 1]  # [2][30]
        self.w = tables["window"].astype(np.float64)
        self.pre = np.exp(2j * np.pi * np.arange(30) / 120)
        self.post = np.exp(2j * np.pi * np.arange(16) / 64)

    def long_frame(self, rows, tail, nch):
        """rows [2][960] coefficients, tail [2][60] (updated in place) -> out [960][2]"""
        lanes = np.arange(32)
        c1, n2 = lanes >> 4, lanes & 15
        n1 = np.arange(30)
        # v[lane][n1] = (X[2i], X[2i+1]), i = 16 n1 + n2
        i = 16 * n1[None, :] + n2[:, None]
        vx = rows[c1[:, None], 2 * i]
        vy = rows[c1[:, None], 2 * i + 1]
        # xb = X[959 - 2i]: the odd element of bin 479-i, i.e. lane^15's v[29-n1].y (the kernel reads it directly)
        xb = vy[lanes ^ 15][:, ::-1]
        g = (xb + 1j * vx) * self.pre[None, :]
        A = _idft(g, 1)                                    # [lane][k1]
        xbuf = np.zeros((2, 16, 31), complex)
        xbuf[c1, n2, :30] = A * self.t_long[n2, :30]
        out = np.zeros((960, 2))
        new_tail = tail.copy()
        for ch in range(nch):
            E = np.zeros((30, 16)); IM = np.zeros((30, 16))
            for k1 in range(30):
                z = xbuf[ch, :, k1]
                Z = _idft(z, 0) * self.post
                E[k1] = -Z.real
                IM[k1] = Z.imag
            O = np.zeros((30, 16))
            for k1 in range(30):
                for k2 in range(16):
                    O[k1, k2] = IM[29 - k1, 15 - k2]
            for k1 in range(30):
                t0, t1 = tail[ch, 59 - 2 * k1], tail[ch, 58 - 2 * k1]
                new_tail[ch, 2 * k1], new_tail[ch, 2 * k1 + 1] = E[k1, 15], O[k1, 15]
                w = self.w
                wA, wB, wC, wD = w[59 - 2 * k1], w[60 + 2 * k1], w[58 - 2 * k1], w[61 + 2 * k1]
                y0, y1 = E[k1, 0], O[k1, 0]
                out[60 + 2 * k1, ch] = wA * t0 + wB * y0
                out[61 + 2 * k1, ch] = wC * t1 + wD * y1
                out[58 - 2 * k1, ch] = wD * t1 - wC * y1
                out[59 - 2 * k1, ch] = wB * t0 - wA * y0
                for k2 in range(1, 15):
                    n = 60 + 2 * (k1 + 30 * k2)
                    out[n, ch] = E[k1, k2]
                    out[n + 1, ch] = O[k1, k2]
        tail[:] = new_tail
        return out

    def short_frame(self, rows, tail, nch):
        lanes = np.arange(32)
        c, b, h = lanes >> 4, (lanes >> 1) & 7, lanes & 1
        n1 = np.arange(30)
        xa = rows[c[:, None], b[:, None] + 32 * n1[None, :] + 16 * h[:, None]]
        xbv = rows[c[:, None], b[:, None] + 952 - 32 * n1[None, :] - 16 * h[:, None]]
        g = (xbv + 1j * xa) * self.pre[None, :]
        A = _idft(g, 1) * self.t_short[h]                  # [lane][k1]
        P = A[lanes ^ 1]
        Z = np.where(h[:, None] == 1, P - A, A + P)
        d = np.where(h == 1, np.exp(1j * np.pi / 4), 1.0)[:, None]
        Y = Z * d
        head = np.where(h[:, None] == 1, Y.imag, -Y.real)
        tl = np.where(h[:, None] == 1, -Y.real, Y.imag)
        out = np.zeros((960, 2))
        new_tail = tail.copy()
        for lane in range(32):
            for k1 in range(30):
                m = 59 - 2 * k1 if h[lane] else 2 * k1
                tp = tl[lane - 2, k1] if b[lane] > 0 else tail[c[lane], 59 - m]
                wlo, whi = self.w[59 - m], self.w[60 + m]
                out[120 * b[lane] + 59 - m, c[lane]] = whi * tp - wlo * head[lane, k1]
                out[120 * b[lane] + 60 + m, c[lane]] = wlo * tp + whi * head[lane, k1]
                if b[lane] == 7:
                    new_tail[c[lane], 59 - m] = tl[lane, k1]
        tail[:] = new_tail
        if nch == 1:
            out[:, 1] = 0
        return out

    def synth(self, coef, transient, tail_in=None):
        nframes, C, _ = coef.shape
        pcm = np.zeros((nframes * 960, C))
        tail_out = np.zeros((C, 60))
        for pair in range((C + 1) // 2):
            cb = 2 * pair
            nch = min(2, C - cb)
            tail = np.zeros((2, 60))
            if tail_in is not None:
                tail[:nch] = tail_in[cb:cb + nch]
            for f in range(nframes):
                rows = np.zeros((2, 968))
                rows[:nch, :960] = coef[f, cb:cb + nch]
                o = self.short_frame(rows, tail, nch) if transient[f] else self.long_frame(rows, tail, nch)
                pcm[f * 960:(f + 1) * 960, cb:cb + nch] = o[:, :nch]
            tail_out[cb:cb + nch] = tail[:nch]
        return pcm, tail_out
