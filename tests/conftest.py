import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

FULL_SCALE = 32768.0          # CELT_SIG_SCALE, celt/arch.h:49
TOL_FS = 1e-5                 # BASELINE.json north_star: max-abs error <= 1e-5 of full scale
MIN_SNR_DB = 100.0            # ... (>= 100 dB SNR)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def load_npz(name):
    return np.load(os.path.join(GOLDEN, name))


def snr_db(ref, got):
    ref = np.asarray(ref, np.float64)
    err = np.asarray(got, np.float64) - ref
    den = float((err ** 2).sum())
    if den == 0.0:
        return float("inf")
    return 10.0 * np.log10(float((ref ** 2).sum()) / den)


def assert_parity(ref, got, what=""):
    """The north_star bar: max|got-ref| <= 1e-5 * 32768 and SNR >= 100 dB."""
    ref = np.asarray(ref)
    got = np.asarray(got)
    assert ref.shape == got.shape, (what, ref.shape, got.shape)
    assert np.isfinite(got).all(), what
    err = float(np.abs(got.astype(np.float64) - ref.astype(np.float64)).max()) if ref.size else 0.0
    assert err <= TOL_FS * FULL_SCALE, f"{what}: max abs err {err} > {TOL_FS * FULL_SCALE}"
    if ref.size and float(np.abs(ref).max()) > 0:
        s = snr_db(ref, got)
        assert s >= MIN_SNR_DB, f"{what}: SNR {s:.1f} dB < {MIN_SNR_DB}"
