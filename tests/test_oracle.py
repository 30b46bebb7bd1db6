"""CPU tests that PIN the oracle (oracle/celt_synth_oracle.c):
  * against the reference's own golden vectors (test_data/ifft_*.bin), bit-exact;
  * against fixtures produced by the compiled reference (tests/golden/*.npz,
    made by tests/golden/make_golden.py), bit-exact;
  * against the compiled reference itself when oracle/_ref/libnq_ref.so exists.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_npz
from oracle import port, ref


def case_names(z):
    return sorted({k.split(".")[0] for k in z.files})


@pytest.mark.parametrize("n,shift", [(480, 0), (60, 3)])
def test_ifft_reference_golden_vectors_bit_exact(n, shift):
    x = np.fromfile(os.path.join(GOLDEN, f"ifft_input_N{n}.bin"), np.float32)
    y = np.fromfile(os.path.join(GOLDEN, f"ifft_output_N{n}.bin"), np.float32)
    assert x.size == 2 * n and y.size == 2 * n
    got = port.opus_ifft(x, shift)
    assert np.array_equal(got.view(np.uint32), y.view(np.uint32))


@pytest.mark.parametrize("n,shift", [(480, 0), (240, 1), (120, 2), (60, 3)])
def test_ifft_is_unnormalised_inverse_dft(n, shift):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(2 * n).astype(np.float32)
    got = port.opus_ifft(x, shift).astype(np.float64)
    z = x[0::2].astype(np.float64) + 1j * x[1::2].astype(np.float64)
    want = np.fft.ifft(z) * n
    got_c = got[0::2] + 1j * got[1::2]
    assert np.abs(got_c - want).max() < 2e-5 * np.abs(want).max()


def test_bitrev_and_tables_match_reference_statics():
    t = load_npz("ref_tables.npz")
    mine = port.tables()
    for k in ("bitrev480", "bitrev240", "bitrev120", "bitrev60"):
        assert np.array_equal(mine[k], t[k]), k
    for k in ("window120", "trig481", "twiddles480"):
        assert np.array_equal(mine[k].view(np.uint32), t[k].view(np.uint32)), k


def test_formula_tables_within_one_ulp_of_reference_literals():
    t = load_npz("ref_tables.npz")
    port.use_default_tables()
    try:
        mine = port.tables()
    finally:
        port._lib = None  # reload with the golden tables for the other tests
    for k in ("window120", "trig481", "twiddles480"):
        a, b = mine[k].astype(np.float64), t[k].astype(np.float64)
        assert np.abs(a - b).max() <= 6.0e-8, k


def test_single_mdct_calls_bit_exact_all_shifts_and_strides():
    z = load_npz("mdct_calls.npz")
    for shift in range(4):
        for stride in (1, 2, 4, 8):
            key = f"s{shift}_st{stride}"
            inp = z[key + ".in"].copy()
            out = z[key + ".out_before"].copy()
            port.clt_mdct_backward(inp, out, shift, stride)
            assert np.array_equal(inp, z[key + ".in"]), "input must be left untouched"
            assert np.array_equal(out.view(np.uint32), z[key + ".out_after"].view(np.uint32)), key


@pytest.mark.parametrize("name", case_names(load_npz("synth_cases.npz")))
def test_synth_batch_bit_exact_vs_compiled_reference_fixture(name):
    z = load_npz("synth_cases.npz")
    tail_in = z[name + ".tail_in"]
    tail_in = None if tail_in.size == 0 else tail_in
    for nthreads in (1, 3):
        pcm, tail, _ = port.synth_batch(z[name + ".coef"], z[name + ".transient"], tail_in, nthreads)
        assert np.array_equal(pcm.view(np.uint32), z[name + ".pcm"].view(np.uint32)), (name, nthreads)
        assert np.array_equal(tail.view(np.uint32), z[name + ".tail_out"].view(np.uint32)), (name, nthreads)


@pytest.mark.parametrize("tag", ["reverie", "reverie60", "short"])
def test_real_recorded_frames_bit_exact(tag):
    """Frames recorded while the reference decoded the bundled .opus files:
    frame 0 of each run is the halo (supplies the tail), frames 1.. are checked."""
    z = load_npz("real_frames.npz")
    coef, tr, out = z[tag + ".coef"], z[tag + ".transient"], z[tag + ".out"]
    assert tr.any(), "run should contain a transient frame"
    pcm, _, _ = port.synth_batch(coef, tr, None)
    pcm = pcm.reshape(coef.shape[0], 960, 2).transpose(0, 2, 1)
    assert np.array_equal(pcm[1:].view(np.uint32), out[1:].view(np.uint32))


def test_frame_mix_of_bundled_file_recorded():
    z = load_npz("real_frames.npz")
    # 11184 frames, 314 transient; 21472602 samples
    assert int(z["reverie.n_records"]) == 11184
    assert int(z["reverie.n_transient"]) == 314
    assert int(z["reverie.pcm_len"]) == 21472602


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libnq_ref.so not built (needs /root/reference)")
def test_live_compiled_reference_bit_exact_random_batches():
    rng = np.random.default_rng(7)
    for C in (1, 2, 5):
        nframes = 40
        coef = (rng.standard_normal((nframes, C, 960)) * 800).astype(np.float32)
        tr = (rng.uniform(size=nframes) < 0.3).astype(np.uint8)
        tail_in = (rng.standard_normal((C, 60)) * 100).astype(np.float32)
        a = ref.synth_batch(coef, tr, tail_in, nthreads=1)
        b = port.synth_batch(coef, tr, tail_in, nthreads=4)
        assert np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32))
        assert np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
    for shift in range(4):
        x = rng.standard_normal(2 * (480 >> shift)).astype(np.float32)
        assert np.array_equal(ref.opus_ifft(x, shift).view(np.uint32), port.opus_ifft(x, shift).view(np.uint32))
