"""GPU tests (-m gpu) of the two-phase decoder end to end:
integration/_build/libnyquist_twophase.so is the REFERENCE's own library (bundled Ogg/Opus +
Common.cpp, compiled in place by integration/Makefile) with the CELT synthesis overlaid away and
integration/OpusDecoderTwoPhase.cpp in place of src/OpusDecoder.cpp.  nqr::NyquistIO::Load on it
must return the AudioData the unmodified reference returns.

Bar: |pcm - reference pcm| <= 1e-5 (north_star), identical length / channel count, and the
reference's own acceptance test (examples/src/Main.cpp:131-149: int(sum), length).
"""
import ctypes as C
import os
import time

import numpy as np
import pytest

from conftest import ROOT, snr_db
from oracle import ref

pytestmark = pytest.mark.gpu

LIB = os.path.join(ROOT, "integration", "_build", "libnyquist_twophase.so")
CHECKSUMS = {"sb-reverie.opus": (403, 21472602), "sb-reverie-60ms-frames.opus": (719, 21472602), "short.opus": (22, 421930)}


@pytest.fixture(scope="module")
def twophase():
    if not os.path.exists(LIB):
        pytest.skip("integration/_build/libnyquist_twophase.so not built (needs /root/reference; make -C integration)")
    L = C.CDLL(LIB)
    L.nq_twophase_load.argtypes = [C.c_char_p, C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.c_size_t),
                                   C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double)]
    L.nq_twophase_free.argtypes = [C.POINTER(C.c_float)]
    return L


def load(L, path):
    p = C.POINTER(C.c_float)()
    n, ch, sr = C.c_size_t(0), C.c_int(0), C.c_int(0)
    tm = (C.c_double * 8)()
    t0 = time.perf_counter()
    rc = L.nq_twophase_load(path.encode(), C.byref(p), C.byref(n), C.byref(ch), C.byref(sr), tm)
    wall = time.perf_counter() - t0
    if rc != 0:
        return None, 0, 0, None, wall
    a = np.ctypeslib.as_array(p, shape=(n.value,)).copy()
    L.nq_twophase_free(p)
    return a.reshape(-1, ch.value), ch.value, sr.value, list(tm), wall


@pytest.mark.parametrize("fname", sorted(CHECKSUMS))
def test_nyquistio_load_two_phase_matches_reference(twophase, fname):
    path = os.path.join(os.path.dirname(ref.LIB_PATH), "test_data", fname)
    if not (ref.available() and os.path.exists(path)):
        pytest.skip("oracle/_ref (compiled reference + staged test_data) not present")
    load(twophase, path)                       # warm-up: CUDA context, pinned pools
    got, ch, sr, tm, wall = load(twophase, path)
    if ref.load_available():     # the unmodified reference's own NyquistIO::Load, same shim, same box
        ref.nyquist_load(path)
        want, sr_ref, t_ref = ref.nyquist_load(path)
        assert sr_ref == sr
    else:
        want, _ = ref.decode_file(path)
        t_ref = float("nan")
    assert got is not None and (ch, sr) == (2, 48000)
    assert got.shape == want.shape
    err = float(np.abs(got.astype(np.float64) - want).max())
    assert err <= 1e-5 and snr_db(want, got) >= 100.0, err
    flat = np.concatenate([got[:, c] for c in range(ch)])
    s = float(np.cumsum(flat, dtype=np.float32)[-1])
    assert (int(s), flat.size) == CHECKSUMS[fname]
    print(f"\n{fname}: Load {wall * 1e3:.0f} ms (phase 1 CPU {tm[0] * 1e3:.0f} ms, phase 2 tail {tm[1] * 1e3:.0f} ms, "
          f"gain {tm[2] * 1e3:.0f} ms) vs the reference's own CPU Load {t_ref * 1e3:.0f} ms; "
          f"max |err| {err:.2e}, sum {s:.4f}")


def test_nyquistio_load_two_phase_8_channel_multistream(twophase):
    """BASELINE config 4 through nqr::NyquistIO::Load: 7.1 file (3 coupled + 2 mono streams, made
    with the reference's own encoder; the reference mount lacks Rachel8ch.opus)."""
    from conftest import GOLDEN
    path = os.path.join(GOLDEN, "surround8.opus")
    load(twophase, path)
    got, ch, sr, tm, wall = load(twophase, path)
    assert got is not None and (ch, sr) == (8, 48000)
    if ref.load_available():
        want, _, t_ref = ref.nyquist_load(path)
    elif ref.available():
        want, _ = ref.decode_file(path)
        t_ref = float("nan")
    else:
        pytest.skip("oracle/_ref (compiled reference) not present")
    assert got.shape == want.shape
    err = float(np.abs(got.astype(np.float64) - want).max())
    assert err <= 1e-5 and snr_db(want, got) >= 100.0, err
    print(f"\nsurround8.opus: Load {wall * 1e3:.0f} ms (phase 1 {tm[0] * 1e3:.0f} ms, phase 2 tail {tm[1] * 1e3:.0f} ms) "
          f"vs reference Load {t_ref * 1e3:.0f} ms; max |err| {err:.2e}")


def test_parallel_phase1_over_streams_equals_sequential(twophase, monkeypatch):
    """the streams of a multistream packet are entropy-decoded on
    helper threads at the same time (opusfile decode callback -> opus_decode_native per stream,
    pushes of different streams racing into the sink).  The PCM must be bit-identical to the
    sequential reference path (NQ_PHASE1_THREADS=1), load after load."""
    from conftest import GOLDEN
    path = os.path.join(GOLDEN, "surround8.opus")
    monkeypatch.setenv("NQ_PHASE1_THREADS", "1")
    load(twophase, path)
    seq, ch, _, tm_seq, wall_seq = load(twophase, path)
    assert seq is not None and ch == 8
    monkeypatch.delenv("NQ_PHASE1_THREADS")
    walls, p1 = [], []
    for _ in range(12):
        par, ch, _, tm, wall = load(twophase, path)
        assert par is not None and par.shape == seq.shape and np.array_equal(par, seq)
        walls.append(wall)
        p1.append(tm[0])
    monkeypatch.setenv("NQ_PHASE1_THREADS", "2")
    two, *_ = load(twophase, path)
    assert np.array_equal(two, seq)
    print(f"\nsurround8.opus phase 1: sequential {tm_seq[0] * 1e3:.1f} ms (Load {wall_seq * 1e3:.1f} ms), "
          f"5 streams in parallel {min(p1) * 1e3:.1f} ms (Load {min(walls) * 1e3:.1f} ms)")


def test_concurrent_loads_from_several_threads(twophase):
    """nqr::NyquistIO::Load is callable from several threads at once in the reference (the CPU
    decoder has no shared state); here every Load leases its own device context."""
    import threading
    from conftest import GOLDEN
    paths = [os.path.join(GOLDEN, "surround8.opus")]
    short = os.path.join(os.path.dirname(ref.LIB_PATH), "test_data", "short.opus")
    if os.path.exists(short):
        paths.append(short)
    want = {p: load(twophase, p)[0] for p in paths}
    assert all(w is not None for w in want.values())
    results = {}

    def work(i):
        p = paths[i % len(paths)]
        out = []
        for _ in range(3):
            out.append(load(twophase, p)[0])
        results[i] = (p, out)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(6)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert len(results) == 6
    for p, outs in results.values():
        for got in outs:
            assert got is not None and np.array_equal(got, want[p])


@pytest.mark.parametrize("gain_q8", [-1536, 768])
def test_header_gain_is_applied_like_the_reference(twophase, tmp_path, gain_q8):
    """OpusHead.output_gain (opusfile OP_HEADER_GAIN -> OPUS_SET_GAIN -> opus_decoder_clean.c:578-588):
    short.opus with the gain field rewritten to -6 dB / +3 dB."""
    src = os.path.join(os.path.dirname(ref.LIB_PATH), "test_data", "short.opus")
    if not (ref.available() and os.path.exists(src)):
        pytest.skip("oracle/_ref (compiled reference + staged test_data) not present")
    path = tmp_path / "gain.opus"
    path.write_bytes(ref.with_output_gain(open(src, "rb").read(), gain_q8))
    got, ch, sr, tm, wall = load(twophase, str(path))
    want, _ = ref.decode_file(str(path))
    plain, _ = ref.decode_file(src)
    assert ref.header_info()[1] == 0 and got is not None and got.shape == want.shape
    scale = float(np.abs(want).max() / np.abs(plain).max())
    assert abs(scale - 10 ** (gain_q8 / 256 / 20)) < 1e-4          # the reference really applied it
    assert float(np.abs(got.astype(np.float64) - want).max()) <= 1e-5 * max(1.0, scale)
    assert snr_db(want, got) >= 100.0


def test_silk_only_file_decodes_through_phase_1(twophase):
    """test_data/ad_hoc/detodos.opus is SILK-only: none of its packets reaches celt_decode_with_ec
    (opus_decoder_clean.c:499-513), so there is no synthesis for phase 2 and the reference's own
    SILK decoder, which phase 1 runs unchanged, produces the final PCM -- bit-identical to the
    unmodified reference.  (Hybrid packets, which mix the two decoders, are refused.)"""
    path = os.path.join(os.path.dirname(ref.LIB_PATH), "test_data", "detodos.opus")
    if not (os.path.exists(path) and ref.available()):
        pytest.skip("detodos.opus / oracle/_ref not staged")
    got, ch, sr, tm, _ = load(twophase, path)
    want, _ = ref.decode_file(path)
    assert want.shape == (139848, 1)
    assert got is not None and (ch, sr) == (1, 48000) and got.shape == want.shape
    assert np.array_equal(got, want)


def test_hybrid_file_celt_layer_on_the_gpu_plus_silk_layer_of_phase_1(twophase):
    """SURVEY.md section 8(f) row 4, end to end: tests/golden/hybrid.opus (the reference's own encoder
    forced into MODE_HYBRID: 150 packets, each a SILK layer + a 20 ms CELT frame starting at band 17,
    47 of them transient).  opus_decode_frame sums the two layers (opus_decoder_clean.c:553-560); here
    the CELT layer is synthesised in phase 2 and the SILK layer is what phase 1 hands back."""
    import hashlib
    import json
    from conftest import GOLDEN
    info = json.load(open(os.path.join(GOLDEN, "modes.json")))
    for name, exact in (("hybrid", False), ("silk_stereo", True)):
        path = os.path.join(GOLDEN, name + ".opus")
        got, ch, sr, tm, wall = load(twophase, path)
        assert got is not None and (ch, sr) == (2, 48000)
        assert got.shape == (info[name]["samples_per_channel"], 2)
        if ref.available():
            want, recs = ref.decode_file(path, record=True)
            assert hashlib.sha256(want.tobytes()).hexdigest() == info[name]["reference_pcm_sha256"]
            assert len(recs) == info[name]["celt_frames"]
            err = float(np.abs(got.astype(np.float64) - want).max())
            assert err <= 1e-5 and snr_db(want, got) >= 100.0, (name, err)
            if exact:
                assert np.array_equal(got, want)
            else:     # the CELT layer is really there: without it the error is orders of magnitude larger
                assert float(np.abs(want).max()) > 0.1
            print(f"\n{name}.opus: Load {wall * 1e3:.1f} ms, max |err| {err:.2e}")
        elif exact:
            assert hashlib.sha256(got.tobytes()).hexdigest() == info[name]["reference_pcm_sha256"]


def _switch_signal(n, channels, seed=1):
    t = np.arange(n) / 48000
    rng = np.random.default_rng(seed)
    sig = (0.2 * np.sin(2 * np.pi * 180 * t) * (1 + 0.5 * np.sin(2 * np.pi * 2.3 * t)) + 0.05 * rng.standard_normal(n)).astype(np.float32)
    sig[::7200] += 0.5                                   # a few clicks: transient frames
    return np.stack([sig, np.roll(sig, 41) * 0.7], 1)[:, :channels].copy()


MODES = {"silk": 1000, "hybrid": 1001, "celt": 1002}     # opus_private.h:90-92


@pytest.mark.parametrize("a,b,channels", [("hybrid", "celt", 2), ("celt", "hybrid", 2), ("silk", "celt", 2), ("celt", "silk", 2),
                                          ("hybrid", "silk", 2), ("silk", "hybrid", 2), ("celt", "hybrid", 1), ("silk", "celt", 1)])
def test_mode_switching_file_matches_the_reference(twophase, tmp_path, a, b, channels):
    """a switch between coding modes makes opus_decode_frame decode a
    5 ms redundancy frame into a side buffer and cross-fade it in (opus_decoder_clean.c:478-487,
    :530-555), reset the CELT decoder (:495-497, :536-538) and, from hybrid to SILK, add a 2.5 ms CELT
    fade-out frame (:507-514).  Phase 2 decodes those frames in sequence (side frames, reset flags,
    frames with their own place in the output), the loader applies the same fades to its PCM."""
    if not ref.available():
        pytest.skip("oracle/_ref (compiled reference) not present")
    data = ref.encode_mode_switch(_switch_signal(960 * 60, channels), MODES[a], MODES[b], 30)
    path = tmp_path / "switch.opus"
    path.write_bytes(data)
    want, recs = ref.decode_bytes(data, record=True)
    sizes = [r["coef"].shape[1] for r in recs]
    if "celt" in (a, b):
        assert 240 in sizes                                  # the redundancy frame is there
    # (hybrid -> SILK with its 2.5 ms fade-out frame: the fixture of the next test; this generator keeps the
    # full bandwidth after the switch, which makes the encoder stay hybrid)
    got, ch, sr, tm, wall = load(twophase, str(path))
    assert got is not None and ch == channels and got.shape == want.shape
    err = float(np.abs(got.astype(np.float64) - want).max())
    assert err <= 1e-5 and snr_db(want, got) >= 100.0, (a, b, err)
    assert float(np.abs(want).max()) > 0.1


@pytest.mark.parametrize("seed", range(10))
def test_random_mode_schedules_match_or_are_refused(twophase, tmp_path, seed):
    """Random walks through the coding modes (random switch points, mono / stereo, bitrates from 12 to
    64 kbit/s, some switches only a frame apart): the loader either returns the reference decoder's PCM
    or refuses the file loudly (switches the encoder makes without a redundancy frame need loss
    concealment, which the bundled decoder does not define) -- never anything else."""
    if not ref.available():
        pytest.skip("oracle/_ref (compiled reference) not present")
    rng = np.random.default_rng(900 + seed)
    channels = int(rng.integers(1, 3))
    nfr = int(rng.integers(40, 90))
    modes = list(MODES.values())
    first = int(rng.choice(modes))
    sched, f, cur = [], 0, first
    while True:
        f += int(rng.choice([1, 2, 3, 7, 15, 25]))
        if f >= nfr - 1:
            break
        cur = int(rng.choice([m for m in modes if m != cur]))
        sched.append((f, cur))
    bitrate = int(rng.choice([12000, 20000, 32000, 48000, 64000]))
    data = ref.encode_mode_schedule(_switch_signal(960 * nfr, channels, seed), first, sched, bitrate)
    path = tmp_path / "walk.opus"
    path.write_bytes(data)
    want, recs = ref.decode_bytes(data, record=True)
    got, ch, sr, tm, wall = load(twophase, str(path))
    if got is None:
        return                       # refused (the message names the reason); the reference alone decodes it
    assert ch == channels and got.shape == want.shape
    err = float(np.abs(got.astype(np.float64) - want).max())
    assert err <= 1e-5 and snr_db(want, got) >= 100.0, (seed, first, sched, bitrate, err)


def test_mode_walk_fixture_matches_the_reference(twophase):
    """tests/golden/modeswitch.opus (make_golden.py): CELT -> hybrid -> SILK -> hybrid -> CELT -> SILK ->
    CELT -> hybrid in one file: 7 redundancy frames, 1 fade-out frame, 110 ordinary CELT frames."""
    import hashlib
    import json
    from conftest import GOLDEN
    info = json.load(open(os.path.join(GOLDEN, "modes.json")))["modeswitch"]
    path = os.path.join(GOLDEN, "modeswitch.opus")
    got, ch, sr, tm, wall = load(twophase, path)
    assert got is not None and (ch, sr) == (2, 48000) and got.shape == (info["samples_per_channel"], 2)
    if not ref.available():
        pytest.skip("oracle/_ref (compiled reference) not present: shape checked only")
    want, recs = ref.decode_file(path, record=True)
    assert hashlib.sha256(want.tobytes()).hexdigest() == info["reference_pcm_sha256"]
    sizes = [r["coef"].shape[1] for r in recs]
    assert (len(recs), sizes.count(240), sizes.count(120)) == (info["celt_frames"], info["redundancy_frames"], info["fade_out_frames"])
    err = float(np.abs(got.astype(np.float64) - want).max())
    assert err <= 1e-5 and snr_db(want, got) >= 100.0, err
    print(f"\nmodeswitch.opus: Load {wall * 1e3:.1f} ms, max |err| {err:.2e}")


def test_two_phase_load_errors_like_the_reference(twophase, tmp_path):
    bad = tmp_path / "noise.opus"
    bad.write_bytes(np.random.default_rng(0).integers(0, 256, 5000, dtype=np.uint8).tobytes())
    got, *_ = load(twophase, str(bad))          # not an Ogg Opus file: Load throws, as the reference's does
    assert got is None
    got, *_ = load(twophase, str(tmp_path / "missing.opus"))
    assert got is None


def load_batch(L, paths, threads=0):
    L.nq_twophase_load_batch.argtypes = [C.POINTER(C.c_char_p), C.c_int, C.c_int, C.POINTER(C.POINTER(C.c_float)),
                                         C.POINTER(C.c_size_t), C.POINTER(C.c_int), C.POINTER(C.c_double)]
    n = len(paths)
    arr = (C.c_char_p * n)(*[p.encode() for p in paths])
    ptrs = (C.POINTER(C.c_float) * n)()
    counts, chans, stats = (C.c_size_t * n)(), (C.c_int * n)(), (C.c_double * 7)()
    t0 = time.perf_counter()
    rc = L.nq_twophase_load_batch(arr, n, threads, ptrs, counts, chans, stats)
    wall = time.perf_counter() - t0
    assert rc == 0
    out = []
    for i in range(n):
        a = np.ctypeslib.as_array(ptrs[i], shape=(counts[i],)).copy().reshape(-1, chans[i])
        L.nq_twophase_free(ptrs[i])
        out.append(a)
    names = ("phase1_s", "phase2_s", "total_s", "files_batched", "files_single", "frames", "launches")
    return out, dict(zip(names, list(stats))), wall


def test_batched_loader_many_files_one_phase2(twophase):
    """nqr::LoadOpusBatch: K files -> phase 1 on K host threads, ONE synthesis launch + ONE post
    launch per channel layout (every file a segment that starts from a reset decoder); the files the
    batch does not cover (hybrid, SILK-only) go through the ordinary loader.  Every file must come
    out exactly as nqr::NyquistIO::Load gives it one by one."""
    from conftest import GOLDEN
    td = os.path.join(os.path.dirname(ref.LIB_PATH), "test_data")
    paths = [os.path.join(td, f) for f in ("sb-reverie.opus", "short.opus", "sb-reverie-60ms-frames.opus", "short.opus")]
    paths += [os.path.join(GOLDEN, f) for f in ("surround8.opus", "hybrid.opus", "surround8.opus")]
    paths = [p for p in paths if os.path.exists(p)]
    if len(paths) < 3:
        pytest.skip("test files not staged")
    singles = [load(twophase, p)[0] for p in paths]
    got, st, wall = load_batch(twophase, paths)
    nhyb = sum(p.endswith("hybrid.opus") for p in paths)
    assert st["files_single"] == nhyb and st["files_batched"] == len(paths) - nhyb
    layouts = len({s.shape[1] for s, p in zip(singles, paths) if not p.endswith("hybrid.opus")})
    assert st["launches"] == 2 * layouts, st          # one synthesis + one post launch per channel layout
    for p, a, b in zip(paths, got, singles):
        assert a.shape == b.shape, p
        assert np.array_equal(a, b), (p, float(np.abs(a - b).max()))
    print(f"\nbatched loader: {len(paths)} files in {wall * 1e3:.0f} ms (phase 1 {st['phase1_s'] * 1e3:.0f} ms on host threads, "
          f"phase 2 {st['phase2_s'] * 1e3:.0f} ms, {int(st['frames'])} frames, {int(st['launches'])} launches)")
