"""CPU tests that PIN the oracle's post stage:
comb_filter (celt/celt.c:114-172) and deemphasis (celt_decoder_clean.c:192-256),
restated in oracle/celt_synth_oracle.c, against

  * tests/golden/post_cases.npz: single calls of the COMPILED reference and the
    reference decoder's final PCM for the last 26 frames of short.opus;
  * the compiled reference live (oracle/_ref), whole bundled files, bit-exact.
"""
import os

import numpy as np
import pytest

from conftest import load_npz
from oracle import port, ref

HIST = port.HIST


def comb_cases(z):
    return sorted({k.split(".")[0] for k in z.files if k.startswith("comb")})


@pytest.mark.parametrize("name", comb_cases(load_npz("post_cases.npz")))
def test_comb_filter_bit_exact_vs_reference_fixture(name):
    z = load_npz("post_cases.npz")
    T0, T1, N, g0, g1, ts0, ts1 = z[name + ".args"]
    buf = z[name + ".before"].copy()
    port.comb_filter(buf, HIST, int(T0), int(T1), int(N), g0, g1, int(ts0), int(ts1))
    assert np.array_equal(buf.view(np.uint32), z[name + ".after"].view(np.uint32))
    assert np.array_equal(buf[:HIST], z[name + ".before"][:HIST]), "history must be left untouched"


def test_deemphasis_bit_exact_vs_reference_fixture():
    z = load_npz("post_cases.npz")
    x = z["deemph.x"]                      # [C][N]
    frames = np.zeros(1, port.POST_FRAME_DTYPE)
    frames["N"] = x.shape[1]
    frames["pitch"] = 15                   # gains 0: comb_filter is the identity (celt.c:124-129)
    pcm, _, mem = port.post_batch(np.ascontiguousarray(x.T), frames, None, z["deemph.mem_in"])
    assert np.array_equal(pcm.view(np.uint32), z["deemph.pcm"].view(np.uint32))
    assert np.array_equal(mem.view(np.uint32), z["deemph.mem_out"].view(np.uint32))


def test_post_batch_reproduces_reference_pcm_of_short_opus_tail():
    z = load_npz("post_cases.npz")
    frames = z["short_tail.frames"]
    assert frames.dtype == port.POST_FRAME_DTYPE
    assert frames["N"][-1] == 120 and (frames["gain"] > 0).any() and z["short_tail.transient"].any()
    pcm, hist, mem = port.post_batch(z["short_tail.sig"], frames, z["short_tail.hist_in"], z["short_tail.mem_in"])
    assert np.array_equal(pcm.view(np.uint32), z["short_tail.pcm"].view(np.uint32))
    # splitting the run anywhere and carrying (hist, mem) is invisible
    k = 11
    n0 = int(frames["N"][:k].sum())
    a, h1, m1 = port.post_batch(z["short_tail.sig"][:n0], frames[:k], z["short_tail.hist_in"], z["short_tail.mem_in"])
    b, h2, m2 = port.post_batch(z["short_tail.sig"][n0:], frames[k:], h1, m1)
    assert np.array_equal(np.concatenate([a, b]).view(np.uint32), pcm.view(np.uint32))
    assert np.array_equal(h2, hist) and np.array_equal(m2, mem)


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libnq_ref.so not built (needs /root/reference)")
def test_live_comb_filter_and_deemphasis_bit_exact_random():
    rng = np.random.default_rng(5)
    for _ in range(40):
        N = int(rng.choice([120, 840]))
        T0, T1 = (int(v) for v in rng.integers(15, 1023, 2))
        g0, g1 = (float(v) for v in rng.choice([0.0, 0.09375, 0.375, 0.75], 2))
        ts0, ts1 = (int(v) for v in rng.integers(0, 3, 2))
        a = rng.uniform(-4000, 4000, HIST + N).astype(np.float32)
        b = a.copy()
        ref.comb_filter(a, HIST, T0, T1, N, g0, g1, ts0, ts1)
        port.comb_filter(b, HIST, T0, T1, N, g0, g1, ts0, ts1)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (T0, T1, N, g0, g1, ts0, ts1)


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libnq_ref.so not built (needs /root/reference)")
def test_live_whole_file_pcm_bit_exact_short_opus():
    """Reference decoder PCM of short.opus == oracle post stage over the recorded out_syn stream,
    shifted by the OpusHead pre-skip (opusfile.c:2711-2721)."""
    path = os.path.join(os.path.dirname(ref.LIB_PATH), "test_data", "short.opus")
    if not os.path.exists(path):
        pytest.skip("bundled short.opus not staged")
    pcm, recs = ref.decode_file(path, record=True)
    pre_skip, gain = ref.header_info()
    assert (pre_skip, gain) == (312, 0)
    sig = np.concatenate([r["out"].T for r in recs], axis=0)
    full, _, _ = port.post_batch(sig, port.post_frames_from_records(recs))
    assert np.array_equal(full[pre_skip:pre_skip + len(pcm)].view(np.uint32), pcm.view(np.uint32))


# ---- BASELINE config 4: 8-channel multistream file (stand-in for the missing Rachel8ch.opus) ----
@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libnq_ref.so not built (needs /root/reference)")
def test_surround8_reference_decode_is_pinned_and_oracle_reproduces_it_bit_exactly():
    """tests/golden/surround8.opus was made with the reference's own surround encoder.  The
    reference decoder's PCM for it is pinned by hash; the oracle (per-stream synthesis + post stage
    + opus_multistream channel routing) reproduces that PCM bit for bit, mono streams included."""
    import hashlib
    import json
    from conftest import GOLDEN
    from ms_helpers import multistream_batch, oracle_decode_multistream
    meta = json.load(open(os.path.join(GOLDEN, "surround8.json")))
    pcm, recs = ref.decode_file(os.path.join(GOLDEN, "surround8.opus"), record=True)
    assert pcm.shape == (meta["samples_per_channel"], meta["channels"]) and len(recs) == meta["records"]
    assert hashlib.sha256(pcm.tobytes()).hexdigest() == meta["reference_pcm_sha256"]
    ch, streams, coupled, mapping = ref.layout_info()
    assert (ch, streams, coupled, mapping) == (8, 5, 3, [0, 6, 1, 2, 3, 4, 5, 7])
    pre_skip, gain = ref.header_info()
    coef, flags, frames = multistream_batch(recs, streams, coupled)
    assert (flags >> 1 == 0).all() and (flags & 1).sum() == meta["transient_records"]
    assert len({tuple(flags[f]) for f in range(len(flags))}) > 4, "streams should switch blocks independently"
    full = oracle_decode_multistream(coef, flags, frames, streams, coupled, mapping)
    assert np.array_equal(full[pre_skip:pre_skip + len(pcm)].view(np.uint32), pcm.view(np.uint32))


# ---- the other coding modes of opus_decode_frame ----
@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libnq_ref.so not built (needs /root/reference)")
def test_hybrid_and_silk_files_reference_decode_is_pinned():
    """tests/golden/hybrid.opus / silk_stereo.opus were made with the reference's own encoder forced into
    a coding mode (tests/golden/make_golden.py).  The reference decoder's PCM is pinned by hash; the
    hybrid file's CELT layer is one 20 ms frame per packet with nothing below band 17 (bin 320), the
    SILK-only file never reaches the CELT synthesis."""
    import hashlib
    import json
    from conftest import GOLDEN
    info = json.load(open(os.path.join(GOLDEN, "modes.json")))
    for name, meta in info.items():
        pcm, recs = ref.decode_file(os.path.join(GOLDEN, name + ".opus"), record=True)
        assert pcm.shape == (meta["samples_per_channel"], meta["channels"]) and len(recs) == meta["celt_frames"]
        assert hashlib.sha256(pcm.tobytes()).hexdigest() == meta["reference_pcm_sha256"]
        assert sum(r["B"] > 1 for r in recs) == meta["transient_frames"]
        if name == "modeswitch":   # a walk through the modes: full-band CELT frames, 5 ms redundancy frames, a fade-out frame
            sizes = [r["coef"].shape[1] for r in recs]
            assert (sizes.count(240), sizes.count(120)) == (meta["redundancy_frames"], meta["fade_out_frames"])
            continue
        for r in recs:
            assert r["coef"].shape == (2, 960) and not r["coef"][:, :320].any() and r["coef"][:, 320:].any()
