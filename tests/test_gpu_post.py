"""GPU parity tests (-m gpu) of the post stage: comb_filter x2 +
deemphasis through the C ABI (nq_celt_post_batch_device, nq_celt_decode_batch_host) against the
oracle's bit-exact restatement and the reference decoder's own PCM.

Bar: |pcm_gpu - pcm_ref| <= 1e-5 (PCM full scale is 1.0, i.e. 1e-5 of full scale) and SNR >= 100 dB.
The GPU evaluates the de-emphasis IIR as lane segments + a warp scan and contracts a*b+c into
FMAs, so parity is tolerance-based, not bit-exact.
"""
import os

import numpy as np
import pytest

from conftest import load_npz, snr_db
import libnyquist_b200 as nq
from oracle import port, ref

pytestmark = pytest.mark.gpu

PCM_TOL = 1e-5


@pytest.fixture(scope="module")
def synth():
    import torch
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    with nq.CeltSynth(0) as s:
        yield s


def assert_pcm(want, got, what):
    assert want.shape == got.shape, (what, want.shape, got.shape)
    assert np.isfinite(got).all(), what
    err = float(np.abs(got.astype(np.float64) - want).max()) if want.size else 0.0
    assert err <= PCM_TOL, f"{what}: max abs err {err:.3e}"
    if want.size and np.abs(want).max() > 0:
        assert snr_db(want, got) >= 100.0, f"{what}: SNR {snr_db(want, got):.1f} dB"
    return err


def rand_frames(rng, nframes, N=960, p_off=0.2):
    fr = np.zeros(nframes, nq.POST_FRAME_DTYPE)
    fr["N"] = N
    pitch = rng.integers(15, 1023, nframes + 2)
    pitch[rng.uniform(size=nframes + 2) < 0.2] = 15                       # shortest period: 13-sample steps
    gain = rng.choice([0.09375 * k for k in range(1, 9)], nframes + 2).astype(np.float32)
    gain[rng.uniform(size=nframes + 2) < p_off] = 0
    tap = rng.integers(0, 3, nframes + 2)
    for i in range(nframes):   # hand-over of celt_decoder_clean.c:672-683 (LM != 0: old <- cur <- new)
        fr["pitch"][i] = pitch[i + 1], pitch[i + 1], pitch[i + 2]
        fr["gain"][i] = gain[i + 1], gain[i + 1], gain[i + 2]
        fr["tapset"][i] = tap[i + 1], tap[i + 1], tap[i + 2]
    fr["pitch"][0][0], fr["gain"][0][0], fr["tapset"][0][0] = pitch[0], gain[0], tap[0]
    return fr


def test_post_fixture_short_opus_tail(synth):
    """Last 26 frames of short.opus: active post-filter, tapset changes, a transient, the final
    LM=0 frame; expected values are the reference decoder's PCM."""
    import torch
    z = load_npz("post_cases.npz")
    pcm = torch.from_numpy(z["short_tail.sig"]).cuda()
    hist, mem = synth.post_batch_torch(pcm, z["short_tail.frames"], torch.from_numpy(z["short_tail.hist_in"]).cuda(),
                                       torch.from_numpy(z["short_tail.mem_in"]).cuda())
    torch.cuda.synchronize()
    assert_pcm(z["short_tail.pcm"], pcm.cpu().numpy(), "short.opus tail")
    _, want_hist, want_mem = port.post_batch(z["short_tail.sig"], z["short_tail.frames"], z["short_tail.hist_in"],
                                             z["short_tail.mem_in"])
    assert np.abs(hist.cpu().numpy() - want_hist).max() <= 1e-5 * 32768
    assert np.abs(mem.cpu().numpy() - want_mem).max() <= 1e-5 * 32768


@pytest.mark.parametrize("C", [1, 2])
@pytest.mark.parametrize("N", [120, 240, 480, 960])
def test_post_random_side_info_vs_oracle(synth, C, N):
    import torch
    rng = np.random.default_rng(10 * N + C)
    nframes = 60
    fr = rand_frames(rng, nframes, N)
    sig = (rng.standard_normal((nframes * N, C)) * 1500).astype(np.float32)
    hist_in = (rng.standard_normal((C, nq.POST_HISTORY)) * 1500).astype(np.float32)
    mem_in = (rng.standard_normal(C) * 500).astype(np.float32)
    want, want_hist, want_mem = port.post_batch(sig, fr, hist_in, mem_in)
    pcm = torch.from_numpy(sig).cuda()
    hist, mem = synth.post_batch_torch(pcm, fr, torch.from_numpy(hist_in).cuda(), torch.from_numpy(mem_in).cuda())
    torch.cuda.synchronize()
    assert_pcm(want, pcm.cpu().numpy(), f"C {C} N {N}")
    assert np.abs(hist.cpu().numpy() - want_hist).max() <= 1e-5 * 32768
    assert np.abs(mem.cpu().numpy() - want_mem).max() <= 1e-5 * 32768
    # chaining two calls through the state is invisible (bit for bit)
    k = 23
    a = torch.from_numpy(sig[:k * N]).cuda()
    b = torch.from_numpy(sig[k * N:]).cuda()
    h1, m1 = synth.post_batch_torch(a, fr[:k], torch.from_numpy(hist_in).cuda(), torch.from_numpy(mem_in).cuda())
    h2, m2 = synth.post_batch_torch(b, fr[k:], h1, m1)
    torch.cuda.synchronize()
    assert torch.equal(torch.cat([a, b]), pcm) and torch.equal(h2, hist) and torch.equal(m2, mem)


def test_post_reset_state_and_zero_gain_is_pure_deemphasis(synth):
    import torch
    rng = np.random.default_rng(3)
    fr = np.zeros(5, nq.POST_FRAME_DTYPE)
    fr["N"] = 960   # gains 0, pitch 0 as celt_decoder_clean.c leaves them for an inactive post-filter
    sig = (rng.standard_normal((5 * 960, 2)) * 1000).astype(np.float32)
    pcm = torch.from_numpy(sig).cuda()
    synth.post_batch_torch(pcm, fr)
    torch.cuda.synchronize()
    x = sig.astype(np.float64)
    y = np.zeros_like(x)
    m = np.zeros(2)
    for i in range(len(x)):
        t = x[i] + m
        m = 0.85000610 * t
        y[i] = t / 32768
    assert_pcm(y.astype(np.float32), pcm.cpu().numpy(), "pure de-emphasis")


def test_post_multistream_mapping(synth):
    """7.1 layout: the post stage runs on the OUTPUT layout, every output channel with the side
    info of the stream that feeds it; a silent channel stays exactly zero."""
    import torch
    rng = np.random.default_rng(8)
    streams, coupled, mapping = 5, 3, [0, 6, 1, 2, 3, 4, 255, 7]
    D, nframes = streams + coupled, 40
    fr = np.stack([rand_frames(rng, nframes) for _ in range(streams)], axis=1)
    dec = (rng.standard_normal((nframes * 960, D)) * 1200).astype(np.float32)
    want = np.zeros((nframes * 960, len(mapping)), np.float32)
    for c, d in enumerate(mapping):
        if d == 255:
            continue
        s = d // 2 if d < 2 * coupled else d - coupled
        want[:, c:c + 1] = port.post_batch(np.ascontiguousarray(dec[:, d:d + 1]), np.ascontiguousarray(fr[:, s]))[0]
    sig = np.zeros_like(want)
    for c, d in enumerate(mapping):
        if d != 255:
            sig[:, c] = dec[:, d]
    pcm = torch.from_numpy(sig).cuda()
    synth.post_batch_torch(pcm, fr, streams=streams, coupled_streams=coupled, mapping=mapping)
    torch.cuda.synchronize()
    got = pcm.cpu().numpy()
    assert_pcm(want, got, "7.1")
    assert not got[:, 6].any()


@pytest.mark.parametrize("C", [1, 2])
def test_batch_of_many_files_reset_flag_and_segments(synth, C):
    """A batch that concatenates independent streams: flag bit 3 on the first frame of each (the
    synthesis zeroes its tail there), one post-stage CTA per segment, each from a reset decoder."""
    import torch
    from test_gpu_parity import rand_batch
    rng = np.random.default_rng(50 + C)
    lens = [700, 1, 2, 1300, 37]
    seg = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    nframes = int(seg[-1])
    coef, tr = rand_batch(rng, nframes, C, 0.1)
    fr = rand_frames(rng, nframes)
    flags = tr.copy()
    flags[seg[:-1]] |= 8
    want = []
    for a, b in zip(seg[:-1], seg[1:]):
        sig, _, _ = port.synth_batch(coef[a:b], tr[a:b], None, nthreads=4)
        want.append(port.post_batch(sig, fr[a:b])[0])
    want = np.concatenate(want)
    d_coef, d_fl = torch.from_numpy(coef).cuda(), torch.from_numpy(flags).cuda()
    pcm, _ = synth.synth_batch_torch(d_coef, d_fl)
    synth.post_segments_torch(pcm, fr, seg)
    torch.cuda.synchronize()
    assert_pcm(want, pcm.cpu().numpy(), f"segments C {C}")
    # the host entry reads the same flag bit: one call, same result bit for bit
    got, _ = synth.decode_batch(coef, flags, fr)
    assert np.array_equal(got, pcm.cpu().numpy())


def test_many_segments_big_batch_threaded_side_info_scan(synth):
    """A batch big enough for the side information to be validated on several host threads
    (>= 200 k frames): segments of uneven lengths incl. empty ones, mixed frame sizes.  Segments are
    independent, so the one big call must equal, bit for bit, separate calls over the two halves
    (each below the threshold: single-threaded scan); a bad record deep inside is still caught."""
    import torch
    rng = np.random.default_rng(77)
    lens = rng.integers(0, 220, 2100)
    lens[[3, 500, 2099]] = 0
    seg = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    nframes = int(seg[-1])
    assert nframes >= 200_000
    fr = rand_frames(rng, nframes)
    small = rng.uniform(size=nframes) < 0.1
    fr["N"][small] = rng.choice([120, 240, 480], int(small.sum()))
    first = np.concatenate([[0], np.cumsum(fr["N"])]).astype(np.int64)
    g = torch.Generator(device="cuda").manual_seed(5)
    sig = (torch.rand((int(first[-1]), 2), generator=g, device="cuda") * 2 - 1) * 3000.0
    whole = sig.clone()
    synth.post_segments_torch(whole, fr, seg)
    half = int(np.searchsorted(seg, nframes // 2))
    parts = sig.clone()
    for a, b in ((0, half), (half, len(seg) - 1)):
        f0, f1 = int(seg[a]), int(seg[b])
        assert f1 - f0 < 200_000
        view = parts[int(first[f0]):int(first[f1])]
        synth.post_segments_torch(view, fr[f0:f1], seg[a:b + 1] - seg[a])
    torch.cuda.synchronize()
    assert torch.equal(whole, parts)
    bad = fr.copy()
    bad["tapset"][nframes - 12345] = 0, 7, 0
    with pytest.raises(nq.NqError):
        synth.post_segments_torch(sig.clone(), bad, seg)


def test_post_rejects_bad_side_info(synth):
    import torch
    pcm = torch.zeros((960, 2), device="cuda")
    fr = np.zeros(1, nq.POST_FRAME_DTYPE)
    fr["N"] = 960
    fr["gain"][0] = 0.5, 0.5, 0.5
    fr["pitch"][0] = 14, 100, 100          # below COMBFILTER_MINPERIOD with a live gain
    with pytest.raises(nq.NqError):
        synth.post_batch_torch(pcm, fr)
    fr["pitch"][0] = 100, 100, 100
    fr["tapset"][0] = 0, 3, 0
    with pytest.raises(nq.NqError):
        synth.post_batch_torch(pcm, fr)
    fr["tapset"][0] = 0, 0, 0
    fr["N"] = 961
    with pytest.raises(nq.NqError):
        synth.post_batch_torch(torch.zeros((961, 2), device="cuda"), fr)


# ---- BASELINE configs 1-3 end to end: decoded PCM of every bundled file ----------------------
CHECKSUMS = {"sb-reverie.opus": (403, 21472602), "sb-reverie-60ms-frames.opus": (719, 21472602), "short.opus": (22, 421930)}


@pytest.mark.parametrize("fname", sorted(CHECKSUMS))
def test_whole_file_pcm_vs_reference_decoder(synth, fname):
    """Phase 1 = the reference's own entropy decoder (oracle/_ref, compiled in place) recording
    freq[] and the post-filter parameters; phase 2 = GPU synthesis + post stage over the whole
    file.  The result must match the reference decoder's PCM (north_star: within 1e-5) and pass
    the reference's own acceptance test, examples/src/Main.cpp:131-149 (int(sum), length)."""
    import torch
    path = os.path.join(os.path.dirname(ref.LIB_PATH), "test_data", fname)
    if not (ref.available() and os.path.exists(path)):
        pytest.skip("oracle/_ref (compiled reference + staged test_data) not present")
    pcm_ref, recs = ref.decode_file(path, record=True)
    pre_skip, gain = ref.header_info()
    assert gain == 0
    frames = port.post_frames_from_records(recs)
    assert all(r["nch"] == 2 for r in recs)
    # short.opus ends with one 2.5 ms (LM=0) frame: rows are padded to 960, the flag byte says LM
    coef = np.zeros((len(recs), 2, 960), np.float32)
    flags = np.zeros(len(recs), np.uint8)
    for i, r in enumerate(recs):
        N = r["coef"].shape[1]
        coef[i, :, :N] = r["coef"]
        LM = {120: 0, 240: 1, 480: 2, 960: 3}[N]
        flags[i] = (1 if r["B"] > 1 else 0) | ((3 - LM) << 1)
    got, _ = synth.decode_batch(coef, flags, frames)          # ONE call: coefficients in, PCM out
    got = got[pre_skip:pre_skip + len(pcm_ref)]     # opusfile: pre-skip and end trim (opusfile.c:2673-2721)
    err = assert_pcm(pcm_ref, got, fname)
    # the reference's acceptance test: de-interleave, float sum in order
    total = np.float32(0)
    flat = np.concatenate([got[:, 0], got[:, 1]])
    s = float(np.cumsum(flat, dtype=np.float32)[-1])
    print(f"{fname}: max |err| {err:.2e}, sum {s:.4f}, len {flat.size}")
    assert (int(s), flat.size) == CHECKSUMS[fname]


def test_surround8_whole_file_multistream_decode(synth):
    """BASELINE config 4 (stand-in for the missing Rachel8ch.opus, made with the reference's own
    surround encoder): phase 1 = the reference decoder recording all 5 streams, phase 2 = ONE GPU
    call -- 5 streams in 4 warps, the 7.1 channel mapping fused into the synthesis store pass, post
    stage per output channel -- against the reference decoder's own 8-channel PCM."""
    from conftest import GOLDEN
    from ms_helpers import multistream_batch
    path = os.path.join(GOLDEN, "surround8.opus")
    if not ref.available():
        pytest.skip("oracle/_ref (compiled reference) not present")
    pcm_ref, recs = ref.decode_file(path, record=True)
    ch, streams, coupled, mapping = ref.layout_info()
    pre_skip, gain = ref.header_info()
    assert (ch, streams, coupled, gain) == (8, 5, 3, 0)
    coef, flags, frames = multistream_batch(recs, streams, coupled)
    got, _ = synth.decode_batch(coef, flags, frames, streams=streams, coupled_streams=coupled, mapping=mapping)
    err = assert_pcm(pcm_ref, got[pre_skip:pre_skip + len(pcm_ref)], "surround8.opus")
    print(f"surround8.opus: {coef.shape[0]} frames x 5 streams, max |err| {err:.2e}")
