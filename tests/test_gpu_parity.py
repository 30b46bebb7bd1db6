"""GPU parity tests (-m gpu): every call goes through the C ABI
(include/nq_celt_synth.h) of the CUDA library and is compared with the CPU
oracle / the compiled-reference fixtures.

Bar (BASELINE.json north_star): max |gpu - reference| <= 1e-5 * 32768 and
SNR >= 100 dB -- float32 path, FFT factorisation differs from kiss_fft's, so
tolerance-based, never bit-exact (conftest.assert_parity).
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, FULL_SCALE, TOL_FS, assert_parity, load_npz, snr_db
import libnyquist_b200 as nq
from oracle import port

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def synth():
    import torch
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    with nq.CeltSynth(0) as s:
        yield s


def case_names(z):
    return sorted({k.split(".")[0] for k in z.files})


def rand_batch(rng, nframes, C, p_transient=0.05, amp=1000.0):
    k = np.arange(960)
    env = amp / (1.0 + k / 60.0)
    coef = (rng.uniform(-1, 1, (nframes, C, 960)) * env).astype(np.float32)
    coef[..., 800:] = 0
    tr = (rng.uniform(size=nframes) < p_transient).astype(np.uint8)
    return coef, tr


# ---- a3.2 opus_ifft against the reference's own golden vectors --------------
@pytest.mark.parametrize("n,shift", [(480, 0), (60, 3)])
def test_opus_ifft_reference_golden_vectors(synth, n, shift):
    x = np.fromfile(os.path.join(GOLDEN, f"ifft_input_N{n}.bin"), np.float32)
    y = np.fromfile(os.path.join(GOLDEN, f"ifft_output_N{n}.bin"), np.float32)
    got = synth.opus_ifft(x, shift)
    # same relative bar as the time-domain one: 1e-5 of the vector's own full scale
    assert np.abs(got - y).max() <= 1e-5 * np.abs(y).max()
    assert snr_db(y, got) >= 100.0


@pytest.mark.parametrize("shift", [0, 1, 2, 3])
def test_opus_ifft_batch_vs_oracle(synth, shift):
    rng = np.random.default_rng(shift)
    n = 2 * (480 >> shift)
    x = (rng.standard_normal((17, n)) * 1000).astype(np.float32)
    got = synth.opus_ifft(x, shift)
    want = np.stack([port.opus_ifft(r, shift) for r in x])
    assert_parity(want, got, f"ifft shift {shift}")


# ---- a3 clt_mdct_backward, all shifts and strides --------------------------
def test_clt_mdct_backward_all_shifts_and_strides():
    z = load_npz("mdct_calls.npz")
    for shift in range(4):
        for stride in (1, 2, 4, 8):
            key = f"s{shift}_st{stride}"
            inp = z[key + ".in"].copy()
            out = z[key + ".out_before"].copy()
            nq.clt_mdct_backward(inp, out, shift, stride)
            assert np.array_equal(inp, z[key + ".in"]), "input must be left untouched"
            assert_parity(z[key + ".out_after"], out, key)


def test_clt_mdct_backward_B1_C2_and_fork_seam():
    import ctypes as C
    rng = np.random.default_rng(11)
    L = nq.load_library()
    for shift, stride in ((0, 1), (3, 8)):
        N2 = 960 >> shift
        ins = [(rng.standard_normal(N2 * stride) * 900).astype(np.float32) for _ in range(2)]
        outs0 = [(rng.standard_normal(N2 + 60) * 300).astype(np.float32) for _ in range(2)]
        want = [o.copy() for o in outs0]
        for c in range(2):
            port.clt_mdct_backward(ins[c], want[c], shift, stride)
        got = [o.copy() for o in outs0]
        nq.clt_mdct_backward_B1_C2(ins, got, shift, stride)
        for c in range(2):
            assert_parity(want[c], got[c], f"B1_C2 shift {shift} ch {c}")
        # the fork's seam, cuda/mdct_cuda.hpp:92-98 (N, sine, trig, window passed as mdct.c:223-243 does)
        got2 = [o.copy() for o in outs0]
        fp = C.POINTER(C.c_float)
        pin = (fp * 2)(*[a.ctypes.data_as(fp) for a in ins])
        pout = (fp * 2)(*[a.ctypes.data_as(fp) for a in got2])
        N = 1920 >> shift
        sine = float(np.float32(2) * np.float32(3.141592653) * np.float32(.125) / np.float32(N))
        L.processMDCTCudaB1C2(pin, pout, None, N, shift, stride, sine, 120, None)
        for c in range(2):
            assert np.array_equal(got2[c], got[c])
        got3 = outs0[0].copy()
        L.processMDCTCuda(ins[0].ctypes.data_as(fp), got3.ctypes.data_as(fp), None, N, shift, stride, sine, 120, None)
        assert np.array_equal(got3, got[0])
    L.cleanupCudaBuffers()
    L.printCudaVersion()


# ---- a1 compute_inv_mdcts, every LM / shortBlocks / C combination ------------
@pytest.mark.parametrize("LM", [0, 1, 2, 3])
@pytest.mark.parametrize("short", [False, True])
@pytest.mark.parametrize("C", [1, 2])
def test_compute_inv_mdcts_all_shapes(synth, LM, short, C):
    rng = np.random.default_rng(100 * LM + 10 * short + C)
    M = 1 << LM
    N = 120 * M
    X = (rng.standard_normal((C, N)) * 700).astype(np.float32)
    outs0 = [(rng.standard_normal(N + 60) * 200).astype(np.float32) for _ in range(C)]
    want = [o.copy() for o in outs0]
    port.compute_inv_mdcts(M if short else 0, X, want, C, LM)
    got = [o.copy() for o in outs0]
    synth.compute_inv_mdcts(M if short else 0, X, got, C, LM)
    for c in range(C):
        assert_parity(want[c], got[c], f"LM {LM} short {short} C {C} ch {c}")


# ---- batched phase 2 against compiled-reference fixtures ---------------------
@pytest.mark.parametrize("name", case_names(load_npz("synth_cases.npz")))
def test_synth_batch_vs_compiled_reference_fixture(synth, name):
    z = load_npz("synth_cases.npz")
    tail_in = z[name + ".tail_in"]
    tail_in = None if tail_in.size == 0 else tail_in
    pcm, tail = synth.synth_batch(z[name + ".coef"], z[name + ".transient"], tail_in)
    assert_parity(z[name + ".pcm"], pcm, name)
    assert_parity(z[name + ".tail_out"], tail, name + " tail")


@pytest.mark.parametrize("tag", ["reverie", "reverie60", "short"])
def test_real_frames_recorded_from_bundled_opus_files(synth, tag):
    """BASELINE configs 1-3: coefficients recorded while the reference decoded
    sb-reverie.opus / sb-reverie-60ms-frames.opus / short.opus (incl. transients)."""
    z = load_npz("real_frames.npz")
    coef, tr, out = z[tag + ".coef"], z[tag + ".transient"], z[tag + ".out"]
    pcm, _ = synth.synth_batch(np.ascontiguousarray(coef), tr, None)
    pcm = pcm.reshape(coef.shape[0], 960, 2).transpose(0, 2, 1)
    assert_parity(out[1:], pcm[1:], tag)


@pytest.mark.parametrize("C", [1, 2, 3, 4, 6, 8])
@pytest.mark.parametrize("p_tr", [0.0, 0.05, 1.0])
def test_synth_batch_many_runs_vs_oracle(synth, C, p_tr):
    """Enough frames for thousands of warp runs: exercises the run-boundary re-synthesis."""
    rng = np.random.default_rng(1000 + C)
    nframes = 5000 if C <= 2 else 1500
    coef, tr = rand_batch(rng, nframes, C, p_tr)
    tail_in = (rng.standard_normal((C, 60)) * 100).astype(np.float32)
    want, want_tail, _ = port.synth_batch(coef, tr, tail_in, nthreads=8)
    pcm, tail = synth.synth_batch(coef, tr, tail_in)
    assert_parity(want, pcm, f"C {C} p {p_tr}")
    assert_parity(want_tail, tail, "tail")


@pytest.mark.parametrize("C", [3, 5, 7])
@pytest.mark.parametrize("pairs", ["0", "1"])
def test_lone_mono_stream_frame_pairs(synth, monkeypatch, C, pairs):
    """Layouts with an odd number of channels end on a lone mono stream; in the group kernel its warp
    may take two consecutive frames of equal block type as its two channels (NQ_LONE_PAIRS forces the
    choice either way; the default depends on the group size).  Both ways must give the oracle's
    samples, and the SAME samples: the pairing only changes which warp lane computes what."""
    rng = np.random.default_rng(4000 + C)
    coef, tr = rand_batch(rng, 1200, C, 0.15)
    tail_in = (rng.standard_normal((C, 60)) * 100).astype(np.float32)
    want, want_tail, _ = port.synth_batch(coef, tr, tail_in, nthreads=8)
    monkeypatch.setenv("NQ_LONE_PAIRS", pairs)
    pcm, tail = synth.synth_batch(coef, tr, tail_in)
    assert_parity(want, pcm, f"C {C} pairs {pairs}")
    assert_parity(want_tail, tail, "tail")
    monkeypatch.setenv("NQ_LONE_PAIRS", "1" if pairs == "0" else "0")
    pcm2, tail2 = synth.synth_batch(coef, tr, tail_in)
    assert np.array_equal(pcm, pcm2) and np.array_equal(tail, tail2)


@pytest.mark.parametrize("C", [1, 2])
def test_hybrid_start_band_17_and_narrow_end_bands(synth, C):
    """the CELT layer of a hybrid frame starts at band 17
    (opus_decoder_clean.c:443-444, CELT_SET_START_BAND) -- denormalise_bands leaves freq[] zero below
    eBands[17] << LM = 40 << LM (celt_decoder_clean.c:620-636, static_modes_float.h eBands) -- and
    narrower audio bandwidths end at band 13 / 17 / 19 (zero from eBands[end] << LM up).  For the
    synthesis these are ordinary frames; long and short blocks, 10 and 20 ms (the hybrid sizes)."""
    rng = np.random.default_rng(170 + C)
    ebands = [0, 1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 14, 16, 20, 24, 28, 34, 40, 48, 60, 78, 100]
    nframes = 240
    lm = rng.choice([3, 2], nframes)
    coef, tr = rand_batch(rng, nframes, C, 0.2)
    for f in range(nframes):
        M = 1 << lm[f]
        start = rng.choice([0, 17])
        end = rng.choice([13, 17, 19, 21]) if start == 0 else rng.choice([19, 21])
        coef[f, :, :ebands[start] * M] = 0
        coef[f, :, ebands[end] * M:] = 0
    flags = (tr | ((3 - lm) << 1)).astype(np.uint8)
    want, want_tail, offs = oracle_any_size(coef, flags, None)
    import torch
    if C == 2:
        pcm, tail = synth.synth_batch_ms_torch(torch.from_numpy(coef).cuda(), torch.from_numpy(flags).cuda().reshape(-1, 1), 1, 1, None,
                                               frame_offset=torch.from_numpy(offs).cuda())
    else:
        pcm, tail = synth.synth_batch_ms_torch(torch.from_numpy(coef).cuda(), torch.from_numpy(flags).cuda().reshape(-1, 1), 1, 0, None,
                                               frame_offset=torch.from_numpy(offs).cuda())
    torch.cuda.synchronize()
    assert_parity(want, pcm.cpu().numpy(), f"hybrid-shaped frames C {C}")
    assert_parity(want_tail, tail.cpu().numpy(), "tail")


# ---- BASELINE config 4: Opus multistream layouts, interleave fused into the store pass ----
def ms_oracle(coef, tr, streams, coupled, mapping, tail_in=None):
    """Per-stream compute_inv_mdcts (oracle) + the channel routing of
    opus_multistream_decoder.c:260-299 / opus_multistream.c:57-91, in numpy."""
    nframes, D, _ = coef.shape
    dec = np.zeros((nframes * 960, D), np.float32)
    tails = np.zeros((D, 60), np.float32)
    for s in range(streams):
        rows = [2 * s, 2 * s + 1] if s < coupled else [s + coupled]
        ti = None if tail_in is None else np.ascontiguousarray(tail_in[rows])
        pcm, tl, _ = port.synth_batch(np.ascontiguousarray(coef[:, rows]), np.ascontiguousarray(tr[:, s]), ti, nthreads=4)
        dec[:, rows] = pcm
        tails[rows] = tl
    out = np.zeros((nframes * 960, len(mapping)), np.float32)
    for c, d in enumerate(mapping):
        if d != 255:
            out[:, c] = dec[:, d]
    return out, tails


MS_LAYOUTS = {
    # (streams, coupled, mapping): the Vorbis-order surround layouts of opus_multistream_encoder.c:60-69
    "mono": (1, 0, [0]),
    "stereo_mapped": (1, 1, [1, 0]),
    "surround_5.1": (4, 2, [0, 4, 1, 2, 3, 5]),
    "surround_7.1": (5, 3, [0, 6, 1, 2, 3, 4, 5, 7]),
    "dual_mono_muted_dup": (3, 1, [2, 255, 0, 0, 1, 3, 3]),
    "fourteen_mono": (14, 0, list(range(14))),
    "dual_mono_identity": (2, 0, [0, 1]),      # D = C = 2 like stereo, but two decoders with their own flags
    "five_mono_odd": (5, 0, [4, 3, 2, 1, 0]),
    "max_warps_13_coupled_2_mono": (15, 13, list(range(28))),
    # 37 output channels fed by 3 decoded ones: no store-thread count divides evenly -> general store loop
    "thirty_seven_outputs_one_stream": (1, 1, [(7 * i) % 2 if i % 5 else 255 for i in range(37)]),
}


@pytest.mark.parametrize("name", sorted(MS_LAYOUTS))
def test_multistream_layouts_fused_interleave(synth, name):
    import torch
    streams, coupled, mapping = MS_LAYOUTS[name]
    D = streams + coupled
    rng = np.random.default_rng(len(name))
    nframes = 700
    coef, _ = rand_batch(rng, nframes, D, 0.0)
    tr = (rng.uniform(size=(nframes, streams)) < 0.15).astype(np.uint8)
    tail_in = (rng.standard_normal((D, 60)) * 100).astype(np.float32)
    want, want_tail = ms_oracle(coef, tr, streams, coupled, mapping, tail_in)
    d_coef, d_tr = torch.from_numpy(coef).cuda(), torch.from_numpy(tr).cuda()
    pcm, tail = synth.synth_batch_ms_torch(d_coef, d_tr, streams, coupled, mapping, tail_in=torch.from_numpy(tail_in).cuda())
    torch.cuda.synchronize()
    assert_parity(want, pcm.cpu().numpy(), name)
    assert_parity(want_tail, tail.cpu().numpy(), name + " tail")
    # a shard that starts mid-stream from a halo frame is bit-identical to the unsharded run
    cut = 333
    part, _ = synth.synth_batch_ms_torch(d_coef[cut:].contiguous(), d_tr[cut:].contiguous(), streams, coupled, mapping,
                                         halo_coef=d_coef[cut - 1].contiguous(), halo_transient=tr[cut - 1])
    torch.cuda.synchronize()
    assert torch.equal(part, pcm[cut * 960:])


def test_multistream_rejects_bad_layouts(synth):
    import torch
    coef = torch.zeros((4, 3, 960), device="cuda")
    tr = torch.zeros((4, 2), dtype=torch.uint8, device="cuda")
    with pytest.raises(nq.NqError):
        synth.synth_batch_ms_torch(coef, tr, 2, 1, [0, 1, 3])        # decoded channel 3 does not exist
    with pytest.raises(nq.NqError) as e:
        # 29 mono streams = 15 warps (mono streams pair up): one more than a CTA has
        synth.synth_batch_ms_torch(torch.zeros((4, 29, 960), device="cuda"),
                                   torch.zeros((4, 29), dtype=torch.uint8, device="cuda"), 29, 0, list(range(29)))
    assert e.value.code == -5


# ---- frames shorter than 20 ms anywhere inside a batch ---------------------
def oracle_any_size(coef, flags, tail_in):
    """compute_inv_mdcts frame by frame (oracle), any LM: coef [nframes][C][960] (first 120<<LM
    valid), flags [nframes] (bit 0 transient, bits 1-2 = 3-LM).  Returns (pcm [nsamples][C], tail, offsets)."""
    nframes, C, _ = coef.shape
    tail = np.zeros((C, 60), np.float32) if tail_in is None else tail_in.copy()
    out, offs, pos = [], [], 0
    for f in range(nframes):
        LM = 3 - ((int(flags[f]) >> 1) & 3)
        N = 120 << LM
        X = np.ascontiguousarray(coef[f, :, :N])
        bufs = [np.concatenate([tail[c], np.zeros(N, np.float32)]) for c in range(C)]
        port.compute_inv_mdcts((1 << LM) if flags[f] & 1 else 0, X, bufs, C, LM)
        out.append(np.stack([b[:N] for b in bufs], axis=1))
        tail = np.stack([b[N:N + 60] for b in bufs])
        offs.append(pos)
        pos += N
    offs.append(pos)
    return np.concatenate(out), tail, np.array(offs, np.int64)


@pytest.mark.parametrize("C", [1, 2])
def test_batch_with_short_frames_everywhere(synth, C):
    import torch
    rng = np.random.default_rng(40 + C)
    nframes = 900
    coef, tr = rand_batch(rng, nframes, C, 0.2)
    lm = rng.choice([3, 3, 3, 2, 1, 0], nframes)
    flags = (tr | ((3 - lm) << 1)).astype(np.uint8)
    for f in range(nframes):
        coef[f, :, 120 << lm[f]:] = 7777.0     # must be ignored
    tail_in = (rng.standard_normal((C, 60)) * 100).astype(np.float32)
    want, want_tail, offs = oracle_any_size(coef, flags, tail_in)
    d_coef, d_fl, d_off = torch.from_numpy(coef).cuda(), torch.from_numpy(flags).cuda().reshape(-1, 1), torch.from_numpy(offs).cuda()
    pcm, tail = synth.synth_batch_ms_torch(d_coef, d_fl, 1, C - 1, None, tail_in=torch.from_numpy(tail_in).cuda(),
                                           frame_offset=d_off)
    torch.cuda.synchronize()
    assert_parity(want, pcm.cpu().numpy(), f"any-size C {C}")
    assert_parity(want_tail, tail.cpu().numpy(), "tail")
    # shard start on a halo frame that is itself a short frame
    cut = int(np.nonzero(lm[1:] < 3)[0][5]) + 2
    assert lm[cut - 1] < 3
    part, _ = synth.synth_batch_ms_torch(d_coef[cut:].contiguous(), d_fl[cut:].contiguous(), 1, C - 1, None,
                                         halo_coef=d_coef[cut - 1].contiguous(), halo_transient=flags[cut - 1:cut],
                                         frame_offset=(d_off[cut:] - d_off[cut]).contiguous())
    torch.cuda.synchronize()
    assert torch.equal(part, pcm[int(offs[cut]):])


def test_multistream_batch_with_short_frames(synth):
    import torch
    streams, coupled, mapping = MS_LAYOUTS["surround_7.1"]
    D = streams + coupled
    rng = np.random.default_rng(77)
    nframes = 300
    coef, _ = rand_batch(rng, nframes, D, 0.0)
    lm = rng.choice([3, 3, 2, 1, 0], nframes)
    tr = (rng.uniform(size=(nframes, streams)) < 0.2).astype(np.uint8)
    flags = (tr | ((3 - lm[:, None]) << 1)).astype(np.uint8)
    dec = []
    for s in range(streams):
        rows = [2 * s, 2 * s + 1] if s < coupled else [s + coupled]
        w, _, offs = oracle_any_size(np.ascontiguousarray(coef[:, rows]), flags[:, s], None)
        dec.append((rows, w))
    want = np.zeros((int(offs[-1]), len(mapping)), np.float32)
    for c, d in enumerate(mapping):
        for rows, w in dec:
            if d in rows:
                want[:, c] = w[:, rows.index(d)]
    pcm, _ = synth.synth_batch_ms_torch(torch.from_numpy(coef).cuda(), torch.from_numpy(flags).cuda(), streams, coupled,
                                        mapping, frame_offset=torch.from_numpy(offs).cuda())
    torch.cuda.synchronize()
    assert_parity(want, pcm.cpu().numpy(), "7.1 any-size")


def test_more_channel_pairs_than_warps_direct_variant(synth):
    """30 plain channels = 15 pairs, one more than a CTA has warps: the scattered-store variant."""
    assert nq.debug_plan(30)["mode"] == nq.MODE_DIRECT
    rng = np.random.default_rng(30)
    coef, tr = rand_batch(rng, 300, 30, 0.2)
    tail_in = (rng.standard_normal((30, 60)) * 100).astype(np.float32)
    want, want_tail, _ = port.synth_batch(coef, tr, tail_in, nthreads=8)
    pcm, tail = synth.synth_batch(coef, tr, tail_in)
    assert_parity(want, pcm, "C 30")
    assert_parity(want_tail, tail, "tail")


def test_edge_cases_and_error_codes(synth):
    rng = np.random.default_rng(5)
    # empty batch: tail passes through
    tail_in = rng.standard_normal((2, 60)).astype(np.float32)
    pcm, tail = synth.synth_batch(np.zeros((0, 2, 960), np.float32), np.zeros(0, np.uint8), tail_in)
    assert pcm.shape == (0, 2) and np.array_equal(tail, tail_in)
    # zero coefficients, zero tail -> exact zeros
    pcm, tail = synth.synth_batch(np.zeros((9, 2, 960), np.float32), np.array([0, 1] * 4 + [0], np.uint8))
    assert not pcm.any() and not tail.any()
    # bad arguments are refused, not crashed on
    L = nq.load_library()
    assert L.nq_celt_synth_batch_host(synth._h, None, None, None, None, None, 4, 2) == -1
    assert L.nq_celt_synth_batch_host(synth._h, None, None, None, None, None, 4, 0) == -1
    assert L.nq_celt_synth_batch_host(synth._h, None, None, None, None, None, -1, 2) == -1
    assert L.nq_compute_inv_mdcts(synth._h, 3, None, None, 2, 3) == -1
    with pytest.raises(nq.NqError):
        synth.compute_inv_mdcts(4, np.zeros((2, 960), np.float32), [np.zeros(1020, np.float32)] * 2, 2, 3)


# ---- device-pointer entry (what bench.py times) -----------------------------
def test_device_entry_tail_and_halo_and_chaining(synth):
    import torch
    rng = np.random.default_rng(21)
    nframes, C = 4000, 2
    coef, tr = rand_batch(rng, nframes, C, 0.1)
    want, want_tail, _ = port.synth_batch(coef, tr, None, nthreads=8)
    d_coef = torch.from_numpy(coef).cuda()
    d_tr = torch.from_numpy(tr).cuda()
    pcm, tail = synth.synth_batch_torch(d_coef, d_tr)
    torch.cuda.synchronize()
    assert_parity(want, pcm.cpu().numpy(), "device whole")
    assert_parity(want_tail, tail.cpu().numpy(), "device tail")
    # two calls chained through tail_out -> tail_in must equal one call BIT FOR BIT
    cut = 1777
    a, ta = synth.synth_batch_torch(d_coef[:cut].contiguous(), d_tr[:cut].contiguous())
    b, tb = synth.synth_batch_torch(d_coef[cut:].contiguous(), d_tr[cut:].contiguous(), tail_in=ta)
    torch.cuda.synchronize()
    assert torch.equal(torch.cat([a, b]), pcm) and torch.equal(tb, tail)
    # the same second half from a halo frame instead of a tail (multi-GPU shard start)
    b2, _ = synth.synth_batch_torch(d_coef[cut:].contiguous(), d_tr[cut:].contiguous(),
                                    halo_coef=d_coef[cut - 1].contiguous(), halo_transient=int(tr[cut - 1]))
    torch.cuda.synchronize()
    assert torch.equal(b2, b)


def test_full_size_properties_on_device(synth):
    """Size-independent checks at a large batch (1M stereo frames, 15 GB of traffic):
    exact homogeneity under power-of-two scaling, shard invariance, and spot
    parity of randomly chosen frames against the oracle."""
    import torch
    nframes, C = 1_000_000, 2
    g = torch.Generator(device="cuda").manual_seed(0x0B200)
    env = (1000.0 / (1.0 + torch.arange(960, device="cuda") / 60.0)).float()
    env[800:] = 0
    coef = (torch.rand((nframes, C, 960), generator=g, device="cuda") * 2 - 1) * env
    tr = (torch.rand(nframes, generator=g, device="cuda") < 0.028).to(torch.uint8)
    pcm, tail = synth.synth_batch_torch(coef, tr)
    pcm2, tail2 = synth.synth_batch_torch(coef * 4.0, tr)
    torch.cuda.synchronize()
    assert torch.isfinite(pcm).all()
    assert torch.equal(pcm2, pcm * 4.0) and torch.equal(tail2, tail * 4.0)
    # spot parity: 48 random windows of 3 frames, each synthesised by the oracle from its halo
    rng = np.random.default_rng(9)
    worst = 0.0
    for f in rng.integers(1, nframes - 3, 48):
        f = int(f)
        c = coef[f - 1:f + 3].cpu().numpy()
        t = tr[f - 1:f + 3].cpu().numpy()
        want, _, _ = port.synth_batch(c, t, None)
        got = pcm[f * 960:(f + 3) * 960].cpu().numpy()
        assert_parity(want[960:], got, f"frame {f}")
        worst = max(worst, float(np.abs(want[960:] - got).max()))
    print(f"worst spot error {worst / FULL_SCALE:.3e} of full scale")
    # shard invariance across 8 contiguous shards with halos (the multi-GPU decomposition)
    for s in nq.shard_plan(nframes, 8)[1:3]:
        part, _ = synth.synth_batch_torch(coef[s.f0:s.f1], tr[s.f0:s.f1], halo_coef=coef[s.halo],
                                          halo_transient=int(tr[s.halo]))
        torch.cuda.synchronize()
        assert torch.equal(part, pcm[s.f0 * 960:s.f1 * 960])


@pytest.mark.parametrize("name", ["surround_7.1", "surround_5.1", "plain_8", "plain_4", "plain_3", "five_mono_odd"])
def test_group_kernels_scheduling_invariance_at_scale(synth, name, monkeypatch):
    """The warp-specialised group kernels on batches large enough for every SM to claim runs
    dynamically (synthesis warps one frame ahead of their store warp, store warps serving several
    groups, claims travelling through the shared-memory ring): the result must not depend on how
    the work was scheduled.  Bit-identical: dynamic runs of 64 == one static run per group ==
    two chained calls (tail handed over) == itself again; plus spot parity against the oracle."""
    import torch
    if name.startswith("plain_"):
        C = int(name.split("_")[1])
        streams, coupled, mapping = (C + 1) // 2, C // 2, None
        D = C
    else:
        streams, coupled, mapping = MS_LAYOUTS[name]
        D = streams + coupled
    nframes = 90_000 if D <= 4 else 50_000
    g = torch.Generator(device="cuda").manual_seed(77)
    coef = (torch.rand((nframes, D, 960), generator=g, device="cuda") * 2 - 1) * 800.0
    ncols = streams if mapping is not None else 1
    tr = (torch.rand((nframes, ncols), generator=g, device="cuda") < 0.05).to(torch.uint8)

    def run(c, t, tail_in=None):
        if mapping is None:
            return synth.synth_batch_torch(c, t.reshape(-1), tail_in=tail_in)
        return synth.synth_batch_ms_torch(c, t, streams, coupled, mapping, tail_in=tail_in)

    pcm, tail = run(coef, tr)
    again, _ = run(coef, tr)
    monkeypatch.setenv("NQ_FRAMES_PER_RUN", "0")          # one long run per group, assigned statically
    static, tail_s = run(coef, tr)
    monkeypatch.delenv("NQ_FRAMES_PER_RUN")
    cut = nframes // 2 + 7
    a, ta = run(coef[:cut], tr[:cut])
    b, tb = run(coef[cut:], tr[cut:], tail_in=ta)
    torch.cuda.synchronize()
    assert torch.equal(again, pcm)
    assert torch.equal(static, pcm) and torch.equal(tail_s, tail)
    assert torch.equal(torch.cat([a, b]), pcm) and torch.equal(tb, tail)
    # spot parity of a few windows against the oracle (per-stream synthesis + channel routing)
    rng = np.random.default_rng(5)
    for f in rng.integers(1, nframes - 3, 4):
        f = int(f)
        c = coef[f - 1:f + 3].cpu().numpy()
        t = tr[f - 1:f + 3].cpu().numpy()
        if mapping is None:
            want, _, _ = port.synth_batch(c, t.reshape(-1), None)
        else:
            want, _ = ms_oracle(c, t, streams, coupled, mapping)
        assert_parity(want[960:], pcm[f * 960:(f + 3) * 960].cpu().numpy(), f"{name} frame {f}")


def test_multi_gpu_in_process_entry(synth):
    import torch
    rng = np.random.default_rng(33)
    coef, tr = rand_batch(rng, 3000, 2, 0.05)
    tail_in = (rng.standard_normal((2, 60)) * 100).astype(np.float32)
    want, want_tail = synth.synth_batch(coef, tr, tail_in)
    ndev = torch.cuda.device_count()
    for devices in ([0], list(range(ndev)), [0] * 3):
        pcm, tail = nq.synth_batch_multi_gpu(coef, tr, tail_in, devices)
        assert np.array_equal(pcm, want) and np.array_equal(tail, want_tail), devices


# ---- BASELINE configs 1-3: every frame of the bundled files -----------------
REAL_FILES = {"sb-reverie.opus": (11184, 314), "sb-reverie-60ms-frames.opus": (11184, 298), "short.opus": (None, None)}


@pytest.mark.parametrize("fname", sorted(REAL_FILES))
def test_whole_bundled_file_coefficient_stream(synth, fname):
    """Phase 1 = the reference's own decoder (oracle/_ref, compiled in place) recording freq[] at
    the inverse-MDCT call sites; phase 2 = ONE batched GPU call over every 20 ms stereo frame of
    the file; compared with the out_syn the reference produced for the same frames."""
    from oracle import ref
    path = os.path.join(os.path.dirname(ref.LIB_PATH), "test_data", fname)
    if not (ref.available() and os.path.exists(path)):
        pytest.skip("oracle/_ref (compiled reference + staged test_data) not present")
    pcm_ref, recs = ref.decode_file(path, record=True)
    # keep the maximal prefix-free set of consecutive LM=3 stereo frames (short.opus has one LM=0 frame)
    idx = [i for i, r in enumerate(recs) if r["nch"] == 2 and r["coef"].shape[1] == 960]
    runs, cur = [], [idx[0]]
    for a, b in zip(idx, idx[1:]):
        if b == a + 1:
            cur.append(b)
        else:
            runs.append(cur)
            cur = [b]
    runs.append(cur)
    total = 0
    for run in runs:
        if len(run) < 2:
            continue
        coef = np.stack([recs[i]["coef"] for i in run])
        tr = np.array([recs[i]["B"] == 8 for i in run], np.uint8)
        want = np.stack([recs[i]["out"] for i in run])
        pcm, _ = synth.synth_batch(np.ascontiguousarray(coef), tr, None)
        got = pcm.reshape(len(run), 960, 2).transpose(0, 2, 1)
        # frame 0 of a run has an unknown previous tail (it is the halo); all later frames are checked
        assert_parity(want[1:], got[1:], fname)
        total += len(run)
    nfr, ntr = REAL_FILES[fname]
    if nfr is not None:
        assert total == nfr and sum(r["B"] == 8 for r in recs) == ntr
