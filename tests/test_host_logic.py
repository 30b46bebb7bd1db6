"""CPU tests of everything in the product that does not need a GPU: the C-ABI
library loads and exports every symbol of include/nq_celt_synth.h, the
host-built tables, the kernel design (lane-accurate model vs the oracle), the
in-register DFT codelets (run on the host), and the shard planner."""
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

from conftest import ROOT, assert_parity, load_npz
import libnyquist_b200 as nq
from oracle import port
from kernel_model import WarpModel


def test_library_builds_loads_and_exports_every_declared_symbol():
    nq.build_library()
    L = nq.load_library()
    hdr = open(os.path.join(ROOT, "include", "nq_celt_synth.h")).read()
    declared = set(re.findall(r"^NQ_API[^;(]*?\b(\w+)\s*\(", hdr, re.M))
    assert declared, "no NQ_API declarations found"
    assert declared == set(nq.EXPORTED_SYMBOLS), declared ^ set(nq.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(L, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", nq.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    assert declared <= exported


def test_library_contains_only_sm_100a_code_and_uses_tma():
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    nq.build_library()
    elf = subprocess.run(["cuobjdump", "-lelf", nq.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in elf and not re.search(r"sm_(?!100a)\d+", elf), elf
    sass = subprocess.run(["cuobjdump", "-sass", nq.LIB_PATH], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass, "fast kernel should stage coefficient rows with TMA bulk copies"
    assert "SHFL" in sass and "SYNCS" in sass


def test_no_compute_without_gpu_is_loud():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(nq.NqError):
        nq.CeltSynth(0)


def test_host_tables_match_reference_formulas():
    t = nq.debug_tables()
    ref = load_npz("ref_tables.npz")
    # window / trig are regenerated from the reference's formulas (modes.c:374, mdct.c:99);
    # the reference ships 8-digit literals, so allow one float32 ulp.
    assert np.abs(t["window"].astype(np.float64) - ref["window120"]).max() <= 6e-8
    assert np.abs(t["trig"].astype(np.float64) - ref["trig481"]).max() <= 6e-8
    assert np.array_equal(t["window"], port.tables()["window120"]) or \
        np.abs(t["window"] - port.tables()["window120"]).max() <= 6e-8
    tl = t["t_long"][..., 0] + 1j * t["t_long"][..., 1]
    assert np.all(np.abs(np.abs(tl[:, :30]) - 1) < 1e-6)     # unit rotations times (1+s^2) ~ 1 + 1.7e-7
    ts = t["t_short"][..., 0] + 1j * t["t_short"][..., 1]
    s = float(np.float32(2) * np.float32(3.141592653) * np.float32(.125) / np.float32(240))
    assert np.allclose(np.abs(ts), 1 + s * s, atol=2e-7)      # the reference's sin(x)~x gain, SURVEY.md 3.3


@pytest.mark.parametrize("name", ["stereo_long", "stereo_short", "stereo_mixed_tail", "mono_mixed_tail",
                                  "three_ch_mixed", "single_frame_long", "single_frame_short"])
def test_kernel_design_model_matches_reference_fixtures(name):
    z = load_npz("synth_cases.npz")
    model = WarpModel(nq.debug_tables())
    tail_in = z[name + ".tail_in"]
    tail_in = None if tail_in.size == 0 else tail_in
    pcm, tail = model.synth(z[name + ".coef"].astype(np.float64), z[name + ".transient"], tail_in)
    assert_parity(z[name + ".pcm"], pcm, name)
    assert_parity(z[name + ".tail_out"], tail, name + " tail")


def test_codelets_on_host():
    nvcc = shutil.which("nvcc")
    if nvcc is None:
        pytest.skip("nvcc not on PATH")
    exe = "/tmp/nq_host_codelet_check"
    subprocess.run([nvcc, "-O2", "-Wno-deprecated-gpu-targets", "-o", exe,
                    os.path.join(ROOT, "tests", "host_codelet_check.cu")], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_shard_plan_is_a_partition_with_one_frame_halos():
    for nframes in (0, 1, 7, 8, 1000, 10_000_000):
        for world in (1, 2, 3, 4, 8):
            plan = nq.shard_plan(nframes, world)
            assert plan[0].f0 == 0 and plan[-1].f1 == nframes
            for a, b in zip(plan, plan[1:]):
                assert a.f1 == b.f0
            sizes = [s.nframes for s in plan]
            assert max(sizes) - min(sizes) <= 1
            for s in plan:
                assert s.halo == (s.f0 - 1 if s.f0 > 0 and s.nframes > 0 else None)


def test_sharded_synthesis_equals_unsharded_oracle():
    """Shard + halo + gather reproduces the unsharded result (oracle as stand-in compute)."""
    rng = np.random.default_rng(3)
    nframes, C = 23, 2
    coef = (rng.standard_normal((nframes, C, 960)) * 500).astype(np.float32)
    tr = (rng.uniform(size=nframes) < 0.4).astype(np.uint8)
    tail_in = (rng.standard_normal((C, 60)) * 50).astype(np.float32)
    want, want_tail, _ = port.synth_batch(coef, tr, tail_in)
    for world in (2, 3, 5):
        parts = []
        tail = None
        for s in nq.shard_plan(nframes, world):
            if s.halo is None:
                pcm, tail, _ = port.synth_batch(coef[s.f0:s.f1], tr[s.f0:s.f1], tail_in)
            else:   # re-synthesise the halo frame only for its tail
                _, halo_tail, _ = port.synth_batch(coef[s.halo:s.halo + 1], tr[s.halo:s.halo + 1], None)
                pcm, tail, _ = port.synth_batch(coef[s.f0:s.f1], tr[s.f0:s.f1], halo_tail)
            parts.append(pcm)
        got = np.concatenate(parts)
        assert np.array_equal(got, want) and np.array_equal(tail, want_tail)


# ---- launch planning per channel layout (pure host code behind nq_celt_debug_plan) -------------
def test_plan_plain_layouts():
    st = nq.debug_plan(2)
    assert st["mode"] == nq.MODE_STEREO and st["post_ctas"] == 1 and st["post_ctas_two_channel"] == 1
    # stereo: short runs claimed dynamically by the 148 x 14 resident warps
    # ... the last half wave's worth of frames (148 x 14 x 64 / 2) in runs of 16, so the launch ends evenly
    assert st["frames_per_run"] == 64 and st["runs"] == 14589 + 4144
    mono = nq.debug_plan(1)
    assert mono["mode"] == nq.MODE_MONO and mono["post_ctas_two_channel"] == 0
    assert mono["frames_per_run"] == 128
    c8 = nq.debug_plan(8)
    assert (c8["mode"], c8["warps_per_group"], c8["groups_per_cta"]) == (nq.MODE_GROUP, 4, 3)
    # warp-specialised groups: 4 synthesis warps + 1 store warp each, 3 groups = 15 warps per CTA
    assert c8["store_threads"] == 32 and c8["store_shape"] == 0 and not c8["paired_mono"]
    assert c8["frames_per_run"] == 64 and c8["post_ctas"] == 4
    c3 = nq.debug_plan(3)
    assert (c3["warps_per_group"], c3["groups_per_cta"]) == (2, 6)
    assert c3["store_threads"] == 30 and c3["store_shape"] == 1      # 4*T2 must be a multiple of C = 3
    c6 = nq.debug_plan(6)
    assert c6["store_threads"] == 30 and c6["store_shape"] == 0      # rows of 6 floats: either half of a float4 is one pair
    # tiny batches: runs never shorter than 8 frames
    small = nq.debug_plan(2, nframes=100)
    assert small["frames_per_run"] == 8 and small["runs"] == 13


@pytest.mark.parametrize("C", [1, 2, 3, 8])
def test_runs_tile_the_batch(C):
    """Every frame belongs to exactly one run, runs are in order, none is empty; a dynamically claimed
    launch ends on quarter-length runs (the last half wave's worth of frames)."""
    for nframes in (1, 7, 8, 9, 100, 2071, 2072 * 8, 2072 * 8 + 1, 123_457, 397_823, 397_824, 1_000_000, 1_250_000, 10_000_000):
        first = nq.debug_runs(C, nframes)
        assert first[0] == 0 and first[-1] == nframes
        lens = np.diff(first)
        assert (lens > 0).all(), (C, nframes)
        big = int(lens[0])
        assert lens.max() == big
        if len(lens) > 1:
            assert (lens[:-1] == big).all() or set(lens[:-1].tolist()) <= {big, big // 4}
            small = np.flatnonzero(lens[:-1] != big)
            if small.size:   # the small runs are the tail of the list, and there is about half a wave of frames in them
                assert small[0] + small.size == len(lens) - 1 or (lens[small[0]:] <= big // 4).all()
                assert (lens[small[0]:] <= big // 4).all()


def test_plan_multistream_layouts():
    s71 = nq.debug_plan(8, 5, 3, [0, 6, 1, 2, 3, 4, 5, 7])
    # 3 coupled streams + the 2 mono streams sharing one warp = 4 synthesis warps, 3 groups (+ 3 store warps) per CTA
    assert (s71["mode"], s71["warps_per_group"], s71["groups_per_cta"], s71["paired_mono"]) == (nq.MODE_GROUP, 4, 3, 1)
    assert s71["decoded_channels"] == 8 and s71["store_shape"] == 1 and not s71["identity"]
    # post stage: (0) (6) (1) (2,3) (4,5) (7): the mapping separates L and R of stream 0
    assert s71["post_ctas"] == 6 and s71["post_ctas_two_channel"] == 2
    s51 = nq.debug_plan(6, 4, 2, [0, 4, 1, 2, 3, 5])
    assert (s51["warps_per_group"], s51["groups_per_cta"], s51["paired_mono"]) == (3, 4, 1)
    muted = nq.debug_plan(7, 3, 1, [2, 255, 0, 0, 1, 3, 3])
    assert muted["store_shape"] == 2 and muted["post_ctas"] == 5      # the silent channel gets no post CTA
    ident = nq.debug_plan(2, 1, 1, [0, 1])
    assert ident["mode"] == nq.MODE_STEREO and ident["identity"]
    dual = nq.debug_plan(2, 2, 0, [0, 1])
    assert (dual["mode"], dual["warps_per_group"], dual["paired_mono"]) == (nq.MODE_GROUP, 1, 1)
    seven = nq.debug_plan(14, 7, 7, list(range(14)))
    assert (seven["warps_per_group"], seven["groups_per_cta"]) == (7, 2)   # 14 synthesis + 2 store warps
    odd = nq.debug_plan(37, 1, 1, [(7 * i) % 2 if i % 5 else 255 for i in range(37)])
    assert (odd["mode"], odd["store_threads"], odd["store_shape"]) == (nq.MODE_GROUP, 0, 3)   # general store loop
    with pytest.raises(nq.NqError) as e:
        nq.debug_plan(29, 29, 0, list(range(29)))                      # 15 warps: one more than a CTA has
    assert e.value.code == -5
    with pytest.raises(nq.NqError) as e:
        nq.debug_plan(3, 2, 1, [0, 1, 3])                               # decoded channel 3 does not exist
    assert e.value.code == -1


# ---- reference-side integration library (integration/, built where /root/reference exists) -----
TWOPHASE = os.path.join(ROOT, "integration", "_build", "libnyquist_twophase.so")


@pytest.mark.skipif(not os.path.exists(TWOPHASE), reason="integration/_build not built (needs /root/reference)")
def test_two_phase_library_loads_and_refuses_to_decode_without_a_gpu():
    """nqr::NyquistIO::Load of the two-phase build has no CPU synthesis to fall back to: without a
    usable B200 it must fail loudly (exception -> -1 from the shim), never return audio."""
    import ctypes as C
    import torch
    L = C.CDLL(TWOPHASE)
    for sym in ("nq_twophase_load", "nq_twophase_free", "nq_twophase_last_timing", "nq_phase1_frame_tap", "nq_phase1_note_silk"):
        assert hasattr(L, sym), sym
    if torch.cuda.is_available():
        pytest.skip("this check is about the GPU-less case")
    path = os.path.join(ROOT, "oracle", "_ref", "test_data", "short.opus")
    if not os.path.exists(path):
        pytest.skip("bundled short.opus not staged")
    L.nq_twophase_load.argtypes = [C.c_char_p, C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.c_size_t),
                                   C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double)]
    p, n, ch, sr = C.POINTER(C.c_float)(), C.c_size_t(0), C.c_int(0), C.c_int(0)
    assert L.nq_twophase_load(path.encode(), C.byref(p), C.byref(n), C.byref(ch), C.byref(sr), None) == -1
    assert not p


def test_overlay_files_contain_no_reference_code():
    """The phase-1 overlays are macro redirections + #include_next, nothing else."""
    for rel in ("integration/overlay/opus/celt/celt_decoder_clean.c", "integration/overlay/opus/libopus/src/opus_decoder_clean.c",
                "oracle/ref_overlay/opus/celt/celt_decoder_clean.c"):
        src = open(os.path.join(ROOT, rel)).read()
        code = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        lines = [l.strip() for l in code.splitlines() if l.strip()]
        joined = []
        for l in lines:           # glue continuation lines
            if joined and joined[-1].endswith("\\"):
                joined[-1] = joined[-1][:-1] + " " + l
            else:
                joined.append(l)
        assert any(l.startswith("#include_next") for l in joined), rel
        for l in joined:
            assert l.startswith("#") or l.startswith("void nqref_tap") or l.startswith("int ") or l.startswith("const float") or l.endswith(";"), (rel, l)
        assert len(joined) < 45, rel   # (preprocessor lines only, see the loop above)


# ---- frame sink: argument checking is host code (no device needed) ------------------------------
def test_sink_push_validation_without_a_device():
    import ctypes as C
    L = nq.load_library()
    L.nq_celt_sink_push_at.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64]
    sink = nq.FrameSink(2, 1, 1, [0, 1])
    freq = np.zeros((2, 960), np.float32)
    post = np.zeros(1, nq.POST_FRAME_DTYPE)
    post["N"] = 960
    # (a valid push page-locks a block of host memory, which needs the CUDA driver: the GPU tests do that)
    with pytest.raises(nq.NqError):                      # stream index
        sink.push(1, freq, 0, post[0])
    with pytest.raises(nq.NqError):                      # a coupled stream pushes two channels
        sink.push(0, freq[:1], 0, post[0])
    with pytest.raises(nq.NqError):                      # shortBlocks must be 0 or 1 << LM
        sink.push(0, freq, 4, post[0])
    bad = post.copy()
    bad["N"] = 480
    with pytest.raises(nq.NqError):                      # post->N must say the frame size
        sink.push(0, freq, 0, bad[0])
    # a frame with a place of its own (files that switch coding modes) needs streaming mode
    rc = L.nq_celt_sink_push_at(sink._h, 0, freq.ctypes.data, 2, 960, 0, post.ctypes.data, 4800)
    assert rc != nq.NQ_OK and b"streaming" in L.nq_celt_sink_last_error(sink._h)
    assert L.nq_celt_sink_side_count(sink._h) == 0
    assert sink.pending_frames == 0
    sink.reset()
    L.nq_celt_sink_reset_stream.argtypes = [C.c_void_p, C.c_int]
    L.nq_celt_sink_reset_stream(sink._h, 0)
    L.nq_celt_sink_reset_stream(sink._h, 7)              # out of range: ignored
    sink.close()
    # multistream sinks take no frames with a place of their own at all
    ms = nq.FrameSink(8, 5, 3, [0, 6, 1, 2, 3, 4, 5, 7])
    rc = L.nq_celt_sink_push_at(ms._h, 0, freq.ctypes.data, 2, 960, 0, post.ctypes.data, 0)
    assert rc != nq.NQ_OK and b"single-stream" in L.nq_celt_sink_last_error(ms._h)
    ms.close()


# ---- bench.py's own checker must be able to fail ------------------------------------------------
def test_bench_spot_check_detects_a_wrong_frame():
    """bench.py compares three frames of every rank's timed output with the oracle (first frame after
    the halo, one in the middle, the last but one).  Fed the oracle's own output it reports ~0; with
    one of those frames damaged, or with the halo ignored, it reports the damage."""
    import torch
    import bench
    from oracle import port
    rng = np.random.default_rng(12)
    n = 40
    coef = (rng.standard_normal((n + 1, 2, 960)) * 300).astype(np.float32)
    tr = (rng.uniform(size=n + 1) < 0.2).astype(np.uint8)
    tr[0] = 0                                            # the halo frame is a long block (halo_transient=0 in bench.py)
    full, _, _ = port.synth_batch(coef, tr, None)        # frame 0 is the halo of the "rank"
    pcm = torch.from_numpy(full[960:].copy())
    c, t, halo = torch.from_numpy(coef[1:]), torch.from_numpy(tr[1:]), torch.from_numpy(coef[0])
    assert bench.oracle_spot_check(np, c, t, pcm, halo, n) <= 1e-7
    no_halo, _, _ = port.synth_batch(coef[1:], tr[1:], None)
    assert bench.oracle_spot_check(np, c, t, torch.from_numpy(no_halo), None, n) <= 1e-7
    assert bench.oracle_spot_check(np, c, t, torch.from_numpy(no_halo), halo, n) > 1e-4   # first frame lacks the halo's tail
    for f in (0, n // 3, n - 2):
        bad = pcm.clone()
        bad[f * 960 + 17, 1] += 40.0
        assert bench.oracle_spot_check(np, c, t, bad, halo, n) > 1e-3
