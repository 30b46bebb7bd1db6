#!/usr/bin/env python
"""Regenerates tests/golden/*.npz|*.bin from the reference.  Runs ONLY where
/root/reference exists (this container); the GPU box uses the committed files.

    make -C oracle ref && python tests/golden/make_golden.py

Outputs (all small):
  ifft_{input,output}_N{480,60}.bin  byte copies of the reference's own golden
                                     vectors (test_data/)
  ref_tables.npz     the reference's static tables (static_modes_float.h)
  synth_cases.npz    compute_inv_mdcts outputs of the COMPILED REFERENCE on
                     seeded synthetic batches: long / short / mixed, mono /
                     stereo / 3ch / 8ch, zero and non-zero initial tail
  mdct_calls.npz     single clt_mdct_backward calls for every (shift, stride)
  real_frames.npz    runs of consecutive frames recorded at the inverse-MDCT
                     call sites while the reference decodes the bundled
                     sb-reverie.opus / sb-reverie-60ms-frames.opus / short.opus
                     (coefficients in, out_syn out), incl. transient frames
  surround8.opus     BASELINE config 4 stand-in for the missing Rachel8ch.opus: a 7.1 Ogg Opus file made
                     with the reference's own surround encoder from seeded synthetic audio
                     (+ surround8.json: what the reference decoder makes of it)
  hybrid.opus, silk_stereo.opus
                     SURVEY.md 8(f) row 4: a hybrid file (every packet SILK below 8 kHz + a CELT
                     layer from band 17 up) and a stereo SILK-only one, made with the reference's
                     own encoder forced into the mode (OPUS_SET_FORCE_MODE) from seeded synthetic
                     speech-like audio (+ modes.json: what the reference decoder makes of them)
  post_cases.npz     SURVEY.md 8(f) row 1: single comb_filter / deemphasis calls of
                     the compiled reference on seeded inputs, and the LAST 26 frames
                     of short.opus (active post-filter, tapset changes, a transient
                     and the final 2.5 ms LM=0 frame): out_syn in, side info, and the
                     reference decoder's final PCM out.  The incoming filter state of
                     that run comes from the oracle, after the generator has checked
                     that the oracle reproduces the WHOLE file's PCM bit-exactly.
"""
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

REF = "/root/reference"


def synth_coef(rng, nframes, C, amp=1000.0):
    """Random spectra shaped like real data: zero above bin 800
    (celt_decoder_clean.c:628-636), band-decaying amplitude."""
    k = np.arange(960)
    env = amp / (1.0 + k / 60.0)
    x = rng.uniform(-1, 1, (nframes, C, 960)) * env
    x[..., 800:] = 0
    return x.astype(np.float32)


def main():
    for n in (480, 60):
        for kind in ("input", "output"):
            shutil.copyfile(f"{REF}/test_data/ifft_{kind}_N{n}.bin", f"{HERE}/ifft_{kind}_N{n}.bin")

    np.savez_compressed(f"{HERE}/ref_tables.npz", **ref.tables())

    rng = np.random.default_rng(0x0B200)
    cases = {}

    def add(name, nframes, C, transient, tail):
        coef = synth_coef(rng, nframes, C)
        tail_in = None if not tail else rng.uniform(-500, 500, (C, 60)).astype(np.float32)
        pcm, tail_out, _ = ref.synth_batch(coef, transient, tail_in)
        cases[f"{name}.coef"] = coef
        cases[f"{name}.transient"] = np.asarray(transient, np.uint8)
        cases[f"{name}.tail_in"] = np.zeros((0,), np.float32) if tail_in is None else tail_in
        cases[f"{name}.pcm"] = pcm
        cases[f"{name}.tail_out"] = tail_out

    add("stereo_long", 6, 2, [0] * 6, False)
    add("stereo_short", 5, 2, [1] * 5, False)
    add("stereo_mixed_tail", 9, 2, [0, 1, 0, 0, 1, 1, 0, 1, 0], True)
    add("mono_mixed_tail", 7, 1, [0, 1, 1, 0, 0, 1, 0], True)
    add("three_ch_mixed", 5, 3, [1, 0, 0, 1, 0], True)
    add("eight_ch_mixed", 4, 8, [0, 1, 0, 1], True)
    add("single_frame_long", 1, 2, [0], True)
    add("single_frame_short", 1, 2, [1], False)
    np.savez_compressed(f"{HERE}/synth_cases.npz", **cases)

    calls = {}
    for shift in range(4):
        N2 = 960 >> shift
        for stride in (1, 2, 4, 8):
            inp = rng.uniform(-2000, 2000, N2 * stride).astype(np.float32)
            out = rng.uniform(-3000, 3000, N2 + 60).astype(np.float32)
            calls[f"s{shift}_st{stride}.in"] = inp.copy()
            calls[f"s{shift}_st{stride}.out_before"] = out.copy()
            ref.clt_mdct_backward(inp, out, shift, stride)
            calls[f"s{shift}_st{stride}.out_after"] = out
    np.savez_compressed(f"{HERE}/mdct_calls.npz", **calls)

    real = {}
    for fname, tag in (("sb-reverie.opus", "reverie"), ("sb-reverie-60ms-frames.opus", "reverie60"),
                       ("short.opus", "short")):
        pcm, recs = ref.decode_file(f"{REF}/test_data/{fname}", record=True)
        # float sum in sample order after de-interleave, as examples/src/Main.cpp:131-144
        real[f"{tag}.n_records"] = np.int64(len(recs))
        real[f"{tag}.pcm_len"] = np.int64(pcm.size)
        stereo20 = [i for i, r in enumerate(recs) if r["nch"] == 2 and r["coef"].shape[1] == 960]
        tr = np.array([recs[i]["B"] == 8 for i in stereo20])
        real[f"{tag}.n_transient"] = np.int64(tr.sum())
        # a run of 12 consecutive LM=3 stereo frames around the first transient
        first = int(np.argmax(tr)) if tr.any() else 0
        lo = max(first - 5, 1)
        idx = stereo20[lo - 1:lo + 12]          # one extra leading frame = the halo
        assert all(b - a == 1 for a, b in zip(idx, idx[1:])), "records not consecutive"
        real[f"{tag}.coef"] = np.stack([recs[i]["coef"] for i in idx])          # [13][2][960]
        real[f"{tag}.transient"] = np.array([recs[i]["B"] == 8 for i in idx], np.uint8)
        real[f"{tag}.out"] = np.stack([recs[i]["out"] for i in idx])            # [13][2][960]
    np.savez_compressed(f"{HERE}/real_frames.npz", **real)

    from oracle import port
    post = {}
    for i, (T0, T1, g0, g1, ts0, ts1, N) in enumerate([(15, 15, 0.75, 0.75, 0, 0, 840), (1022, 15, 0.5625, 0.09375, 2, 1, 840),
                                                      (163, 650, 0.28125, 0.0, 0, 2, 840), (300, 77, 0.0, 0.375, 1, 0, 120),
                                                      (40, 41, 0.0, 0.0, 0, 0, 120), (100, 100, 0.46875, 0.46875, 1, 1, 120)]):
        buf = rng.uniform(-3000, 3000, 1026 + N).astype(np.float32)
        post[f"comb{i}.args"] = np.array([T0, T1, N, g0, g1, ts0, ts1], np.float64)
        post[f"comb{i}.before"] = buf.copy()
        ref.comb_filter(buf, 1026, T0, T1, N, g0, g1, ts0, ts1)
        post[f"comb{i}.after"] = buf
    x = rng.uniform(-6000, 6000, (2, 960)).astype(np.float32)
    mem = np.array([123.5, -77.25], np.float32)
    post["deemph.x"], post["deemph.mem_in"] = x, mem.copy()
    post["deemph.pcm"] = ref.deemphasis(x, mem)
    post["deemph.mem_out"] = mem
    for fname in ("short.opus", "sb-reverie.opus", "sb-reverie-60ms-frames.opus"):
        pcm, recs = ref.decode_file(f"{REF}/test_data/{fname}", record=True)
        pre_skip, gain = ref.header_info()
        assert gain == 0
        sig = np.concatenate([r["out"].T for r in recs], axis=0)
        frames = port.post_frames_from_records(recs)
        full, _, _ = port.post_batch(sig, frames)
        # opusfile drops pre_skip samples at the head and trims the tail (opusfile.c:2673-2721)
        assert np.array_equal(full[pre_skip:pre_skip + len(pcm)].view(np.uint32), pcm.view(np.uint32)), fname
        if fname == "short.opus":
            k = len(recs) - 26
            n0 = int(frames["N"][:k].sum())
            _, hist, mem = port.post_batch(sig[:n0], frames[:k])
            post["short_tail.sig"] = sig[n0:]
            post["short_tail.frames"] = frames[k:]
            post["short_tail.hist_in"], post["short_tail.mem_in"] = hist, mem
            post["short_tail.pcm"] = full[n0:]                      # == reference PCM where the file has it
            post["short_tail.n_in_file"] = np.int64(pre_skip + len(pcm) - n0)
            post["short_tail.transient"] = np.array([r["B"] == 8 for r in recs[k:]], np.uint8)
            post["short_tail.coef_first"] = recs[k]["coef"]
    np.savez_compressed(f"{HERE}/post_cases.npz", **post)

    # BASELINE config 4: the reference mount lacks test_data/Rachel8ch.opus (.MISSING_LARGE_BLOBS), so an
    # 8-channel (7.1: 3 coupled + 2 mono streams) Ogg Opus file is made with the reference's own surround
    # ENCODER + libogg (oracle/ref_harness.c nqref_encode_surround) from seeded synthetic audio: tones,
    # noise and percussive bursts at channel-dependent times, so every stream switches blocks on its own.
    fs, secs, ch = 48000, 6, 8
    n = (fs * secs // 960) * 960
    t = np.arange(n) / fs
    r8 = np.random.default_rng(8)
    pcm8 = np.zeros((n, ch), np.float32)
    for c in range(ch):
        f0 = 110.0 * (c + 2)
        x = 0.25 * np.sin(2 * np.pi * f0 * t) + 0.1 * np.sin(2 * np.pi * f0 * 2.01 * t) + 0.02 * r8.standard_normal(n)
        for k in range(12):
            p0 = int((0.3 + 0.45 * k + 0.037 * c) * fs)
            if p0 + 2000 < n:
                x[p0:p0 + 2000] += 0.6 * r8.standard_normal(2000) * np.exp(-np.arange(2000) / 300.0)
        pcm8[:, c] = x
    data = ref.encode_surround(np.clip(pcm8, -0.95, 0.95).astype(np.float32), 512000)
    open(f"{HERE}/surround8.opus", "wb").write(data)
    out8, recs8 = ref.decode_bytes(data, record=True)
    assert ref.layout_info() == (8, 5, 3, [0, 6, 1, 2, 3, 4, 5, 7]) and out8.shape == (n - 120, 8)
    import hashlib
    import json
    json.dump({"samples_per_channel": int(out8.shape[0]), "channels": 8, "records": len(recs8),
               "transient_records": int(sum(r["B"] == 8 for r in recs8)),
               "reference_pcm_sha256": hashlib.sha256(out8.tobytes()).hexdigest()},
              open(f"{HERE}/surround8.json", "w"), indent=1)
    make_mode_files()
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))


def make_mode_files():
    import hashlib
    import json
    """hybrid.opus, silk_stereo.opus, modeswitch.opus + modes.json (python make_golden.py --modes-only
    regenerates just these)."""
    fs = 48000
    # ---- hybrid / SILK-only files (the coding modes of opus_decode_frame besides CELT-only) ----
    n = 960 * 150
    t = np.arange(n) / fs
    rs = np.random.default_rng(0)
    f0 = 140 + 30 * np.sin(2 * np.pi * 0.7 * t)
    ph = np.cumsum(2 * np.pi * f0 / fs)
    voiced = sum(np.sin(k * ph) / k for k in range(1, 30))
    env = (0.5 + 0.5 * np.sin(2 * np.pi * 2.1 * t)) ** 2
    sig = 0.25 * env * voiced / np.abs(voiced).max() + 0.03 * rs.standard_normal(n) * (1 - env)
    sig += 0.02 * env * rs.standard_normal(n)       # content above 8 kHz for the CELT layer
    speech = np.stack([sig, np.roll(sig, 37) * 0.8], 1).astype(np.float32)
    info = {}
    for name, mode, rate in (("hybrid", ref.MODE_HYBRID, 40000), ("silk_stereo", ref.MODE_SILK_ONLY, 24000)):
        data = ref.encode_forced_mode(speech, mode, rate)
        open(f"{HERE}/{name}.opus", "wb").write(data)
        out, recs = ref.decode_bytes(data, record=True)
        if mode == ref.MODE_HYBRID:   # one 20 ms CELT frame per packet, nothing below band 17 (bin 320)
            assert len(recs) == 150 and all(r["coef"].shape[1] == 960 and not r["coef"][:, :320].any() for r in recs)
        else:
            assert len(recs) == 0
        info[name] = {"samples_per_channel": int(out.shape[0]), "channels": int(out.shape[1]), "celt_frames": len(recs),
                      "transient_frames": int(sum(r["B"] > 1 for r in recs)),
                      "reference_pcm_sha256": hashlib.sha256(out.tobytes()).hexdigest()}
    # ---- a file that walks through the coding modes: CELT -> hybrid -> SILK -> hybrid -> CELT -> SILK ->
    # CELT -> hybrid.  The encoder puts a 5 ms redundancy frame next to every switch from / to CELT-only
    # and the decoder cross-fades it in (opus_decoder_clean.c:478-487, :530-555); hybrid -> SILK makes the
    # decoder add a 2.5 ms CELT fade-out frame (:507-514) ----
    S, H, Cm = ref.MODE_SILK_ONLY, ref.MODE_HYBRID, ref.MODE_CELT_ONLY
    sched = [(20, H), (40, S), (60, H), (80, Cm), (100, S), (120, Cm), (135, H)]
    data = ref.encode_mode_schedule(speech, Cm, sched, 40000)
    open(f"{HERE}/modeswitch.opus", "wb").write(data)
    out, recs = ref.decode_bytes(data, record=True)
    sizes = [int(r["coef"].shape[1]) for r in recs]
    assert sizes.count(240) >= 5 and sizes.count(120) >= 1, sizes
    info["modeswitch"] = {"samples_per_channel": int(out.shape[0]), "channels": int(out.shape[1]), "celt_frames": len(recs),
                          "transient_frames": int(sum(r["B"] > 1 for r in recs)),
                          "redundancy_frames": sizes.count(240), "fade_out_frames": sizes.count(120),
                          "schedule": [[f, m] for f, m in sched],
                          "reference_pcm_sha256": hashlib.sha256(out.tobytes()).hexdigest()}
    json.dump(info, open(f"{HERE}/modes.json", "w"), indent=1)


if __name__ == "__main__":
    if "--modes-only" in sys.argv:
        make_mode_files()
    else:
        main()
