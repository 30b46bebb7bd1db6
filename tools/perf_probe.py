#!/usr/bin/env python
"""Device-resident throughput of the batched synthesis for several shapes
(channels x transient fraction).  Diagnostic only; bench.py is the contract."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import libnyquist_b200 as nq

if os.environ.get("NQ_PROBE_LIB"):   # A/B runs: another build of the library
    nq.LIB_PATH = os.environ["NQ_PROBE_LIB"]

def sm_clock():
    """Current SM clock in MHz (the issue-bound variants follow it under the power cap)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        return int(pynvml.nvmlDeviceGetClockInfo(pynvml.nvmlDeviceGetHandleByIndex(0), pynvml.NVML_CLOCK_SM))
    except Exception:
        return None


def run(synth, frames, C, p_tr, steps=5):
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(1)
    coef = torch.empty((frames, C, 960), dtype=torch.float32, device=dev).uniform_(-1, 1, generator=g)
    tr = (torch.rand(frames, generator=g, device=dev) < p_tr).to(torch.uint8)
    pcm = torch.empty((frames * 960, C), dtype=torch.float32, device=dev)
    for _ in range(3):
        synth.synth_batch_torch(coef, tr, out=pcm, want_tail=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        synth.synth_batch_torch(coef, tr, out=pcm, want_tail=False)
    e1.record()
    mhz = sm_clock()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    gbs = frames * C * 960 * 8 / (ms * 1e-3) / 1e9
    return dict(frames=frames, C=C, p_transient=p_tr, ms=round(ms, 3), GBps=round(gbs, 1),
                Mframes_per_s=round(frames / ms / 1e3, 2), frac_of_6527=round(gbs / 6527.5, 3), sm_mhz=mhz)

def run_ms(synth, frames, streams, coupled, p_tr, steps=5):
    """Multistream layout (own transient flag per stream), identity channel mapping."""
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(1)
    D = streams + coupled
    coef = torch.empty((frames, D, 960), dtype=torch.float32, device=dev).uniform_(-1, 1, generator=g)
    tr = (torch.rand((frames, streams), generator=g, device=dev) < p_tr).to(torch.uint8)
    pcm = torch.empty((frames * 960, D), dtype=torch.float32, device=dev)
    for _ in range(3):
        synth.synth_batch_ms_torch(coef, tr, streams, coupled, list(range(D)), out=pcm, want_tail=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        synth.synth_batch_ms_torch(coef, tr, streams, coupled, list(range(D)), out=pcm, want_tail=False)
    e1.record()
    mhz = sm_clock()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    gbs = frames * D * 960 * 8 / (ms * 1e-3) / 1e9
    return dict(frames=frames, streams=streams, coupled=coupled, p_transient=p_tr, ms=round(ms, 3), GBps=round(gbs, 1),
                frac_of_6527=round(gbs / 6527.5, 3), sm_mhz=mhz)


def run_lm(synth, frames, LM, p_tr, steps=5):
    """Stereo frames of 2.5 / 5 / 10 ms (LM = 0 / 1 / 2) inside a batch: rows keep their 960-float stride, the
    flag byte carries 3 - LM in bits 1-2, every frame names its first output sample (frame_offset)."""
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(1)
    N = 120 << LM
    coef = torch.zeros((frames, 2, 960), dtype=torch.float32, device=dev)
    coef[:, :, :N].uniform_(-1, 1, generator=g)
    tr = ((torch.rand((frames, 1), generator=g, device=dev) < p_tr).to(torch.uint8) | ((3 - LM) << 1)).to(torch.uint8)
    offs = (torch.arange(frames + 1, device=dev, dtype=torch.int64) * N)
    pcm = torch.empty((frames * N, 2), dtype=torch.float32, device=dev)
    call = lambda: synth.synth_batch_ms_torch(coef, tr, 1, 1, None, out=pcm, want_tail=False, frame_offset=offs)
    for _ in range(2):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        call()
    e1.record()
    mhz = sm_clock()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    gbs = frames * 2 * N * 8 / (ms * 1e-3) / 1e9
    return dict(frames=frames, LM=LM, frame_ms=2.5 * (1 << LM), p_transient=p_tr, ms=round(ms, 3), GBps_algorithmic=round(gbs, 1),
                Mframes_per_s=round(frames / ms / 1e3, 2), frac_of_6527=round(gbs / 6527.5, 4), sm_mhz=mhz,
                note="LM < 3 frames go through the compact warp routine small_frame_planes (correctness path)")


if __name__ == "__main__":
    # --case frames,C,p_transient[,steps] (repeatable): run only these (used under ncu)
    cases = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--case=")]
    with nq.CeltSynth(0) as s:
        for c in cases:
            fr, C, p, *st = c.split(",")
            print(json.dumps(run(s, int(fr), int(C), float(p), int(st[0]) if st else 5)), flush=True)
        ms_cases = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--ms=")]
        for c in ms_cases:   # --ms frames,streams,coupled,p_transient[,steps]
            fr, st, cp, p, *sx = c.split(",")
            print(json.dumps(run_ms(s, int(fr), int(st), int(cp), float(p), int(sx[0]) if sx else 5)), flush=True)
        lm_cases = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--lm=")]
        for c in lm_cases:   # --lm frames,LM,p_transient[,steps]
            fr, lm, p, *sx = c.split(",")
            print(json.dumps(run_lm(s, int(fr), int(lm), float(p), int(sx[0]) if sx else 5)), flush=True)
        if cases or ms_cases or lm_cases:
            sys.exit(0)
        for fr, st, cp in ((500_000, 5, 3), (600_000, 4, 2)):   # 7.1 / 5.1 multistream, own flags per stream
            print(json.dumps(run_ms(s, fr, st, cp, 0.028)), flush=True)
        for fr, lm in ((200_000, 2), (200_000, 1), (200_000, 0)):   # 10 / 5 / 2.5 ms frames
            print(json.dumps(run_lm(s, fr, lm, 0.028, 2)), flush=True)
        for frames, C, p in [(2_000_000, 2, 0.0), (2_000_000, 2, 0.028), (2_000_000, 2, 0.2), (2_000_000, 2, 1.0),
                             (4_000_000, 1, 0.028), (500_000, 8, 0.028), (500_000, 8, 0.2), (1_300_000, 3, 0.028), (700_000, 6, 0.028), (1_000_000, 4, 0.028),
                             (20_000, 2, 0.028), (2_000, 2, 0.028)]:
            print(json.dumps(run(s, frames, C, p)), flush=True)
