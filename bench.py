#!/usr/bin/env python
"""Benchmark of the CELT synthesis hot path (inverse MDCT + TDAC overlap-add +
stereo interleave) on B200, metric and workload as BASELINE.json names them.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--frames F] [--impl reference]

One step = one pass of the hot path over one batch of synthetic stereo 20 ms
CELT frames (BASELINE.json configs[4]: 10 M frames; coefficients uniform with
a band-decaying envelope, zero above bin 800, 2.8 % transient frames).  At
N > 1 (launched under torchrun, one rank per GPU) every rank owns its own
contiguous frame range of the same size (weak scaling, no data-path
collective: each shard only needs the previous frame as a halo).  The same
run also times the STRONG-scaling split BASELINE.md section 4.4 describes
(the same 10 M frames in all, 10 M / N per GPU, wall = max over ranks) and
reports it under extras.strong_scaling; `--scaling strong` makes that split
the headline line instead.  Every rank checks three frames of its timed
output against the oracle (incl. the first frame after its halo); the line
carries the max over ranks as parity_max_err.

Prints ONE JSON line:
  value      whole-job frames/s with the inputs resident in HBM (CUDA events
             on the launching stream, max over ranks)
  e2e        same metric through the C ABI with HOST (pinned) buffers: H2D of
             the coefficients, kernel, D2H of the PCM inside the timed region
  roofline   achieved algorithmic GB/s (15 360 B per stereo frame) of the
             dominant kernel vs the measured HBM peak (MEASURED_PEAKS.json)
  cpu_baseline  the reference's own compute_inv_mdcts (oracle/_ref, else the
             C oracle port) on the box's host cores, bounded sample
--impl reference times that CPU path alone (rank 0 only).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_FRAME = 2 * 960 * 4 * 2          # stereo: 7680 B read + 7680 B written
METRIC = "celt_imdct_ola_frames_per_s"
UNIT = "frames/s"
DEFAULT_FRAMES = 10_000_000               # BASELINE.json configs[4]
P_TRANSIENT = 0.028                       # measured frame mix of sb-reverie.opus
E2E_FRAMES = 262_144                      # host-buffer leg: 2 GB in + 2 GB out of pinned memory per step
FALLBACK_HBM_GBS = 6650.0                 # /opt/skills/guides/B200_PROFILING.md, used only without MEASURED_PEAKS.json
# roofline.traffic: dram__bytes_read.sum + dram__bytes_write.sum of the bench's own 10 M-frame launch of
# celt_synth_kernel<kModeStereo> from the committed `ncu --set full` capture, newest first (a capture
# cannot be taken inside a timed run; the files are written by tools/profile_round.sh + summarize_profiles.py)
NCU_CAPTURES = ("profiles/r2_full_stereo10M.csv", "profiles/r1e_full_stereo10M.csv")
NCU_CAPTURE_FRAMES = 10_000_000


def ncu_traffic_bytes_per_frame():
    """(bytes per frame, source file) from the newest committed capture, or (None, None)."""
    import csv
    for rel in NCU_CAPTURES:
        path = os.path.join(ROOT, rel)
        if not os.path.exists(path):
            continue
        try:
            total = 0.0
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
            for row in csv.reader(open(path)):
                if len(row) == 3 and row[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    total += float(row[2].replace(",", "")) * scale[row[1]]
            if total > 0:
                return total / NCU_CAPTURE_FRAMES, rel
        except Exception:
            continue
    return None, None


# Only the JSON line may reach stdout: NCCL (and anything else that writes to the C-level stdout,
# e.g. "NCCL version ..." at communicator creation) is sent to stderr for the whole run.
# (Done by main() / the probes that import this module, not at import time: tests import it too.)
_REAL_STDOUT = None


def protect_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def workload_name(frames):
    return (f"synthetic batch of {frames} stereo 20 ms CELT frames per GPU (2x960 f32 coefficients -> 2x960 f32 "
            f"samples, {P_TRANSIENT * 100:.1f}% transient), BASELINE.json configs[4]")


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU during the timed region (NVML, 20 ms period)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ------------------------------------------------------------------ CPU arm ---
def cpu_reference_run(frames_per_pass, passes, steps, warmup):
    """The reference's compute_inv_mdcts over stereo frames on all host cores.
    Returns (frames/s, cores, kind, sample description, ms per step)."""
    import numpy as np
    from oracle import port, ref
    use_ref = ref.available()
    mod = ref if use_ref else port
    cores = os.cpu_count() or 1
    rng = np.random.default_rng(0x0B200)
    env = (1000.0 / (1.0 + np.arange(960) / 60.0)).astype(np.float32)
    env[800:] = 0
    coef = rng.uniform(-1, 1, (frames_per_pass, 2, 960)).astype(np.float32) * env
    tr = (rng.uniform(size=frames_per_pass) < P_TRANSIENT).astype(np.uint8)
    best = None
    times = []
    for it in range(warmup + steps):
        t = 0.0
        for _ in range(passes):
            _, _, sec = mod.synth_batch(coef, tr, None, nthreads=cores)
            t += sec
        if it >= warmup:
            times.append(t)
            best = t if best is None else min(best, t)
    mean = sum(times) / len(times)
    kind = "reference" if use_ref else "port"
    sample = (f"{passes} x {frames_per_pass} stereo frames per step on {cores} threads "
              f"({'oracle/_ref: reference sources compiled in place' if use_ref else 'oracle/ C restatement'}; "
              f"gcc -O3, contiguous frame ranges per thread)")
    return frames_per_pass * passes / mean, cores, kind, sample, mean * 1e3


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    value, cores, kind, sample, ms = cpu_reference_run(65536, 2, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.frames), "note": "CPU arm: bounded sample of the same workload"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ------------------------------------------------------------------ GPU arm ---
def make_device_batch(torch, frames, device, seed):
    """Coefficients generated on the device chunk by chunk (no host copy of 77 GB)."""
    g = torch.Generator(device=device).manual_seed(seed)
    env = (1000.0 / (1.0 + torch.arange(960, device=device) / 60.0)).float()
    env[800:] = 0
    coef = torch.empty((frames, 2, 960), dtype=torch.float32, device=device)
    step = 262_144
    for f0 in range(0, frames, step):
        v = coef[f0:f0 + step]
        v.uniform_(-1.0, 1.0, generator=g)
        v.mul_(env)
    tr = (torch.rand(frames, generator=g, device=device) < P_TRANSIENT).to(torch.uint8)
    return coef, tr


def extras(torch, np, nq, synth, dev, peak, with_cpu_baseline):
    """Informational legs outside the contract (N = 1 only, a few seconds): the other channel
    layouts of BASELINE.json's configs on the same kernel family, the PCIe ceiling the e2e leg runs
    against, and the end-to-end file decode (nqr::NyquistIO::Load) of the two-phase build next to
    -- a second cpu_baseline -- the unmodified reference's own Load (oracle/_ref), when both
    libraries travelled with the repo."""
    out = {}

    def timed(fn, steps=5):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    g = torch.Generator(device=dev).manual_seed(7)
    shapes = {}
    for name, frames, streams, coupled, mapping, p_tr in (
            ("mono", 2_000_000, 1, 0, [0], P_TRANSIENT),
            ("stereo_all_transient", 1_000_000, 1, 1, [0, 1], 1.0),
            ("surround_5.1_multistream", 1_200_000, 4, 2, [0, 4, 1, 2, 3, 5], P_TRANSIENT),
            ("surround_7.1_multistream", 1_000_000, 5, 3, [0, 6, 1, 2, 3, 4, 5, 7], P_TRANSIENT)):
        D = streams + coupled
        c = torch.empty((frames, D, 960), dtype=torch.float32, device=dev).uniform_(-1, 1, generator=g)
        t = (torch.rand((frames, streams), generator=g, device=dev) < p_tr).to(torch.uint8)
        o = torch.empty((frames * 960, len(mapping)), dtype=torch.float32, device=dev)
        ms = timed(lambda: synth.synth_batch_ms_torch(c, t, streams, coupled, mapping, out=o, want_tail=False))
        gbs = frames * 960 * 4 * (D + len(mapping)) / (ms * 1e-3) / 1e9
        shapes[name] = {"frames": frames, "ms": ms, "frames_per_s": frames / (ms * 1e-3), "GBps": gbs, "frac_of_hbm_peak": gbs / peak}
        del c, t, o
    out["other_layouts_device_resident"] = shapes

    # whole phase 2 (synthesis + post-filter + de-emphasis) over a batch of many independent stereo
    # streams: the post stage's filters are recurrences along time, so it fills the GPU with streams
    try:
        nseg, per = 4096, 256
        frames = nseg * per
        c = torch.empty((frames, 2, 960), dtype=torch.float32, device=dev).uniform_(-1, 1, generator=g).mul_(1000.0)
        fl = (torch.rand(frames, generator=g, device=dev) < P_TRANSIENT).to(torch.uint8)
        fl[::per] |= 8                                   # decoder reset at the head of every stream
        o = torch.empty((frames * 960, 2), dtype=torch.float32, device=dev)
        rng = np.random.default_rng(3)
        # side info in pinned host memory, like the sink keeps it (40 bytes per stream-frame)
        fr = torch.zeros(frames * nq.POST_FRAME_DTYPE.itemsize, dtype=torch.uint8, pin_memory=True).numpy().view(nq.POST_FRAME_DTYPE)
        fr["N"] = 960
        pitch = rng.integers(15, 1023, frames + 1)
        gain = (rng.integers(0, 9, frames + 1) * 0.09375).astype(np.float32)
        tap = rng.integers(0, 3, frames + 1)
        fr["pitch"] = np.stack([pitch[:-1], pitch[:-1], pitch[1:]], 1)
        fr["gain"] = np.stack([gain[:-1], gain[:-1], gain[1:]], 1)
        fr["tapset"] = np.stack([tap[:-1], tap[:-1], tap[1:]], 1)
        seg = np.arange(nseg + 1, dtype=np.int64) * per

        def phase2():
            synth.synth_batch_torch(c, fl, out=o, want_tail=False)
            synth.post_segments_torch(o, fr, seg)

        ms = timed(phase2, steps=3)
        ms_synth = timed(lambda: synth.synth_batch_torch(c, fl, out=o, want_tail=False), steps=3)
        gbs = frames * 2 * BYTES_PER_FRAME / (ms * 1e-3) / 1e9
        out["phase2_many_streams"] = {
            "what": f"{nseg} independent stereo streams x {per} frames: synthesis kernel + post kernel (one CTA per stream), "
                    "side info (40 B per frame) validated and uploaded from pinned host memory inside the call",
            "frames": frames, "ms": ms, "ms_synthesis_only": ms_synth, "frames_per_s": frames / (ms * 1e-3),
            "GBps_algorithmic_4_passes": gbs, "frac_of_hbm_peak": gbs / peak}
        del c, fl, o
    except Exception as e:   # informational leg: never fail the bench line
        out["phase2_many_streams_error"] = repr(e)

    # PCIe ceiling of the e2e leg: concurrent pinned H2D + D2H copies of 256 MB each
    n = 64 << 20
    h_a = torch.empty(n, dtype=torch.float32, pin_memory=True)
    h_b = torch.empty(n, dtype=torch.float32, pin_memory=True)
    d_a = torch.empty(n, dtype=torch.float32, device=dev)
    d_b = torch.empty(n, dtype=torch.float32, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def both():
        with torch.cuda.stream(s1):
            d_a.copy_(h_a, non_blocking=True)
        with torch.cuda.stream(s2):
            h_b.copy_(d_b, non_blocking=True)

    for _ in range(2):
        both()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        both()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    out["pcie_concurrent_copy_GBps_each_way"] = n * 4 / dt / 1e9
    del h_a, h_b, d_a, d_b

    # end-to-end file decode, BASELINE.json configs[0]
    if not with_cpu_baseline:
        return out
    try:
        import ctypes as C
        from oracle import ref
        lib = os.path.join(ROOT, "integration", "_build", "libnyquist_twophase.so")
        path = os.path.join(ROOT, "oracle", "_ref", "test_data", "sb-reverie.opus")
        if os.path.exists(lib) and os.path.exists(path) and ref.load_available():
            L = C.CDLL(lib)
            L.nq_twophase_load.argtypes = [C.c_char_p, C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.c_size_t),
                                           C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double)]
            L.nq_twophase_free.argtypes = [C.POINTER(C.c_float)]
            def timed_load(path, reps):
                best, got = None, None
                for _ in range(reps):
                    p, cnt, ch, sr = C.POINTER(C.c_float)(), C.c_size_t(0), C.c_int(0), C.c_int(0)
                    tm = (C.c_double * 8)()
                    t0 = time.perf_counter()
                    rc = L.nq_twophase_load(path.encode(), C.byref(p), C.byref(cnt), C.byref(ch), C.byref(sr), tm)
                    dt = time.perf_counter() - t0
                    assert rc == 0
                    got = np.ctypeslib.as_array(p, shape=(cnt.value,)).copy().reshape(-1, ch.value)
                    L.nq_twophase_free(p)
                    if best is None or dt < best[0]:
                        best = (dt, list(tm))
                ref_best = None
                for _ in range(2):
                    want, _, dt = ref.nyquist_load(path)
                    ref_best = dt if ref_best is None else min(ref_best, dt)
                tm = best[1]
                return {"two_phase_ms": best[0] * 1e3, "phase1_cpu_entropy_decode_ms": tm[0] * 1e3,
                        "phase2_gpu_tail_after_phase1_ms": tm[1] * 1e3, "reference_cpu_ms": ref_best * 1e3,
                        "speedup": ref_best / best[0], "max_abs_pcm_err": float(np.abs(got - want).max()),
                        # where the rest of the Load goes (OpusDecoderTwoPhase.cpp): open + header, context lease +
                        # sink, waiting for the zero-filled samples vector (its resize overlaps phase 1), gain / SILK
                        # sum; "outside" = ReadFile + the AudioData bookkeeping around the decoder
                        "budget_ms": {"open_header": tm[3] * 1e3, "lease_sink_attach": tm[5] * 1e3, "phase1": tm[0] * 1e3,
                                      "wait_for_samples_resize": tm[4] * 1e3, "phase2_tail": tm[1] * 1e3,
                                      "gain_silk_sum": tm[2] * 1e3, "samples_resize_overlapped": tm[7] * 1e3,
                                      "outside_decoder": (best[0] - tm[6]) * 1e3}}

            out["file_decode_sb_reverie_opus"] = dict(
                what="nqr::NyquistIO::Load, 223.7 s of stereo audio, 11184 CELT frames (BASELINE.json configs[0])",
                **timed_load(path, 3))
            path8 = os.path.join(ROOT, "tests", "golden", "surround8.opus")
            if os.path.exists(path8):
                out["file_decode_surround8_opus"] = dict(
                    what="nqr::NyquistIO::Load, 7.1 multistream file (3 coupled + 2 mono streams, BASELINE.json configs[3]); "
                         "phase 1 decodes the five streams of a packet in parallel",
                    **timed_load(path8, 5))
            # many files at once: nqr::LoadOpusBatch (phase 1 on one host thread per file, ONE synthesis +
            # ONE post launch for all of them) next to K loads of the unmodified reference on K threads
            try:
                K = min(os.cpu_count() or 1, 16)
                L.nq_twophase_load_batch.argtypes = [C.POINTER(C.c_char_p), C.c_int, C.c_int, C.POINTER(C.POINTER(C.c_float)),
                                                     C.POINTER(C.c_size_t), C.POINTER(C.c_int), C.POINTER(C.c_double)]
                arr = (C.c_char_p * K)(*[path.encode()] * K)
                best = None
                for _ in range(3):
                    ptrs, counts, chans, st = (C.POINTER(C.c_float) * K)(), (C.c_size_t * K)(), (C.c_int * K)(), (C.c_double * 7)()
                    t0 = time.perf_counter()
                    rc = L.nq_twophase_load_batch(arr, K, K, ptrs, counts, chans, st)
                    dt = time.perf_counter() - t0
                    assert rc == 0
                    first = np.ctypeslib.as_array(ptrs[0], shape=(counts[0],)).copy()
                    for i in range(K):
                        L.nq_twophase_free(ptrs[i])
                    if best is None or dt < best[0]:
                        best = (dt, list(st))
                R = C.CDLL(ref.LOAD_LIB_PATH)
                R.nqref_load_many.argtypes = [C.POINTER(C.c_char_p), C.c_int, C.c_int]
                R.nqref_load_many.restype = C.c_double
                ref_dt = min(R.nqref_load_many(arr, K, K) for _ in range(2))
                want1, _, _ = ref.nyquist_load(path)
                out["file_decode_batch"] = {
                    "what": f"nqr::LoadOpusBatch of {K} x sb-reverie.opus: phase 1 on {K} host threads, then ONE synthesis launch + ONE post "
                            f"launch over all files; next to {K} loads of the unmodified reference on {K} threads of the same host",
                    "files": K, "two_phase_ms": best[0] * 1e3, "phase1_ms": best[1][0] * 1e3, "phase2_ms": best[1][1] * 1e3,
                    "kernel_launches": int(best[1][6]), "frames": int(best[1][5]), "files_per_s": K / best[0],
                    "reference_ms": ref_dt * 1e3, "reference_files_per_s": K / ref_dt, "speedup": ref_dt / best[0],
                    "max_abs_pcm_err": float(np.abs(first - want1.ravel()).max())}
            except Exception as e:
                out["file_decode_batch_error"] = repr(e)
    except Exception as e:   # informational leg: never fail the bench line
        out["file_decode_error"] = repr(e)
    return out


def near_gpu_cpus(local):
    """CPUs NVML reports as local to the GPU (its NUMA node), or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        return cpus or None
    except Exception:
        return None


class near_gpu:
    """Pinned host buffers are page-locked where the allocating thread runs: allocate them with the
    thread bound to the GPU's own NUMA node, then give the thread its CPUs back."""

    def __init__(self, local):
        self.cpus = near_gpu_cpus(local)
        self.before = None

    def __enter__(self):
        if self.cpus:
            try:
                self.before = os.sched_getaffinity(0)
                os.sched_setaffinity(0, self.cpus)
            except Exception:
                self.before = None
        return self

    def __exit__(self, *a):
        if self.before:
            os.sched_setaffinity(0, self.before)


def pcie_ceiling(torch, dev, h_in, h_out, nbytes, reps, sync):
    """What the platform gives raw copies of the e2e leg's own traffic: `nbytes` host->device and,
    concurrently on a second stream, `nbytes` device->host, pinned buffers, 64 MB per cudaMemcpyAsync.
    `sync` = barrier + device synchronize (every rank copies at the same time).  Seconds per rep."""
    n = nbytes // 4
    d_a = torch.empty(n, dtype=torch.float32, device=dev)
    d_b = torch.empty(n, dtype=torch.float32, device=dev)
    ha, hb = h_in.view(-1)[:n], h_out.view(-1)[:n]
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    step = (64 << 20) // 4

    def both():
        for o in range(0, n, step):
            with torch.cuda.stream(s1):
                d_a[o:o + step].copy_(ha[o:o + step], non_blocking=True)
            with torch.cuda.stream(s2):
                hb[o:o + step].copy_(d_b[o:o + step], non_blocking=True)

    both()
    sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        both()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    del d_a, d_b
    return dt


def oracle_spot_check(np, coef, tr, pcm, halo, frames):
    """Three frames of the timed step's output against the oracle (the checker, not the product):
    the first frame (which follows the rank's halo frame, or a reset decoder on rank 0), one in the
    middle, the last but one.  Max abs error as a fraction of full scale (32768)."""
    from oracle import port
    worst = 0.0
    for f in (0, frames // 3, frames - 2):
        if f < 0 or f >= frames:
            continue
        if f == 0 and halo is not None:
            c2 = np.stack([halo.cpu().numpy(), coef[0].cpu().numpy()])
            t2 = np.array([0, int(tr[0])], np.uint8)
        elif f == 0:
            c2, t2 = coef[0:1].cpu().numpy(), tr[0:1].cpu().numpy()
        else:
            c2, t2 = coef[f - 1:f + 1].cpu().numpy(), tr[f - 1:f + 1].cpu().numpy()
        want, _, _ = port.synth_batch(np.ascontiguousarray(c2), np.ascontiguousarray(t2), None)
        got = pcm[f * 960:(f + 1) * 960].cpu().numpy()
        worst = max(worst, float(np.abs(got - want[-960:]).max()) / 32768.0)
    return worst


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--frames", type=int, default=DEFAULT_FRAMES, help="stereo frames per GPU per step (weak) / in all (strong)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --frames per GPU (the driver's line); strong: --frames in all, split evenly over the GPUs")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the informational other-shapes / file-decode legs")
    args = ap.parse_args()
    protect_stdout()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference_arm(args)

    import numpy as np
    import torch
    import libnyquist_b200 as nq

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libnyquist_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    cpu_group = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)   # plumbing only: barriers + max-over-ranks of the times
        # a second, CPU-side group: ranks that wait while rank 0 drives every GPU from one process
        # (the host_multi leg) must not spin in an NCCL kernel on the GPU they wait for
        cpu_group = dist.new_group(backend="gloo")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def min_over_ranks(x):
        return -max_over_ranks(-x)

    synth = nq.CeltSynth(local)
    strong_main = args.scaling == "strong"
    frames = args.frames // world if strong_main else args.frames
    # Resident workload: the whole batch in HBM (10 M frames = 76.8 GB in + 76.8 GB out).  If it does
    # not fit, halve until it does and say so -- never silently.
    free, _total = torch.cuda.mem_get_info()
    resident_note = "whole batch resident in HBM"
    while frames * BYTES_PER_FRAME + (6 << 30) > free and frames > 1024:
        frames //= 2
        resident_note = f"batch reduced to {frames} frames to fit {free >> 30} GiB free HBM"
    coef, tr = make_device_batch(torch, frames, dev, 0x0B200 + rank)
    pcm = torch.empty((frames * 960, 2), dtype=torch.float32, device=dev)
    # every rank but the first starts mid-stream: hand it a halo frame like a real shard gets
    halo = coef[frames // 2].clone() if rank > 0 else None
    stream = torch.cuda.current_stream(dev)

    def timed_device_leg(nfr):
        """args.steps launches over the first nfr frames of the resident batch: (per-step ms of this
        rank, total ms max over ranks, clocks, launches)."""
        c, t, o = coef[:nfr], tr[:nfr], pcm[:nfr * 960]

        def step():
            synth.synth_batch_torch(c, t, halo_coef=halo, halo_transient=0, out=o, want_tail=False, stream=stream)

        for _ in range(args.warmup):
            step()
        barrier()
        l0 = synth.launch_count
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
        with ClockSampler(local) as clocks:
            ev[0].record(stream)
            for i in range(args.steps):
                step()
                ev[i + 1].record(stream)
            barrier()
        per = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
        return per, max_over_ranks(ev[0].elapsed_time(ev[args.steps])), clocks, synth.launch_count - l0

    per_step_ms, total_ms_max, clocks, launches = timed_device_leg(frames)
    value = world * frames * args.steps / (total_ms_max * 1e-3)

    # ---- parity of the timed output, every rank (the oracle as the checker) ----
    parity = None
    if not args.no_cpu_baseline:
        parity = max_over_ranks(oracle_spot_check(np, coef, tr, pcm, halo, frames))
        if not parity <= 1e-5:
            raise SystemExit(f"bench.py: output of the timed step fails parity ({parity:.3e} of full scale, max over {world} ranks)")

    # ---- the other scaling mode (the same total batch split evenly; wall = max over ranks) ----
    other = None
    if world > 1 and not strong_main:
        nfr = args.frames // world
        if nfr <= frames:
            _per, tot, _clk, _l = timed_device_leg(nfr)
            ms = tot / args.steps
            other = {"what": f"strong scaling: {nfr * world} stereo frames in all, {nfr} per GPU, device-resident, "
                             "CUDA events, max over ranks",
                     "total_frames": nfr * world, "frames_per_gpu": nfr, "ms_per_step": ms,
                     "frames_per_s": nfr * world / (ms * 1e-3),
                     "frac_of_hbm_peak_per_gpu": nfr * BYTES_PER_FRAME / (ms * 1e-3) / 1e9 / hbm_peak()[0],
                     # against this run's own one-GPU time for the whole batch (= the weak leg's per-rank time)
                     "efficiency_vs_one_gpu_same_run": (total_ms_max / args.steps) * (nfr * world / frames) / (world * ms)}

    # ---- e2e: host (pinned) buffers through the C ABI, copies inside the timed region ----
    e2e = None
    multi = None
    if not args.no_e2e:
        ef = min(E2E_FRAMES, frames)
        with near_gpu(local) as ng:   # page-lock the host buffers on the GPU's own NUMA node
            h_coef = torch.empty((ef, 2, 960), dtype=torch.float32, pin_memory=True)
            h_tr = torch.empty(ef, dtype=torch.uint8, pin_memory=True)
            h_pcm = torch.empty((ef * 960, 2), dtype=torch.float32, pin_memory=True)
            h_tail = torch.empty((2, 60), dtype=torch.float32, pin_memory=True)
            h_pcm.zero_()
        h_coef.copy_(coef[:ef])
        h_tr.copy_(tr[:ef])

        def e2e_step():
            synth.synth_batch_host_ptr(h_coef.data_ptr(), h_tr.data_ptr(), 0, h_pcm.data_ptr(), h_tail.data_ptr(), ef, 2)

        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()          # synchronous: returns when the PCM is back in host memory
        torch.cuda.synchronize()
        dt_own = time.perf_counter() - t0
        dt = max_over_ranks(dt_own)
        gb = ef * 7680 / 1e9
        got = h_pcm[:960 * 4].numpy().copy()
        assert np.isfinite(got).all()
        if parity is not None:   # the host-buffer path against the device-resident one: same kernel, same bits
            same = bool(torch.equal(h_pcm[:960 * 64], pcm[:960 * 64].cpu())) if rank == 0 else True
            if not same:
                raise SystemExit("bench.py: the host-buffer leg and the device-resident leg disagree")
        # the platform's ceiling for exactly this traffic, measured the same way on every rank at the
        # same time (same bytes each way, same pinned buffers, raw cudaMemcpyAsync, nothing else)
        ceil_dt_own = pcie_ceiling(torch, dev, h_coef, h_pcm, ef * 7680, args.steps, barrier)
        ceil_dt = max_over_ranks(ceil_dt_own)
        e2e = {"value": world * ef * args.steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": ef * (7680 + 1), "d2h_bytes_per_step": ef * 7680 + 480,
               "frames_per_step": ef, "ms_per_step": dt * 1e3 / args.steps,
               "GBps_each_way": world * gb * args.steps / dt,
               "pcie_ceiling_GBps_each_way": world * gb / ceil_dt,
               "e2e_frac_of_ceiling": (args.steps / dt) * ceil_dt,
               "per_rank_GBps_each_way_min": min_over_ranks(gb * args.steps / dt_own),
               "per_rank_ceiling_GBps_each_way_min": min_over_ranks(gb / ceil_dt_own),
               "pinned_near_gpu_cpus": len(ng.cpus) if ng.cpus else None,
               "note": "nq_celt_synth_batch_host: pinned host buffers, chunked H2D/kernel/D2H pipeline; bounded batch; "
                       "pcie_ceiling = raw concurrent pinned H2D + D2H copies of the same bytes on every rank at the same time"}
        del h_coef, h_pcm, h_tr

        # second e2e arm: the repo's own multi-GPU entry, ONE process feeding all N GPUs from one host
        # buffer (nq_celt_synth_batch_host_multi: a host thread and a cached context per device); the
        # other ranks wait on the CPU meanwhile
        if dist is not None:
            dist.barrier(group=cpu_group)
        if rank == 0:
            try:
                mf = 65536 * world
                m_coef = torch.empty((mf, 2, 960), dtype=torch.float32, pin_memory=True)
                m_pcm = torch.empty((mf * 960, 2), dtype=torch.float32, pin_memory=True)
                m_tr = torch.zeros(mf, dtype=torch.uint8, pin_memory=True)
                for o in range(0, mf, 65536):
                    m_coef[o:o + 65536].copy_(coef[:65536])
                    m_tr[o:o + 65536].copy_(tr[:65536])
                m_tr &= 1
                L = nq.load_library()
                import ctypes as C
                devs = (C.c_int * world)(*range(world))

                def multi_step():
                    rc = L.nq_celt_synth_batch_host_multi(devs, world, C.c_void_p(m_coef.data_ptr()), C.c_void_p(m_tr.data_ptr()),
                                                          None, C.c_void_p(m_pcm.data_ptr()), None, mf, 2)
                    assert rc == 0, rc

                multi_step()
                multi_step()
                t0 = time.perf_counter()
                reps = max(2, args.steps // 2)
                for _ in range(reps):
                    multi_step()
                dtm = (time.perf_counter() - t0) / reps
                # shard invariance: the single-process multi-GPU result equals the one-GPU result of the same frames
                ok = bool(torch.equal(m_pcm[:960 * 64], pcm[:960 * 64].cpu()))
                multi = {"value": mf / dtm, "unit": UNIT, "frames_per_step": mf, "ms_per_step": dtm * 1e3,
                         "GBps_each_way": mf * 7680 / 1e9 / dtm, "devices": world, "matches_one_gpu_output": ok,
                         "note": "nq_celt_synth_batch_host_multi from rank 0: one process, one host thread + cached context per GPU, "
                                 "one pinned host buffer; the other ranks idle"}
                L.nq_celt_multi_release()
                del m_coef, m_pcm
            except Exception as e:   # informational arm
                multi = {"error": repr(e)}
        if dist is not None:
            dist.barrier(group=cpu_group)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    peak, peak_src = hbm_peak()
    kernel_ms = sum(per_step_ms) / len(per_step_ms)          # one launch per step: this IS the kernel's launch duration
    achieved = frames * BYTES_PER_FRAME / (kernel_ms * 1e-3) / 1e9
    if achieved > 1.3 * peak:
        raise SystemExit(f"bench.py: {achieved:.0f} GB/s is far above the HBM peak: the timed region missed the kernel")
    traffic_bpf, traffic_src = ncu_traffic_bytes_per_frame()
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": None if traffic_bpf is None else traffic_bpf * frames, "traffic_source": traffic_src,
        "kernel": "nq::celt_synth_kernel<kModeStereo, 14 warps, 20 ms frames>", "algorithmic_bytes_per_launch": frames * BYTES_PER_FRAME,
        "launch_ms": kernel_ms, "peak_source": peak_src,
    }
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms_max / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(frames), "frames_per_gpu": frames, "channels": 2,
                   "residency": resident_note,
                   "l2": "inputs (7680 B/frame x frames) far larger than the 126 MB L2; no flush needed",
                   "hbm_gbs_aggregate": value * BYTES_PER_FRAME / 1e9},
        "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "clocks": clocks.summary(),
    }
    if parity is not None:
        line["parity_max_err"] = parity     # of full scale, max over all ranks (bar: 1e-5)
    if not args.no_cpu_baseline and world == 1:
        # cpu_baseline leg: the reference's CPU path timed on a bounded sample of the same workload
        v, cores, kind, sample, _ = cpu_reference_run(65536, 2, 3, 1)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                                "timed_output_max_err_of_full_scale": parity}
    ex = {}
    if other is not None:
        ex["strong_scaling"] = other
    if multi is not None:
        ex["e2e_host_multi"] = multi
    if not args.no_extras and world == 1:
        del coef, pcm
        torch.cuda.empty_cache()
        try:   # informational legs: they must never cost the bench line
            ex.update(extras(torch, np, nq, synth, dev, peak, not args.no_cpu_baseline))
        except Exception as e:
            line["extras_error"] = repr(e)
    if ex:
        line["extras"] = ex
    emit(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
