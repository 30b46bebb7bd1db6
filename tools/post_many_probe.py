#!/usr/bin/env python
"""Post stage over a batch of many independent stereo streams (one CTA each). Diagnostic."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import libnyquist_b200 as nq

if os.environ.get("NQ_PROBE_LIB"):   # A/B runs: another build of the library
    nq.LIB_PATH = os.environ["NQ_PROBE_LIB"]

nseg, per = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, int(sys.argv[2]) if len(sys.argv) > 2 else 256
frames = nseg * per
rng = np.random.default_rng(3)
fr = np.zeros(frames, nq.POST_FRAME_DTYPE)
fr["N"] = 960
pitch = rng.integers(15, 1023, frames + 1)
gain = (rng.integers(0, 9, frames + 1) * 0.09375).astype(np.float32)
tap = rng.integers(0, 3, frames + 1)
fr["pitch"] = np.stack([pitch[:-1], pitch[:-1], pitch[1:]], 1)
fr["gain"] = np.stack([gain[:-1], gain[:-1], gain[1:]], 1)
fr["tapset"] = np.stack([tap[:-1], tap[:-1], tap[1:]], 1)
seg = np.arange(nseg + 1, dtype=np.int64) * per
with nq.CeltSynth(0) as s:
    o = torch.randn((frames * 960, 2), device="cuda") * 1000
    for _ in range(2):
        s.post_segments_torch(o, fr, seg)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s.post_segments_torch(o, fr, seg)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps(dict(nseg=nseg, per=per, frames=frames, wall_ms=round(dt * 1e3, 2), Mframes_per_s=round(frames / dt / 1e6, 1))))
