#!/usr/bin/env python
"""Latency of the post stage (comb filter + de-emphasis) per frame for ONE stereo stream -- the
recurrence-bound case (one CTA) -- and for a wide multistream batch.  Diagnostic only."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import libnyquist_b200 as nq


def frames_for(rng, n, pitch_lo=15, pitch_hi=1023, p_off=0.1):
    fr = np.zeros(n, nq.POST_FRAME_DTYPE)
    fr["N"] = 960
    pitch = rng.integers(pitch_lo, pitch_hi, n + 2)
    gain = rng.choice([0.09375 * k for k in range(1, 9)], n + 2).astype(np.float32)
    gain[rng.uniform(size=n + 2) < p_off] = 0
    tap = rng.integers(0, 3, n + 2)
    for i in range(n):
        fr["pitch"][i] = pitch[i + 1], pitch[i + 1], pitch[i + 2]
        fr["gain"][i] = gain[i + 1], gain[i + 1], gain[i + 2]
        fr["tapset"][i] = tap[i + 1], tap[i + 1], tap[i + 2]
    return fr


def run(s, n, label, **kw):
    rng = np.random.default_rng(1)
    fr = frames_for(rng, n, **kw)
    pcm = torch.randn((n * 960, 2), device="cuda") * 1000
    for _ in range(2):
        s.post_batch_torch(pcm, fr, want_state=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s.post_batch_torch(pcm, fr, want_state=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(json.dumps(dict(case=label, frames=n, ms=round(ms, 3), us_per_frame=round(ms * 1e3 / n, 3))), flush=True)


if __name__ == "__main__":
    with nq.CeltSynth(0) as s:
        run(s, 4000, "one stereo stream, random pitch 15..1022")
        run(s, 4000, "one stereo stream, pitch 15 (13-sample steps)", pitch_lo=15, pitch_hi=16)
        run(s, 4000, "one stereo stream, pitch 500..1022", pitch_lo=500)
        run(s, 4000, "one stereo stream, post-filter off", p_off=1.0)
