#!/usr/bin/env python
"""Small run of every kernel variant, meant to be run under compute-sanitizer
(memcheck / racecheck / synccheck / initcheck):
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py
Sizes are tiny; results are checked against the oracle so a sanitizer-clean run is also a correct one."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch

import libnyquist_b200 as nq
from oracle import port


def check(want, got, what, tol):
    err = float(np.abs(want.astype(np.float64) - got).max())
    assert err <= tol, (what, err)
    print(f"ok {what}: max err {err:.2e}", flush=True)


def main():
    rng = np.random.default_rng(0)
    with nq.CeltSynth(0) as s:
        # stereo / mono / 8ch plain, with transients, a tail in and enough frames for several runs + dynamic claims
        for C, nframes in ((2, 300), (1, 300), (8, 80), (3, 60)):
            coef = (rng.standard_normal((nframes, C, 960)) * 500).astype(np.float32)
            tr = (rng.uniform(size=nframes) < 0.2).astype(np.uint8)
            tail_in = (rng.standard_normal((C, 60)) * 100).astype(np.float32)
            want, wt, _ = port.synth_batch(coef, tr, tail_in, nthreads=4)
            pcm, tail = s.synth_batch_torch(torch.from_numpy(coef).cuda(), torch.from_numpy(tr).cuda(),
                                            tail_in=torch.from_numpy(tail_in).cuda())
            torch.cuda.synchronize()
            check(want, pcm.cpu().numpy(), f"synth C={C}", 0.33)
            check(wt, tail.cpu().numpy(), f"tail C={C}", 0.33)
        # 7.1 multistream: paired mono streams with their own flags (split passes), channel mapping
        from test_gpu_parity import ms_oracle, oracle_any_size
        streams, coupled, mapping = 5, 3, [0, 6, 1, 2, 3, 4, 5, 7]
        D = streams + coupled
        nframes = 60
        coef = (rng.standard_normal((nframes, D, 960)) * 500).astype(np.float32)
        tr = (rng.uniform(size=(nframes, streams)) < 0.3).astype(np.uint8)
        want, _ = ms_oracle(coef, tr, streams, coupled, mapping)
        pcm, _ = s.synth_batch_ms_torch(torch.from_numpy(coef).cuda(), torch.from_numpy(tr).cuda(), streams, coupled, mapping)
        torch.cuda.synchronize()
        check(want, pcm.cpu().numpy(), "synth 7.1", 0.33)
        # frames shorter than 20 ms, resets
        nframes = 50
        lm = rng.choice([3, 2, 1, 0], nframes)
        coef = (rng.standard_normal((nframes, 2, 960)) * 500).astype(np.float32)
        flags = ((rng.uniform(size=nframes) < 0.3).astype(np.uint8) | ((3 - lm) << 1)).astype(np.uint8)
        want, _, offs = oracle_any_size(coef, flags, None)
        pcm, _ = s.synth_batch_ms_torch(torch.from_numpy(coef).cuda(), torch.from_numpy(flags).cuda().reshape(-1, 1), 1, 1, None,
                                        frame_offset=torch.from_numpy(offs).cuda())
        torch.cuda.synchronize()
        check(want, pcm.cpu().numpy(), "synth any-size", 0.33)
        # post stage: one stream, segments, mono
        from test_gpu_post import rand_frames
        for C in (2, 1):
            nframes = 12
            fr = rand_frames(rng, nframes)
            sig = (rng.standard_normal((nframes * 960, C)) * 1500).astype(np.float32)
            want = np.concatenate([port.post_batch(sig[:5 * 960], fr[:5])[0], port.post_batch(sig[5 * 960:], fr[5:])[0]])
            d = torch.from_numpy(sig).cuda()
            s.post_segments_torch(d, fr, np.array([0, 5, nframes], np.int64))
            torch.cuda.synchronize()
            check(want, d.cpu().numpy(), f"post C={C}", 1e-5)
        # generic single-call kernel
        out = np.zeros(1020, np.float32)
        nq.clt_mdct_backward(coef[0, 0], out, 0, 1)
        ref1 = np.zeros(1020, np.float32)
        port.clt_mdct_backward(coef[0, 0], ref1, 0, 1)
        check(ref1, out, "clt_mdct_backward", 0.33)
    print("sanitize_smoke: all variants ran")


if __name__ == "__main__":
    main()
