"""Test helper: the records the reference decoder leaves behind for a multistream file
(oracle.ref.decode_file(..., record=True): one per CELT frame per decoder, in decode order, tagged
with the decoder's stream index) -> the batch arrays of include/nq_celt_synth.h."""
import numpy as np

from oracle import port


def multistream_batch(recs, streams, coupled):
    """Returns (coef [nframes][D][960], flags [nframes][streams] u8, frames [nframes][streams])."""
    D = streams + coupled
    per = [[r for r in recs if r["stream"] == s] for s in range(streams)]
    nframes = len(per[0])
    assert all(len(p) == nframes for p in per), [len(p) for p in per]
    coef = np.zeros((nframes, D, 960), np.float32)
    flags = np.zeros((nframes, streams), np.uint8)
    frames = np.zeros((nframes, streams), port.POST_FRAME_DTYPE)
    for s in range(streams):
        rows = [2 * s, 2 * s + 1] if s < coupled else [s + coupled]
        assert all(r["nch"] == len(rows) for r in per[s])
        frames[:, s] = port.post_frames_from_records(per[s])
        for f, r in enumerate(per[s]):
            N = r["coef"].shape[1]
            coef[f, rows, :N] = r["coef"]
            LM = {120: 0, 240: 1, 480: 2, 960: 3}[N]
            flags[f, s] = (1 if r["B"] > 1 else 0) | ((3 - LM) << 1)
    return coef, flags, frames


def oracle_decode_multistream(coef, flags, frames, streams, coupled, mapping):
    """Per-stream oracle synthesis + post stage, then the channel routing of
    opus_multistream_decoder.c:260-299.  20 ms frames only."""
    nframes = coef.shape[0]
    out = np.zeros((nframes * 960, len(mapping)), np.float32)
    for s in range(streams):
        rows = [2 * s, 2 * s + 1] if s < coupled else [s + coupled]
        sig, _, _ = port.synth_batch(np.ascontiguousarray(coef[:, rows]), np.ascontiguousarray(flags[:, s] & 1), None, nthreads=4)
        pcm, _, _ = port.post_batch(sig, np.ascontiguousarray(frames[:, s]))
        for c, d in enumerate(mapping):
            if d in rows:
                out[:, c] = pcm[:, rows.index(d)]
    return out
