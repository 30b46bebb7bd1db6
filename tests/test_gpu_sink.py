"""GPU tests (-m gpu) of the frame sink (include/nq_celt_synth.h, nq_celt_sink_*): the phase-1 ->
phase-2 hand-over of the restructured decoder, in both its synchronous (flush) and streaming
(attach / finish: phase 2 on a worker thread, block by block) forms."""
import numpy as np
import pytest

import libnyquist_b200 as nq
from test_gpu_post import rand_frames, assert_pcm
from test_gpu_parity import rand_batch, MS_LAYOUTS, ms_oracle
from oracle import port

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def synth():
    import torch
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    with nq.CeltSynth(0) as s:
        yield s


def push_all(sink, coef, tr, fr, streams, coupled, per_packet=3):
    """Push in the order a multistream decoder produces frames: per packet, stream by stream,
    `per_packet` frames each (60 ms packets), like opus_multistream_decode_native."""
    nframes = coef.shape[0]
    for f0 in range(0, nframes, per_packet):
        for s in range(streams):
            rows = [2 * s, 2 * s + 1] if s < coupled else [s + coupled]
            for f in range(f0, min(f0 + per_packet, nframes)):
                sink.push(s, np.ascontiguousarray(coef[f, rows]), 8 if tr[f, s] else 0, fr[f, s])


def test_sink_flush_equals_decode_batch_and_oracle(synth):
    rng = np.random.default_rng(1)
    nframes = 2600          # more than one 2048-frame block
    coef, tr = rand_batch(rng, nframes, 2, 0.1)
    fr = rand_frames(rng, nframes)
    sink = nq.FrameSink(2, 1, 1, [0, 1])
    push_all(sink, coef, tr[:, None], fr[:, None], 1, 1)
    assert sink.pending_frames == nframes
    got = sink.flush(synth)
    want, _ = synth.decode_batch(coef, tr, fr)
    assert np.array_equal(got, want)
    sig, _, _ = port.synth_batch(coef, tr, None, nthreads=8)
    ref_pcm, _, _ = port.post_batch(sig, fr)
    assert_pcm(ref_pcm, got, "sink vs oracle")
    # the sink carries the decoder state across flushes: two halves == one piece, bit for bit
    sink2 = nq.FrameSink(2, 1, 1, [0, 1])
    push_all(sink2, coef[:1000], tr[:1000, None], fr[:1000, None], 1, 1)
    a = sink2.flush(synth)
    push_all(sink2, coef[1000:], tr[1000:, None], fr[1000:, None], 1, 1)
    b = sink2.flush(synth)
    assert np.array_equal(np.concatenate([a, b]), got)
    # ... and reset() forgets it, between flushes or in the middle of one
    sink2.reset()
    push_all(sink2, coef[:50], tr[:50, None], fr[:50, None], 1, 1)
    assert np.array_equal(sink2.flush(synth), got[:50 * 960])
    push_all(sink2, coef[50:90], tr[50:90, None], fr[50:90, None], 1, 1)
    sink2.reset()
    push_all(sink2, coef[:30], tr[:30, None], fr[:30, None], 1, 1)
    mid = sink2.flush(synth)
    assert np.array_equal(mid[:40 * 960], got[50 * 960:90 * 960]) and np.array_equal(mid[40 * 960:], got[:30 * 960])


def test_sink_streaming_equals_flush_with_window(synth):
    rng = np.random.default_rng(2)
    nframes = 4500          # two full blocks on the worker + a partial one at finish
    coef, tr = rand_batch(rng, nframes, 2, 0.05)
    fr = rand_frames(rng, nframes)
    sink = nq.FrameSink(2, 1, 1, [0, 1])
    push_all(sink, coef, tr[:, None], fr[:, None], 1, 1)
    want = sink.flush(synth)
    skip, keep = 312, nframes * 960 - 312 - 777      # pre-skip at the head, trim at the end
    dst = np.full((keep, 2), np.nan, np.float32)
    sink.attach(synth, dst, skip)
    push_all(sink, coef, tr[:, None], fr[:, None], 1, 1)
    assert sink.finish() == nframes * 960
    # same decoder state as after the flush (the sink was NOT reset): continue from there
    sink3 = nq.FrameSink(2, 1, 1, [0, 1])
    push_all(sink3, np.concatenate([coef, coef]), np.concatenate([tr, tr])[:, None], np.concatenate([fr, fr])[:, None], 1, 1)
    both = sink3.flush(synth)
    assert np.array_equal(both[:nframes * 960], want)
    assert np.array_equal(dst, both[nframes * 960 + skip:nframes * 960 + skip + keep])


def test_sink_multistream_7_1(synth):
    streams, coupled, mapping = MS_LAYOUTS["surround_7.1"]
    D = streams + coupled
    rng = np.random.default_rng(3)
    nframes = 2100
    coef, _ = rand_batch(rng, nframes, D, 0.0)
    tr = (rng.uniform(size=(nframes, streams)) < 0.1).astype(np.uint8)
    fr = np.stack([rand_frames(rng, nframes) for _ in range(streams)], axis=1)
    sink = nq.FrameSink(len(mapping), streams, coupled, mapping)
    dst = np.zeros((nframes * 960, len(mapping)), np.float32)
    sink.attach(synth, dst, 0)
    push_all(sink, coef, tr, fr, streams, coupled)
    assert sink.finish() == nframes * 960
    sig, _ = ms_oracle(coef, tr, streams, coupled, mapping)
    want = np.zeros_like(sig)
    for c, d in enumerate(mapping):
        s = d // 2 if d < 2 * coupled else d - coupled
        want[:, c:c + 1] = port.post_batch(np.ascontiguousarray(sig[:, c:c + 1]), np.ascontiguousarray(fr[:, s]))[0]
    assert_pcm(want, dst, "7.1 through the sink")


def test_sink_rejects_inconsistent_pushes(synth):
    sink = nq.FrameSink(3, 2, 1, [0, 1, 2])
    fr = np.zeros(1, nq.POST_FRAME_DTYPE)
    fr["N"] = 960
    with pytest.raises(nq.NqError):
        sink.push(0, np.zeros((1, 960), np.float32), 0, fr)     # coupled stream needs 2 channels
    with pytest.raises(nq.NqError):
        sink.push(2, np.zeros((1, 960), np.float32), 0, fr)     # no such stream
    with pytest.raises(nq.NqError):
        sink.push(1, np.zeros((1, 960), np.float32), 4, fr)     # shortBlocks must be 0 or 8 at LM=3
    sink.push(0, np.zeros((2, 960), np.float32), 0, fr)
    with pytest.raises(nq.NqError):
        sink.flush(synth)                                       # stream 1 has not pushed its frame


def test_sink_5_1_mixed_frame_sizes_streaming(synth):
    """Everything at once: 5.1 multistream (2 coupled + 2 mono streams sharing a warp), frames of
    2.5 / 5 / 10 / 20 ms, per-stream block switching, post-filter, streaming phase 2, pre-skip window."""
    from test_gpu_parity import oracle_any_size
    streams, coupled, mapping = MS_LAYOUTS["surround_5.1"]
    D = streams + coupled
    rng = np.random.default_rng(5)
    nframes = 2300
    lm = rng.choice([3, 3, 3, 2, 1, 0], nframes)
    coef, _ = rand_batch(rng, nframes, D, 0.0)
    tr = (rng.uniform(size=(nframes, streams)) < 0.15).astype(np.uint8)
    fr = np.stack([rand_frames(rng, nframes) for _ in range(streams)], axis=1)
    fr["N"] = (120 << lm)[:, None]
    sink = nq.FrameSink(len(mapping), streams, coupled, mapping)
    total = int((120 << lm).sum())
    skip = 100
    dst = np.zeros((total - skip, len(mapping)), np.float32)
    sink.attach(synth, dst, skip)
    for f in range(nframes):
        N = 120 << lm[f]
        for s in range(streams):
            rows = [2 * s, 2 * s + 1] if s < coupled else [s + coupled]
            sink.push(s, np.ascontiguousarray(coef[f, rows, :N]), (1 << lm[f]) if tr[f, s] else 0, fr[f, s])
    assert sink.finish() == total
    want = np.zeros((total, len(mapping)), np.float32)
    for s in range(streams):
        rows = [2 * s, 2 * s + 1] if s < coupled else [s + coupled]
        flags = (tr[:, s] | ((3 - lm) << 1)).astype(np.uint8)
        sig, _, _ = oracle_any_size(np.ascontiguousarray(coef[:, rows]), flags, None)
        pcm, _, _ = port.post_batch(sig, np.ascontiguousarray(fr[:, s]))
        for c, d in enumerate(mapping):
            if d in rows:
                want[:, c] = pcm[:, rows.index(d)]
    assert_pcm(want[skip:], dst, "5.1 mixed sizes through the sink")
