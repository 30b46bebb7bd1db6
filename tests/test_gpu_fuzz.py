"""Randomised end-to-end check (-m gpu) of phase 2 through the C ABI: random multistream layouts
(coupled / mono streams, duplicated and silent output channels), random frame sizes (2.5 - 20 ms),
per-stream block switching, decoder resets at random frames, random flush points of the sink --
against the oracle composed the way the reference composes it (per-stream synthesis + post stage,
then opus_multistream channel routing)."""
import numpy as np
import pytest

import libnyquist_b200 as nq
from oracle import port
from test_gpu_parity import oracle_any_size, rand_batch
from test_gpu_post import rand_frames, assert_pcm

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def synth():
    import torch
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    with nq.CeltSynth(0) as s:
        yield s


def oracle_phase2(coef, flags, frames, resets, streams, coupled, mapping):
    """coef [n][D][960], flags [n][streams] (bit 0 transient, bits 1-2 = 3-LM), frames [n][streams],
    resets: sorted frame indices where every decoder is reset (0 included)."""
    n = coef.shape[0]
    total = int(frames["N"][:, 0].sum())
    out = np.zeros((total, len(mapping)), np.float32)
    bounds = list(resets) + [n]
    first = np.concatenate([[0], np.cumsum(frames["N"][:, 0])]).astype(np.int64)
    for s in range(streams):
        rows = [2 * s, 2 * s + 1] if s < coupled else [s + coupled]
        pcm_s = np.zeros((total, len(rows)), np.float32)
        for a, b in zip(bounds[:-1], bounds[1:]):
            if a == b:
                continue
            sig, _, _ = oracle_any_size(np.ascontiguousarray(coef[a:b][:, rows]), flags[a:b, s] & 7, None)
            pcm_s[first[a]:first[b]] = port.post_batch(sig, np.ascontiguousarray(frames[a:b, s]))[0]
        for c, d in enumerate(mapping):
            if d != 255 and d in rows:
                out[:, c] = pcm_s[:, rows.index(d)]
    return out


@pytest.mark.parametrize("seed", range(12))
def test_random_layouts_sizes_resets_through_host_entry_and_sink(synth, seed):
    rng = np.random.default_rng(1000 + seed)
    coupled = int(rng.integers(0, 4))
    mono = int(rng.integers(0 if coupled else 1, 4))
    streams, D = coupled + mono, 2 * coupled + mono
    # output channels: a random gather of the decoded channels, with a duplicate and a silent one sometimes
    mapping = list(rng.permutation(D))
    if rng.uniform() < 0.5:
        mapping.insert(int(rng.integers(0, len(mapping) + 1)), 255)
    if rng.uniform() < 0.5:
        mapping.append(int(rng.integers(0, D)))
    mapping = [int(m) for m in mapping]
    nframes = int(rng.integers(40, 400))
    any_size = rng.uniform() < 0.6
    lm = rng.choice([3, 3, 3, 2, 1, 0], nframes) if any_size else np.full(nframes, 3)
    coef, _ = rand_batch(rng, nframes, D, 0.0)
    tr = (rng.uniform(size=(nframes, streams)) < rng.choice([0.0, 0.05, 0.3])).astype(np.uint8)
    frames = np.stack([rand_frames(rng, nframes) for _ in range(streams)], axis=1)
    frames["N"] = (120 << lm)[:, None]
    resets = sorted({0, *[int(v) for v in rng.integers(1, nframes, int(rng.integers(0, 4)))]})
    flags = (tr | ((3 - lm)[:, None] << 1)).astype(np.uint8)
    flags_r = flags.copy()
    flags_r[resets[1:]] |= 8            # frame 0 starts from a reset decoder anyway
    want = oracle_phase2(coef, flags, frames, resets, streams, coupled, mapping)

    # 1. the host entry, one call
    got, _ = synth.decode_batch(coef, flags_r, frames, streams=streams, coupled_streams=coupled, mapping=mapping)
    assert_pcm(want, got, f"seed {seed}: decode_batch")
    for c, d in enumerate(mapping):
        if d == 255:
            assert not got[:, c].any()

    # 2. the sink: pushes in packet order, flushed at random points, resets through nq_celt_sink_reset
    sink = nq.FrameSink(len(mapping), streams, coupled, mapping)
    cuts = sorted({int(v) for v in rng.integers(1, nframes, 3)} | {nframes})
    parts, f0 = [], 0
    for cut in cuts:
        for f in range(f0, cut):
            if f in resets[1:]:
                sink.reset()
            N = 120 << lm[f]
            for s in range(streams):
                rows = [2 * s, 2 * s + 1] if s < coupled else [s + coupled]
                sink.push(s, np.ascontiguousarray(coef[f, rows, :N]), (1 << lm[f]) if tr[f, s] else 0, frames[f, s])
        parts.append(sink.flush(synth))
        f0 = cut
    assert np.array_equal(np.concatenate(parts), got), f"seed {seed}: sink != host entry"
