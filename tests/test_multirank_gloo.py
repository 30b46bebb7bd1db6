"""world_size-2 (and 3) gloo test of the N > 1 host logic on CPU: each rank
takes its contiguous shard + one-frame halo (libnyquist_b200.sharding), the
synthesis itself is stood in for by the oracle (no GPU here), rank 0 gathers by
concatenation and must reproduce the unsharded result bit for bit; the
max-over-ranks timing reduction is exercised too.  No data-path collective is
used by the product -- the gather here is the *host* gather of disjoint outputs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import libnyquist_b200 as nq
from libnyquist_b200.sharding import max_over_ranks, take_shard
from oracle import port


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port_no, nframes, C, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(42)           # same stream on every rank
        coef = (rng.standard_normal((nframes, C, 960)) * 400).astype(np.float32)
        tr = (rng.uniform(size=nframes) < 0.3).astype(np.uint8)
        tail_in = (rng.standard_normal((C, 60)) * 50).astype(np.float32)
        c, t, halo, halo_tr = take_shard(coef, tr, world, rank)
        if halo is None:
            start_tail = tail_in
        else:
            _, start_tail, _ = port.synth_batch(halo[None], np.array([halo_tr], np.uint8), None)
        pcm, tail, _ = port.synth_batch(c, t, start_tail)
        parts = [None] * world
        dist.all_gather_object(parts, (pcm, tail))
        slowest = max_over_ranks(float(rank + 1), dist)
        if rank == 0:
            want, want_tail, _ = port.synth_batch(coef, tr, tail_in)
            got = np.concatenate([p[0] for p in parts])
            ok = np.array_equal(got, want) and np.array_equal(parts[-1][1], want_tail) and slowest == float(world)
            q.put(bool(ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nframes,C", [(2, 17, 2), (3, 10, 1)])
def test_sharded_ranks_reproduce_unsharded_stream(world, nframes, C):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port_no, nframes, C, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
