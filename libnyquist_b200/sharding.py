"""Host-side shard planning for the multi-GPU path (one process per GPU).

The synthesis of frame f depends on earlier data only through the 60-sample
raw tail of frame f-1 (mdct.c:361-377; SURVEY.md section 5 "sequence
parallelism"), and that tail is a function of frame f-1's coefficients alone.
So a stream shards into contiguous frame ranges with a ONE-FRAME halo and no
data-path collective: rank r synthesises [f0, f1) and, when f0 > 0, is handed
the coefficients of frame f0-1 (`halo`) which it re-synthesises only for the
tail.  Outputs are disjoint; the host gathers them by concatenation.
"""
from __future__ import annotations

from typing import List, NamedTuple, Optional


class Shard(NamedTuple):
    rank: int
    f0: int            # first frame (inclusive)
    f1: int            # last frame (exclusive)
    halo: Optional[int]  # index of the halo frame (f0 - 1) or None for the shard that opens the stream

    @property
    def nframes(self) -> int:
        return self.f1 - self.f0


def shard_range(nframes: int, world: int, rank: int) -> Shard:
    """Contiguous, balanced split: sizes differ by at most one frame."""
    if world < 1 or not (0 <= rank < world) or nframes < 0:
        raise ValueError("bad shard request")
    f0 = nframes * rank // world
    f1 = nframes * (rank + 1) // world
    return Shard(rank, f0, f1, f0 - 1 if (f0 > 0 and f1 > f0) else None)


def shard_plan(nframes: int, world: int) -> List[Shard]:
    return [shard_range(nframes, world, r) for r in range(world)]


def take_shard(coef, transient, world: int, rank: int):
    """Slices one rank's work out of a whole stream held as arrays/tensors
    indexed [frame, ...]: returns (coef_shard, transient_shard, halo_coef or
    None, halo_transient)."""
    s = shard_range(len(transient), world, rank)
    halo_coef = None if s.halo is None else coef[s.halo]
    halo_tr = 0 if s.halo is None else int(transient[s.halo])
    return coef[s.f0:s.f1], transient[s.f0:s.f1], halo_coef, halo_tr


def max_over_ranks(value: float, dist=None, device=None) -> float:
    """Multi-GPU times are reported as the max over ranks (bench contract)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
