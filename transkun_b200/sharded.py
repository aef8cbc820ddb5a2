"""Track sharding across the GPUs of one box.

The CRF instances (tracks, the innermost axis of score[T,T,N]) are independent,
so the path shards with NO data-path collective: rank r owns a contiguous slice
of tracks and runs the whole sweep + backtrack locally.  The only exchange is the
gather of the packed decoded intervals (and the [N] log-partitions), one NCCL
all-gather of fixed-size records over NVLink; message sizes are latency-bound.

Shard upstream: a slice of an existing dense [T,T,N] tensor along its innermost
axis is strided, so each rank should produce (scorer) or receive its own
contiguous [T,T,N_r] block.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def track_shard(n_tracks: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced slice [lo, hi) of the track axis owned by `rank` (first ranks get the remainder)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(n_tracks, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(n_tracks: int, world: int) -> List[int]:
    return [track_shard(n_tracks, world, r)[1] - track_shard(n_tracks, world, r)[0] for r in range(world)]


def gather_decoded(pairs: torch.Tensor, counts: torch.Tensor, n_tracks: int, max_pairs: Optional[int] = None,
                   group=None):
    """All-gather packed decode results of a track-sharded problem.

    pairs [n_local, P, 2] int32, counts [n_local] int32 on every rank (n_local may differ by one between
    ranks).  Returns (pairs [n_tracks, P', 2], counts [n_tracks]) identical on every rank, tracks in global
    order.  P' = max_pairs (records are truncated to it before the exchange) or P.
    """
    world = dist.get_world_size(group)
    sizes = shard_sizes(n_tracks, world)
    nmax = max(sizes)
    P = pairs.shape[1] if max_pairs is None else min(max_pairs, pairs.shape[1])
    rec = torch.zeros((nmax, P * 2 + 1), dtype=torch.int32, device=pairs.device)  # fixed-size padded records
    n_local = pairs.shape[0]
    rec[:n_local, 0] = counts
    rec[:n_local, 1:] = pairs[:, :P].reshape(n_local, P * 2)
    out = torch.empty((world * nmax, P * 2 + 1), dtype=torch.int32, device=pairs.device)
    dist.all_gather_into_tensor(out, rec, group=group)
    out = out.view(world, nmax, P * 2 + 1)
    keep = torch.cat([out[r, :sizes[r]] for r in range(world)], 0)
    return keep[:, 1:].reshape(n_tracks, P, 2), keep[:, 0].contiguous()


def gather_vector(v: torch.Tensor, n_tracks: int, group=None) -> torch.Tensor:
    """All-gather a per-track vector ([n_local] -> [n_tracks]), e.g. log-partitions or log-probabilities."""
    world = dist.get_world_size(group)
    sizes = shard_sizes(n_tracks, world)
    nmax = max(sizes)
    buf = torch.zeros((nmax,), dtype=v.dtype, device=v.device)
    buf[: v.shape[0]] = v
    out = torch.empty((world * nmax,), dtype=v.dtype, device=v.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    out = out.view(world, nmax)
    return torch.cat([out[r, :sizes[r]] for r in range(world)], 0)


def gather_records(rec: torch.Tensor, n_tracks: int, group=None) -> torch.Tensor:
    """All-gather fixed-size per-track records ([n_local, R] -> [n_tracks, R]) with ONE collective and no
    packing copy when every rank owns the same number of tracks (the usual case: 88/8 = 11)."""
    world = dist.get_world_size(group)
    sizes = shard_sizes(n_tracks, world)
    nmax = max(sizes)
    if rec.shape[0] != nmax:  # uneven shard: pad this rank's block
        pad = torch.zeros((nmax, rec.shape[1]), dtype=rec.dtype, device=rec.device)
        pad[: rec.shape[0]] = rec
        rec = pad
    out = torch.empty((world * nmax, rec.shape[1]), dtype=rec.dtype, device=rec.device)
    dist.all_gather_into_tensor(out, rec.contiguous(), group=group)
    if min(sizes) == nmax:
        return out
    out = out.view(world, nmax, rec.shape[1])
    return torch.cat([out[r, : sizes[r]] for r in range(world)], 0)


def split_records(rec: torch.Tensor):
    """(counts [N] int32, logZ [N] fp32, pairs [N, 2T, 2] int32) views of gathered records."""
    T4 = rec.shape[1] - 2
    return rec[:, 0], rec[:, 1].view(torch.float32), rec[:, 2:].view(rec.shape[0], T4 // 2, 2)


class PushGather:
    """All-gather of the per-track records by copy engines, overlapped with the next step's sweep.

    The NCCL all-gather is latency-bound (a few MB) but its kernel cannot share the GPU with the cooperative sweep
    launch, which wants 143 of the 148 SMs at once: at 8 GPUs the exchange costs a third of a step.  Here every rank
    PUSHES its block into a symmetric (peer-mapped, NVLink) buffer of every other rank with plain device-to-device
    copies on a side stream -- DMA engines, no SMs -- followed by one signal-pad barrier (a one-CTA kernel), while the
    compute stream is already running the next sweep.  Buffers are double-buffered by step parity.  Requires every
    rank to own the same number of tracks and torch's symmetric memory (same node, P2P): `available()` says so, and
    callers fall back to `gather_records` otherwise."""

    def __init__(self, n_local: int, record_len: int, device: torch.device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.shape = (2, self.world, n_local, record_len)
        self.buf = symm_mem.empty(self.shape, dtype=torch.int32, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        peers = [self.hdl.get_buffer(r, self.shape, torch.int32) for r in range(self.world)]
        # destination views, built once (the per-step host work is on the critical path of a 0.3 ms step): my block in
        # the buffer of rank (me + k) % world, for both slots
        self.dst = [[peers[(self.rank + k) % self.world][slot, self.rank] for k in range(self.world)] for slot in (0, 1)]
        self.comm = torch.cuda.Stream(device=device)
        self.done = [None, None]  # per slot: event after which the slot holds a complete gather
        self.step = 0
        self.hdl.barrier(channel=0)

    def submit(self, rec: torch.Tensor) -> int:
        """Start the exchange of this step's records; returns the slot that will hold [world, n_local, R]."""
        slot = self.step & 1
        self.step += 1
        cur = torch.cuda.current_stream(rec.device)
        ready = torch.cuda.Event()
        ready.record(cur)
        rec.record_stream(self.comm)
        with torch.cuda.stream(self.comm):
            self.comm.wait_event(ready)
            # everyone has reached this submit, i.e. is done with the result this slot held two steps ago
            self.hdl.barrier(channel=2 + slot)
            for d in self.dst[slot]:  # my block into everyone's buffer (including mine)
                d.copy_(rec, non_blocking=True)
            self.hdl.barrier(channel=slot)  # every rank's pushes of this step have landed everywhere
            ev = torch.cuda.Event()
            ev.record(self.comm)
        self.done[slot] = ev
        return slot

    def result(self, slot: int) -> torch.Tensor:
        """[world * n_local, R] records of the step submitted into `slot` (valid once the current stream passed wait())."""
        return self.buf[slot].view(self.world * self.shape[2], self.shape[3])

    def wait(self) -> None:
        cur = torch.cuda.current_stream()
        for ev in self.done:
            if ev is not None:
                cur.wait_event(ev)
