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


def _parse_cpulist(text: str) -> set:
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(device_index: int, sysfs: str = "/sys") -> Optional[int]:
    """One process per GPU: restrict the calling process to the CPUs of the NUMA node its GPU hangs off, BEFORE it
    allocates pinned host buffers (first touch then places them on that node, so the rank's host<->device copies do not
    cross the socket interconnect and do not share a memory controller with the other ranks' uploads).  Returns the node,
    or None when the topology cannot be read or the node's CPUs are not available to the process (nothing is changed)."""
    import os
    try:
        props = torch.cuda.get_device_properties(device_index)
        addr = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        node = int(open(os.path.join(sysfs, "bus/pci/devices", addr, "numa_node")).read())
        if node < 0:
            return None
        cpus = _parse_cpulist(open(os.path.join(sysfs, f"devices/system/node/node{node}/cpulist")).read())
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except (AttributeError, OSError, ValueError):
        return None


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


class FusedPushGather:
    """Decode + all-gather in ONE kernel: `tkb_semicrf_backtrack_push` back-tracks this rank's tracks and stores every
    track's record {count, logZ, pairs} straight into the symmetric (peer-mapped, NVLink) record buffer of every rank,
    then publishes a per-rank step flag; `result()` enqueues a one-thread kernel that waits for all ranks' flags.  No
    collective kernel, no copy engines, no barrier kernels: the next sweep starts as soon as the back-track has issued
    its stores.  Two record buffers alternate by step parity; the records of step k stay valid until submit(k+2), and a
    caller must have enqueued its reads of step k before it calls submit(k+1) (the kernel of step k+2 waits for every
    rank's flag of step k+1 before it overwrites the buffer).  Requires the same number of tracks on every rank and
    torch symmetric memory (one node, P2P)."""

    def __init__(self, n_local: int, T: int, device: torch.device, group=None):
        import ctypes

        import torch.distributed._symmetric_memory as symm_mem

        from . import _lib
        self._lib, self._ct = _lib, ctypes
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.n_local, self.T, self.R = n_local, T, 2 + 4 * T
        self.device = device
        slot_ints = self.world * n_local * self.R
        self.flag_off = 2 * slot_ints                      # int32 index of the flags inside the symmetric buffer
        total = self.flag_off + 64
        self.buf = symm_mem.empty((total,), dtype=torch.int32, device=device)
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        torch.cuda.synchronize(device)
        self.hdl.barrier(channel=0)                        # everybody's flags are zero before anybody pushes
        bases = [int(p) for p in self.hdl.buffer_ptrs]
        vp = ctypes.c_void_p
        self.rec_ptrs = [(vp * self.world)(*[b + 4 * s * slot_ints for b in bases]) for s in (0, 1)]
        self.flag_ptrs = (vp * self.world)(*[b + 4 * self.flag_off for b in bases])
        self.local = torch.zeros((16,), dtype=torch.int32, device=device)   # [0] ticket, [8] status
        self.step = 0
        self.slot_ints = slot_ints

    def submit(self, code: torch.Tensor, forced, direction: int, logz=None) -> int:
        """Back-track `code` [n_local, T] and push the records of this step; returns the step number."""
        assert code.shape == (self.n_local, self.T)
        self.step += 1
        stream = torch.cuda.current_stream(self.device).cuda_stream
        if logz is not None:
            logz = logz.contiguous()
        with torch.cuda.device(self.device):
            rc = self._lib.load().tkb_semicrf_backtrack_push(
                code.data_ptr(), self.T, self.n_local, None if forced is None else forced.data_ptr(), direction,
                None if logz is None else logz.data_ptr(), self.rec_ptrs[self.step & 1], self.flag_ptrs, self.world,
                self.rank, self.R, self.step, self.local.data_ptr(), self.local.data_ptr() + 32, stream)
        self._lib.check(rc, "tkb_semicrf_backtrack_push")
        self._keep = (code, forced, logz)   # alive until the kernel has run
        return self.step

    def result(self, step: int) -> torch.Tensor:
        """[world * n_local, 2 + 4T] records of `step` (tracks in global order), valid on the current stream after the
        flag wait enqueued here."""
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            rc = self._lib.load().tkb_wait_flags(self.buf.data_ptr() + 4 * self.flag_off, self.world, step,
                                                 self.local.data_ptr() + 32, stream)
        self._lib.check(rc, "tkb_wait_flags")
        s = step & 1
        return self.buf[s * self.slot_ints:(s + 1) * self.slot_ints].view(self.world * self.n_local, self.R)

    def timed_out(self) -> bool:
        """True if a flag wait gave up (synchronises)."""
        return int(self.local[8].item()) != 0
