"""The neural semi-CRF output layer, B200-native.

Drop-in for the reference class `NeuralSemiCRFInterval`
(/root/reference/transkun/CRF/NeuralSemiCRFInterval.py:553-588): same
constructor, same four methods, same return types.  All arithmetic runs in
hand-written sm_100a CUDA (libtranskun_b200.so, include/transkun_b200.h); this
file only validates arguments, allocates outputs/workspaces with torch and turns
packed device results into the Python objects the reference returns.  There is
no CPU path: tensors must live on a CUDA device.

    score      [T, T, N] : score of every closed interval [begin, end], laid out
                           [end, begin, track]; only end >= begin is read
    noiseScore [T-1, N]  : score of "no event between t and t+1"
"""
from __future__ import annotations

import gc
import itertools
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .. import _lib
from .._lib import BACKWARD, FORWARD, SWEEP_LOGSUM, SWEEP_VITERBI

Intervals = List[List[Tuple[int, int]]]


# ---------------------------------------------------------------------------
# plumbing
# ---------------------------------------------------------------------------
def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class _Workspace:
    """Mailbox of the persistent sweep kernel: zero-filled once, then reused with a growing epoch."""

    def __init__(self, T: int, N: int, device: torch.device):
        nbytes = _lib.load().tkb_sweep_workspace_bytes(T, N)
        self.buf = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        self.epoch = 0

    def next_epoch(self) -> int:
        self.epoch += 1
        if self.epoch >= 0xFFFFFFFF:
            self.buf.zero_()
            self.epoch = 1
        return self.epoch


_workspaces: Dict[tuple, _Workspace] = {}


def _workspace(T: int, N: int, device: torch.device, slot: int = 0) -> _Workspace:
    key = (device.index, _stream(device), T, N, slot)
    ws = _workspaces.get(key)
    if ws is None:
        if len(_workspaces) > 64:
            _workspaces.clear()
        ws = _workspaces[key] = _Workspace(T, N, device)
    return ws


def _raise_if_flagged(ws: Optional["_Workspace"]):
    """The status word of a sweep workspace holds the epoch of a launch whose inter-CTA wait timed out (0 if none
    ever did): only the launch that failed is flagged, the workspace stays usable.  Synchronises."""
    if ws is not None and ws.epoch != 0 and (int(ws.buf[:4].view(torch.int32).item()) & 0xFFFFFFFF) == ws.epoch:
        raise _lib.TkbError("semi-CRF sweep: an inter-CTA wait timed out; results are invalid")


def _poison_if_flagged(ws: "_Workspace", out: torch.Tensor) -> torch.Tensor:
    """Same check without a synchronisation, for results that stay on the device (log Z, log-probabilities): if
    the launch was flagged the values become NaN, so a timed-out sweep cannot pass for a valid loss."""
    epoch = ws.epoch if ws.epoch < 0x80000000 else ws.epoch - 0x100000000
    bad = ws.buf[:4].view(torch.int32) == epoch
    return torch.where(bad, torch.full_like(out, float("nan")), out)


def _check_inputs(score: torch.Tensor, noiseScore: torch.Tensor):
    # the reference asserts these in TorchScript (:17-18, :111-112, :209-215, :377-382)
    assert score.dim() == 3, "score must be [T, T, nBatch]"
    assert score.shape[0] == score.shape[1], "score must be square in its first two dims"
    T, N = score.shape[0], score.shape[2]
    assert noiseScore.dim() == 2 and noiseScore.shape[0] == T - 1 and noiseScore.shape[1] == N, \
        "noiseScore must be [T-1, nBatch]"
    if not score.is_cuda or not noiseScore.is_cuda:
        raise RuntimeError("transkun_b200 has no CPU path: score/noiseScore must be CUDA tensors")
    if score.device != noiseScore.device:
        raise RuntimeError("score and noiseScore must be on the same device")
    return T, N


def _prep(t: torch.Tensor) -> torch.Tensor:
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()  # the reference's DP tables are fp32 regardless of the input dtype (:22, :116)
    return t.contiguous()


def _pitched(t: torch.Tensor) -> bool:
    """[T, T, N] with the track axis contiguous and padded: strides (T * P, P, 1), P >= N."""
    return t.dim() == 3 and t.stride(2) == 1 and t.stride(1) >= t.shape[2] and t.stride(0) == t.shape[0] * t.stride(1)


_PAD_MIN_ELEMS = 1 << 22  # below this a sweep is launch-bound and the padding pass does not pay


def _prep_score_for_sweep(t: torch.Tensor) -> torch.Tensor:
    """The score tensor as the sweep wants it: fp32, track axis contiguous.  A track count that is not a multiple
    of 4 (the model's 90) only allows 8- or 4-byte copies from a dense tensor; a tensor that already comes padded
    (what our scorer emits) is used as it is, a dense one is re-laid out once with the track axis padded to a
    multiple of 4 when it is large enough for the extra pass to pay."""
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    N = t.shape[2]
    if _pitched(t) and (t.stride(1) % 4 == 0 or t.is_contiguous()):
        if t.is_contiguous() and N % 4 != 0 and t.numel() >= _PAD_MIN_ELEMS:
            buf = torch.zeros((t.shape[0], t.shape[1], (N + 3) // 4 * 4), dtype=torch.float32, device=t.device)
            buf[:, :, :N].copy_(t)
            return buf[:, :, :N]
        return t
    return t.contiguous()


def sweep(score: torch.Tensor, noise: torch.Tensor, direction: int, flags: int, want_vit: bool = False,
          slot: int = 0):
    """One pass of the persistent DP kernel.  Returns (code[N,T] int32 | None, vit[T,N] | None, lse[T,N] | None).
    `score` may have a padded track axis (strides (T*P, P, 1))."""
    T, N = score.shape[0], score.shape[2]
    dev = score.device
    L = _lib.load()
    assert _pitched(score), "score must be [T, T, N] with a contiguous (possibly padded) track axis"
    ws = _workspace(T, N, dev, slot)
    code = torch.empty((N, T), dtype=torch.int32, device=dev) if flags & SWEEP_VITERBI else None
    vit = torch.empty((T, N), dtype=torch.float32, device=dev) if (flags & SWEEP_VITERBI and want_vit) else None
    lse = torch.empty((T, N), dtype=torch.float32, device=dev) if flags & SWEEP_LOGSUM else None
    with torch.cuda.device(dev):
        rc = L.tkb_semicrf_sweep_pitched(_ptr(score), score.stride(1), _ptr(noise) if T > 1 else None, T, N, direction,
                                         flags, _ptr(ws.buf), ws.next_epoch(), _ptr(code), _ptr(vit), _ptr(lse),
                                         _stream(dev))
    _lib.check(rc, "tkb_semicrf_sweep_pitched")
    return code, vit, lse, ws


def backtrack(code: torch.Tensor, forced: Optional[torch.Tensor], direction: int):
    N, T = code.shape
    dev = code.device
    pairs = torch.empty((N, 2 * T, 2), dtype=torch.int32, device=dev)
    counts = torch.empty((N,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().tkb_semicrf_backtrack(_ptr(code), T, N, _ptr(forced), direction, _ptr(pairs), _ptr(counts),
                                               _stream(dev))
    _lib.check(rc, "tkb_semicrf_backtrack")
    return pairs, counts


def backtrack_records(code: torch.Tensor, forced: Optional[torch.Tensor], direction: int,
                      logz: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Back-tracking into ONE fixed-size int32 record per track: [count, logZ bits, 4*T pair slots].
    This is the unit the multi-GPU gather exchanges (transkun_b200.sharded.gather_records)."""
    N, T = code.shape
    dev = code.device
    rec = torch.empty((N, 2 + 4 * T), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().tkb_semicrf_backtrack_strided(_ptr(code), T, N, _ptr(forced), direction,
                                                       rec.data_ptr() + 8, 2 + 4 * T, rec.data_ptr(), 2 + 4 * T,
                                                       _stream(dev))
    _lib.check(rc, "tkb_semicrf_backtrack_strided")
    if logz is not None:
        rec[:, 1] = logz.contiguous().view(torch.int32)
    return rec


def _forced_tensor(forcedStartPos: Optional[Sequence[int]], N: int, T: int, dev) -> Optional[torch.Tensor]:
    if forcedStartPos is None:
        return None
    if torch.is_tensor(forcedStartPos):
        f = forcedStartPos.to(device=dev, dtype=torch.int32)
    else:
        f = np.asarray(list(forcedStartPos), dtype=np.int64)
        if f.shape != (N,):
            raise ValueError(f"forcedStartPos must have one entry per track ({N}), got shape {f.shape}")
        if (f < 0).any():
            raise IndexError("forcedStartPos must be >= 0")
        f = torch.from_numpy(np.minimum(f, T - 1).astype(np.int32)).to(dev, non_blocking=True)
    assert f.shape == (N,)
    return f.contiguous()


class PackedIntervals:
    """Interval lists as two arrays: pairs [M, 2] int32 (begin, end) and offsets [N + 1] int64 (CSR).  Build it once
    (e.g. in the data loader) with `pack_intervals` and pass it wherever the reference takes the list of lists
    (`evalPath`, `logProb`): the per-call flattening of ~10^5 Python tuples is the largest cost of a training step."""

    def __init__(self, pairs: torch.Tensor, offsets: torch.Tensor):
        assert pairs.dim() == 2 and pairs.shape[1] == 2 and offsets.dim() == 1
        self.pairs = pairs.to(torch.int32).contiguous()
        self.offsets = offsets.to(torch.int64).contiguous()
        # validated once, here: the kernels index score[end, begin] and the dense gradient with these values (the
        # reference's gather would raise an index error; an unchecked kernel would read and atomicAdd out of bounds)
        if self.pairs.numel():
            lo, hi = int(self.pairs.min()), int(self.pairs.max())
            if lo < 0:
                raise IndexError("interval endpoints must be >= 0")
            if bool((self.pairs[:, 0] > self.pairs[:, 1]).any()):
                raise ValueError("intervals must satisfy begin <= end")
            self.max_endpoint = hi
        else:
            self.max_endpoint = -1
        if self.offsets.numel() < 1 or int(self.offsets[0]) != 0 or int(self.offsets[-1]) != self.pairs.shape[0] or \
                bool((self.offsets[1:] < self.offsets[:-1]).any()):
            raise ValueError("offsets must be a CSR row-pointer array over pairs")


def pack_intervals(intervals: Intervals, T: Optional[int] = None) -> PackedIntervals:
    N = len(intervals)
    lens = np.fromiter((len(c) for c in intervals), dtype=np.int64, count=N)
    offsets = np.zeros(N + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    # one C-level pass over all endpoints (a list comprehension of tuples + np.asarray is 3x slower)
    flat = np.fromiter(itertools.chain.from_iterable(itertools.chain.from_iterable(intervals)), dtype=np.int32,
                       count=2 * int(offsets[-1])).reshape(-1, 2)
    if T is not None and flat.size and (flat.min() < 0 or flat.max() >= T):
        raise IndexError("interval endpoints must lie in [0, T)")
    return PackedIntervals(torch.from_numpy(np.ascontiguousarray(flat)), torch.from_numpy(offsets))


def _csr(intervals, N: int, T: int, dev):
    if isinstance(intervals, PackedIntervals):
        if intervals.offsets.numel() != N + 1:
            raise ValueError(f"intervals must have one list per track ({N}), got {intervals.offsets.numel() - 1}")
        if intervals.max_endpoint >= T:
            raise IndexError("interval endpoints must lie in [0, T)")
        pairs = intervals.pairs.to(dev, non_blocking=True)
        if pairs.numel() == 0:
            pairs = torch.zeros((1, 2), dtype=torch.int32, device=dev)
        return pairs, intervals.offsets.to(dev, non_blocking=True)
    if len(intervals) != N:
        raise ValueError(f"intervals must have one list per track ({N}), got {len(intervals)}")
    packed = pack_intervals(intervals, T)
    return _csr(packed, N, T, dev)


_pylists = None


def _load_pylists():
    """csrc/_tkb_pylists.so (built by transkun_b200.build from csrc/pylists.c): host-side list construction in C.
    Not a compute path: without it the same lists are built with zip() below, three times slower."""
    global _pylists
    if _pylists is None:
        import importlib.util
        import os
        path = os.path.join(os.path.dirname(_lib.lib_path()), "_tkb_pylists.so")
        _pylists = False
        if os.path.exists(path):
            try:
                spec = importlib.util.spec_from_file_location("_tkb_pylists", path)
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
                _pylists = mod
            except Exception:
                _pylists = False
    return _pylists


def _pairs_to_lists(pairs: torch.Tensor, counts: torch.Tensor) -> Intervals:
    counts_h = counts.cpu()  # synchronises (the reference synchronises at ptr.cpu(), :56)
    maxc = int(counts_h.max()) if counts_h.numel() else 0
    if maxc == 0:
        return [[] for _ in range(pairs.shape[0])]
    pairs_h = pairs[:, :maxc].cpu().numpy()
    # ~5e5 small objects are created here; the cyclic collector would run a young-generation pass every 700 of them
    # for nothing (tuples of ints cannot form cycles): 40 % of the time of this function
    gc_was_on = gc.isenabled()
    gc.disable()
    try:
        return _build_lists(pairs_h, counts_h, maxc)
    finally:
        if gc_was_on:
            gc.enable()


def _build_lists(pairs_h, counts_h, maxc) -> Intervals:
    helper = _load_pylists()
    if helper:
        return helper.pairs_to_lists(np.ascontiguousarray(pairs_h), counts_h.numpy(), maxc)
    # two flat int lists per track zipped into tuples: ~5x faster than tuple() over a list of 2-lists
    begins, ends = pairs_h[:, :, 0], pairs_h[:, :, 1]
    return [list(zip(begins[n, :c].tolist(), ends[n, :c].tolist())) for n, c in enumerate(counts_h.tolist())]


# ---------------------------------------------------------------------------
# autograd
# ---------------------------------------------------------------------------
def _alpha_beta(score: torch.Tensor, noise: torch.Tensor):
    """alpha (forward) and beta (backward) log-sum tables; two independent sweeps."""
    _, _, alpha, wsa = sweep(score, noise, FORWARD, SWEEP_LOGSUM, slot=1)
    _, _, beta, wsb = sweep(score, noise, BACKWARD, SWEEP_LOGSUM, slot=0)
    return alpha, beta, (wsa, wsb)


def _poison_all(wss, out: torch.Tensor) -> torch.Tensor:
    for ws in wss:
        out = _poison_if_flagged(ws, out)
    return out


def _marginals(score, noise, alpha, beta, gscale, want_score: bool, want_noise: bool):
    T, N = score.shape[0], score.shape[2]
    dev = score.device
    grad = torch.empty_like(score) if want_score else None
    gnoise = torch.empty_like(noise) if (want_noise and T > 1) else None
    with torch.cuda.device(dev):
        rc = _lib.load().tkb_semicrf_marginals(_ptr(score), _ptr(noise) if T > 1 else None, T, N, _ptr(alpha),
                                               _ptr(beta), _ptr(gscale), _ptr(grad), _ptr(gnoise), _stream(dev))
    _lib.check(rc, "tkb_semicrf_marginals")
    if want_noise and gnoise is None:
        gnoise = torch.zeros_like(noise)
    return grad, gnoise


class _LogZFn(torch.autograd.Function):
    """log Z with the closed-form marginal gradient (reference ComputeLogZFasterGrad, :459-475).

    Unlike the reference, the dense [T,T,N] gradient is not materialised in forward and
    kept alive; forward keeps only alpha/beta ([T,N] each) and backward writes
    grad * grad_output in a single pass over the triangle."""

    @staticmethod
    def forward(ctx, score, noiseScore):
        s, z = _prep(score), _prep(noiseScore)
        alpha, beta, wss = _alpha_beta(s, z)
        ctx.save_for_backward(s, z, alpha, beta)
        ctx.in_dtypes = (score.dtype, noiseScore.dtype)
        return _poison_all(wss, alpha[-1].clone())  # logZ = v[-1] (:417)

    @staticmethod
    def backward(ctx, grad_output):
        s, z, alpha, beta = ctx.saved_tensors
        g = grad_output.detach().to(torch.float32).contiguous()
        assert g.shape[-1] == s.shape[-1]  # (:471)
        grad, gnoise = _marginals(s, z, alpha, beta, g, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        if grad is not None and ctx.in_dtypes[0] != torch.float32:
            grad = grad.to(ctx.in_dtypes[0])
        if gnoise is not None and ctx.in_dtypes[1] != torch.float32:
            gnoise = gnoise.to(ctx.in_dtypes[1])
        return grad, gnoise


def _evalpath_forward(s, z, pairs, offsets):
    T, N = s.shape[0], s.shape[2]
    dev = s.device
    cum = torch.empty((T, N), dtype=torch.float32, device=dev)
    out = torch.empty((N,), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().tkb_semicrf_evalpath(_ptr(s), _ptr(z) if T > 1 else None, T, N, _ptr(pairs), _ptr(offsets),
                                              _ptr(cum), _ptr(out), _stream(dev))
    _lib.check(rc, "tkb_semicrf_evalpath")
    return out


def _evalpath_backward_into(T, N, pairs, offsets, g, sign, grad, gnoise, dev):
    with torch.cuda.device(dev):
        rc = _lib.load().tkb_semicrf_evalpath_grad(T, N, _ptr(pairs), _ptr(offsets), _ptr(g), float(sign), _ptr(grad),
                                                   _ptr(gnoise), _stream(dev))
    _lib.check(rc, "tkb_semicrf_evalpath_grad")


class _EvalPathFn(torch.autograd.Function):
    """Un-normalised path score (reference evalPath, :508-550); gradient = the gather's adjoint."""

    @staticmethod
    def forward(ctx, score, noiseScore, pairs, offsets):
        s, z = _prep(score), _prep(noiseScore)
        ctx.save_for_backward(pairs, offsets)
        ctx.shape = (s.shape[0], s.shape[2])
        ctx.in_dtypes = (score.dtype, noiseScore.dtype)
        return _evalpath_forward(s, z, pairs, offsets)

    @staticmethod
    def backward(ctx, grad_output):
        pairs, offsets = ctx.saved_tensors
        T, N = ctx.shape
        dev = grad_output.device
        g = grad_output.detach().to(torch.float32).contiguous()
        grad = torch.zeros((T, T, N), dtype=torch.float32, device=dev) if ctx.needs_input_grad[0] else None
        gnoise = torch.zeros((max(T - 1, 0), N), dtype=torch.float32, device=dev) if ctx.needs_input_grad[1] else None
        _evalpath_backward_into(T, N, pairs, offsets, g, 1.0, grad, gnoise, dev)
        if grad is not None:
            grad = grad.to(ctx.in_dtypes[0])
        if gnoise is not None:
            gnoise = gnoise.to(ctx.in_dtypes[1])
        return grad, gnoise, None, None


class _LogProbFn(torch.autograd.Function):
    """evalPath - logZ with ONE dense gradient buffer: path indicators minus marginals."""

    @staticmethod
    def forward(ctx, score, noiseScore, pairs, offsets):
        s, z = _prep(score), _prep(noiseScore)
        alpha, beta, wss = _alpha_beta(s, z)
        path = _evalpath_forward(s, z, pairs, offsets)
        ctx.save_for_backward(s, z, alpha, beta, pairs, offsets)
        ctx.in_dtypes = (score.dtype, noiseScore.dtype)
        return _poison_all(wss, path - alpha[-1])

    @staticmethod
    def backward(ctx, grad_output):
        s, z, alpha, beta, pairs, offsets = ctx.saved_tensors
        T, N = s.shape[0], s.shape[2]
        g = grad_output.detach().to(torch.float32).contiguous()
        grad, gnoise = _marginals(s, z, alpha, beta, -g, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        _evalpath_backward_into(T, N, pairs, offsets, g, 1.0, grad, gnoise, s.device)
        if grad is not None:
            grad = grad.to(ctx.in_dtypes[0])
        if gnoise is not None:
            gnoise = gnoise.to(ctx.in_dtypes[1])
        return grad, gnoise, None, None


# ---------------------------------------------------------------------------
# the reference's class
# ---------------------------------------------------------------------------
class NeuralSemiCRFInterval:
    def __init__(self, score, noiseScore):
        """The output layer for multiple tracks of non-overlapping intervals

        arguments:
        score -- the score matrix for all possible [begin, end] pairs, shape [T, T, nBatch]
        noiseScore -- the non-event score for the interval [t, t+1], shape [T-1, nBatch]
        (reference :554-564; the object only borrows the two tensors)
        """
        self.score = score
        self.noiseScore = noiseScore

    @classmethod
    def fromHost(cls, score: torch.Tensor, noiseScore: torch.Tensor, device, rows_per_chunk: int = 64):
        """Build the object from HOST tensors (the reference's users call `.cuda()` on both first,
        crfMinimalExample.py:13-14).  Only the part of `score` the semi-CRF reads (end >= begin) is uploaded -- a
        staircase of strided 2-D copies, about half the bytes of the dense tensor; the rest of the device tensor is
        left uninitialised (nothing reads it).  Asynchronous on the current stream when the host tensors are pinned."""
        device = torch.device(device)
        assert score.dim() == 3 and score.shape[0] == score.shape[1], "score must be [T, T, nBatch]"
        if score.is_cuda or noiseScore.is_cuda:
            raise RuntimeError("fromHost expects host tensors")
        T, N = score.shape[0], score.shape[2]
        s = score.detach().to(torch.float32).contiguous()
        # cells above the staircase are never read by the semi-CRF (tests/test_crf_gpu.py::test_from_host_uploads_only_what_is_read):
        # no 1.5 GB zero fill
        dev_score = torch.empty((T, T, N), dtype=torch.float32, device=device)
        with torch.cuda.device(device):
            rc = _lib.load().tkb_upload_lower_triangle(s.data_ptr(), dev_score.data_ptr(), T, N, rows_per_chunk,
                                                       _stream(device))
        _lib.check(rc, "tkb_upload_lower_triangle")
        obj = cls(dev_score, noiseScore.detach().to(torch.float32).to(device, non_blocking=True))
        obj._host_keepalive = s  # the copies are asynchronous
        return obj

    @staticmethod
    def lowerTriangleUploadBytes(T: int, N: int, rows_per_chunk: int = 64) -> int:
        return sum((min(e0 + rows_per_chunk, T) - e0) * min(e0 + rows_per_chunk, T) * N * 4
                   for e0 in range(0, T, rows_per_chunk))

    # -- packed device-side results (what a fused caller should use) -----------------------
    def decode_packed(self, forcedStartPos=None, forward: bool = False, with_logz: bool = False):
        """Viterbi on the device.  Returns (pairs[N, 2T, 2] int32, counts[N] int32, logZ[N] | None), all on
        score.device, nothing synchronised.  with_logz=True also evaluates the log-partition in the SAME
        pass over the score tensor (BACKWARD direction only, where both tables walk the triangle alike)."""
        T, N = _check_inputs(self.score, self.noiseScore)
        direction = FORWARD if forward else BACKWARD
        code, lse, ws = self._swept(direction, with_logz)
        forced = _forced_tensor(forcedStartPos, N, T, code.device)
        pairs, counts = backtrack(code, forced, direction)
        logz = None
        if with_logz:
            logz = _poison_if_flagged(ws, lse[T - 1 if forward else 0].clone())
        self._last_ws = ws
        return pairs, counts, logz

    def _swept(self, direction: int, with_logz: bool):
        """The DP tables do not depend on forcedStartPos (only back-tracking does, reference :61-71): the sweep of an
        object is run once per direction and reused by later decode() calls -- the segment loop of the model decodes
        every segment of a batch with a different forced start (transkun_b200.batched).  Keyed by the tensors'
        version counters, so an in-place update of score / noiseScore invalidates it."""
        key = (direction, self.score.data_ptr(), self.score._version, self.noiseScore.data_ptr(), self.noiseScore._version)
        cache = getattr(self, "_sweep_cache", None)
        if cache is not None and cache[0] == key and (cache[2] is not None or not with_logz):
            return cache[1], cache[2], cache[3]
        s, z = _prep_score_for_sweep(self.score), _prep(self.noiseScore)
        flags = SWEEP_VITERBI | (SWEEP_LOGSUM if with_logz else 0)
        code, _, lse, ws = sweep(s, z, direction, flags)
        self._sweep_cache = (key, code, lse, ws)
        return code, lse, ws

    def _raise_if_timed_out(self):
        ws = getattr(self, "_last_ws", None)
        _raise_if_flagged(ws)

    # -- reference surface -------------------------------------------------------------------
    def decode(self, forcedStartPos=None, forward=False) -> Intervals:
        """Viterbi decoding (reference :567-571 -> viterbi :107 / viterbiBackward :13)."""
        pairs, counts, _ = self.decode_packed(forcedStartPos, forward)
        out = _pairs_to_lists(pairs, counts)
        self._raise_if_timed_out()
        return out

    def decodeWithLogZ(self, forcedStartPos=None):
        """decode() and computeLogZ(noBackward=True) from one read of the score tensor."""
        pairs, counts, logz = self.decode_packed(forcedStartPos, False, with_logz=True)
        out = _pairs_to_lists(pairs, counts)
        self._raise_if_timed_out()
        return out, logz

    def evalPath(self, intervals: Intervals):
        """compute the unnormalized score (reference :574-577)"""
        T, N = _check_inputs(self.score, self.noiseScore)
        pairs, offsets = _csr(intervals, N, T, self.score.device)
        return _EvalPathFn.apply(self.score, self.noiseScore, pairs, offsets)

    def computeLogZ(self, noBackward=False):
        """compute the log normalization factor (reference :580-585)"""
        T, N = _check_inputs(self.score, self.noiseScore)
        needs_grad = torch.is_grad_enabled() and (self.score.requires_grad or self.noiseScore.requires_grad)
        if needs_grad:
            # noBackward=True differentiates the same value through autograd in the reference (:583);
            # the closed-form marginals are that gradient
            return _LogZFn.apply(self.score, self.noiseScore)
        s, z = _prep_score_for_sweep(self.score), _prep(self.noiseScore)
        _, _, alpha, ws = sweep(s, z, FORWARD, SWEEP_LOGSUM, slot=1)
        return _poison_if_flagged(ws, alpha[-1].clone())

    def logProb(self, intervals: Intervals, noBackward=False):
        """evalPath - computeLogZ (reference :587-588)"""
        T, N = _check_inputs(self.score, self.noiseScore)
        pairs, offsets = _csr(intervals, N, T, self.score.device)
        needs_grad = torch.is_grad_enabled() and (self.score.requires_grad or self.noiseScore.requires_grad)
        if needs_grad:
            return _LogProbFn.apply(self.score, self.noiseScore, pairs, offsets)
        s, z = _prep(self.score), _prep(self.noiseScore)
        _, _, alpha, ws = sweep(s, z, FORWARD, SWEEP_LOGSUM, slot=1)
        return _poison_if_flagged(ws, _evalpath_forward(s, z, pairs, offsets) - alpha[-1])
