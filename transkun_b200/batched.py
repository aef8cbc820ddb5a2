"""Batched segment pipeline for the reference's `TransKun.transcribe`.

The reference transcribes a file segment by segment (16 s windows, 8 s hop), one `transcribeFrames` call with batch 1
per segment (/root/reference/transkun/ModelTransformer.py:758-828): frontend, backbone, scorer and the semi-CRF of every
segment are separate small launches, and at the model shape (T=691, N=90) none of them can fill a B200.  The ONLY
dependency between segments is the forced start position of the back-tracking (:789-791: next startPos = last decoded
position - hop) -- the DP tables do not depend on it (NeuralSemiCRFInterval.py:61-71).

`transcribe_batched(model, x)` therefore
  1. cuts the padded signal into the same segments the reference's loop would (:741-774) and stacks their frames,
  2. runs the model's own `processFramesBatch` (:151-225) on chunks of `max_batch` segments -- one frontend launch, one
     backbone pass, one scorer launch, and ONE semi-CRF sweep over N = 90 * nSeg tracks per chunk,
  3. calls the reference's own, unmodified `transcribe` (:729-848) with `processFramesBatch` temporarily answering
     from those results: the sequential part that remains is the per-segment back-track (a 15 us kernel), the attribute
     heads and the host-side event merging, all executed by the reference's code.
The result is the Note list `model.transcribe(x)` returns.  Works on a reference model with the B200-native modules
installed (transkun_b200.transcribe.install_into); the CRF objects must be ours (they cache the sweep).
"""
from __future__ import annotations

import math
from typing import List

import torch
import torch.nn.functional as F

from .CRF.NeuralSemiCRFInterval import NeuralSemiCRFInterval, _forced_tensor, _pairs_to_lists, backtrack
from ._lib import BACKWARD, FORWARD


class _SegmentCRF:
    """The tracks [lo, hi) of a batched NeuralSemiCRFInterval whose sweep has already run: decode() is a back-track of
    this segment's rows of the shared back-pointer table (same surface as the reference object, :553-588)."""

    def __init__(self, parent: NeuralSemiCRFInterval, lo: int, hi: int):
        self.parent, self.lo, self.hi = parent, lo, hi

    @property
    def score(self):
        return self.parent.score[:, :, self.lo:self.hi]

    @property
    def noiseScore(self):
        return self.parent.noiseScore[:, self.lo:self.hi]

    def decode(self, forcedStartPos=None, forward=False):
        direction = FORWARD if forward else BACKWARD
        code, _, ws = self.parent._swept(direction, False)
        T = code.shape[1]
        sub = code[self.lo:self.hi]
        forced = _forced_tensor(forcedStartPos, self.hi - self.lo, T, sub.device)
        pairs, counts = backtrack(sub, forced, direction)
        out = _pairs_to_lists(pairs, counts)
        self.parent._last_ws = ws
        self.parent._raise_if_timed_out()
        return out

    def _sub(self):
        return NeuralSemiCRFInterval(self.score.contiguous(), self.noiseScore.contiguous())

    def evalPath(self, intervals):
        return self._sub().evalPath(intervals)

    def computeLogZ(self, noBackward=False):
        return self._sub().computeLogZ(noBackward)

    def logProb(self, intervals, noBackward=False):
        return self._sub().logProb(intervals, noBackward)


def segment_frames(model, x: torch.Tensor, stepInSecond=None, segmentSizeInSecond=None) -> torch.Tensor:
    """Frames of every segment the reference's loop visits, stacked: [nSeg, C, nFrame, windowSize] (:729-774)."""
    from .Util import makeFrame
    if stepInSecond is None and segmentSizeInSecond is None:
        stepInSecond, segmentSizeInSecond = model.segmentHopSizeInSecond, model.segmentSizeInSecond
    x = x.transpose(-1, -2)
    pad = segmentSizeInSecond - stepInSecond
    x = F.pad(x, (math.ceil(pad * model.fs), math.ceil(model.fs * pad)))
    nSample = x.shape[-1]
    stepSize = math.ceil(stepInSecond * model.fs / model.hopSize) * model.hopSize
    segmentSize = math.ceil(segmentSizeInSecond * model.fs)
    frames = []
    for i in range(0, nSample, stepSize):
        cur = x[:, i:min(i + segmentSize, nSample)]
        if cur.shape[-1] < segmentSize:
            cur = F.pad(cur, (0, segmentSize - cur.shape[-1]))
        frames.append(makeFrame(cur, model.hopSize, model.windowSize))
    return torch.stack(frames, 0)


def transcribe_batched(model, x: torch.Tensor, max_batch: int = 8, **kwargs) -> List:
    """Drop-in for `model.transcribe(x, **kwargs)` with the per-segment network and semi-CRF work batched."""
    frames = segment_frames(model, x, kwargs.get("stepInSecond"), kwargs.get("segmentSizeInSecond"))
    nSeg = frames.shape[0]
    nSym = len(model.targetMIDIPitch)
    original = model.processFramesBatch
    results = []
    with torch.no_grad():
        for s0 in range(0, nSeg, max_batch):
            crf, ctx = original(frames[s0:s0 + max_batch])
            if not isinstance(crf, NeuralSemiCRFInterval):
                raise RuntimeError("transcribe_batched needs the B200-native CRF installed (transkun_b200.transcribe.install)")
            crf._swept(BACKWARD, False)  # ONE sweep for all segments of the chunk
            for k in range(ctx.shape[0]):
                results.append((_SegmentCRF(crf, k * nSym, (k + 1) * nSym), ctx[k:k + 1]))
    it = iter(results)

    def answer(framesBatch):
        # the reference's loop asks for the segments in order, one at a time (:785)
        assert framesBatch.shape[0] == 1
        return next(it)

    model.processFramesBatch = answer
    try:
        with torch.no_grad():
            return model.transcribe(x, **kwargs)
    finally:
        del model.processFramesBatch  # the instance attribute shadows the class method only during this call
