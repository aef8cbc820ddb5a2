"""STFT / log-mel frontend, B200-native.

Mirrors the part of `transkun.Util` the model uses (/root/reference/transkun/Util.py:21-170):
`makeFrame`, `GaussianWindows`, `Spectrum`, `MelSpectrum`, with the same constructor arguments and
the same parameter / buffer names (`freq2mels`, `spectrogramExtractor.win`,
`spectrogramExtractor.winGen.sigma`, `.center`), so the shipped checkpoint's
`framewiseFeatureExtractor.*` entries load unchanged.

`MelSpectrum.forward(frames)` runs `tkb_logmel` (fused framing+window kernel -> cached batched cuFFT ->
fused power / channel-mean / banded-mel / log kernel); the window functions themselves are a few
thousand floats computed from the two learnable vectors with torch.  Inference path only: gradients
w.r.t. the window parameters are not implemented yet (raises under autograd).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib


def makeFrame(x, hopSize, windowSize, leftPaddingHalfFrame=True):
    """Overlapping frames as a strided VIEW of the padded signal: [..., nFrame, windowSize] (Util.py:21-43)."""
    assert hopSize < windowSize
    n = x.shape[-1]
    nFrame = math.ceil(n / hopSize) + 1
    lead = windowSize // 2 if leftPaddingHalfFrame else 0
    total = (nFrame - 1) * hopSize + windowSize
    padded = F.pad(x, (lead, total - lead - n))
    frames = padded.unfold(-1, windowSize, hopSize)
    assert frames.shape[-2] == nFrame, (frames.shape[-2], nFrame)
    return frames


def mel_filterbank_htk(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int) -> torch.Tensor:
    """Triangular HTK-scale filterbank without normalisation, [n_freqs, n_mels] -- what
    torchaudio.functional.melscale_fbanks returns with its defaults (Util.py:135-141)."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + f_min / 700.0)
    m_max = 2595.0 * math.log10(1.0 + f_max / 700.0)
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.clamp(torch.min(down, up), min=0.0)


class GaussianWindows(nn.Module):
    def __init__(self, n, nWin):
        super().__init__()
        self.n, self.nWin = n, nWin
        self.sigma = nn.Parameter(-torch.ones(n))
        self.center = nn.Parameter(torch.logit(torch.arange(1, n + 1) / (n + 1)))

    def get(self):
        """[nWin, n] Gaussian windows exp(-0.5 ((x - nWin*c) / (s*nWin/2))^2), c,s = sigmoid(params) (Util.py:62-69)."""
        s, c = torch.sigmoid(self.sigma), torch.sigmoid(self.center)
        x = torch.arange(self.nWin, device=self.sigma.device)
        return (-0.5 * ((x.unsqueeze(1) - self.nWin * c) / (s * self.nWin / 2)) ** 2).exp()


class Spectrum(nn.Module):
    """Holds the Hann window and the learnable Gaussian windows (Util.py:78-124); MelSpectrum drives the kernels."""

    def __init__(self, windowSize, nExtraWins=0, log=False):
        super().__init__()
        self.outputDim = windowSize // 2 + 1
        self.nChannel = nExtraWins + 1
        self.log = log
        self.register_buffer("win", torch.hann_window(windowSize))
        if nExtraWins > 0:
            self.winGen = GaussianWindows(nExtraWins, windowSize)
        self.nExtraWins = nExtraWins

    def windows(self) -> torch.Tensor:
        """[nExtraWins+1, windowSize]: Hann first, then the Gaussians (Util.py:99-102)."""
        if self.nExtraWins > 0:
            return torch.cat([self.win.unsqueeze(0), self.winGen.get().t()], dim=0).contiguous()
        return self.win.unsqueeze(0).contiguous()


class MelSpectrum(nn.Module):
    def __init__(self, windowSize, f_min, f_max, n_mels, fs, nExtraWins=0, log=False, eps=1e-5, toMono=False):
        super().__init__()
        self.outputDim = n_mels
        self.nChannel = nExtraWins + 1
        self.register_buffer("freq2mels", mel_filterbank_htk(windowSize // 2 + 1, f_min, f_max, n_mels, fs))
        self.log, self.eps, self.toMono = log, eps, toMono
        self.spectrogramExtractor = Spectrum(windowSize, nExtraWins)
        self._bands = None

    def _band_tables(self):
        fb = self.freq2mels
        key = (fb.data_ptr(), fb._version, fb.device)
        if self._bands is None or self._bands[0] != key:
            nz = fb > 0
            any_nz = nz.any(0)
            lo = torch.where(any_nz, nz.float().argmax(0), torch.zeros_like(any_nz, dtype=torch.long))
            hi = fb.shape[0] - nz.flip(0).float().argmax(0)  # one past the last non-zero row
            cnt = torch.where(any_nz, hi - lo, torch.zeros_like(lo))
            self._bands = (key, lo.to(torch.int32).contiguous(), cnt.to(torch.int32).contiguous())
        return self._bands[1], self._bands[2]

    def forward(self, frames):
        # output format: (., #frame, #mel, #window)   (Util.py:151-170)
        if not self.log:
            raise NotImplementedError("only the shipped configuration log=True")
        if not frames.is_cuda:
            raise RuntimeError("transkun_b200 has no CPU path: frames must be a CUDA tensor")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()) and self.training:
            raise NotImplementedError("frontend training path (gradients of the window parameters) not implemented")
        # the reference accepts any leading shape [..., nFrame, windowSize]; with toMono it averages dim -4 of the
        # [..., nFrame, nFreq, nWin] spectrum whenever that has >= 4 dims (Util.py:158-159), i.e. the dim right before
        # nFrame of the input whenever the input has >= 3 dims (keepdim=True)
        lead = frames.shape[:-2]
        if frames.dim() == 2:
            frames = frames[None, None]
        elif frames.dim() == 3:
            frames = frames.unsqueeze(0)
        elif frames.dim() > 4:
            frames = frames.reshape(-1, *frames.shape[-3:])
        assert frames.dim() == 4
        if frames.dtype != torch.float32 or frames.stride(-1) != 1:
            frames = frames.float().contiguous()
        B, C, Fr, W = frames.shape
        with torch.no_grad():
            wins = self.spectrogramExtractor.windows().float()
        nWin, nMel = wins.shape[0], self.freq2mels.shape[1]
        lo, cnt = self._band_tables()
        mono = 1 if (self.toMono and len(lead) >= 1) else 0
        L = _lib.load()
        ws = torch.empty(L.tkb_logmel_workspace_bytes(B, C, Fr, W, nWin), dtype=torch.uint8, device=frames.device)
        out = torch.empty((B, 1 if mono else C, Fr, nMel, nWin), dtype=torch.float32, device=frames.device)
        with torch.cuda.device(frames.device):
            rc = L.tkb_logmel(frames.data_ptr(), frames.stride(0), frames.stride(1), frames.stride(2), B, C, Fr, W,
                              wins.data_ptr(), nWin, self.freq2mels.data_ptr(), lo.data_ptr(), cnt.data_ptr(), nMel, mono,
                              float(self.eps), out.data_ptr(), ws.data_ptr(),
                              torch.cuda.current_stream(frames.device).cuda_stream)
        _lib.check(rc, "tkb_logmel")
        if len(lead) >= 1:
            lead = lead[:-1] + ((1,) if mono else (lead[-1],))
        return out.reshape(*lead, Fr, nMel, nWin)
