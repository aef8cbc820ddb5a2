"""Generates tests/golden/frontend_*.npz from the UNMODIFIED reference frontend (run in the build container)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
from transkun.Util import MelSpectrum, makeFrame  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
for name, B, C, n, hop, W, nmel, nextra, fs, fmin, fmax, seed in [
        ("frontend_small", 2, 2, 3000, 64, 256, 20, 2, 8000, 30, 3500, 0),
        ("frontend_shipped_cfg", 1, 2, 9000, 1024, 4096, 229, 5, 44100, 30, 8000, 1)]:
    torch.manual_seed(seed)
    m = MelSpectrum(W, f_min=fmin, f_max=fmax, n_mels=nmel, fs=fs, nExtraWins=nextra, log=True, toMono=True).eval()
    with torch.no_grad():
        m.spectrogramExtractor.winGen.sigma.add_(torch.randn(nextra) * 0.3)
        m.spectrogramExtractor.winGen.center.add_(torch.randn(nextra) * 0.3)
    audio = torch.randn(B, C, n) * 0.3
    frames = makeFrame(audio, hop, W)
    with torch.no_grad():
        out = m(frames)
    sd = {k: v.numpy() for k, v in m.state_dict().items()}
    np.savez_compressed(os.path.join(HERE, name + ".npz"), audio=audio.numpy(), out=out.numpy(), hop=hop, W=W, nmel=nmel,
                        nextra=nextra, fs=fs, fmin=fmin, fmax=fmax, nframe=frames.shape[-2],
                        **{"sd_" + k: v for k, v in sd.items()})
    print(name, tuple(frames.shape), tuple(out.shape), float(out.min()), float(out.max()))
