"""Tiny driver: the log-mel frontend on one 16 s stereo segment at the shipped configuration (timing + ncu)."""
import sys
import time
import torch
sys.path.insert(0, ".")
from transkun_b200.Util import MelSpectrum, makeFrame
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
m = MelSpectrum(4096, f_min=30, f_max=8000, n_mels=229, fs=44100, nExtraWins=5, log=True, toMono=True).cuda().eval()
audio = torch.randn(B, 2, 705600, device="cuda") * 0.1
frames = makeFrame(audio, 1024, 4096)
with torch.no_grad():
    for _ in range(3):
        out = m(frames)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        out = m(frames)
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
alg = audio.numel() * 4 + out.numel() * 4
print(f"frontend B={B}: {ms*1e3:.0f} us per call, out {tuple(out.shape)}, algorithmic bytes {alg/1e6:.1f} MB -> {alg/ms/1e6:.1f} GB/s; "
      f"{B*16/ (ms*1e-3):.0f}x real time")
