"""Diagnostics for the default solver/helper sweep (semicrf_sweep.cu, -DTKB_TIMELINE build: scripts/build_all.sh): per-block globaltimer
stamps of the first chain warp of a solver CTA (0 block start, 1 far partial merged, 2 chain done) and of the helper CTAs
(0 start, 1 far field done, 3 partial published).  usage: TKB_LIBRARY=<timeline .so> python scripts/timeline.py [T] [N]"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "tests")
sys.path.insert(0, ".")
from golden_util import make_inputs  # noqa: E402
from transkun_b200 import _lib  # noqa: E402
from transkun_b200.CRF.NeuralSemiCRFInterval import sweep  # noqa: E402
from transkun_b200._lib import BACKWARD, SWEEP_LOGSUM, SWEEP_VITERBI  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
N = int(sys.argv[2]) if len(sys.argv) > 2 else 88
flags = int(sys.argv[3]) if len(sys.argv) > 3 else (SWEEP_VITERBI | SWEEP_LOGSUM)
ND = 2
L = _lib.load()
score, noise = make_inputs("randn", T, N, 1234)
s, z = torch.from_numpy(score).cuda(), torch.from_numpy(noise).cuda()
tl = torch.zeros((148 * 64 * 8,), dtype=torch.int64, device="cuda")
L.tkb_debug_set_timeline.argtypes = [ctypes.c_void_p]
L.tkb_debug_set_timeline(tl.data_ptr())
for _ in range(3):
    tl.zero_()
    *_, ws = sweep(s, z, BACKWARD, flags)
    torch.cuda.synchronize()
if os.environ.get("TKB_REPLAY") == "1":
    ws.epoch -= 1
    tl.zero_()
    sweep(s, z, BACKWARD, flags)
    torch.cuda.synchronize()
    print("REPLAY launch (same epoch, no waits)")
t = tl.cpu().numpy().astype(np.float64)[: 148 * 64 * 4].reshape(148, 64, 4)
G, nb = (N + 7) // 8, (T + 31) // 32
H = max(1, min(148 // G - 2, nb - ND - 1))
per = 2 + H
t0 = t[t > 0].min()
print(f"T={T} N={N} G={G} H={H}; span {(t.max() - t0) / 1e3:.1f} us")
for g in (0, G // 2):
    sol = (t[g * per] - t0) / 1e3
    its = range(3, min(nb, 64) - 1)
    wait = np.mean([sol[i, 1] - sol[i, 0] for i in its])
    chain = np.mean([sol[i, 2] - sol[i, 1] for i in its])
    step = np.mean([sol[i + 1, 0] - sol[i, 0] for i in its])
    print(f"group {g}: per block {step:.2f} us = far-partial wait {wait:.2f} + chain {chain:.2f} + rest {step - wait - chain:.2f}")
hs = []
for g in range(G):
    for h in range(H):
        a = t[g * per + 2 + h]
        m = a[:, 0] > 0
        if m.any():
            hs.append(((a[m, 3] - a[m, 0]).sum() / 1e3, (a[m, 1] - a[m, 0]).sum() / 1e3))
hs = np.array(hs)
print(f"helpers: busy mean {hs[:, 0].mean():.1f} us, max {hs[:, 0].max():.1f}; far-field part mean {hs[:, 1].mean():.1f}")
