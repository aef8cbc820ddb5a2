"""Diagnostics: per-block timeline of the persistent sweep kernel (needs the -DTKB_TIMELINE build:
python -c 'from transkun_b200.build import build_timeline; build_timeline()' ; run with
TKB_LIBRARY=transkun_b200/csrc/libtranskun_b200_timeline.so python scripts/timeline.py [T] [N])."""
import ctypes
import os
import sys

os.environ.setdefault("TKB_SWEEP", "strip")  # these diagnostics target the strip design

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
grid_max = 148
tl = torch.zeros((grid_max * 64 * 8,), dtype=torch.int64, device="cuda")
L.tkb_debug_set_timeline.argtypes = [ctypes.c_void_p]
L.tkb_debug_set_timeline(tl.data_ptr())
import os  # noqa: E402
for _ in range(3):
    tl.zero_()
    *_, ws = sweep(s, z, BACKWARD, flags)
    torch.cuda.synchronize()
if os.environ.get("TKB_REPLAY") == "1":  # same epoch again: nobody waits (free-running roles)
    ws.epoch -= 1
    tl.zero_()
    sweep(s, z, BACKWARD, flags)
    torch.cuda.synchronize()
    print("REPLAY launch (same epoch, no waits)")
raw = tl.cpu().numpy().astype(np.float64).reshape(grid_max, 64, 8)
t = raw[:, :, :4].copy()
G = (N + 7) // 8
nb = (T + 31) // 32
NSOLV = 4
F = min(nb, 16)
H = max(1, min((148 - F) // G - NSOLV, nb - ND - 1))
per = NSOLV + H
# solver clock64 stamps of the Viterbi chain warp of track 0 (cycles): 3 block start, 4 prep slot ready + far partial merged,
# 5 band ready, 7 chain done
for g in (0, G // 2):
    c = raw[g * NSOLV]
    its = [it for it in range(3, min(nb, 64) - 1)]
    prep_wait = np.mean([c[it, 4] - c[it, 3] for it in its])
    band_wait = np.mean([c[it, 5] - c[it, 4] for it in its])
    chain = np.mean([c[it, 7] - c[it, 5] for it in its])
    block = np.mean([c[it + 1, 3] - c[it, 3] for it in its])
    print(f"group {g} V chain cycles per block: {block:.0f} | prep wait+merge {prep_wait:.0f} | band wait {band_wait:.0f} | "
          f"32 columns {chain:.0f} ({chain / 32:.1f}/column) | rest {block - prep_wait - band_wait - chain:.0f}")
    print("   V band wait per block:", " ".join(f"{int(c[it, 5] - c[it, 4])}" for it in range(min(nb, 64))))
    print("   V prep wait per block:", " ".join(f"{int(c[it, 4] - c[it, 3])}" for it in range(min(nb, 64))))
    print("   V columns  per block:", " ".join(f"{int(c[it, 7] - c[it, 5])}" for it in range(min(nb, 64))))
    if flags & 2:
        prep_wait = np.mean([c[it, 1] - c[it, 0] for it in its])
        band_wait = np.mean([c[it, 2] - c[it, 1] for it in its])
        chain = np.mean([c[it, 6] - c[it, 2] for it in its])
        block = np.mean([c[it + 1, 0] - c[it, 0] for it in its])
        print(f"group {g} L chain cycles per block: {block:.0f} | prep wait+merge {prep_wait:.0f} | band wait {band_wait:.0f} | "
              f"32 columns {chain:.0f} ({chain / 32:.1f}/column) | rest {block - prep_wait - band_wait - chain:.0f}")

# strip CTAs (globaltimer ns): per unit 0 start, 1 near tiles done, 2 far field done, 3 partial published
nsolv = ((N + 1) // 2 + 3) // 4 * 4
nstrip = 148 - nsolv
K = 8
tt = raw[nsolv:nsolv + nstrip, :, :4]
valid = tt[:, :, 3] > 0
t00 = tt[:, :, 0][tt[:, :, 0] > 0].min()
near = (tt[:, :, 1] - tt[:, :, 0])[valid] / 1e3
far = (tt[:, :, 2] - tt[:, :, 1])[valid] / 1e3
pub = (tt[:, :, 3] - tt[:, :, 2])[valid] / 1e3
print(f"strips: {nstrip} CTAs, {valid.sum()} units with a far field; per unit: near {near.mean():.2f} us, far {far.mean():.2f} us, "
      f"partial {pub.mean():.2f} us; last publish at {(tt[:, :, 3].max() - t00) / 1e3:.1f} us")
for hh in (0, 50, nstrip - 1):
    line = []
    for uo in range(8):
        if tt[hh, uo, 0] <= 0:
            break
        u = hh + uo * nstrip
        J = nb - 1 - u // K
        R = T - (J * 32 + (ND + 1) * 32)
        stages = 0 if R < 1 else max(0, (((R + 3) // 4) - (u % K) + K - 1) // K)
        line.append(f"[J={J} k={u % K} stages={stages}: start {(tt[hh, uo, 0] - t00) / 1e3:.1f} near {(tt[hh, uo, 1] - tt[hh, uo, 0]) / 1e3:.1f} "
                    f"far {(tt[hh, uo, 2] - tt[hh, uo, 1]) / 1e3:.1f} pub {(tt[hh, uo, 3] - tt[hh, uo, 2]) / 1e3:.1f}]")
    print(f"  strip {hh}:", " ".join(line))
