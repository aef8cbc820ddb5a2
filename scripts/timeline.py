"""Diagnostics: per-block timeline of the persistent sweep kernel (needs the -DTKB_TIMELINE build:
python -c 'from transkun_b200.build import build_timeline; build_timeline()' ; run with
TKB_LIBRARY=transkun_b200/csrc/libtranskun_b200_timeline.so python scripts/timeline.py [T] [N])."""
import ctypes
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
grid_max = 148
tl = torch.zeros((grid_max * 64 * 4,), dtype=torch.int64, device="cuda")
L.tkb_debug_set_timeline.argtypes = [ctypes.c_void_p]
L.tkb_debug_set_timeline(tl.data_ptr())
for _ in range(3):
    tl.zero_()
    sweep(s, z, BACKWARD, flags)
    torch.cuda.synchronize()
t = tl.cpu().numpy().astype(np.float64).reshape(grid_max, 64, 4)
G = (N + 7) // 8
nb = (T + 31) // 32
H = max(1, min(148 // G - 2, nb - ND - 1))
per = 2 + H
t0 = t[t > 0].min()
print(f"T={T} N={N} G={G} H={H} nb={nb}; kernel span {(t.max() - t0) / 1e3:.1f} us")
for g in (0, G // 2):
    sol = (t[g * per] - t0) / 1e3  # solver of quad 0: [block it][start, far partial merged, chain done]
    print(f"group {g}: solver blocks (us since kernel start): start | wait for far partial | chain | block time")
    prev = None
    for it in range(min(nb, 64)):
        st, mg, dn = sol[it, 0], sol[it, 1], sol[it, 2]
        j = nb - 1 - it
        own = None
        if j <= nb - ND - 2:
            h = (nb - ND - 2 - j) % H
            idx = (nb - ND - 2 - j) // H
            if idx < 64:
                own = (t[g * per + 2 + h, idx] - t0) / 1e3  # helper: start, far done, merged-sync, published
        line = f"  j={j:3d} start {st:7.2f} | wait {mg - st:5.2f} | chain {dn - mg:5.2f} | step {0.0 if prev is None else st - prev:5.2f}"
        if own is not None:
            line += f" || helper start {own[0]:7.2f} far_done {own[1]:7.2f} published {own[3]:7.2f} (slack {st - own[3]:6.2f})"
        print(line)
        prev = st
hs = []
for g in range(G):
    for h in range(H):
        a = t[g * per + 2 + h]
        m = a[:, 0] > 0
        if m.any():
            hs.append(((a[m, 3] - a[m, 0]).sum() / 1e3, (a[m, 1] - a[m, 0]).sum() / 1e3))
hs = np.array(hs)
print(f"helpers: busy (start->published) mean {hs[:, 0].mean():.1f} us, max {hs[:, 0].max():.1f}; far-field part mean {hs[:, 1].mean():.1f}")
