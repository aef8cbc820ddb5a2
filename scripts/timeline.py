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
L = _lib.load()
score, noise = make_inputs("randn", T, N, 1234)
s, z = torch.from_numpy(score).cuda(), torch.from_numpy(noise).cuda()
grid_max = 148
tl = torch.zeros((grid_max, 64, 4), dtype=torch.int64, device="cuda")
L.tkb_debug_set_timeline.argtypes = [ctypes.c_void_p]
L.tkb_debug_set_timeline(tl.data_ptr())
for _ in range(3):
    tl.zero_()
    sweep(s, z, BACKWARD, flags)
    torch.cuda.synchronize()
t = tl.cpu().numpy().astype(np.float64)
G = (N + 7) // 8
nb = (T + 31) // 32
K = min(148 // G, nb)
t0 = t[t > 0].min()
print(f"T={T} N={N} G={G} K={K} nb={nb}; kernel span {(t.max() - t0) / 1e3:.1f} us")
# group 0: blocks in chain order
rows = []
for J in range(nb - 1, -1, -1):
    k = (nb - 1 - J) % K
    idx = (nb - 1 - J) // K
    st = (t[k, idx] - t0) / 1e3
    rows.append((J, k, *st))
rows = np.array(rows)
print("  J  cta  far_start  far_end   solve_start solve_end | far_us wait_sync_us solve_us | chain_gap_us")
prev_end = None
for J, k, a, b, c, d in rows[:: max(1, nb // 32)]:
    print(f"{int(J):4d} {int(k):3d} {a:10.1f} {b:9.1f} {c:11.1f} {d:9.1f} | {b - a:6.1f} {c - b:8.1f} {d - c:8.1f}")
solve = rows[:, 5] - rows[:, 4]
gap = rows[1:, 4] - rows[:-1, 5]   # solve start of next block minus solve end of previous block
print(f"solve_us mean {solve.mean():.2f} min {solve.min():.2f} max {solve.max():.2f}")
print(f"handoff gap (next solve start - prev solve end) mean {gap.mean():.2f} median {np.median(gap):.2f} max {gap.max():.2f}")
print(f"far_us mean {np.mean(rows[:, 3] - rows[:, 2]):.2f}; chain per block {(rows[-1, 5] - rows[0, 4]) / (nb - 1):.2f} us")
