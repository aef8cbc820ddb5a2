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
tl = torch.zeros((grid_max * 64 * 8 + grid_max * 64 * 16 * 8,), dtype=torch.int64, device="cuda")
L.tkb_debug_set_timeline.argtypes = [ctypes.c_void_p]
L.tkb_debug_set_timeline(tl.data_ptr())
for _ in range(3):
    tl.zero_()
    sweep(s, z, BACKWARD, flags)
    torch.cuda.synchronize()
tall = tl.cpu().numpy().astype(np.float64)
t = tall[: grid_max * 64 * 8].reshape(grid_max, 64, 8)
wt = tall[grid_max * 64 * 8:].reshape(grid_max, 64, 16, 8)
G = (N + 7) // 8
nb = (T + 31) // 32
K = min(148 // G, nb)
t0 = t[t > 0].min()
print(f"T={T} N={N} G={G} K={K} nb={nb}; kernel span {(t.max() - t0) / 1e3:.1f} us")
# stamps (thread 0 = Viterbi warp of track 0): 0 block start | 1 far field done | 2 partials merged (near tile
# starts) | 3 near tile done (diagonal solve starts) | 4 solve done
rows = []
for J in range(nb - 1, -1, -1):
    k = (nb - 1 - J) % K
    idx = (nb - 1 - J) // K
    rows.append((J, k, *((t[k, idx, :8] - t0) / 1e3)))
rows = np.array(rows)
print("   J cta    start  far_done | V:merged near_done solve_done | L:setup  near_done solve_done | V solve  L solve | V step  L step")
pv = pl = None
for i, (J, k, a, b, c, d, e, f, g2, h) in enumerate(rows):
    if i % max(1, nb // 32) == 0:
        sv = e - pv if pv is not None else 0.0
        sl = h - pl if pl is not None else 0.0
        print(f"{int(J):4d} {int(k):3d} {a:8.1f} {b:9.1f} | {c:8.1f} {d:9.1f} {e:10.1f} | {f:8.1f} {g2:9.1f} {h:10.1f} | {e-d:6.2f} {h-g2:7.2f} | {sv:6.2f} {sl:6.2f}")
    pv, pl = e, h
print(f"V: solve mean {np.mean(rows[:,6]-rows[:,5]):.2f} us, handoff (prev solve_done -> my solve start) mean {np.mean(rows[1:,5]-rows[:-1,6]):.2f}, "
      f"chain step mean {np.mean(np.diff(rows[:,6])):.2f}")
print(f"L: solve mean {np.mean(rows[:,9]-rows[:,8]):.2f} us, handoff mean {np.mean(rows[1:,8]-rows[:-1,9]):.2f}, "
      f"chain step mean {np.mean(np.diff(rows[:,9])):.2f};  L setup(after sync) - far_done mean {np.mean(rows[:,7]-rows[:,3]):.2f}")
print(f"L: prev L solve_done -> my far_done mean {np.mean(rows[1:,3]-rows[:-1,9]):.2f}; my far_done -> L setup done {np.mean(rows[:,7]-rows[:,3]):.2f}; "
      f"L setup done -> near done {np.mean(rows[:,8]-rows[:,7]):.2f}")

# per-warp stamps for a few mid-run blocks of group 0:
# far loop done | set-up done, at the barrier | merged (solve starts) | solve done
print("\nper-warp stamps relative to the PREVIOUS block's latest solve_done (us); warps 0-7 Viterbi, 8-15 log-sum")
for J in (40, 39, 20, 8):
    if J + 1 > nb - 1:
        continue
    k, idx = (nb - 1 - J) % K, (nb - 1 - J) // K
    kp, idxp = (nb - 1 - (J + 1)) % K, (nb - 1 - (J + 1)) // K
    kpp, idxpp = (nb - 1 - (J + 2)) % K, (nb - 1 - (J + 2)) // K
    ref = wt[kp, idxp, :, 5].max()
    refV, refL = wt[kp, idxp, :8, 5].max(), wt[kp, idxp, 8:, 5].max()
    print(f"block J={J}: prev block solve_done V {0.0:.2f} L {(refL - refV) / 1e3:.2f} (rel. to prev V done); "
          f"block J+2 solve_done V {(wt[kpp, idxpp, :8, 5].max() - refV) / 1e3:.2f} L {(wt[kpp, idxpp, 8:, 5].max() - refV) / 1e3:.2f}")
    for w in range(16):
        print(f"  warp {w:2d}: " + " ".join(f"{(wt[k, idx, w, sidx] - refV) / 1e3:8.2f}" for sidx in (0, 1, 4, 5)))
