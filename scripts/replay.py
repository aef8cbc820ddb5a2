"""Diagnostics: free-running time of the sweep's two roles.

A launch that reuses the previous launch's epoch finds every mailbox word and every far partial already
valid, so nobody ever waits: its duration is max(chain alone, far-field streaming alone) -- which of the two
bounds the real (dependent) launch.  Results are identical (same values rewritten).
usage: python scripts/replay.py [T] [N] [flags]"""
import os
import sys

os.environ.setdefault("TKB_SWEEP", "strip")  # these diagnostics target the strip design

import torch

sys.path.insert(0, "tests")
sys.path.insert(0, ".")
from golden_util import make_inputs  # noqa: E402
from transkun_b200 import _lib  # noqa: E402
from transkun_b200._lib import BACKWARD, SWEEP_LOGSUM, SWEEP_VITERBI  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
N = int(sys.argv[2]) if len(sys.argv) > 2 else 88
flags = int(sys.argv[3]) if len(sys.argv) > 3 else (SWEEP_VITERBI | SWEEP_LOGSUM)
L = _lib.load()
score, noise = make_inputs("randn", T, N, 1234)
s, z = torch.from_numpy(score).cuda(), torch.from_numpy(noise).cuda()
ws = torch.zeros(L.tkb_sweep_workspace_bytes(T, N), dtype=torch.uint8, device="cuda")
code = torch.empty((N, T), dtype=torch.int32, device="cuda")
lse = torch.empty((T, N), dtype=torch.float32, device="cuda")
st = torch.cuda.current_stream().cuda_stream


def run(epoch):
    rc = L.tkb_semicrf_sweep(s.data_ptr(), z.data_ptr(), T, N, BACKWARD, flags, ws.data_ptr(), epoch,
                             code.data_ptr(), None, lse.data_ptr(), st)
    _lib.check(rc, "sweep")


def timed(fn, reps=10):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn(i)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(reps))
    return ts[len(ts) // 2], ts[0]


for e in range(1, 4):
    run(e)
torch.cuda.synchronize()
ref_code, ref_lse = code.clone(), lse.clone()
med, best = timed(lambda i: run(10 + i))
print(f"dependent launches : median {med:.1f} us, best {best:.1f} us")
med, best = timed(lambda i: run(19))
print(f"replay (same epoch): median {med:.1f} us, best {best:.1f} us  (no waits: max of chain-only / stream-only)")
# (a replayed launch reads recycled ring slots: its results are not meaningful, only its duration)
alg = 4.0 * N * T * (T + 1) / 2
print(f"algorithmic bytes {alg / 1e6:.1f} MB -> replay {alg / med / 1e3:.0f} GB/s")
