"""Free-running time of the sweep kernel: a launch that re-uses the previous launch's epoch finds every mailbox word and
far partial already valid, so nobody waits -- the duration is the slower role's own pace.  CUDA events around the bare
C-ABI call (outputs preallocated).  usage: python scripts/replay.py [T] [N] [flags]"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, "tests")
sys.path.insert(0, ".")
from golden_util import make_inputs  # noqa: E402
from transkun_b200 import _lib  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
N = int(sys.argv[2]) if len(sys.argv) > 2 else 88
flags = int(sys.argv[3]) if len(sys.argv) > 3 else 3
L = _lib.load()
score, noise = make_inputs("randn", T, N, 1234)
s, z = torch.from_numpy(score).cuda(), torch.from_numpy(noise).cuda()
ws = torch.zeros(L.tkb_sweep_workspace_bytes(T, N), dtype=torch.uint8, device="cuda")
code = torch.empty((N, T), dtype=torch.int32, device="cuda")
lse = torch.empty((T, N), dtype=torch.float32, device="cuda")
stream = torch.cuda.current_stream().cuda_stream


def launch(epoch):
    rc = L.tkb_semicrf_sweep(s.data_ptr(), z.data_ptr(), T, N, 0, flags, ws.data_ptr(), epoch, code.data_ptr(), None,
                             lse.data_ptr(), stream)
    _lib.check(rc, "sweep")


def timed(epoch, reps=5):
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        launch(epoch)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3)
    return best


for e in (1, 2, 3):
    launch(e)
torch.cuda.synchronize()
dep = []
for e in range(4, 9):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    launch(e)
    e1.record()
    torch.cuda.synchronize()
    dep.append(e0.elapsed_time(e1) * 1e3)
print(f"T={T} N={N} flags={flags}: dependent launch {min(dep):.1f} us (best of 5), replay {timed(8):.1f} us", flush=True)
