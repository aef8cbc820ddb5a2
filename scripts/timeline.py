"""Diagnostics for the sweep kernel (-DTKB_TIMELINE build: python -c "from transkun_b200 import build; build.build_timeline()").
globaltimer stamps: chain warp 0 of every solver CTA per 32-column block (0 block start, 1 far partial merged, 2 chain done);
thread 0 of every streaming CTA per batch of 8 rows (0 top of the loop, 1 score rows + mailbox rows present, 2 batch done),
fetch warp (3 = validated mailbox rows handed to the consumers).
usage: TKB_LIBRARY=transkun_b200/csrc/libtranskun_b200_timeline.so python scripts/timeline.py [T] [N] [flags]"""
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
L = _lib.load()
score, noise = make_inputs("randn", T, N, 1234)
s, z = torch.from_numpy(score).cuda(), torch.from_numpy(noise).cuda()
NS = 256
tl = torch.zeros((148 * NS * 8,), dtype=torch.int64, device="cuda")
L.tkb_debug_set_timeline.argtypes = [ctypes.c_void_p]
L.tkb_debug_set_timeline(tl.data_ptr())
if os.environ.get("TKB_DBG"):
    L.tkb_debug_set_flags(int(os.environ["TKB_DBG"]))
    print("debug flags", os.environ["TKB_DBG"], "(1 = no arithmetic, 2 = no score prefetch): RESULTS ARE GARBAGE")
for _ in range(3):
    tl.zero_()
    *_, ws = sweep(s, z, BACKWARD, flags)
    torch.cuda.synchronize()
if os.environ.get("TKB_REPLAY") == "1":  # same epoch again: every mailbox word / partial is already valid, nobody waits
    ws.epoch -= 1
    tl.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sweep(s, z, BACKWARD, flags)
    e1.record()
    torch.cuda.synchronize()
    print(f"REPLAY launch (same epoch, no waits): {e0.elapsed_time(e1) * 1e3:.1f} us by CUDA events")
t = tl.cpu().numpy().astype(np.float64).reshape(148, NS, 8)
S = (N + 3) // 4
nb = (T + 31) // 32
S = (N + 3) // 4
ck = t[S:, 248:254, :2]
print("consumer warp 0 cycles per batch (mean over CTAs): stamp/top %.0f, issue %.0f, cp.async wait %.0f, mailbox wait %.0f, compute %.0f, release %.0f" % tuple(ck[:, k, 0].mean() / max(1, (T - 1) // 8 - 11) for k in range(6)))
print("consumer warp 1 cycles per batch (mean over CTAs): stamp/top %.0f, issue %.0f, cp.async wait %.0f, mailbox wait %.0f, compute %.0f, release %.0f" % tuple(ck[:, k, 1].mean() / max(1, (T - 1) // 8 - 11) for k in range(6)))
pk = t[S:, 248:254, 2:6]
for w in range(4):
    print("mailbox producer warp %d cycles per batch: top %.0f, copy issue %.0f, buffer-free wait %.0f, copy-landed wait %.0f, validate+store %.0f, hand-over %.0f" % ((w,) + tuple(pk[:, k, w].mean() / max(1, (T - 1) // 8 - 11) for k in range(6))))
t[:, 248:254, :] = 0
if not (t[:, :248] > 0).any():
    sys.exit(0)
t0 = t[:, :248][t[:, :248] > 0].min()
print(f"T={T} N={N} solvers={S}; span {(t.max() - t0) / 1e3:.1f} us")
end = (t[:, 255, :3] - t0) / 1e3
def mx(a):
    a = a[a > 0]
    return f"{a.max():.1f} (min {a.min():.1f}, n={a.size})" if a.size else "none"
print(f"role end times (us): chain warp 0 {mx(end[:S, 0])}; publisher warp 0 {mx(end[:S, 1])}; TMA thread {mx(end[:S, 2])}; "
      f"consumer thread 0 {mx(end[S:, 0])}; mailbox producer 0 {mx(end[S:, 1])}")
sol = (t[:S] - t0) / 1e3
its = list(range(3, min(nb, NS) - 1))
for sidx in (0, S // 2, S - 1):
    a = sol[sidx]
    wait = np.mean([a[i, 1] - a[i, 0] for i in its])
    chain = np.mean([a[i, 2] - a[i, 1] for i in its])
    step = np.mean([a[i + 1, 0] - a[i, 0] for i in its])
    band = np.mean([a[i, 3] - a[i, 1] for i in its])
    print(f"solver {sidx}: per block {step:.2f} us = far-partial wait {wait:.2f} + band wait {band:.2f} + steps {chain - band:.2f} + rest {step - wait - chain:.2f}")
a = sol[0]
print("solver 0 block starts (us), every 4th block:", " ".join(f"{a[i, 0]:.0f}" for i in range(0, min(nb, NS), 4)))
print("solver 0 far-partial wait per block (us):", " ".join(f"{a[i, 1] - a[i, 0]:.1f}" for i in range(0, min(nb, NS), 2)))
for h in (0, 60, 125):
    a = (t[S + h] - t0) / 1e3
    m = a[:, 0] > 0
    nbat = int(m.sum())
    if nbat == 0:
        continue
    w = a[:nbat, 1] - a[:nbat, 0]
    c = a[:nbat, 2] - a[:nbat, 1]
    per = np.diff(a[:nbat, 0])
    print(f"streaming CTA {h}: {nbat} batches; per batch mean {per.mean():.2f} us (wait {w.mean():.2f}, compute {c.mean():.2f}); "
          f"first {a[0, 0]:.0f} us last {a[nbat - 1, 2]:.0f} us")
    print("   per-batch period every 16th:", " ".join(f"{per[i]:.2f}" for i in range(0, nbat - 1, 16)))
    print("   wait     every 16th:", " ".join(f"{w[i]:.2f}" for i in range(0, nbat, 16)))
    print("   compute  every 16th:", " ".join(f"{c[i]:.2f}" for i in range(0, nbat, 16)))
    f = a[:nbat]
    ok = f[:, 4] > 0
    print("   producer warp 0 per batch: period %.2f us; loads issued + buffer free %.2f, validated + stored %.2f after its loop top; batches with a retry: %d" % (
        np.mean(np.diff(f[ok, 4])), np.mean(f[ok, 5] - f[ok, 4]), np.mean(f[ok, 7] - f[ok, 4]), int((f[:, 6] > 0).sum())))
    print("   producer hand-over minus consumer loop top, every 16th:", " ".join(f"{f[i, 7] - a[i, 0]:.2f}" for i in range(0, nbat, 16)))
