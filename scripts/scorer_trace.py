"""Per-CTA phase times of the scorer kernel (diagnostics build, libtranskun_b200_timeline.so): setup, the first work item's main loop and
epilogue, and the average per work item of the persistent loop.  usage: python scripts/scorer_trace.py [NT T]"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(os.path.join(ROOT, "transkun_b200", "csrc", "libtranskun_b200_timeline.so"))
NT, T, D = (int(sys.argv[1]), int(sys.argv[2]), 256) if len(sys.argv) > 2 else (88, 2048, 256)
g = torch.Generator().manual_seed(0)
q, k, d = (torch.randn(NT, T, D, generator=g).cuda(), torch.randn(NT, T, D, generator=g).cuda(),
           torch.randn(NT, T, generator=g).cuda())
P = (NT + 7) // 8 * 8
S = torch.empty((T, T, P), device="cuda")
trace = torch.zeros((1 << 12, 16), dtype=torch.int64, device="cuda")
lib.tkb_sip_score_scaled.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 3 + [ctypes.c_float, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
lib.tkb_debug_set_scorer_trace.argtypes = [ctypes.c_void_p]


def call():
    rc = lib.tkb_sip_score_scaled(q.data_ptr(), k.data_ptr(), d.data_ptr(), NT, T, D, 1.0 / 16.0, S.data_ptr(), P, None)
    if rc != 0:
        lib.tkb_last_error.restype = ctypes.c_char_p
        raise RuntimeError(lib.tkb_last_error().decode())


for _ in range(2):
    call()
lib.tkb_debug_set_scorer_trace(trace.data_ptr())
call()
torch.cuda.synchronize()
t = trace.cpu().numpy()
t = t[t[:, 0] > 0].astype(np.float64)
tiles = sum(-(-T // 64) - 2 * c for c in range(-(-T // 128)))
items = tiles * -(-NT // 8)
print(f"NT={NT} T={T}: {len(t)} persistent CTAs, {items} work items ({items / len(t):.1f} per CTA), kernel span {(t[:, 5].max() - t[:, 0].min()) / 1e3:.0f} us")
for name, i, j in (("setup (TMEM alloc, barriers, sync)", 0, 1), ("first item: operands + MMAs", 1, 2), ("first item: epilogue", 2, 3),
                   ("all items", 1, 4), ("whole CTA", 0, 5)):
    dt = (t[:, j] - t[:, i]) / 1e3
    print(f"  {name:40s} mean {dt.mean():7.2f} us   p10 {np.percentile(dt, 10):7.2f}   p90 {np.percentile(dt, 90):7.2f}")
print(f"  per work item: {((t[:, 4] - t[:, 1]) / 1e3).mean() / (items / len(t)):.2f} us")
steps = t[:, 12].mean()
for name, i in (("producer: waiting for a free stage", 8), ("producer: expect_tx + TMA issue", 9), ("MMA thread: waiting for operands", 10),
                ("MMA thread: fence + MMAs + commit", 11), ("MMA thread: waiting for TMEM (per item)", 13)):
    per = t[:, i].mean() / (items / len(t) if i == 13 else steps)
    print(f"  {name:42s} {per:8.0f} cycles per {'item' if i == 13 else 'step'}")
print(f"  SM clock during the kernel: {((t[:, 15] - t[:, 14]) / (t[:, 5] - t[:, 0])).mean() * 1e3:.0f} MHz")
