"""Per-CTA phase times of the scorer kernel (diagnostics build, libtranskun_b200_timeline.so): start -> setup done -> first
operands landed -> last operands landed -> accumulators complete -> epilogue done.  usage: python scripts/scorer_trace.py [NT T]"""
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
trace = torch.zeros((1 << 16, 8), dtype=torch.int64, device="cuda")
lib.tkb_sip_score_scaled.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 3 + [ctypes.c_float, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
lib.tkb_debug_set_scorer_trace.argtypes = [ctypes.c_void_p]


def call():
    rc = lib.tkb_sip_score_scaled(q.data_ptr(), k.data_ptr(), d.data_ptr(), NT, T, D, 1.0 / 16.0, S.data_ptr(), P, None)
    assert rc == 0


for _ in range(2):
    call()
lib.tkb_debug_set_scorer_trace(trace.data_ptr())
call()
torch.cuda.synchronize()
t = trace.cpu().numpy()
t = t[t[:, 0] > 0]
alive = t[:, 4] > 0
print(f"NT={NT} T={T}: {len(t)} CTAs ({int(alive.sum())} with a tile), kernel span {(t[:, 5].max() - t[:, 0].min()) / 1e3:.0f} us")
a = t[alive].astype(np.float64)
for name, i, j in (("setup (TMEM alloc, barriers, sync)", 0, 1), ("first operands", 1, 2), ("main loop (first -> last operands landed)", 2, 3),
                   ("last MMAs", 3, 4), ("epilogue", 4, 5), ("whole CTA", 0, 5)):
    dt = (a[:, j] - a[:, i]) / 1e3
    print(f"  {name:45s} mean {dt.mean():7.2f} us   p10 {np.percentile(dt, 10):7.2f}   p90 {np.percentile(dt, 90):7.2f}")
