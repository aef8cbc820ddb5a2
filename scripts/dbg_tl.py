import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from golden_util import make_inputs
from transkun_b200 import _lib
from transkun_b200.CRF.NeuralSemiCRFInterval import sweep
from transkun_b200._lib import BACKWARD
T, N, flags = 2048, 88, int(sys.argv[1]) if len(sys.argv) > 1 else 3
L = _lib.load()
score, noise = make_inputs("randn", T, N, 1234)
s, z = torch.from_numpy(score).cuda(), torch.from_numpy(noise).cuda()
tl = torch.zeros((148 * 256 * 8,), dtype=torch.int64, device="cuda")
L.tkb_debug_set_timeline.argtypes = [ctypes.c_void_p]
L.tkb_debug_set_timeline(tl.data_ptr())
L.tkb_debug_set_flags(int(os.environ.get("TKB_DBG", "0")))
for _ in range(3):
    *_, ws = sweep(s, z, BACKWARD, flags)
torch.cuda.synchronize()
ws.epoch -= 1
tl.zero_()
sweep(s, z, BACKWARD, flags)
torch.cuda.synchronize()
t = tl.cpu().numpy().reshape(148, 256, 8)
nz = np.argwhere(t[0] > 0)
print("block 0 nonzero idx range", nz[:, 0].min(), nz[:, 0].max(), "slots", sorted(set(nz[:, 1])))
print("block 0 idx 255:", t[0, 255], "idx 248..253 slot0:", t[0, 248:254, 0])
print("block 30 idx 255:", t[30, 255], "idx 248..253 slot0/1:", t[30, 248:254, :2].tolist())
t0 = t[:, :244][t[:, :244] > 0].min()
e = t[:22, 255, :3].astype(np.float64)
print("solver ends (us): chain", ((e[:, 0] - t0) / 1e3).round(1).tolist())
print("publisher", ((e[:, 1] - t0) / 1e3).round(1).tolist())
print("tma", ((e[:, 2] - t0) / 1e3).round(1).tolist())

c = (t[:22, 128:192, :].astype(np.float64) - t0) / 1e3   # [solver][block][warp]
pb = (t[:22, 64:128, :4].astype(np.float64) - t0) / 1e3
for sidx in (0, 11):
    print(f"solver {sidx}: chain warps start of block 60 (us): V", c[sidx, 60, :4].round(1).tolist(), "L", c[sidx, 60, 4:].round(1).tolist(),
          "| publishers end of block 60:", pb[sidx, 60].round(1).tolist(), "block 63:", pb[sidx, 63].round(1).tolist())
    print("   V warp 1 block starts every 8th:", c[sidx, ::8, 1].round(0).tolist(), " L warp 1:", c[sidx, ::8, 5].round(0).tolist())
