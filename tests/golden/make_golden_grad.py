"""Gradient (marginal) fixtures at model-size shapes, from the UNMODIFIED reference on CPU (forward_backward,
CRF/NeuralSemiCRFInterval.py:374-456 = what ComputeLogZFasterGrad saves and returns, :459-475).

The dense [T,T,N] gradient of T=691, N=90 is 172 MB, so a fixture stores: logZ, the full gradNoise [T-1,N], the row
masses sum_b grad[e,b,n] ([T,N]: every cell contributes), 30000 sampled lower-triangle cells and 2000 cells above the
diagonal (which must be zero).  Inputs are regenerated from the seed (SHA-256 recorded).  N % 4 covers 2, 3 and 1
(the marginals kernel has a vector path per remainder).  Run in the build container:
    python tests/golden/make_golden_grad.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
from baseline import ref_loader  # noqa: E402
from golden_util import input_sha, make_inputs  # noqa: E402

CASES = [("model", 691, 90, 1234), ("model", 300, 90, 1234), ("randn", 300, 90, 5), ("randn", 200, 7, 6), ("randn", 129, 5, 7)]


def main():
    ref_loader.import_reference()
    import importlib
    fb = importlib.import_module("transkun.CRF.NeuralSemiCRFInterval").forward_backward
    torch.set_num_threads(os.cpu_count() or 1)
    for kind, T, N, seed in CASES:
        score, noise = make_inputs(kind, T, N, seed)
        with torch.no_grad():
            logz, grad, gnoise = fb(torch.from_numpy(score), torch.from_numpy(noise))
        grad = grad.numpy()
        # the same function in float64: the yardstick for "how far is an fp32 implementation from the exact marginals"
        with torch.no_grad():
            logz64, grad64, gnoise64 = fb(torch.from_numpy(score).double(), torch.from_numpy(noise).double())
        grad64 = grad64.numpy()
        rs = np.random.RandomState(seed + T)
        e = rs.randint(0, T, size=30000)
        b = (rs.rand(30000) * (e + 1)).astype(np.int64)      # b <= e
        n = rs.randint(0, N, size=30000)
        ue = rs.randint(0, T - 1, size=2000)
        ub = ue + 1 + (rs.rand(2000) * (T - 1 - ue)).astype(np.int64)   # b > e
        un = rs.randint(0, N, size=2000)
        out = dict(kind=kind, T=T, N=N, seed=seed, sha=input_sha(score, noise), logz=logz.numpy(), grad_noise=gnoise.numpy(),
                   row_mass=grad.sum(axis=1), idx=np.stack([e, b, n], 1).astype(np.int32), val=grad[e, b, n],
                   uidx=np.stack([ue, ub, un], 1).astype(np.int32), uval=grad[ue, ub, un],
                   logz64=logz64.numpy(), grad_noise64=gnoise64.numpy(), row_mass64=grad64.sum(axis=1), val64=grad64[e, b, n])
        name = f"gradbig_{kind}_T{T}_N{N}.npz"
        np.savez_compressed(os.path.join(HERE, name), **out)
        print(name, "logz[0] =", float(logz[0]), "max grad", float(grad.max()), "upper nonzero", int((out["uval"] != 0).sum()),
              "| reference fp32 vs float64: cells max abs", float(np.abs(out["val"] - out["val64"]).max()),
              "gradNoise", float(np.abs(out["grad_noise"] - out["grad_noise64"]).max()))


if __name__ == "__main__":
    main()
