"""Shared input generators and golden-fixture loader for tests/, bench.py and smoke().

Inputs follow SURVEY.md section 8(d):
  randn0 : crfMinimalExample.py style -- torch.manual_seed(seed); randn(T,T,N); randn(T-1,N)
  randn  : (A) g=Generator().manual_seed(seed); randn(T,T,N,generator=g); randn(T-1,N,generator=g)
  ties   : small integers in [-2,2] -- sums are exact in fp32, so argmax ties are frequent and
           the reference's first-index rule (skip first, then nearest end) is exercised
  model  : (B) q,k~N(0,1)[N,T,64]; S=(q.k^T/8)*|e-b| + diag_embed(N(-2,2)); noise=0 -- heavy
           negative tail like the shipped checkpoint's scores
  zeros  : all-zero inputs (every candidate ties; decode must be empty)
Generated on CPU so the oracle and the CUDA path see identical bits.
"""
from __future__ import annotations

import glob
import hashlib
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def make_inputs(kind: str, T: int, N: int, seed: int):
    if kind == "randn0":
        torch.manual_seed(seed)
        score = torch.randn(T, T, N)
        noise = torch.randn(T - 1, N)
    elif kind == "randn":
        g = torch.Generator().manual_seed(seed)
        score = torch.randn(T, T, N, generator=g)
        noise = torch.randn(T - 1, N, generator=g)
    elif kind == "ties":
        g = torch.Generator().manual_seed(seed)
        score = torch.randint(-2, 3, (T, T, N), generator=g).float()
        noise = torch.randint(-2, 3, (T - 1, N), generator=g).float()
    elif kind == "model":
        g = torch.Generator().manual_seed(seed)
        q = torch.randn(N, T, 64, generator=g)
        k = torch.randn(N, T, 64, generator=g)
        s = torch.einsum("ned,nbd->neb", q, k) / 8.0
        t = torch.arange(T, dtype=torch.float32)
        s = s * (t[:, None] - t[None, :]).abs()[None]
        diag = torch.randn(N, T, generator=g) * 2.0 - 2.0
        s = s + torch.diag_embed(diag)
        score = s.permute(1, 2, 0).contiguous()
        noise = torch.zeros(T - 1, N)
    elif kind == "zeros":
        score = torch.zeros(T, T, N)
        noise = torch.zeros(T - 1, N)
    else:
        raise ValueError(kind)
    return score.numpy().astype(np.float32, copy=False), noise.numpy().astype(np.float32, copy=False)


def random_intervals(T: int, N: int, seed: int):
    """Random valid paths of the grammar, one list per track."""
    rs = np.random.RandomState(seed)
    out = []
    for _ in range(N):
        cur, p = [], 0
        while p < T:
            if rs.rand() < 0.15:
                cur.append((p, p))
            if p == T - 1:
                break
            if rs.rand() < 0.7:
                p += 1
            else:
                e = min(T - 1, p + 1 + int(rs.geometric(0.2)) - 1)
                cur.append((p, e))
                p = e
        out.append(cur)
    return out


def input_sha(score, noise) -> str:
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(score).tobytes())
    h.update(np.ascontiguousarray(noise).tobytes())
    return h.hexdigest()


def unpack(pairs, off):
    return [[(int(b), int(e)) for b, e in pairs[off[n]:off[n + 1]]] for n in range(len(off) - 1)]


class Golden:
    def __init__(self, path):
        self.name = os.path.splitext(os.path.basename(path))[0]
        self.z = np.load(path, allow_pickle=False)
        self.T, self.N = int(self.z["T"]), int(self.z["N"])
        self.kind, self.seed = str(self.z["kind"]), int(self.z["seed"])

    def inputs(self):
        """(score, noise) or None when regenerated inputs do not hash to the recorded SHA."""
        if "score" in self.z:
            return self.z["score"], self.z["noise"]
        score, noise = make_inputs(self.kind, self.T, self.N, self.seed)
        if input_sha(score, noise) != str(self.z["sha"]):
            return None
        return score, noise

    def lists(self, key):
        return unpack(self.z[key + "_pairs"], self.z[key + "_off"])

    def has(self, key):
        return key in self.z

    def __getitem__(self, key):
        return self.z[key]


def golden_cases(prefix: str = ""):
    """Semi-CRF fixtures only (the scorer_* / frontend_* files belong to their own tests)."""
    return sorted(p for p in glob.glob(os.path.join(GOLDEN_DIR, prefix + "*.npz"))
                  if not os.path.basename(p).startswith(("scorer_", "frontend_", "config3_", "gradbig_")))
