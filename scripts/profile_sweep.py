"""Tiny driver for ncu: a few fused sweeps (+ backtrack) at the benchmark shape."""
import sys

import torch

sys.path.insert(0, "tests")
sys.path.insert(0, ".")
from golden_util import make_inputs  # noqa: E402
from transkun_b200.CRF.NeuralSemiCRFInterval import backtrack, sweep  # noqa: E402
from transkun_b200._lib import BACKWARD, SWEEP_LOGSUM, SWEEP_VITERBI  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
N = int(sys.argv[2]) if len(sys.argv) > 2 else 88
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
score, noise = make_inputs("randn", T, N, 1234)
s, z = torch.from_numpy(score).cuda(), torch.from_numpy(noise).cuda()
for _ in range(reps):
    code, _, lse, _ = sweep(s, z, BACKWARD, SWEEP_VITERBI | SWEEP_LOGSUM)
    pairs, counts = backtrack(code, None, BACKWARD)
torch.cuda.synchronize()
print("done", float(lse[0, 0]), int(counts.sum()))
