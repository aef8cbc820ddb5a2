"""Fused / single-semiring sweep timing over shapes (CUDA events).  usage: python scripts/sweep_shapes.py [T,N ...]
(negative N: the same N with the track axis padded to a multiple of 4)."""
import sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import torch
from golden_util import make_inputs
from transkun_b200.CRF.NeuralSemiCRFInterval import sweep
from transkun_b200._lib import BACKWARD, FORWARD, SWEEP_LOGSUM, SWEEP_VITERBI
SHAPES = [(691, 88), (691, 90), (691, -90), (691, 360)]  # negative N: padded track axis
if len(sys.argv) > 1:
    SHAPES = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
for T, N in SHAPES:
    padded, N = N < 0, abs(N)
    score, noise = make_inputs("randn", T, N, 3)
    s, z = torch.from_numpy(score).cuda(), torch.from_numpy(noise).cuda()
    if padded:
        buf = torch.zeros((T, T, (N + 3) // 4 * 4), device="cuda")
        buf[:, :, :N] = s
        s = buf[:, :, :N]
    for name, d, fl in (("bwd V+L", BACKWARD, 3), ("fwd L", FORWARD, 2), ("bwd V", BACKWARD, 1)):
        for _ in range(3): sweep(s, z, d, fl)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): sweep(s, z, d, fl)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100
        print(f"T={T} N={N}{' padded' if padded else ''} {name}: {us:.0f} us, {4.0*N*T*(T+1)/2/us/1e3:.0f} GB/s", flush=True)
    del s, z
