"""Training-path timing: logProb forward + backward (two log-sum sweeps, path score, marginals writer) at the
MAESTRO-shape batch of BASELINE config 4 (T=691, N=360) and at the benchmark shape.  CUDA events, no profiler."""
import sys

import torch

sys.path.insert(0, "tests")
sys.path.insert(0, ".")
from golden_util import make_inputs  # noqa: E402
from transkun_b200.CRF import NeuralSemiCRFInterval, pack_intervals  # noqa: E402


def run(T, N, reps=5, packed=True):
    score, noise = make_inputs("randn", T, N, 11)
    s = torch.from_numpy(score).cuda().requires_grad_()
    z = torch.from_numpy(noise).cuda().requires_grad_()
    crf = NeuralSemiCRFInterval(s, z)
    with torch.no_grad():
        iv = crf.decode()  # some valid path per track
    iv = pack_intervals(iv, T) if packed else iv
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf, tb = [], []
    for i in range(reps + 2):
        s.grad = z.grad = None
        ev[0].record()
        lp = crf.logProb(iv)
        ev[1].record()
        lp.sum().backward()
        ev[2].record()
        torch.cuda.synchronize()
        if i >= 2:
            tf.append(ev[0].elapsed_time(ev[1]))
            tb.append(ev[1].elapsed_time(ev[2]))
    tri = 4.0 * N * T * (T + 1) / 2
    dense = 4.0 * N * T * T
    f, b = sorted(tf)[len(tf) // 2], sorted(tb)[len(tb) // 2]
    print(f"T={T} N={N} {'packed intervals' if packed else 'list-of-lists intervals'}: logProb forward {f * 1e3:.0f} us (2 sweeps = 2 triangle reads, {2 * tri / f / 1e6:.0f} GB/s), "
          f"backward {b * 1e3:.0f} us (1 triangle read + dense grad write = {(tri + dense) / 1e6:.0f} MB, "
          f"{(tri + dense) / b / 1e6:.0f} GB/s); algorithmic total {(3 * tri + dense) / 1e6:.0f} MB")


if __name__ == "__main__":
    run(691, 360, packed=False)
    run(691, 360)
    run(2048, 88)
