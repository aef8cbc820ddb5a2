"""GPU diagnostics: runs the CUDA path on a ladder of shapes and prints where it differs from the
CPU oracle instead of asserting (one gpurun call -> as much information as possible)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, "tests")
sys.path.insert(0, ".")
from golden_util import make_inputs  # noqa: E402
from oracle.semicrf_oracle import SemiCRFOracle  # noqa: E402
from transkun_b200.CRF.NeuralSemiCRFInterval import NeuralSemiCRFInterval, sweep  # noqa: E402
from transkun_b200._lib import BACKWARD, FORWARD, SWEEP_LOGSUM, SWEEP_VITERBI  # noqa: E402


def run(T, N, kind, seed=7):
    score, noise = make_inputs(kind, T, N, seed)
    o = SemiCRFOracle(score, noise)
    s, z = torch.from_numpy(score).cuda(), torch.from_numpy(noise).cuda()
    ok = True
    for direction, forward in ((BACKWARD, False), (FORWARD, True)):
        t0 = time.time()
        code, vit, lse, ws = sweep(s, z, direction, SWEEP_VITERBI | SWEEP_LOGSUM, want_vit=True)
        torch.cuda.synchronize()
        dt = time.time() - t0
        status = int(ws.buf[:4].view(torch.int32).item())
        q, sel = o.viterbi_dp(forward)
        ref = o.alpha() if forward else o.beta()
        vit_h, lse_h = vit.cpu().numpy(), lse.cpu().numpy()
        code_h = code.cpu().numpy().view(np.uint32).T
        bad_q = np.argwhere(vit_h.view(np.uint32) != q.view(np.uint32))
        bad_s = np.argwhere((code_h >> 1).astype(np.int64) - 1 != sel)
        rel = np.abs(lse_h - ref) / np.maximum(1.0, np.abs(ref))
        line = (f"T={T} N={N} {kind} dir={direction} status={status} t={dt*1e3:.1f}ms "
                f"q_bad={len(bad_q)} sel_bad={len(bad_s)} lse_maxrel={rel.max():.2e}")
        if len(bad_q):
            t, n = bad_q[-1] if direction == BACKWARD else bad_q[0]
            line += f" first_bad_q(pos={t},n={n}) got={vit_h[t, n]} want={q[t, n]}"
        if not np.isfinite(lse_h).all():
            line += " lse_nonfinite"
        good = status == 0 and len(bad_q) == 0 and len(bad_s) == 0 and rel.max() < 1e-4
        ok &= good
        print(("OK   " if good else "FAIL ") + line, flush=True)
    crf = NeuralSemiCRFInterval(s, z)
    for fwd in (False, True):
        d, dref = crf.decode(forward=fwd), o.decode(forward=fwd)
        if d != dref:
            ok = False
            nb = [n for n in range(N) if d[n] != dref[n]]
            print(f"FAIL decode forward={fwd}: {len(nb)} tracks differ, e.g. n={nb[0]} got={d[nb[0]][:6]} want={dref[nb[0]][:6]}")
    return ok


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), flush=True)
    allok = True
    shapes = [(2, 1, "randn"), (5, 3, "randn"), (31, 8, "randn"), (32, 8, "randn"), (33, 8, "randn"),
                       (64, 8, "ties"), (65, 9, "ties"), (100, 4, "model"), (200, 88, "randn"), (256, 90, "ties"),
                       (512, 88, "randn"), (691, 90, "model"), (300, 180, "randn"), (1024, 88, "randn"), (2048, 88, "randn")]
    if len(sys.argv) > 1:
        shapes = [(int(a.split(",")[0]), int(a.split(",")[1]), a.split(",")[2]) for a in sys.argv[1:]]
    for T, N, kind in shapes:
        allok &= run(T, N, kind)
    print("ALL OK" if allok else "SOME FAILED")
