import sys, time
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import torch
from golden_util import make_inputs
from transkun_b200.CRF import NeuralSemiCRFInterval
T, N = 2048, 88
score, noise = make_inputs("randn", T, N, 1234)
sp, zp = torch.from_numpy(score).pin_memory(), torch.from_numpy(noise).pin_memory()
dev = torch.device("cuda")
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
print("dense .to(dev)      : %.1f ms" % t(lambda: sp.to(dev, non_blocking=True)))
for r in (16, 64, 256):
    print("fromHost rows=%3d   : %.1f ms" % (r, t(lambda: NeuralSemiCRFInterval.fromHost(sp, zp, dev, rows_per_chunk=r))))
print("zeros 1.48 GB       : %.2f ms" % t(lambda: torch.zeros((T, T, N), device=dev)))
crf = NeuralSemiCRFInterval(sp.to(dev), zp.to(dev))
with torch.no_grad():
    print("decodeWithLogZ      : %.1f ms" % t(lambda: crf.decodeWithLogZ()))
    print("decode_packed       : %.2f ms" % t(lambda: crf.decode_packed(None, False, with_logz=True)))
