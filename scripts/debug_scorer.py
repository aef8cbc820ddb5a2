"""Scorer timing (CUDA events): one-pass TF32 and 3xTF32, at the benchmark shape and the model shape; TKB_SCORER_CX (cluster width 1/2/4:
CTAs sharing one multicast q tile) and TKB_SCORER_BAND (tile rows per band of the block order) select the variants.  usage: python scripts/debug_scorer.py"""
import sys
import torch
sys.path.insert(0, ".")
from transkun_b200.LayersTransformer import sip_score


def run(NT, T, D, precise):
    g = torch.Generator().manual_seed(0)
    q, k, d = torch.randn(NT, T, D, generator=g).cuda(), torch.randn(NT, T, D, generator=g).cuda(), torch.randn(NT, T, generator=g).cuda()
    P = (NT + 7) // 8 * 8
    S = torch.empty((T, T, P), device="cuda")[:, :, :NT]
    for _ in range(3):
        sip_score(q, k, d, out=S, precise=precise)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        sip_score(q, k, d, out=S, precise=precise)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    out_bytes = 4.0 * NT * T * (T + 1) / 2
    flops = 2.0 * D * (3 if precise else 1) * NT * T * (T + 1) / 2
    print(f"NT={NT} T={T} D={D} {'3xTF32' if precise else 'TF32  '}: {us:7.0f} us  {out_bytes / us / 1e3:6.0f} GB/s of output  {flops / us / 1e6:6.1f} TFLOP/s"
          f"  (incl. operand split/concat on the host side for 3xTF32)", flush=True)


for NT, T in ((88, 2048), (90, 691), (360, 691)):
    for precise in (False, True):
        run(NT, T, 256, precise)
