"""BASELINE.json config 4: one training step of the hot path on a synthetic MAESTRO-shape batch, pitch-sharded over the
GPUs of one box (reference: ModelTransformer.py:228-332 log_prob -> :263-264 crf.evalPath / crf.computeLogZ, train.py:186-189
loss = -logp.sum(-1).mean(); (loss/50).backward()).

    ctx [B=4, P=90, T=691, D=256] (what the backbone hands to the scorer; random here, replicated on every rank)
    rank r owns the symbols p in its slice of the 90: scorer(ctx[:, p_r]) -> S [T,T,B,P_r] -> CRF.logProb(intervals_r)
    -> loss -> backward: marginals (dense dS), scorer adjoint, d ctx[:, p_r]; the slices of d ctx are all-gathered over
    NVLink (NCCL) so that every rank holds the full d ctx for the (replicated) backbone, and the scorer's weight gradients
    are all-reduced.
The reference trains with --allow_tf32 (train.py:41-43): the scorer runs its one-pass TF32 mode here.
Rank 0 prints one JSON line: ms per step (CUDA events, median over the steps, max over ranks), cells/s (T^2 * B * P per step), and the share of
the semi-CRF part against its HBM roofline (4N[3T(T+1)/2 + T^2] bytes per step).
usage: [torchrun ...] python scripts/config4_train.py [--steps 10]"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from golden_util import random_intervals  # noqa: E402
from transkun_b200.CRF import NeuralSemiCRFInterval, pack_intervals  # noqa: E402
from transkun_b200.LayersTransformer import ScaledInnerProductIntervalScorer  # noqa: E402
from transkun_b200.sharded import track_shard  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=30)
ap.add_argument("--B", type=int, default=4)
ap.add_argument("--T", type=int, default=691)
args = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
torch.backends.cuda.matmul.allow_tf32 = True
B, P, T, D = args.B, 90, args.T, 256
torch.manual_seed(0)  # same ctx and weights on every rank
ctx_full = torch.randn(B, P, T, D, device=dev) * 0.5
scorer = ScaledInnerProductIntervalScorer(D, 1).to(dev)
lo, hi = track_shard(P, world, rank)
Pr, Pmax = hi - lo, -(-P // world)
# ground-truth intervals of my symbols, track index = b * Pr + p (the flatten(-2,-1) order of ModelTransformer.py:215)
iv_all = random_intervals(T, B * P, 17)
iv = [iv_all[b * P + p] for b in range(B) for p in range(lo, hi)]
packed = pack_intervals(iv, T)
stream = torch.cuda.current_stream(dev)
gather_buf = torch.empty((world, B, Pmax, T, D), device=dev) if world > 1 else None


def step(ev=None):
    ctx = ctx_full[:, lo:hi].detach().requires_grad_()
    scorer.zero_grad(set_to_none=True)
    if ev:
        ev[0].record(stream)
    S, Sskip = scorer(ctx)                                   # [T,T,B,Pr], [T-1,B,Pr]
    if ev:
        ev[1].record(stream)
    crf = NeuralSemiCRFInterval(S.flatten(-2, -1), Sskip.flatten(-2, -1))
    logp = crf.logProb(packed)                               # [B*Pr]
    loss = -logp.sum() / B
    if ev:
        ev[2].record(stream)
    (loss / 50).backward()
    if ev:
        ev[3].record(stream)
    if world > 1:
        g = ctx.grad
        if Pr < Pmax:
            g = torch.nn.functional.pad(g, (0, 0, 0, 0, 0, Pmax - Pr))
        dist.all_gather_into_tensor(gather_buf, g.contiguous())   # d ctx of all symbols on every rank (NVLink)
        for prm in scorer.parameters():
            dist.all_reduce(prm.grad)
    if ev:
        ev[4].record(stream)
    return loss


for _ in range(3):
    step()
if world > 1:
    dist.barrier()
torch.cuda.synchronize(dev)
parts = []
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record(stream)
for _ in range(args.steps):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    loss = float(step(ev))
    parts.append(ev)
t1.record(stream)
if world > 1:
    dist.barrier()
torch.cuda.synchronize(dev)
import statistics
# median over the steps: the step is short enough (a few ms) for host scheduling noise to show up in single steps
ms = statistics.median(e[0].elapsed_time(e[4]) for e in parts)
seg = [statistics.median(e[i].elapsed_time(e[i + 1]) for e in parts) for i in range(4)]
tt = torch.tensor([ms] + seg, device=dev)
if world > 1:
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
ms, seg = float(tt[0]), [float(v) for v in tt[1:]]
if rank == 0:
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    n_loc = B * Pmax
    alg = 4.0 * n_loc * (3 * T * (T + 1) / 2 + T * T)          # semi-CRF bytes per step on the busiest rank
    crf_ms = seg[1] + seg[2]   # logProb forward + the whole backward (marginals + scorer adjoint)
    print(json.dumps({
        "config": f"training step: ctx[{B},{P},{T},{D}] -> scorer -> CRF.logProb -> backward, symbols sharded over {world} GPU(s)",
        "n_gpus": world, "steps": args.steps, "ms_per_step": ms, "cells_per_s": float(T) * T * B * P / (ms * 1e-3),
        "parts_ms": {"scorer_forward": seg[0], "logProb_forward": seg[1], "backward": seg[2], "exchange": seg[3]},
        "loss": loss,
        "semicrf_roofline": {"algorithmic_bytes_busiest_rank": alg, "forward_plus_backward_ms": crf_ms,
                             "GBps": alg / (crf_ms * 1e-3) / 1e9, "frac_of_hbm_peak": alg / (crf_ms * 1e-3) / 1e9 / peak,
                             "note": "backward also contains the scorer adjoint (torch.bmm), so this understates the CRF kernels"},
    }), flush=True)
if world > 1:
    dist.destroy_process_group()
