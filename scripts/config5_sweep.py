"""BASELINE.json config 5: T sweep 256..4096 at N=88 tracks, strong scaling over the GPUs of one box (88 tracks split over
the ranks, no data-path collective; the decoded records are exchanged by the fused back-track + NVLink push).  One torchrun
launch per GPU count; rank 0 prints ONE JSON object with a row per T: cells/s of the whole job (CUDA events, max over ranks),
the sweep kernel's share, achieved GB/s per GPU and its fraction of the measured HBM peak.  Inputs are generated on the
device (timing only; parity lives in tests/).
usage: [torchrun ...] python scripts/config5_sweep.py [--steps 20] [--cpu]   (--cpu: rank 0 also times the reference on CPU)"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from transkun_b200.CRF.NeuralSemiCRFInterval import backtrack_records, sweep  # noqa: E402
from transkun_b200._lib import BACKWARD, SWEEP_LOGSUM, SWEEP_VITERBI  # noqa: E402
from transkun_b200.sharded import FusedPushGather, track_shard  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--tracks", type=int, default=88)
ap.add_argument("--cpu", action="store_true")
args = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    peak = 6650.0
rows = []
for T in (256, 512, 1024, 2048, 4096):
    lo, hi = track_shard(args.tracks, world, rank)
    n_local = hi - lo
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    score = torch.randn((T, T, n_local), device=dev, generator=g)
    noise = torch.randn((T - 1, n_local), device=dev, generator=g)
    fused = FusedPushGather(n_local, T, dev) if (world > 1 and args.tracks % world == 0) else None
    stream = torch.cuda.current_stream(dev)
    ev = []

    def step(record=False):
        if record:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
        code, _, lse, _ = sweep(score, noise, BACKWARD, SWEEP_VITERBI | SWEEP_LOGSUM)
        if record:
            b.record(stream)
            ev.append((a, b))
        if fused is not None:
            st = fused.submit(code, None, BACKWARD, lse[0])
            if st > 1:
                fused.result(st - 1)
        else:
            backtrack_records(code, None, BACKWARD, lse[0])

    for _ in range(5):
        step()
    if fused is not None:
        fused.result(fused.step)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(stream)
    for _ in range(args.steps):
        step(True)
    if fused is not None:
        fused.result(fused.step)
    t1.record(stream)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    ms = torch.tensor([t0.elapsed_time(t1) / args.steps, sum(a.elapsed_time(b) for a, b in ev) / len(ev)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    step_ms, sweep_ms = float(ms[0]), float(ms[1])
    gbs = 4.0 * n_local * T * (T + 1) / 2.0 / (sweep_ms * 1e-3) / 1e9
    rows.append({"T": T, "tracks_per_gpu": n_local, "ms_per_step": step_ms, "sweep_ms": sweep_ms,
                 "cells_per_s": float(T) * T * args.tracks / (step_ms * 1e-3), "sweep_GBps_per_gpu": gbs,
                 "frac_of_hbm_peak": gbs / peak})
    del score, noise, fused
    torch.cuda.empty_cache()
if rank == 0:
    out = {"config": "T sweep at N=%d tracks, strong scaling" % args.tracks, "n_gpus": world, "steps": args.steps,
           "hbm_peak_GBps": peak, "rows": rows}
    if args.cpu:
        sys.argv = [sys.argv[0]]
        import bench
        cpu = []
        for T in (256, 512, 1024, 2048, 4096):
            base, sec = bench.time_reference(T, args.tracks, 2 if T >= 2048 else 3, 2, budget_s=60.0)
            cpu.append({"T": T, "cells_per_s": base["value"], "cores": base["cores"], "sample": base["sample"][:60]})
        out["reference_cpu"] = cpu
    print(json.dumps(out), flush=True)
if world > 1:
    dist.destroy_process_group()
