#!/usr/bin/env python
"""bench.py -- semi-CRF forward (log-partition) + Viterbi throughput, BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--T 2048] [--tracks 88]

A "step" is one pass of the hot path over one synthetic batch: the persistent sweep kernel
(Viterbi + log-partition from ONE read of the score triangle) followed by the device backtrack.
Workload (config.workload): T=2048, N=88 fp32 score tensor (1.48 GB, > the 126 MB L2, so every step
streams it from HBM; no explicit L2 flush needed) -- the configuration the metric is quoted on.
cells = T^2 * N per step.

value    : whole-job cells/s with inputs resident in HBM (CUDA events, max over ranks).
e2e      : same metric through the public API from HOST inputs (NeuralSemiCRFInterval.fromHost(...).decodeWithLogZ()):
           pinned H2D of the part of score the semi-CRF reads (end >= begin) and of noise, D2H of the decoded
           intervals + logZ, and the construction of the Python interval lists, all inside the timed region.
roofline : dominant kernel (sweep) algorithmic bytes 4*N*T(T+1)/2 per launch / its CUDA-event duration,
           against MEASURED_PEAKS.json's HBM copy bandwidth; traffic = DRAM bytes of the committed ncu capture.
cpu_baseline / --impl reference : the reference is pure Python/PyTorch and cannot travel to the GPU box,
           so the CPU arm is the C/OpenMP oracle port (oracle/, checked against the reference's golden
           outputs) on all host cores.
Multi-GPU: tracks shard with no data-path collective (weak scaling: every rank owns its own 88 tracks); the only
           exchange is the all-gather of the packed intervals: copy-engine pushes into symmetric NVLink peer memory on a
           side stream, overlapped with the next step's sweep (transkun_b200.sharded.PushGather; NCCL all-gather
           per step as the fallback, TKB_GATHER=nccl to force it).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "semi-CRF fwd+Viterbi cells/s (T^2*N)"
UNIT = "cells/s"
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md, used only if MEASURED_PEAKS.json is absent


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--T", type=int, default=2048)
    ap.add_argument("--tracks", type=int, default=88, help="tracks per GPU (weak scaling) ")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s; MEASURED_PEAKS.json absent)"


def ncu_traffic(T, N):
    """DRAM bytes per sweep launch from the committed ncu --set full capture of this workload, if any."""
    path = os.path.join(ROOT, "profiles", "sweep_traffic.json")
    try:
        rec = json.load(open(path))
        if rec.get("T") == T and rec.get("N") == N:
            return rec.get("dram_bytes_per_launch")
    except Exception:
        pass
    return None


# --------------------------------------------------------------------------------------------
# CPU arm: the oracle port (kind "port")
# --------------------------------------------------------------------------------------------
def cpu_pass(oracle_mod, score, noise):
    o = oracle_mod.SemiCRFOracle(score, noise)
    logz = o.computeLogZ()           # reference computeLogZ(noBackward=True)
    pairs, counts = o.decode_packed()  # reference decode()
    return logz, counts


def time_cpu(T, N, steps, warmup, budget_s=60.0):
    from golden_util import make_inputs
    from oracle import semicrf_oracle
    semicrf_oracle.build()
    score, noise = make_inputs("randn", T, N, 1234)
    semicrf_oracle.lib().tko_set_threads(os.cpu_count() or 1)  # torchrun exports OMP_NUM_THREADS=1
    cores = semicrf_oracle.lib().tko_max_threads()
    times = []
    t_begin = time.perf_counter()
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        cpu_pass(semicrf_oracle, score, noise)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if time.perf_counter() - t_begin > budget_s and len(times) >= 1:
            break
    sec = statistics.median(times)
    return {"value": T * T * N / sec, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{len(times)} timed full passes of T={T} N={N} (logZ forward + Viterbi backward + backtrack), "
                      f"C/OpenMP oracle port, median"}, sec, len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = min(args.steps, 8)
    warm = min(args.warmup, 2)
    base, sec, done = time_cpu(args.T, args.tracks, steps, warm, budget_s=150.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": done, "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"semi-CRF logZ+Viterbi decode, T={args.T}, N={args.tracks} fp32 randn seed 1234",
                   "note": "reference is pure PyTorch and cannot run on the GPU box; CPU arm = C/OpenMP oracle port "
                           "pinned to the reference's golden outputs"},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for name, v in zip(names, r[3:7]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s in sm if s > 0.5 * max(sm)]
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from golden_util import make_inputs
    from transkun_b200 import _lib
    from transkun_b200.CRF.NeuralSemiCRFInterval import NeuralSemiCRFInterval, backtrack_records, sweep
    from transkun_b200._lib import BACKWARD, SWEEP_LOGSUM, SWEEP_VITERBI
    from transkun_b200.sharded import PushGather, gather_records, track_shard

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.check(_lib.load().tkb_device_check(), "tkb_device_check")

    T = args.T
    if args.scaling == "weak":
        n_local, n_total = args.tracks, args.tracks * world
    else:
        lo, hi = track_shard(args.tracks, world, rank)
        n_local, n_total = hi - lo, args.tracks
    # synthetic inputs generated on CPU (same bits the oracle sees), one seed per rank
    score_h, noise_h = make_inputs("randn", T, n_local, 1234 + rank)
    score_pin = torch.from_numpy(score_h).pin_memory()
    noise_pin = torch.from_numpy(noise_h).pin_memory()
    score = score_pin.to(dev)
    noise = noise_pin.to(dev)
    stream = torch.cuda.current_stream(dev)

    ev_sweep = []
    # multi-GPU: the only exchange is the all-gather of the packed records.  Preferred: copy-engine pushes into
    # symmetric memory on a side stream, overlapped with the next step's sweep (transkun_b200.sharded.PushGather);
    # fallback: one NCCL all-gather per step on the compute stream.
    push = None
    if world > 1 and args.scaling == "weak" and os.environ.get("TKB_GATHER", "push") == "push":
        try:
            push = PushGather(n_local, 2 + 4 * T, dev)
        except Exception as exc:  # no symmetric memory / P2P: NCCL path
            if rank == 0:
                print(f"[bench] symmetric-memory gather unavailable ({type(exc).__name__}: {exc}); using NCCL", file=sys.stderr)
            push = None
    if world > 1:  # every rank must take the same path
        agree = torch.tensor([1 if push is not None else 0], device=dev)
        dist.all_reduce(agree, op=dist.ReduceOp.MIN)
        if int(agree.item()) == 0:
            push = None

    def step(record=False):
        if record:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
        code, _, lse, _ = sweep(score, noise, BACKWARD, SWEEP_VITERBI | SWEEP_LOGSUM)
        if record:
            e1.record(stream)
            ev_sweep.append((e0, e1))
        rec = backtrack_records(code, None, BACKWARD, lse[0])  # [count, logZ, pairs] per track
        if push is not None:
            return push.result(push.submit(rec))  # complete once the stream has passed push.wait()
        if world > 1:  # one NCCL all-gather of the packed records over NVLink
            rec = gather_records(rec, n_total)
        return rec

    def fence():
        if push is not None:
            push.wait()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    if push is not None:  # once, untimed: the pushed gather equals the NCCL gather
        code0, _, lse0, _ = sweep(score, noise, BACKWARD, SWEEP_VITERBI | SWEEP_LOGSUM)
        rec0 = backtrack_records(code0, None, BACKWARD, lse0[0])
        want = gather_records(rec0, n_total)
        got = push.result(push.submit(rec0))
        push.wait()
        torch.cuda.synchronize(dev)
        assert torch.equal(got, want), "pushed gather differs from the NCCL gather"

    for _ in range(max(args.warmup, 3)):
        step()
    fence()
    sampler = ClockSampler(local) if rank == 0 else None
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(stream)
    for _ in range(args.steps):
        step(record=True)
    if push is not None:
        push.wait()  # the timed region ends when the last exchange has landed everywhere
    t1.record(stream)
    fence()
    ms = t0.elapsed_time(t1)
    # keep the GPU under the same load a little longer so that nvidia-smi (20 ms period) sees it
    t_end = time.time() + 0.6
    while time.time() < t_end:
        step()
        if push is not None:
            push.wait()
        torch.cuda.synchronize(dev)
    clocks = sampler.stop() if sampler else None
    if world > 1:
        tmax = torch.tensor([ms], device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
    sweep_ms = statistics.mean(a.elapsed_time(b) for a, b in ev_sweep)
    cells_per_step = float(T) * T * n_total
    value = cells_per_step * args.steps / (ms * 1e-3)

    # ---- e2e: public API, host inputs, H2D + D2H inside the timed region -----------------------
    fence()
    e2e_steps = max(1, args.e2e_steps)
    d2h_bytes = 0
    with torch.no_grad():  # one untimed pass: first-use allocations of the e2e path
        NeuralSemiCRFInterval.fromHost(score_pin, noise_pin, dev).decodeWithLogZ()
    torch.cuda.synchronize(dev)
    tw0 = time.perf_counter()
    for _ in range(e2e_steps):
        with torch.no_grad():
            # public API from HOST tensors: uploads the part of the score tensor the semi-CRF reads (end >= begin)
            crf = NeuralSemiCRFInterval.fromHost(score_pin, noise_pin, dev)
            dec, logz = crf.decodeWithLogZ()
            logz_h = logz.cpu()
        maxc = max((len(d) for d in dec), default=0)
        d2h_bytes = n_local * 4 + n_local * maxc * 8 + logz_h.numel() * 4
    torch.cuda.synchronize(dev)
    e2e_sec = (time.perf_counter() - tw0) / e2e_steps
    if world > 1:
        tmax = torch.tensor([e2e_sec], device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e_sec = float(tmax.item())
    h2d_bytes = NeuralSemiCRFInterval.lowerTriangleUploadBytes(T, n_local) + noise_pin.numel() * 4

    if rank == 0:
        peak, peak_src = peak_hbm()
        alg_bytes = 4.0 * n_local * T * (T + 1) / 2.0
        achieved = alg_bytes / (sweep_ms * 1e-3) / 1e9
        traffic = ncu_traffic(T, n_local)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"semi-CRF logZ+Viterbi decode (one fused sweep + device backtrack), T={T}, "
                                   f"N={n_local} tracks/GPU fp32 randn seed 1234",
                       "T": T, "tracks_per_gpu": n_local, "tracks_total": n_total, "parallelism": f"track-sharded x{world}",
                       "exchange": ("none" if world == 1 else ("copy-engine pushes into symmetric memory, overlapped with the "
                                    "next sweep" if push is not None else "NCCL all-gather per step")),
                       "l2": "score tensor 1.48 GB per GPU >> 126 MB L2: inputs larger than L2, no flush needed",
                       "timing": "CUDA events on the launching stream, barrier+synchronize both sides, max over ranks"},
            "e2e": {"value": cells_per_step / e2e_sec, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": d2h_bytes, "steps": e2e_steps,
                    "api": "NeuralSemiCRFInterval.fromHost(score, noise, device).decodeWithLogZ() from pinned host tensors "
                           "(uploads the lower-triangle staircase of score, the part the semi-CRF reads)"},
            "gpu_launches": 2 * args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "tkb::sweep_kernel<BACKWARD, A16, VITERBI|LOGSUM>",
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": sweep_ms, "peak_source": peak_src},
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            base, _, _ = time_cpu(T, n_local, 3, 1, budget_s=45.0)
            line["cpu_baseline"] = base
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
