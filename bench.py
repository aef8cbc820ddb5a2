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
cpu_baseline / --impl reference : the reference's OWN PyTorch-CPU path -- the unmodified package installed under
           baseline/_ref (baseline/ref_loader.py): NeuralSemiCRFInterval(score, noise).computeLogZ(noBackward=True) +
           .decode() on all host cores (kind "reference").  Only if baseline/_ref is absent does it fall back to the
           C/OpenMP oracle port (kind "port", ~16x faster than the reference on the same CPU).
scorer   : (1 GPU) the interval scorer in front of the same step, q, k [N, T, 256] -> score -> logZ + decode: time of
           tkb_sip_score in both precision modes, of the chain, and GB/s of its algorithmic bytes (output written
           once + operands read once) against the same HBM peak.  --workload scorer+crf makes that chain the timed step
           (roofline block = the scorer kernel; e2e uploads q / k / diag instead of the score tensor).
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
    ap.add_argument("--tsweep", action="store_true", help="add a T sweep 256..4096 (config 5) to the JSON line")
    ap.add_argument("--no-extras", action="store_true", help="skip the other-shapes and scorer blocks (clean ncu launch lists)")
    ap.add_argument("--workload", default="crf", choices=["crf", "scorer+crf"],
                    help="crf: BASELINE.json's metric, score[T,T,N] resident (default).  scorer+crf: the same step fed from "
                         "the scorer's operands q, k [N,T,256] (tkb_sip_score -> sweep -> back-track), roofline block = "
                         "the scorer kernel, e2e uploads q/k/diag instead of the score tensor")
    return ap.parse_args()


def workload_config(T, n_local, n_total, world):
    """The `config` object: identical in both arms (the driver compares them)."""
    return {"workload": f"semi-CRF logZ + Viterbi decode on score[T,T,N] fp32 randn (seed 1234 + rank), T={T}, "
                        f"N={n_local} tracks per GPU",
            "T": T, "tracks_per_gpu": n_local, "tracks_total": n_total, "parallelism": f"track-sharded x{world}",
            "l2": "score tensor 1.48 GB per GPU >> 126 MB L2: inputs larger than L2, no flush needed",
            "timing": "ours: CUDA events on the launching stream, barrier+synchronize both sides, max over ranks; "
                      "reference arm: perf_counter around K full passes on the host"}


def peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s; MEASURED_PEAKS.json absent)"


def ncu_traffic(T, N, name="sweep_traffic.json"):
    """DRAM bytes per launch of the kernel from the committed ncu --set full capture of this workload, if any."""
    path = os.path.join(ROOT, "profiles", name)
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
def reference_available():
    from baseline import ref_loader
    return ref_loader.available()


def time_reference(T, N, steps, warmup, budget_s=240.0):
    """The unmodified reference on the host cores: computeLogZ(noBackward=True) + decode() per step.  If K+W full
    passes would not fit the budget, a step is the same pass over the first n tracks (tracks are independent, so
    cells/s is the per-track rate; the sample is stated)."""
    import torch
    from golden_util import make_inputs
    from baseline import ref_loader
    RefCRF = ref_loader.reference_crf()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)  # torchrun exports OMP_NUM_THREADS=1
    score, noise = make_inputs("randn", T, N, 1234)
    score_t, noise_t = torch.from_numpy(score), torch.from_numpy(noise)

    def one(s, z):
        with torch.no_grad():
            crf = RefCRF(s, z)
            logz = crf.computeLogZ(noBackward=True)
            dec = crf.decode()
        return logz, dec

    t0 = time.perf_counter()
    one(score_t, noise_t)          # TorchScript profiling run 1 (also the size probe)
    t_probe = time.perf_counter() - t0
    n_used = N
    if t_probe * (steps + warmup) > budget_s:
        n_used = max(4, int(N * budget_s / (t_probe * (steps + warmup))) // 4 * 4)
        score_t, noise_t = score_t[:, :, :n_used].contiguous(), noise_t[:, :n_used].contiguous()
    for _ in range(max(warmup, 2) - (1 if n_used == N else 0)):
        one(score_t, noise_t)
    t0 = time.perf_counter()
    for _ in range(steps):
        one(score_t, noise_t)
    sec = (time.perf_counter() - t0) / steps
    sample = (f"{steps} timed full passes of T={T} N={N}" if n_used == N else
              f"{steps} timed passes of T={T} over the first {n_used} of {N} tracks (full passes would exceed {budget_s:.0f} s)")
    return {"value": float(T) * T * n_used / sec, "unit": UNIT, "cores": cores, "kind": "reference",
            "sample": sample + ": unmodified reference (baseline/_ref) NeuralSemiCRFInterval.computeLogZ(noBackward=True) "
                               f"+ .decode(), torch {torch.__version__} CPU, {cores} threads"}, sec


def cpu_pass(oracle_mod, score, noise):
    o = oracle_mod.SemiCRFOracle(score, noise)
    logz = o.computeLogZ()           # reference computeLogZ(noBackward=True)
    pairs, counts = o.decode_packed()  # reference decode()
    return logz, counts


def time_port(T, N, steps, warmup, budget_s=60.0):
    from golden_util import make_inputs
    from oracle import semicrf_oracle
    semicrf_oracle.build()
    score, noise = make_inputs("randn", T, N, 1234)
    semicrf_oracle.lib().tko_set_threads(os.cpu_count() or 1)  # torchrun exports OMP_NUM_THREADS=1
    cores = semicrf_oracle.lib().tko_max_threads()
    for _ in range(warmup):
        cpu_pass(semicrf_oracle, score, noise)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_pass(semicrf_oracle, score, noise)
    sec = (time.perf_counter() - t0) / steps
    return {"value": T * T * N / sec, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{steps} timed full passes of T={T} N={N} (logZ forward + Viterbi backward + backtrack), "
                      f"C/OpenMP oracle port"}, sec


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    if reference_available():
        base, sec = time_reference(args.T, args.tracks, args.steps, args.warmup)
    else:
        base, sec = time_port(args.T, args.tracks, args.steps, args.warmup)
    n_total = args.tracks * world if args.scaling == "weak" else args.tracks
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.T, args.tracks, n_total, world),
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "notes": "CPU arm: runs on rank 0's host cores only, whatever --gpus says",
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
    from transkun_b200.sharded import FusedPushGather, PushGather, bind_to_gpu_numa_node, gather_records, track_shard

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # one process per GPU: keep this rank's pinned host buffers and its Python work on the GPU's own NUMA node
    numa_node = bind_to_gpu_numa_node(local) if world > 1 and os.environ.get("TKB_NUMA_BIND", "1") != "0" else None
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
    fused = None
    # default: copy-engine pushes (overlap the next sweep: 0.286 ms/step at 8 GPUs); TKB_GATHER=fused selects the one-kernel
    # back-track + NVLink store exchange (lowest latency for a single decode, 0.321 ms/step in this pipelined loop); nccl
    mode = os.environ.get("TKB_GATHER", "push")
    # (needs every rank to own the same number of tracks: 88 splits evenly over 2, 4 and 8 GPUs)
    if world > 1 and n_total % world == 0 and mode in ("fused", "push"):
        try:
            if mode == "fused":
                fused = FusedPushGather(n_local, T, dev)
            else:
                push = PushGather(n_local, 2 + 4 * T, dev)
        except Exception as exc:  # no symmetric memory / P2P: NCCL path
            if rank == 0:
                print(f"[bench] symmetric-memory gather unavailable ({type(exc).__name__}: {exc}); using NCCL", file=sys.stderr)
            push = fused = None
    if world > 1:  # every rank must take the same path
        agree = torch.tensor([1 if (push is not None or fused is not None) else 0], device=dev)
        dist.all_reduce(agree, op=dist.ReduceOp.MIN)
        if int(agree.item()) == 0:
            push = fused = None

    with_scorer = args.workload == "scorer+crf"
    ev_scorer = []
    D_MODEL = 256
    if with_scorer:
        from transkun_b200.LayersTransformer import sip_score
        gq = torch.Generator().manual_seed(4321 + rank)
        q_pin = torch.randn((n_local, T, D_MODEL), generator=gq).pin_memory()
        k_pin = torch.randn((n_local, T, D_MODEL), generator=gq).pin_memory()
        dg_pin = torch.randn((n_local, T), generator=gq).pin_memory()
        q_d, k_d, dg_d = q_pin.to(dev), k_pin.to(dev), dg_pin.to(dev)
        score.zero_()  # the scorer writes end >= begin; the rest of the buffer is defined as zero

    def step(record=False):
        if with_scorer:  # one-pass TF32: the regime the reference trains in (train.py:41-43)
            if record:
                s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s0.record(stream)
            sip_score(q_d, k_d, dg_d, out=score, precise=False)
            if record:
                s1.record(stream)
                ev_scorer.append((s0, s1))
        if record:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
        code, _, lse, _ = sweep(score, noise, BACKWARD, SWEEP_VITERBI | SWEEP_LOGSUM)
        if record:
            e1.record(stream)
            ev_sweep.append((e0, e1))
        if fused is not None:
            # back-track + NVLink push of the records in one kernel.  The gather of step k is complete when all ranks'
            # flags of step k are in; waiting for them one step later (step k-1 here) keeps the ranks out of lock-step:
            # the exchange of a step overlaps the next sweep, as with the copy-engine variant
            st = fused.submit(code, None, BACKWARD, lse[0])
            return fused.result(st - 1) if st > 1 else None
        rec = backtrack_records(code, None, BACKWARD, lse[0])  # [count, logZ, pairs] per track
        if push is not None:
            return push.result(push.submit(rec))  # complete once the stream has passed push.wait()
        if world > 1:  # one NCCL all-gather of the packed records over NVLink
            rec = gather_records(rec, n_total)
        return rec

    def fence():
        if fused is not None and fused.step > 0:
            fused.result(fused.step)  # the last step's exchange has landed everywhere
        if push is not None:
            push.wait()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    if push is not None or fused is not None:  # once, untimed: the pushed gather equals the NCCL gather
        code0, _, lse0, _ = sweep(score, noise, BACKWARD, SWEEP_VITERBI | SWEEP_LOGSUM)
        rec0 = backtrack_records(code0, None, BACKWARD, lse0[0])
        want = gather_records(rec0, n_total)
        if fused is not None:
            got = fused.result(fused.submit(code0, None, BACKWARD, lse0[0]))
        else:
            got = push.result(push.submit(rec0))
            push.wait()
        torch.cuda.synchronize(dev)
        live = torch.arange(want.shape[1], device=dev)[None, :] < (2 + 2 * want[:, :1])   # count, logZ, live pairs
        assert torch.equal(torch.where(live, got, 0), torch.where(live, want, 0)), "pushed gather differs from the NCCL gather"

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
    if fused is not None:
        fused.result(fused.step)
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
    def e2e_crf():
        if with_scorer:  # HOST q, k, diag, noise -> scorer -> semi-CRF -> the reference's Python lists
            qd, kd, dgd = q_pin.to(dev, non_blocking=True), k_pin.to(dev, non_blocking=True), dg_pin.to(dev, non_blocking=True)
            sip_score(qd, kd, dgd, out=score, precise=False)
            return NeuralSemiCRFInterval(score, noise_pin.to(dev, non_blocking=True))
        # public API from HOST tensors: uploads the part of the score tensor the semi-CRF reads (end >= begin)
        return NeuralSemiCRFInterval.fromHost(score_pin, noise_pin, dev)

    with torch.no_grad():  # one untimed pass: first-use allocations of the e2e path
        e2e_crf().decodeWithLogZ()
    torch.cuda.synchronize(dev)
    tw0 = time.perf_counter()
    for _ in range(e2e_steps):
        with torch.no_grad():
            crf = e2e_crf()
            dec, logz = crf.decodeWithLogZ()
            logz_h = logz.cpu()
        maxc = max((len(d) for d in dec), default=0)
        d2h_bytes = n_local * 4 + n_local * maxc * 8 + logz_h.numel() * 4
    torch.cuda.synchronize(dev)
    e2e_sec = (time.perf_counter() - tw0) / e2e_steps
    # where an e2e step goes (untimed extra pass, a synchronisation between the phases)
    bd = {}
    with torch.no_grad():
        torch.cuda.synchronize(dev)
        t_a = time.perf_counter()
        crf = e2e_crf()  # scorer+crf: includes the scorer launch
        torch.cuda.synchronize(dev)
        t_b = time.perf_counter()
        pairs_d, counts_d, logz_d = crf.decode_packed(None, False, with_logz=True)
        torch.cuda.synchronize(dev)
        t_c = time.perf_counter()
        counts_h = counts_d.cpu()
        pairs_h = pairs_d[:, : int(counts_h.max())].cpu()
        logz_d.cpu()
        t_d = time.perf_counter()
        from transkun_b200.CRF.NeuralSemiCRFInterval import _pairs_to_lists
        _pairs_to_lists(pairs_d, counts_d)
        t_e = time.perf_counter()
        bd = {"h2d_ms": (t_b - t_a) * 1e3, "gpu_ms": (t_c - t_b) * 1e3, "d2h_ms": (t_d - t_c) * 1e3,
              "host_lists_ms": (t_e - t_d) * 1e3 - (t_d - t_c) * 1e3,
              "note": "h2d = pinned staircase upload of the lower triangle (PCIe); host_lists = building the reference's "
                      "List[List[Tuple[int,int]]] result (~1.6e5 Python tuples), which the reference's decode() also returns"}
        del pairs_h
    if world > 1:
        tmax = torch.tensor([e2e_sec], device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e_sec = float(tmax.item())
    h2d_bytes = NeuralSemiCRFInterval.lowerTriangleUploadBytes(T, n_local) + noise_pin.numel() * 4
    if with_scorer:
        h2d_bytes = (q_pin.numel() + k_pin.numel() + dg_pin.numel() + noise_pin.numel()) * 4

    # ---- other shapes of the same kernel (device-generated inputs; CUDA events) -------------------------------
    def time_shape(Ts, Ns, pad_to=None, reps=10):
        P = pad_to or Ns
        buf = torch.randn((Ts, Ts, P), device=dev)
        sc = buf[:, :, :Ns]
        nz = torch.randn((Ts - 1, Ns), device=dev)
        for _ in range(3):
            sweep(sc, nz, BACKWARD, SWEEP_VITERBI | SWEEP_LOGSUM)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            sweep(sc, nz, BACKWARD, SWEEP_VITERBI | SWEEP_LOGSUM)
        e1.record(stream)
        torch.cuda.synchronize(dev)
        us = e0.elapsed_time(e1) * 1e3 / reps
        gbs = 4.0 * Ns * Ts * (Ts + 1) / 2.0 / us / 1e3
        del buf, sc, nz
        return {"T": Ts, "N": Ns, "track_pitch": P, "sweep_us": us, "GBps": gbs, "cells_per_s": float(Ts) * Ts * Ns / us * 1e6}

    # ---- the interval scorer (tkb_sip_score) in front of the same sweep: q, k [N, T, 256] -> score -> logZ + decode ----
    def time_scorer(Ts, Ns, reps=10):
        from transkun_b200.LayersTransformer import sip_score as score_fn
        P = (Ns + 7) // 8 * 8
        gsc = torch.Generator(device=dev).manual_seed(99)
        qs = torch.randn((Ns, Ts, D_MODEL), device=dev, generator=gsc)
        ks = torch.randn((Ns, Ts, D_MODEL), device=dev, generator=gsc)
        ds = torch.randn((Ns, Ts), device=dev, generator=gsc)
        nz = torch.randn((Ts - 1, Ns), device=dev, generator=gsc)
        sc = torch.zeros((Ts, Ts, P), device=dev)[:, :, :Ns]

        def timed(fn):
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(reps):
                fn()
            e1.record(stream)
            torch.cuda.synchronize(dev)
            return e0.elapsed_time(e1) * 1e3 / reps

        def chain():
            score_fn(qs, ks, ds, out=sc, precise=False)
            code, _, lse, _ = sweep(sc, nz, BACKWARD, SWEEP_VITERBI | SWEEP_LOGSUM)
            backtrack_records(code, None, BACKWARD, lse[0])

        one = timed(lambda: score_fn(qs, ks, ds, out=sc, precise=False))
        three = timed(lambda: score_fn(qs, ks, ds, out=sc, precise=True))
        both = timed(chain)
        alg = 4.0 * Ns * Ts * (Ts + 1) / 2.0 + 2.0 * Ns * Ts * D_MODEL * 4.0   # output written once + operands read once
        del qs, ks, ds, nz, sc
        return {"T": Ts, "N": Ns, "D": D_MODEL, "track_pitch": P, "scorer_tf32_us": one, "scorer_3xtf32_us": three,
                "scorer_plus_crf_us": both, "cells_per_s_from_qk": float(Ts) * Ts * Ns / both * 1e6,
                "algorithmic_bytes": alg, "GBps": alg / one / 1e3,
                "note": "3xtf32 includes the operand split (tkb_sip_split3); scorer_plus_crf = tf32 scorer + fused sweep + "
                        "back-track, CUDA events"}

    shapes = None
    scorer_shapes = None
    if world == 1 and not args.no_extras:
        del score, noise
        torch.cuda.empty_cache()
        scorer_shapes = [time_scorer(T, n_local), time_scorer(691, 90)]
        torch.cuda.empty_cache()
        shapes = [time_shape(691, 90, pad_to=96), time_shape(1024, 88)]
        if args.tsweep:
            shapes += [time_shape(t_, 88, reps=5) for t_ in (256, 512, 2048, 4096)]

    if rank == 0:
        peak, peak_src = peak_hbm()
        alg_bytes = 4.0 * n_local * T * (T + 1) / 2.0
        achieved = alg_bytes / (sweep_ms * 1e-3) / 1e9
        traffic = ncu_traffic(T, n_local)
        for sh in (shapes or []) + (scorer_shapes or []):
            sh["frac_of_hbm_peak"] = sh["GBps"] / peak
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "tkb::sweep_kernel<BACKWARD, A16, VITERBI|LOGSUM>",
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": sweep_ms, "peak_source": peak_src}
        if with_scorer:  # the scorer kernel dominates this workload: output written once + operands read once
            sc_ms = statistics.mean(a.elapsed_time(b) for a, b in ev_scorer)
            sc_bytes = alg_bytes + 2.0 * n_local * T * D_MODEL * 4.0
            roof = {"bound": "hbm", "achieved": sc_bytes / (sc_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                    "frac": sc_bytes / (sc_ms * 1e-3) / 1e9 / peak, "traffic": ncu_traffic(T, n_local, "scorer_traffic.json"),
                    "kernel": "tkb::sip_scorer_kernel",
                    "algorithmic_bytes_per_launch": sc_bytes, "kernel_ms": sc_ms, "peak_source": peak_src,
                    "sweep": {"achieved": achieved, "frac": achieved / peak, "kernel_ms": sweep_ms}}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": (dict(workload_config(T, n_local, n_total, world), workload_variant="scorer+crf: each step computes "
                            "score[T,T,N] from q, k [N,T,256] fp32 randn (tkb_sip_score, one TF32 pass) before the semi-CRF step")
                       if with_scorer else workload_config(T, n_local, n_total, world)),
            "exchange": ("none" if world == 1 else (
                "fused: the back-track kernel stores every record into all ranks' symmetric (NVLink peer) buffers and "
                "publishes a step flag" if fused is not None else (
                    "copy-engine pushes into symmetric memory, overlapped with the next sweep" if push is not None
                    else "NCCL all-gather per step"))),
            "e2e": {"value": cells_per_step / e2e_sec, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": d2h_bytes, "steps": e2e_steps, "ms_per_step": e2e_sec * 1e3, "breakdown": bd,
                    "api": ("sip_score(q, k, diag from pinned host tensors) -> NeuralSemiCRFInterval(score, noise).decodeWithLogZ(); "
                            "returns the reference's Python interval lists + logZ" if with_scorer else
                            "NeuralSemiCRFInterval.fromHost(score, noise, device).decodeWithLogZ() from pinned host tensors "
                            "(uploads the lower-triangle staircase of score, the part the semi-CRF reads); returns the "
                            "reference's Python interval lists + logZ")},
            "gpu_launches": (3 if with_scorer else 2) * args.steps,
            "roofline": roof,
            "other_shapes": shapes,
            "scorer": scorer_shapes,
            "numa_node_rank0": numa_node,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            if reference_available():
                base, _ = time_reference(T, n_local, 3, 2, budget_s=30.0)
                port, _ = time_port(T, n_local, 3, 1)
                line["cpu_baseline"] = base
                line["cpu_baseline_port"] = port
            else:
                base, _ = time_port(T, n_local, 3, 1)
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
