"""Two GPUs, torchrun over NCCL: the fused back-track + NVLink push exchange (tkb_semicrf_backtrack_push,
transkun_b200.sharded.FusedPushGather) and the copy-engine exchange (PushGather) against the NCCL all-gather of the same
records, over several steps (both record buffers, the step-flag protocol) with forced starts and both directions."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.join(%r, "tests")); sys.path.insert(0, %r)
from golden_util import make_inputs
from transkun_b200.CRF.NeuralSemiCRFInterval import backtrack_records, sweep
from transkun_b200._lib import BACKWARD, FORWARD, SWEEP_LOGSUM, SWEEP_VITERBI
from transkun_b200.sharded import FusedPushGather, PushGather, gather_records
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
ok = True
for T, n_local in ((97, 5), (300, 11)):
    fused = FusedPushGather(n_local, T, dev)
    push = PushGather(n_local, 2 + 4 * T, dev)
    for step in range(5):
        score, noise = make_inputs("randn", T, n_local, 100 * step + rank)
        s, z = torch.from_numpy(score).to(dev), torch.from_numpy(noise).to(dev)
        direction = FORWARD if step == 3 else BACKWARD
        forced = None if step %% 2 == 0 else torch.randint(0, T, (n_local,), dtype=torch.int32, device=dev)
        code, _, lse, _ = sweep(s, z, direction, SWEEP_VITERBI | SWEEP_LOGSUM)
        logz = lse[T - 1 if direction == FORWARD else 0]
        want = gather_records(backtrack_records(code, forced, direction, logz), world * n_local)
        got = fused.result(fused.submit(code, forced, direction, logz))
        got2 = push.result(push.submit(backtrack_records(code, forced, direction, logz)))
        push.wait()
        torch.cuda.synchronize(dev)
        cnt = want[:, 0]
        for n in range(world * n_local):
            c = int(cnt[n])
            ok &= bool(torch.equal(got[n, : 2 + 2 * c], want[n, : 2 + 2 * c]))
            ok &= bool(torch.equal(got2[n, : 2 + 2 * c], want[n, : 2 + 2 * c]))
    ok &= not fused.timed_out()
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("FUSED_PUSH_OK" if int(flag.item()) == 1 else "FUSED_PUSH_MISMATCH")
dist.destroy_process_group()
''' % (ROOT, ROOT)


@pytest.mark.gpu
def test_fused_push_exchange_two_gpus(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
                         cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert "FUSED_PUSH_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
