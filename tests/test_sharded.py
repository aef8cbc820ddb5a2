"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU path (track sharding and the
gather of packed decode records).  The decode itself is exercised with the oracle standing in for the
kernel -- here it is the checker of the gather, not a product path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from golden_util import make_inputs


def test_track_shard_partitions():
    from transkun_b200.sharded import shard_sizes, track_shard
    for n in (1, 7, 88, 90, 360):
        for w in (1, 2, 3, 4, 8):
            spans = [track_shard(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(shard_sizes(n, w)) - min(shard_sizes(n, w)) <= 1
    assert shard_sizes(88, 8) == [11] * 8
    with pytest.raises(ValueError):
        track_shard(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, T, N, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.semicrf_oracle import SemiCRFOracle
        from transkun_b200.sharded import gather_decoded, gather_records, gather_vector, split_records, track_shard
        score, noise = make_inputs("randn", T, N, 21)
        lo, hi = track_shard(N, world, rank)
        o = SemiCRFOracle(score[:, :, lo:hi], noise[:, lo:hi])
        pairs, counts = o.decode_packed()
        gp, gc = gather_decoded(torch.from_numpy(pairs), torch.from_numpy(counts), N)
        gz = gather_vector(torch.from_numpy(o.computeLogZ()), N)
        full = SemiCRFOracle(score, noise)
        fp, fc = full.decode_packed()
        ok = bool(np.array_equal(gc.numpy(), fc))
        for n in range(N):
            ok &= bool(np.array_equal(gp[n, : fc[n]].numpy(), fp[n, : fc[n]]))
        ok &= bool(np.allclose(gz.numpy(), full.computeLogZ(), rtol=1e-6))
        gp2, _ = gather_decoded(torch.from_numpy(pairs), torch.from_numpy(counts), N, max_pairs=int(fc.max()))
        ok &= gp2.shape == (N, int(fc.max()), 2)
        # single-collective record exchange: [count, logZ bits, pairs...] per track
        rec = torch.zeros((hi - lo, 2 + 4 * T), dtype=torch.int32)
        rec[:, 0] = torch.from_numpy(counts)
        rec[:, 1] = torch.from_numpy(o.computeLogZ()).view(torch.int32)
        rec[:, 2:] = torch.from_numpy(pairs).reshape(hi - lo, 4 * T)
        rc, rz, rp = split_records(gather_records(rec, N))
        ok &= bool(np.array_equal(rc.numpy(), fc)) and bool(np.allclose(rz.numpy(), full.computeLogZ(), rtol=1e-6))
        for n in range(N):
            ok &= bool(np.array_equal(rp[n, : fc[n]].numpy(), fp[n, : fc[n]]))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_gather_decoded_world2_gloo():
    world, T, N = 2, 24, 7  # uneven shard: 4 + 3 tracks
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, T, N, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
