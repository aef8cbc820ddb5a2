"""CPU: host-side logic of the Python mirror that needs no GPU (packing of interval lists, the staircase upload
size, the padded-track-axis predicate)."""
import numpy as np
import pytest
import torch

from transkun_b200.CRF import NeuralSemiCRFInterval, PackedIntervals, pack_intervals
from transkun_b200.CRF.NeuralSemiCRFInterval import _pitched, _prep_score_for_sweep


def test_pack_intervals_is_csr_of_the_lists():
    iv = [[(0, 0), (0, 3), (3, 3)], [], [(2, 5)], [(1, 1), (4, 4), (4, 6), (6, 6)]]
    p = pack_intervals(iv, T=7)
    assert isinstance(p, PackedIntervals)
    assert p.offsets.tolist() == [0, 3, 3, 4, 8]
    assert p.pairs.dtype == torch.int32 and p.pairs.tolist() == [list(x) for cur in iv for x in cur]
    back = [[tuple(x) for x in p.pairs[p.offsets[n]:p.offsets[n + 1]].tolist()] for n in range(len(iv))]
    assert back == iv


def test_pack_intervals_checks_range_and_empty_input():
    with pytest.raises(IndexError):
        pack_intervals([[(0, 9)]], T=9)
    with pytest.raises(IndexError):
        pack_intervals([[(-1, 2)]], T=9)
    p = pack_intervals([[], []], T=5)
    assert p.pairs.shape == (0, 2) and p.offsets.tolist() == [0, 0, 0]


def test_lower_triangle_upload_bytes():
    T, N, R = 200, 12, 64
    want = sum((min(e0 + R, T) - e0) * min(e0 + R, T) * N * 4 for e0 in range(0, T, R))
    assert NeuralSemiCRFInterval.lowerTriangleUploadBytes(T, N, R) == want
    dense = T * T * N * 4
    tri = T * (T + 1) // 2 * N * 4
    assert tri <= want < 0.75 * dense  # a staircase: at least the triangle, well below the square
    assert NeuralSemiCRFInterval.lowerTriangleUploadBytes(2048, 88) < 0.52 * 2048 * 2048 * 88 * 4


def test_padded_track_axis_predicate():
    T, N, P = 6, 10, 12
    dense = torch.zeros(T, T, N)
    padded = torch.zeros(T, T, P)[:, :, :N]
    assert _pitched(dense) and _pitched(padded) and not padded.is_contiguous()
    assert not _pitched(dense.permute(1, 0, 2)[:, ::2])       # rows no longer T * pitch apart
    assert not _pitched(torch.zeros(T, N, T).permute(0, 2, 1))  # track axis not innermost
    # small tensors are used as they are (dense stays dense, padded stays padded: no copy)
    assert _prep_score_for_sweep(dense).data_ptr() == dense.data_ptr()
    assert _prep_score_for_sweep(padded).data_ptr() == padded.data_ptr()
    # anything else is made contiguous
    odd = torch.zeros(T, N, T).permute(0, 2, 1)
    assert _prep_score_for_sweep(odd).is_contiguous()
    assert _prep_score_for_sweep(dense.double()).dtype == torch.float32


def test_packed_intervals_are_validated_once():
    """A PackedIntervals object is checked when it is built (endpoints >= 0, begin <= end, CSR offsets) and against T when
    it is used: the kernels index score[end, begin] with these values."""
    import pytest
    import torch
    from transkun_b200.CRF.NeuralSemiCRFInterval import PackedIntervals, _csr, pack_intervals
    ok = pack_intervals([[(0, 2), (4, 6)], [], [(3, 3)]], T=8)
    assert ok.max_endpoint == 6
    with pytest.raises(IndexError):
        _csr(ok, 3, 6, torch.device("cpu"))          # endpoint 6 with T = 6
    with pytest.raises(ValueError):
        _csr(ok, 2, 8, torch.device("cpu"))          # three lists for two tracks
    with pytest.raises(ValueError):
        PackedIntervals(torch.tensor([[5, 2]]), torch.tensor([0, 1]))      # begin > end
    with pytest.raises(IndexError):
        PackedIntervals(torch.tensor([[-1, 2]]), torch.tensor([0, 1]))
    with pytest.raises(ValueError):
        PackedIntervals(torch.tensor([[1, 2]]), torch.tensor([0, 2]))      # offsets past the pairs


def test_numa_binding_reads_the_topology(tmp_path, monkeypatch):
    """bind_to_gpu_numa_node: PCI address of the device -> numa_node -> cpulist -> sched_setaffinity on the intersection
    with the CPUs the process may use; unreadable topology or an empty intersection changes nothing."""
    import os
    import types
    import torch
    from transkun_b200 import sharded
    assert sharded._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    props = types.SimpleNamespace(pci_domain_id=0, pci_bus_id=0x1b, pci_device_id=0)
    monkeypatch.setattr(torch.cuda, "get_device_properties", lambda i: props)
    dev = tmp_path / "bus/pci/devices/0000:1b:00.0"
    dev.mkdir(parents=True)
    (dev / "numa_node").write_text("1\n")
    node = tmp_path / "devices/system/node/node1"
    node.mkdir(parents=True)
    mine = sorted(os.sched_getaffinity(0))
    (node / "cpulist").write_text(f"{mine[0]}\n")
    calls = []
    monkeypatch.setattr(os, "sched_setaffinity", lambda pid, cpus: calls.append(set(cpus)))
    assert sharded.bind_to_gpu_numa_node(0, sysfs=str(tmp_path)) == 1 and calls == [{mine[0]}]
    (node / "cpulist").write_text("100000\n")          # a node whose CPUs this process cannot use
    assert sharded.bind_to_gpu_numa_node(0, sysfs=str(tmp_path)) is None and len(calls) == 1
    (dev / "numa_node").write_text("-1\n")             # no NUMA information
    assert sharded.bind_to_gpu_numa_node(0, sysfs=str(tmp_path)) is None
    assert sharded.bind_to_gpu_numa_node(0, sysfs=str(tmp_path / "nowhere")) is None
