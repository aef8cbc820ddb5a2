"""CPU: the C-ABI library loads and exports every symbol include/transkun_b200.h declares;
argument validation works without a GPU (no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from transkun_b200 import build, _lib
    build.build()
    return _lib.load()


def _declared():
    text = open(os.path.join(ROOT, "include", "transkun_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tkb_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    from transkun_b200 import _lib
    names = _declared()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/transkun_b200.h but not exported"
    assert set(names) == set(_lib.EXPORTS)


def test_version_and_sizes(lib):
    assert lib.tkb_version() == 1
    # sweep workspace: header + row mailbox (2 semirings * T * ceil8(N) words) + far partials
    # ([groups of 8 tracks][blocks][2 semirings][8][32 columns][2] words)
    def ws(T, N):
        npad, nb, G = (N + 7) // 8 * 8, (T + 31) // 32, (N + 7) // 8
        return 256 + 2 * T * npad * 8 + G * nb * 2 * 8 * 32 * 2 * 8
    assert lib.tkb_sweep_workspace_bytes(2048, 88) == ws(2048, 88)
    assert lib.tkb_sweep_workspace_bytes(10, 9) == ws(10, 9)
    assert lib.tkb_sweep_workspace_bytes(300, 1200) == ws(300, 1200)
    assert lib.tkb_sweep_workspace_bytes(0, 4) == 0


def test_argument_validation_without_gpu(lib):
    from transkun_b200 import _lib
    rc = lib.tkb_semicrf_sweep(None, None, 8, 4, 0, 1, None, 1, None, None, None, None)
    assert rc == -1
    assert b"invalid argument" in lib.tkb_last_error()
    with pytest.raises(_lib.TkbError):
        _lib.check(rc, "tkb_semicrf_sweep")
    assert lib.tkb_semicrf_backtrack(None, 8, 4, None, 0, None, None, None) == -1
    assert lib.tkb_semicrf_marginals(None, None, 8, 4, None, None, None, None, None, None) == -1
    assert lib.tkb_semicrf_evalpath(None, None, 8, 4, None, None, None, None, None) == -1


def test_no_cpu_fallback():
    import torch
    from transkun_b200.CRF import NeuralSemiCRFInterval
    crf = NeuralSemiCRFInterval(torch.zeros(4, 4, 2), torch.zeros(3, 2))
    with pytest.raises(RuntimeError, match="no CPU path"):
        crf.decode()
    with pytest.raises(RuntimeError, match="no CPU path"):
        crf.computeLogZ()
    with pytest.raises(AssertionError):
        NeuralSemiCRFInterval(torch.zeros(4, 3, 2), torch.zeros(3, 2)).decode()


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under transkun_b200/ may import, link or execute it."""
    pkg = os.path.join(ROOT, "transkun_b200")
    pat = re.compile(r"(import\s+oracle|from\s+oracle|oracle/|semicrf_oracle|tko_)")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not pat.search(src), f"{os.path.join(dirpath, f)} references the oracle"
