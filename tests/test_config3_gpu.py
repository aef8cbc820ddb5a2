"""BASELINE.json config 3 on the GPU: the reference's own V2 model (unmodified, from baseline/_ref, shipped checkpoint) with
the B200-native frontend + scorer + CRF installed, against (a) the same reference model untouched on the same GPU and (b) the
committed Note list the reference produced on CPU (tests/golden/make_golden_config3.py).  ModelTransformer.py:151-225,
:537-549, :729-848 stay the reference's code; only the three hot-path objects are swapped (transkun_b200.transcribe)."""
import copy
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from baseline import ref_loader  # noqa: E402

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config3_notes.npz")


def _key(n):
    return (n.pitch, round(n.start, 4), round(n.end, 4), n.velocity)


@pytest.fixture(scope="module")
def reference_model():
    if not ref_loader.available():
        pytest.skip("baseline/_ref (the installed reference + checkpoint) is not present")
    model, _ = ref_loader.load_model("cuda")
    return model


def _reference_crf_class():
    """The reference's CRF class loaded straight from its file: `transkun.CRF` in sys.modules may already be ours."""
    import importlib.util
    path = os.path.join(ref_loader.REF_DIR, "transkun", "CRF", "NeuralSemiCRFInterval.py")
    spec = importlib.util.spec_from_file_location("_tkb_reference_crf", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.NeuralSemiCRFInterval


def _transcribe(model, seconds, seed):
    x = torch.from_numpy(ref_loader.synthetic_audio(seconds, seed=seed)).cuda()
    with torch.no_grad():
        return model.transcribe(x)


@pytest.mark.parametrize("seconds,seed", [(16.0, 0), (30.0, 3)])
def test_notes_identical_to_reference_on_the_same_gpu(reference_model, seconds, seed):
    """Everything installed (frontend + 3xTF32 scorer + CRF): same notes in the same order -- pitch and velocity exact,
    onset / offset within 0.1 ms (a frame is 23 ms; the sub-frame refinement heads see log-mel features that differ from
    torch's in the last fp32 bits).  With only the scorer and the CRF installed the lists are identical to the last digit."""
    from transkun_b200.transcribe import install_into
    import transkun.ModelTransformer as MT
    ref_crf = _reference_crf_class()

    class _RefCRF:  # module-like holder: ModelTransformer.py:222 calls CRF.NeuralSemiCRFInterval(...)
        NeuralSemiCRFInterval = ref_crf

    saved = MT.CRF
    MT.CRF = _RefCRF  # the reference's own CRF, whatever an earlier test installed
    try:
        want = _transcribe(reference_model, seconds, seed)
    finally:
        MT.CRF = saved
    ours = install_into(copy.deepcopy(reference_model))
    assert type(ours.framewiseFeatureExtractor).__module__ == "transkun_b200.Util"
    assert type(ours.scorer).__module__ == "transkun_b200.LayersTransformer"
    assert MT.CRF.NeuralSemiCRFInterval.__module__ == "transkun_b200.CRF.NeuralSemiCRFInterval"
    got = _transcribe(ours, seconds, seed)
    assert len(want) > 20 and len(got) == len(want)
    assert [(n.pitch, n.velocity, n.hasOnset, n.hasOffset) for n in got] == \
        [(n.pitch, n.velocity, n.hasOnset, n.hasOffset) for n in want]
    np.testing.assert_allclose([n.start for n in got], [n.start for n in want], rtol=0, atol=1e-4)
    np.testing.assert_allclose([n.end for n in got], [n.end for n in want], rtol=0, atol=1e-4)
    # scorer (3xTF32) + CRF only, the reference's own torch frontend: identical to the last digit
    partial = install_into(copy.deepcopy(reference_model), patch_scorer=True, patch_frontend=False)
    assert type(partial.framewiseFeatureExtractor).__module__ == "transkun.Util"
    got2 = _transcribe(partial, seconds, seed)
    assert [_key(n) for n in got2] == [_key(n) for n in want]


@pytest.mark.parametrize("name,seconds,seed", [("s16_seed0", 16.0, 0), ("s30_seed3", 30.0, 3)])
def test_notes_against_the_reference_cpu_fixture(reference_model, name, seconds, seed):
    """Against the Note list the reference produced on CPU.  The backbone runs in different libraries there (MKL vs
    cuBLAS/cuDNN), so timing attributes are compared to 2 ms and at most 2 % of the notes may differ."""
    from transkun_b200.transcribe import install_into
    g = np.load(GOLDEN)
    ours = install_into(copy.deepcopy(reference_model))
    got = _transcribe(ours, seconds, seed)
    want = set(zip(g[name + "_pitch"].tolist(), np.round(g[name + "_start"] / 2e-3).astype(int).tolist(),
                   np.round(g[name + "_end"] / 2e-3).astype(int).tolist()))
    have = set((n.pitch, int(round(n.start / 2e-3)), int(round(n.end / 2e-3))) for n in got)
    missing, extra = want - have, have - want
    assert len(missing) <= 0.02 * len(want) + 1 and len(extra) <= 0.02 * len(want) + 1, (len(want), sorted(missing), sorted(extra))


def test_batched_segments_give_the_reference_notes(reference_model):
    """transkun_b200.batched.transcribe_batched: the network, scorer and ONE semi-CRF sweep over all segments of a chunk,
    the reference's own loop for the sequential rest (forced start positions, event merging) -- same notes as the
    untouched reference transcribing segment by segment (40 s of audio = 7 segments; chunks of 3 and of 8)."""
    from transkun_b200.batched import transcribe_batched
    from transkun_b200.transcribe import install_into
    import transkun.ModelTransformer as MT
    ref_crf = _reference_crf_class()

    class _RefCRF:
        NeuralSemiCRFInterval = ref_crf

    x = torch.from_numpy(ref_loader.synthetic_audio(40.0, seed=5)).cuda()
    saved = MT.CRF
    MT.CRF = _RefCRF
    try:
        with torch.no_grad():
            want = reference_model.transcribe(x)
    finally:
        MT.CRF = saved
    ours = install_into(copy.deepcopy(reference_model))
    for max_batch in (3, 8):
        got = transcribe_batched(ours, x, max_batch=max_batch)
        assert len(got) == len(want) > 50
        assert [(n.pitch, n.velocity, n.hasOnset, n.hasOffset) for n in got] == \
            [(n.pitch, n.velocity, n.hasOnset, n.hasOffset) for n in want]
        np.testing.assert_allclose([n.start for n in got], [n.start for n in want], rtol=0, atol=1e-4)
        np.testing.assert_allclose([n.end for n in got], [n.end for n in want], rtol=0, atol=1e-4)
    assert "processFramesBatch" not in ours.__dict__   # the temporary override is gone
