"""STFT / log-mel frontend: oracle and host-side mirror vs the reference's golden outputs (CPU), the
cuFFT-backed CUDA path vs the oracle (GPU, through the C ABI).  Tolerance: outputs are normalised logs in
[0, 1]; fp32 FFT + fp32 accumulation give ~1e-6, asserted at 2e-5 absolute / 1e-4 relative."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle.frontend_oracle import logmel

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "frontend_*.npz")))


def _module(z, device="cpu"):
    from transkun_b200.Util import MelSpectrum
    m = MelSpectrum(int(z["W"]), f_min=float(z["fmin"]), f_max=float(z["fmax"]), n_mels=int(z["nmel"]), fs=int(z["fs"]),
                    nExtraWins=int(z["nextra"]), log=True, toMono=True).eval()
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd_")}
    # our filterbank is computed without torchaudio: it must equal the reference's buffer
    torch.testing.assert_close(m.freq2mels, sd["freq2mels"], rtol=1e-5, atol=1e-6)
    missing, unexpected = m.load_state_dict(sd, strict=True)
    return m.to(device)


@pytest.mark.parametrize("path", GOLD, ids=lambda p: os.path.basename(p)[:-4])
def test_oracle_and_host_mirror_match_reference(path):
    from transkun_b200.Util import makeFrame
    z = np.load(path)
    m = _module(z)
    frames = makeFrame(torch.from_numpy(z["audio"]), int(z["hop"]), int(z["W"]))
    assert frames.shape[-2] == int(z["nframe"])
    wins = m.spectrogramExtractor.windows().detach().numpy()
    ref = logmel(frames.numpy(), wins, m.freq2mels.numpy(), 1e-5, True)
    np.testing.assert_allclose(ref, z["out"], rtol=1e-4, atol=2e-5)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(frames)


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLD, ids=lambda p: os.path.basename(p)[:-4])
def test_kernel_matches_reference(path):
    from transkun_b200.Util import makeFrame
    z = np.load(path)
    m = _module(z, "cuda")
    frames = makeFrame(torch.from_numpy(z["audio"]).cuda(), int(z["hop"]), int(z["W"]))  # strided view, read in place
    assert not frames.is_contiguous()
    with torch.no_grad():
        out = m(frames)
        out2 = m(frames.contiguous())
    assert out.shape == z["out"].shape
    np.testing.assert_allclose(out.cpu().numpy(), z["out"], rtol=1e-4, atol=2e-5)
    assert torch.equal(out, out2)


@pytest.mark.gpu
def test_kernel_full_segment_shape_vs_oracle():
    """One 16 s stereo segment at the shipped configuration: [1,2,691,4096] -> [1,1,691,229,6]."""
    from transkun_b200.Util import MelSpectrum, makeFrame
    torch.manual_seed(0)
    m = MelSpectrum(4096, f_min=30, f_max=8000, n_mels=229, fs=44100, nExtraWins=5, log=True, toMono=True).cuda().eval()
    t = torch.arange(705600) / 44100.0
    audio = torch.stack([torch.sin(2 * math.pi * 440 * t) * torch.exp(-t), torch.sin(2 * math.pi * 660 * t) * 0.5])
    audio = (audio + 1e-3 * torch.randn(2, 705600))[None]
    frames = makeFrame(audio.cuda(), 1024, 4096)
    with torch.no_grad():
        out = m(frames)
        mono_off = MelSpectrum(4096, 30, 8000, 229, 44100, nExtraWins=5, log=True, toMono=False).cuda().eval()
        out_st = mono_off(frames)
    assert out.shape == (1, 1, 691, 229, 6) and out_st.shape == (1, 2, 691, 229, 6)
    sel = [0, 1, 345, 689, 690]
    wins = m.spectrogramExtractor.windows().detach().cpu().numpy()
    ref = logmel(frames[:, :, sel].cpu().numpy(), wins, m.freq2mels.cpu().numpy(), 1e-5, True)
    np.testing.assert_allclose(out[:, :, sel].cpu().numpy(), ref, rtol=1e-4, atol=2e-5)
    ref_st = logmel(frames[:, :, sel].cpu().numpy(), wins, m.freq2mels.cpu().numpy(), 1e-5, False)
    np.testing.assert_allclose(out_st[:, :, sel].cpu().numpy(), ref_st, rtol=1e-4, atol=2e-5)


import math  # noqa: E402


@pytest.mark.gpu
@pytest.mark.parametrize("lead", [(), (2,), (3, 2), (2, 1, 2)])
def test_leading_shapes_follow_the_reference(lead):
    """The reference accepts any leading shape [..., nFrame, W]; toMono averages the dim right before nFrame whenever
    the input has one (Util.py:158-161, keepdim=True): 2-D frames stay un-averaged, 3-D frames [C, nFrame, W] become
    [1, nFrame, nMel, nWin]."""
    from transkun_b200.Util import MelSpectrum
    torch.manual_seed(len(lead))
    m = MelSpectrum(256, f_min=30, f_max=8000, n_mels=24, fs=44100, nExtraWins=2, log=True, toMono=True).cuda().eval()
    frames = torch.randn(*lead, 9, 256)
    with torch.no_grad():
        out = m(frames.cuda())
    wins = m.spectrogramExtractor.windows().detach().cpu().numpy()
    ref = logmel(frames.numpy(), wins, m.freq2mels.cpu().numpy(), 1e-5, True)
    assert out.shape == ref.shape
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=1e-4, atol=2e-5)
