"""Loader for the UNMODIFIED reference (Yujia-Yan/Transkun) installed under baseline/_ref.

Install recipe (run once in the build container; recorded in DESIGN.md, also done by __graft_entry__.build() when
/root/reference is present):

    python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
        --target baseline/_ref /root/reference

baseline/_ref is git-ignored but NOT gpurun-ignored: it travels to the GPU box with the shipped checkpoint
(transkun/pretrained/2.0.pt, 2.0.conf).  Nothing here reads /root/reference.

Who may use this: bench.py's reference arm and cpu_baseline leg, and tests/ (parity of config 3 against the
reference's own model).  The product (transkun_b200/) never imports it.

The reference imports pretty_midi / mir_eval at module level (Data.py:11, Evaluation.py:1) and moduleconf in
transcribe.py:5; none is installed in this image and none is touched on the paths used here, so
empty stand-in modules are registered for them.
"""
from __future__ import annotations

import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")
_SHIMS = ("pretty_midi", "mir_eval", "mir_eval.transcription", "mir_eval.transcription_velocity", "mir_eval.util",
          "moduleconf", "ncls", "pydub", "soxr")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "transkun", "CRF", "NeuralSemiCRFInterval.py"))


def import_reference():
    """Returns the reference's `transkun` package (imported from baseline/_ref)."""
    if not available():
        raise RuntimeError(f"the reference is not installed under {REF_DIR}; see baseline/ref_loader.py for the recipe")
    for name in _SHIMS:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import transkun  # noqa: F401
    if not os.path.abspath(transkun.__file__).startswith(REF_DIR):
        raise RuntimeError(f"`transkun` resolved to {transkun.__file__}, not to the reference under {REF_DIR}")
    return transkun


def reference_crf():
    """The reference's CRF class (transkun/CRF/NeuralSemiCRFInterval.py:553)."""
    import_reference()
    import importlib
    return importlib.import_module("transkun.CRF.NeuralSemiCRFInterval").NeuralSemiCRFInterval


def checkpoint_paths():
    d = os.path.join(REF_DIR, "transkun", "pretrained")
    return os.path.join(d, "2.0.conf"), os.path.join(d, "2.0.pt")


def load_model(device="cpu", fresh_modules: bool = False):
    """TransKun V2 with the shipped weights, built the way transcribe.py:44-64 does (the 37-line conf JSON is parsed
    here because moduleconf is not installed).  Returns (model, conf)."""
    import torch
    import_reference()
    import importlib
    mt = importlib.import_module("transkun.ModelTransformer")
    conf_path, weight_path = checkpoint_paths()
    conf = mt.Config()
    conf.__dict__.update(json.load(open(conf_path))["Model"]["config"])
    model = mt.TransKun(conf=conf)
    ckpt = torch.load(weight_path, map_location="cpu")
    state = ckpt["best_state_dict"] if "best_state_dict" in ckpt else ckpt["state_dict"]
    model.load_state_dict(state, strict=False)
    return model.to(device).eval(), conf


def synthetic_audio(seconds: float = 16.0, fs: int = 44100, seed: int = 0):
    """SURVEY.md section 8d input (D): decaying harmonic piano-like tones + 1e-3 noise, stereo [nSample, 2] fp32."""
    import numpy as np
    rs = np.random.RandomState(seed)
    n = int(seconds * fs)
    t = np.arange(n) / fs
    x = np.zeros((n, 2), dtype=np.float64)
    onset = 0.3
    while onset < seconds - 1.0:
        for _ in range(rs.randint(1, 4)):
            pitch = rs.randint(40, 88)
            f0 = 440.0 * 2.0 ** ((pitch - 69) / 12.0)
            dur = rs.uniform(0.3, 1.5)
            m = (t >= onset) & (t < onset + dur + 1.0)
            tt = t[m] - onset
            env = np.exp(-3.0 * tt) * (tt < dur) + np.exp(-3.0 * dur) * np.exp(-20.0 * (tt - dur)) * (tt >= dur)
            tone = sum((0.6 ** h) * np.sin(2 * np.pi * f0 * (h + 1) * tt) for h in range(5))
            amp = rs.uniform(0.05, 0.2)
            pan = rs.uniform(0.3, 0.7)
            x[m, 0] += amp * pan * env * tone
            x[m, 1] += amp * (1 - pan) * env * tone
        onset += rs.uniform(0.15, 0.6)
    x += 1e-3 * rs.randn(n, 2)
    return x.astype(np.float32)
