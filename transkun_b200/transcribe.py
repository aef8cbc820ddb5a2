"""Drop-in for the reference's transcription entry point.

    python -m transkun_b200.transcribe <audioPath> <outPath> [--weight W] [--conf C] [--device D]
                                       [--segmentHopSize S] [--segmentSize S]

Same positional arguments and flags as `python -m transkun.transcribe` / `transkun.transcribe:main`
(/root/reference/transkun/transcribe.py:19-36) -- they are parsed by the reference's own argparse: this entry only
installs the B200-native hot path into the reference package and delegates to its `main()`.  Everything that is out
of scope here (audio decoding, the backbone, note assembly, MIDI writing) stays the reference's own code, so the
reference package must be importable (`pip install transkun`).

What gets installed (all three by default; TKB_PATCH_SCORER=0 / TKB_PATCH_FRONTEND=0 keep the reference's torch code):
  * `transkun.CRF` -> `transkun_b200.CRF`: `TransKun.processFramesBatch` constructs `CRF.NeuralSemiCRFInterval`
    (ModelTransformer.py:222) after `from . import CRF` (:14), which resolves to the module registered in
    `sys.modules` here;
  * `ScaledInnerProductIntervalScorer` (LayersTransformer.py:381) and `MelSpectrum` (Util.py:126): same constructor
    arguments, same parameter / buffer names, so the shipped checkpoint loads unchanged.  The scorer follows the
    reference's precision switch (fp32-grade 3xTF32 unless torch.backends.cuda.matmul.allow_tf32 is set).
`install_into(model)` swaps the two modules of an already constructed reference model.
The reference defaults to `--device cpu`; there is no CPU path here, so `--device cuda` is appended when the caller
gives none.
"""
from __future__ import annotations

import importlib
import os
import sys


_MODEL_MODULES = ("transkun.ModelTransformer", "transkun.Model_ablation")


def install(patch_scorer: bool | None = None, patch_frontend: bool | None = None) -> None:
    """Register the B200-native modules under the reference's names.  Call before the reference model is constructed;
    idempotent."""
    from . import CRF as tkb_crf
    from .CRF import NeuralSemiCRFInterval as tkb_crf_impl

    sys.modules["transkun.CRF"] = tkb_crf
    sys.modules["transkun.CRF.NeuralSemiCRFInterval"] = tkb_crf_impl
    pkg = sys.modules.get("transkun")
    if pkg is not None:
        setattr(pkg, "CRF", tkb_crf)
    for name in _MODEL_MODULES:  # already imported: rebind their global
        mod = sys.modules.get(name)
        if mod is not None and hasattr(mod, "CRF"):
            mod.CRF = tkb_crf
    if patch_scorer is None:
        patch_scorer = os.environ.get("TKB_PATCH_SCORER", "1") != "0"
    if patch_frontend is None:
        patch_frontend = os.environ.get("TKB_PATCH_FRONTEND", "1") != "0"
    if patch_scorer:
        from .LayersTransformer import ScaledInnerProductIntervalScorer
        _rebind("transkun.LayersTransformer", "ScaledInnerProductIntervalScorer", ScaledInnerProductIntervalScorer)
    if patch_frontend:
        from .Util import MelSpectrum
        _rebind("transkun.Util", "MelSpectrum", MelSpectrum)


def _rebind(home: str, name: str, obj) -> None:
    """The reference's model modules do `from .Util import *` / `from .LayersTransformer import *`: the class has to
    be replaced in its home module (for later imports) and in every model module that already copied the name."""
    try:
        setattr(importlib.import_module(home), name, obj)
    except ImportError:
        return  # the reference is not importable: main() reports that
    for mod_name in _MODEL_MODULES:
        mod = sys.modules.get(mod_name)
        if mod is not None and hasattr(mod, name):
            setattr(mod, name, obj)


def install_into(model, patch_scorer: bool = True, patch_frontend: bool = True):
    """Swap the frontend and the scorer of an already constructed reference `TransKun` (V2) for the B200-native
    modules, carrying the parameters over; the CRF class is installed as in install().  Returns the model."""
    install(patch_scorer, patch_frontend)
    if patch_frontend and hasattr(model, "framewiseFeatureExtractor"):
        from .Util import MelSpectrum
        old = model.framewiseFeatureExtractor
        if not isinstance(old, MelSpectrum):
            ext = old.spectrogramExtractor
            new = MelSpectrum(ext.win.numel(), 0.0, 1.0, old.freq2mels.shape[1], 2, nExtraWins=ext.nChannel - 1,
                              log=old.log, eps=old.eps, toMono=old.toMono)
            new.freq2mels = old.freq2mels  # the stored buffer is the filterbank (f_min / f_max / fs above are unused)
            new.load_state_dict(old.state_dict())
            model.framewiseFeatureExtractor = new.to(old.freq2mels.device)
    if patch_scorer and getattr(model, "useInnerProductScorer", False):
        from .LayersTransformer import ScaledInnerProductIntervalScorer
        old = model.scorer
        if not isinstance(old, ScaledInnerProductIntervalScorer):
            new = ScaledInnerProductIntervalScorer(old.size, expansionFactor=old.expansionFactor,
                                                   dropoutProb=old.dropout.p)
            new.load_state_dict(old.state_dict())
            model.scorer = new.to(next(old.parameters()).device)
    return model


def main() -> None:
    if not any(a == "--device" or a.startswith("--device=") for a in sys.argv[1:]):
        sys.argv += ["--device", "cuda"]
    install()
    try:
        ref = importlib.import_module("transkun.transcribe")
    except ImportError as exc:  # the reference (or one of its dependencies) is not installed
        raise SystemExit("transkun_b200.transcribe delegates to the reference package, which failed to import: "
                         f"{exc}.  Install it (`pip install transkun`) and retry.") from exc
    ref.main()


if __name__ == "__main__":
    main()
