"""Drop-in for the reference's transcription entry point.

    python -m transkun_b200.transcribe <audioPath> <outPath> [--weight W] [--conf C] [--device D]
                                       [--segmentHopSize S] [--segmentSize S]

Same positional arguments and flags as `python -m transkun.transcribe` / `transkun.transcribe:main`
(/root/reference/transkun/transcribe.py:19-36) -- they are parsed by the reference's own argparse: this entry only
installs the B200-native hot path into the reference package and delegates to its `main()`.  Everything that is out
of scope here (audio decoding, the backbone, note assembly, MIDI writing) stays the reference's own code, so the
reference package must be importable (`pip install transkun`).

What gets installed:
  * `transkun.CRF` -> `transkun_b200.CRF` (always): `TransKun.processFramesBatch` constructs
    `CRF.NeuralSemiCRFInterval` (ModelTransformer.py:222) after `from . import CRF` (:14), which resolves to the module
    registered in `sys.modules` here;
  * with TKB_PATCH_SCORER=1 also `ScaledInnerProductIntervalScorer` (LayersTransformer.py:381), whose parameters have
    the same names, so the shipped checkpoint loads unchanged.
The reference defaults to `--device cpu`; there is no CPU path here, so `--device cuda` is appended when the caller
gives none.
"""
from __future__ import annotations

import importlib
import os
import sys


def install(patch_scorer: bool | None = None) -> None:
    """Register the B200-native modules under the reference's names.  Call before the reference model is imported
    (or constructed); idempotent."""
    from . import CRF as tkb_crf
    from .CRF import NeuralSemiCRFInterval as tkb_crf_impl

    sys.modules["transkun.CRF"] = tkb_crf
    sys.modules["transkun.CRF.NeuralSemiCRFInterval"] = tkb_crf_impl
    pkg = sys.modules.get("transkun")
    if pkg is not None:
        setattr(pkg, "CRF", tkb_crf)
    for name in ("transkun.ModelTransformer", "transkun.Model_ablation"):  # already imported: rebind their global
        mod = sys.modules.get(name)
        if mod is not None and hasattr(mod, "CRF"):
            mod.CRF = tkb_crf
    if patch_scorer is None:
        patch_scorer = os.environ.get("TKB_PATCH_SCORER", "0") == "1"
    if patch_scorer:
        from .LayersTransformer import ScaledInnerProductIntervalScorer
        layers = importlib.import_module("transkun.LayersTransformer")
        layers.ScaledInnerProductIntervalScorer = ScaledInnerProductIntervalScorer
        for name in ("transkun.ModelTransformer",):
            mod = sys.modules.get(name)
            if mod is not None and hasattr(mod, "ScaledInnerProductIntervalScorer"):
                mod.ScaledInnerProductIntervalScorer = ScaledInnerProductIntervalScorer


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
