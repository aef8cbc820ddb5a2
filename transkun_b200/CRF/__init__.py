# mirrors reference transkun/CRF/__init__.py:1
from .NeuralSemiCRFInterval import NeuralSemiCRFInterval, PackedIntervals, pack_intervals  # noqa: F401
