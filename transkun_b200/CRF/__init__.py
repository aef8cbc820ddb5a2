# mirrors reference transkun/CRF/__init__.py:1
from .NeuralSemiCRFInterval import NeuralSemiCRFInterval  # noqa: F401
