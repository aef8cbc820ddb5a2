"""transkun_b200 -- Blackwell-native (sm_100a) implementation of Transkun's neural
semi-CRF hot path behind the reference's own Python surface.

    from transkun_b200.CRF import NeuralSemiCRFInterval

mirrors `transkun.CRF.NeuralSemiCRFInterval` (reference transkun/CRF/__init__.py:1).
"""
__version__ = "0.1.0"
