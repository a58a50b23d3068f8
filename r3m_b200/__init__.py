"""r3m_b200 — B200-native (sm_100a) implementation of the R3M pretraining hot path.

Public surface mirrors the reference package (r3m/__init__.py:5,44): ``R3M``, ``load_r3m``; the update step lives in
``r3m_b200.trainer.Trainer`` (reference r3m/trainer.py:21).
"""
__all__ = []
