"""Overlay for `src.models.dare` (reference src/models/dare.py)."""
from b200ctr.modules import DARE, DAREEncoderLayer, PositionalBias, RMSNorm, make_norm  # noqa: F401
