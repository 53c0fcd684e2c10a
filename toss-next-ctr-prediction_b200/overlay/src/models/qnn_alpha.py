"""Overlay for `src.models.qnn_alpha` (reference src/models/qnn_alpha.py)."""
from b200ctr.modules import QNNAlphaDetailed, SEBlock, RMSNorm, make_norm  # noqa: F401
