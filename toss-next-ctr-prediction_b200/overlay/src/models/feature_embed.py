"""Overlay for `src.models.feature_embed` (reference src/models/feature_embed.py)."""
from b200ctr.modules import NumericFeatureEmbedding, BinaryFeatureEmbedding  # noqa: F401
