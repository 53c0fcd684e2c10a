"""Namespace-package overlay for `src.models.wrapper` (reference src/models/wrapper.py).
PYTHONPATH=<repo>/toss-next-ctr-prediction_b200/overlay:<repo>/toss-next-ctr-prediction_b200:<reference> python -m src.train --cfg ...
resolves `src.models.*` here and everything else (`src.train`, `src.data`, `src.utils`) from the reference."""
from b200ctr.modules import CTRModel  # noqa: F401
from .dare import DARE  # noqa: F401
from .qnn_alpha import QNNAlphaDetailed  # noqa: F401
from .feature_embed import NumericFeatureEmbedding, BinaryFeatureEmbedding  # noqa: F401
