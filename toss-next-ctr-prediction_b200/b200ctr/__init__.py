"""b200ctr — B200-native (sm_100a) DARE + QNN-alpha CTR hot path behind the reference's nn.Module surfaces.

    from b200ctr import CTRModel, FusedTrainer        # drop-in model + fused training step
    import b200ctr.ops                                # one Python wrapper per C-ABI entry point

All compute runs in libb200ctr.so (hand-written CUDA, built by `make -C csrc` / __graft_entry__.build()).
"""
from ._lib import B200CtrError, LIB_PATH, build, load  # noqa: F401
from .hotpath import DropCtl, ModelSpec  # noqa: F401
from .modules import (BinaryFeatureEmbedding, CTRModel, DARE, DAREEncoderLayer, NumericFeatureEmbedding,  # noqa: F401
                      PositionalBias, QNNAlphaDetailed, RMSNorm, SEBlock, make_norm)
from .trainer import FusedTrainer, cosine_warmup_lr, ema_decay_at  # noqa: F401
from .data import BatchedShardReader  # noqa: F401
from .infer import FoldEnsemble  # noqa: F401

__all__ = ["CTRModel", "DARE", "DAREEncoderLayer", "PositionalBias", "RMSNorm", "make_norm", "QNNAlphaDetailed",
           "SEBlock", "NumericFeatureEmbedding", "BinaryFeatureEmbedding", "FusedTrainer", "BatchedShardReader", "FoldEnsemble", "ModelSpec", "DropCtl",
           "build", "load", "B200CtrError"]
