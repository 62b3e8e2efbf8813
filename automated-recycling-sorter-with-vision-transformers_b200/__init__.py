"""vitk - B200-native (sm_100a) ViT/DeiT encoder hot path behind the reference's nn.Module API.

The directory name carries hyphens (it mirrors the upstream repository name), so import it with
`importlib.import_module("automated-recycling-sorter-with-vision-transformers_b200")` or through
the `vitk` alias module at the repository root.
"""
from . import _lib, ops  # noqa: F401
from ._lib import VitkError, launch_count  # noqa: F401
from .modules import (DataEfficientImageTransformer, MLPBlock,  # noqa: F401
                      MultiHeadSelfAttention, PatchEmbedding, TransformerBlock, ViTClassifier,
                      VisionTransformer)
from .detection import (DeiTObjectDetector, ObjectDetectionHead, ViTObjectDetector,  # noqa: F401
                        post_process_predictions, weighted_cross_entropy)

from .pipeline import HostBatchRunner  # noqa: E402,F401
from .trainer import FineTuner, TrainState  # noqa: E402,F401

__all__ = [
    "HostBatchRunner", "FineTuner", "TrainState",
    "PatchEmbedding", "MultiHeadSelfAttention", "MLPBlock", "TransformerBlock",
    "VisionTransformer", "DataEfficientImageTransformer", "ViTClassifier", "VitkError",
    "ObjectDetectionHead", "ViTObjectDetector", "DeiTObjectDetector", "post_process_predictions", "weighted_cross_entropy",
    "launch_count", "ops",
]
