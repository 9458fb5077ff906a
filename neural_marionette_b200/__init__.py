"""neural_marionette_b200 — B200-native (sm_100a) implementation of Neural Marionette's volumetric
keypoint-detection hot path behind the reference's own module API.

    from neural_marionette_b200 import NeuralMarionette, KyptDetector, HSVRNNBVH, voxelize, FusedAdam

The sub-package layout mirrors the reference (`model/`, `modules/`, `utils/`) so that
`from model.neural_marionette import NeuralMarionette` becomes
`from neural_marionette_b200.model.neural_marionette import NeuralMarionette`.
"""
from .model.hsvrnn_bvh import HSVRNNBVH
from .model.kypt_detector import KyptDetector, KyptToVoxNet, VoxToKyptNet
from .model.neural_marionette import NeuralMarionette
from .optim import FusedAdam
from .utils.dataset_utils import (crop_sequence, episodic_normalization, voxelize, voxelize_clip,
                                  voxelize_raw_clips)

__all__ = ["NeuralMarionette", "KyptDetector", "VoxToKyptNet", "KyptToVoxNet", "HSVRNNBVH", "FusedAdam", "voxelize",
           "voxelize_clip", "voxelize_raw_clips", "episodic_normalization", "crop_sequence"]
