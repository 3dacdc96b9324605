"""6d-pose-estimation_b200 -- B200-native pose-geometry hot path.

The directory name is not a Python identifier; import it with
``importlib.import_module("6d-pose-estimation_b200")`` or, for a drop-in replacement of
the reference's modules, put this directory at the front of ``sys.path`` and keep the
reference's own imports (``from models.add_loss import ADDLoss`` ...).
"""
from . import _lib as core  # noqa: F401  (registers p6d_b200_core)
from .models.add_loss import ADDLoss, SYMMETRIC_OBJECT_IDS  # noqa: F401
from .models.pose_loss import PoseLoss  # noqa: F401
from .utils.camera import (DEFAULT_K, depth_backproject, depth_crop_backproject, detection_backproject,  # noqa: F401
                           get_gt_and_K, pinhole_translation)
from . import workloads  # noqa: F401
from . import sweep  # noqa: F401
from .sweep import PoseEvaluator, evaluate_sweep, reference_batch_means, shard_range  # noqa: F401

__all__ = ["ADDLoss", "PoseLoss", "SYMMETRIC_OBJECT_IDS", "DEFAULT_K", "get_gt_and_K",
           "pinhole_translation", "depth_backproject", "depth_crop_backproject", "detection_backproject", "core", "workloads", "sweep", "PoseEvaluator",
           "evaluate_sweep", "reference_batch_means", "shard_range"]
