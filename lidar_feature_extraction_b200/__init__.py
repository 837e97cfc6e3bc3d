"""lidar_feature_extraction_b200 — B200-native (sm_100a) implementation of the extraction hot path of
tier4/lidar_feature_extraction behind a C ABI (include/lfx.h).

Python here is host-side glue only: a ctypes loader (`_native`), a mirror of the reference node's
interface (`extraction.FeatureExtraction`) and the synthetic scan generator bindings (`synth`).
The compute path is hand-written CUDA in ``csrc/``; there is no CPU fallback.
"""
from .extraction import (  # noqa: F401
    POINT_STEP,
    ExtractionError,
    FeatureExtraction,
    HyperParameters,
    PipelinedExtraction,
    PointCloud2,
    PointField,
    default_params,
    label_to_color,
    launch_yaml_params,
)
from .converter import ConvertError, PointTypeConverter  # noqa: F401
from .mapping import MapBuilder, make_pose, pose_diff_is_sufficiently_small  # noqa: F401
from .localization import LoamProblem  # noqa: F401
from . import synth  # noqa: F401

__all__ = [
    "FeatureExtraction", "PipelinedExtraction", "HyperParameters", "PointCloud2", "PointField", "ExtractionError", "default_params",
    "launch_yaml_params", "label_to_color", "synth", "POINT_STEP", "PointTypeConverter", "ConvertError", "MapBuilder", "make_pose",
    "pose_diff_is_sufficiently_small", "LoamProblem",
]
