"""viddet_b200 -- B200-native (sm_100a) implementation of VidDet's detection-head hot path:
temporal tip conv -> 1x1 prediction conv -> YOLOOutputV3 decode -> per-class box_nms / top-k, plus
the YOLOV3PrefetchTargetGenerator anchor-IoU matching.  Python here is a thin ctypes layer over
the C ABI in include/viddet_b200.h; torch tensors are carriers only.  There is no CPU fallback.
"""
from ._lib import VidDetError, load, SO_PATH  # noqa: F401
from .blocks import (  # noqa: F401
    ClassTree, ConvBNLReLU, DEFAULT_ANCHORS, DEFAULT_CHANNELS, DEFAULT_STRIDES, HeadPipeline, HeadSession, NeckSession, TemporalPooling, TemporalTipConv, TimeDistributed,
    YOLODetectionBlockV3, YOLOOutputV3, YOLOV3DynamicTargetGeneratorSimple, YOLOV3Head, YOLOV3Loss, YOLOV3Neck, YOLOV3PrefetchTargetGenerator, YOLOV3TargetMerger,
    ClipWindows, SplitF32, materialise_windows, window_frame_indices, box_nms, hierarchical_nms, postprocess_detections, to_nhwc_bf16, to_nhwc_split, upsample_concat,
)

from .io import FeatureStream, feature_paths  # noqa: F401

__version__ = "0.1.0"
