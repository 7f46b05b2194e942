"""video_segment_b200 -- B200-native (sm_100a) dense video over-segmentation behind the
reference's DenseSegmentationUnit plug point.  Host side mirrors the reference interface;
all per-pixel work runs in hand-written CUDA kernels through the C ABI in include/vsb200.h."""
from ._lib import DenseOpts, FrameResult, lib  # noqa: F401

__all__ = ["DenseOpts", "FrameResult", "lib"]
