"""drnb200 — B200-native inference path for block-pruned DRN semantic segmentation.

Host-side mirror of the reference's two interfaces for this path (``DRNSeg`` of semantic_seg.py:126-164 and
the ``pruners`` package) over the C-ABI library ``libdrnb200.so`` (include/drnb200.h).
"""
from . import ffi
from .model import DRNSeg, fill_up_weights
from . import drn
from . import pruners
from .evalops import ConfusionMeter, fast_hist, per_class_iu, shard_frames
from .engine import ingest_lut
from . import frameio
from . import checkpoint
from .checkpoint import load_checkpoint, normalize_state_dict, masks_from_zeros
from .frameio import CITYSCAPE_PALETTE, colorize, overlay, load_info, resize_frames, HostBuffer
from . import multiscale
from .pipeline import FramePipeline
from .multiscale import predict_ms, test_ms

__all__ = ["DRNSeg", "fill_up_weights", "drn", "pruners", "ffi", "ConfusionMeter", "fast_hist",
           "per_class_iu", "shard_frames", "frameio", "CITYSCAPE_PALETTE", "colorize", "overlay", "load_info", "resize_frames", "HostBuffer", "FramePipeline",
           "ingest_lut", "multiscale", "predict_ms", "test_ms", "checkpoint", "load_checkpoint", "normalize_state_dict", "masks_from_zeros"]
