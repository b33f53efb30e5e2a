"""Evaluation glue of the callers: confusion matrix, mIoU and the data-parallel sharding of frames.

Reference: ``fast_hist`` / ``per_class_iu`` (semantic_seg.py:293-300), accumulation and rounding in
``test()`` / ``val_miou()`` (semantic_seg.py:435, :455, :466-468).  The confusion matrix is accumulated
on the device by libdrnb200 (int64 counts); with several ranks (one process per GPU, frames sharded on
dim 0 exactly like nn.DataParallel does, semantic_seg.py:811-812) the only collective on the path is one
all-reduce(sum) of the classes x classes int64 matrix over NCCL.
"""
import numpy as np
import torch

from . import ffi


def shard_frames(n_frames, rank, world_size):
    """contiguous slice [lo, hi) of a stream of `n_frames` frames owned by `rank` (balanced to +-1)"""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    base, extra = divmod(n_frames, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def per_class_iu(hist):
    """IoU per class: diag / (rowsum + colsum - diag)   (semantic_seg.py:299-300)"""
    hist = np.asarray(hist, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.diag(hist) / (hist.sum(1) + hist.sum(0) - np.diag(hist))


def fast_hist(pred, label, n):
    """device version of semantic_seg.py:293-296 for CUDA tensors: returns an int64 [n, n] CUDA tensor"""
    meter = ConfusionMeter(n, pred.device)
    meter.update(pred, label)
    return meter.hist.clone()


class ConfusionMeter:
    """running confusion matrix on the device; ``hist[label, pred]`` like the reference"""

    def __init__(self, classes, device):
        self.classes = int(classes)
        self.hist = torch.zeros((self.classes, self.classes), dtype=torch.int64, device=device)

    def update(self, pred, label):
        """pred: uint8 CUDA tensor (DRNSeg.predict); label: uint8 or int64 CUDA tensor of the same numel"""
        if not (pred.is_cuda and label.is_cuda):
            raise ffi.Drnb200Error("ConfusionMeter.update needs CUDA tensors (no CPU path)")
        if pred.dtype != torch.uint8:
            raise ffi.Drnb200Error("pred must be uint8 (got %s)" % pred.dtype)
        if label.dtype not in (torch.uint8, torch.int64):
            raise ffi.Drnb200Error("label must be uint8 or int64 (got %s)" % label.dtype)
        if pred.numel() != label.numel():
            raise ffi.Drnb200Error("pred and label differ in size")
        if pred.numel() == 0:
            return
        pred, label = pred.contiguous(), label.contiguous()
        ffi.check(ffi.lib().drnb200_confusion(ffi.ptr(pred), ffi.ptr(label), int(label.dtype == torch.int64),
                                              pred.numel(), self.classes, ffi.ptr(self.hist),
                                              ffi.stream_ptr()), "confusion")

    def all_reduce(self, group=None):
        """sum the matrix over all ranks (NCCL over NVLink on the GPU box; gloo in the CPU tests)"""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.hist, op=dist.ReduceOp.SUM, group=group)
        return self.hist

    def ious(self):
        return per_class_iu(self.hist.cpu().numpy()) * 100

    def miou(self):
        """``round(np.nanmean(per_class_iu(hist) * 100), 2)``   (semantic_seg.py:466-468)"""
        return round(float(np.nanmean(self.ious())), 2)
