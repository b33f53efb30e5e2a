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


def _save_vis(pred, names, output_dir, num_classes):
    """save_output_images / save_colorful_images (semantic_seg.py:84-112): label PNG + palette PNG per frame"""
    import os
    from PIL import Image
    from .frameio import CITYSCAPE_PALETTE, colorize
    palette = np.asarray([[0, 0, 0], [217, 83, 79], [91, 192, 222]], np.uint8) if num_classes == 3 \
        else CITYSCAPE_PALETTE                       # TRIPLET_PALETTE (RGB part) / CITYSCAPE_PALETTE, :52-77
    color = colorize(pred, palette).cpu().numpy()
    lab = pred.cpu().numpy()
    for ind in range(len(names)):
        for arr, root in ((lab[ind], output_dir), (color[ind], output_dir + "_color")):
            fn = os.path.join(root, names[ind][:-4] + ".png")
            os.makedirs(os.path.split(fn)[0] or ".", exist_ok=True)
            Image.fromarray(arr).save(fn)


def test(eval_data_loader, model, num_classes, output_dir="pred", has_gt=True, save_vis=False, device=None,
         all_reduce=True):
    """signature and return value of semantic_seg.test (semantic_seg.py:429-468): batches `(image, label, name)`,
    `model(image)[0]` + `torch.max(final, 1)` replaced by `model.predict(image)`, the confusion matrix accumulated
    on the device and (with several ranks, each iterating its own shard of the loader) summed over NCCL before the
    mIoU `round(nanmean(per_class_iu(hist)) * 100, 2)` is taken.  Returns None without ground truth."""
    model.eval()
    device = device if device is not None else next(model.parameters()).device
    meter = ConfusionMeter(num_classes, device)
    for batch in eval_data_loader:
        image, label, name = batch[0], (batch[1] if has_gt else None), batch[-1]
        pred = model.predict(image.to(device, non_blocking=True))
        if save_vis:
            _save_vis(pred, name, output_dir, num_classes)
        if has_gt:
            meter.update(pred, label.to(device, non_blocking=True))
    if has_gt:
        if all_reduce:
            meter.all_reduce()
        return meter.miou()


def val_miou(val_loader, model, num_classes, args=None, has_gt=True, device=None, all_reduce=True):
    """semantic_seg.val_miou (semantic_seg.py:638-671): as test() for loaders yielding `(image, label)`"""
    return test(((im, lb, None) for im, lb in val_loader), model, num_classes, has_gt=has_gt, save_vis=False,
                device=device, all_reduce=all_reduce)
