"""Multi-scale test on the device: mirror of the reference's `resize_4d_tensor` / `test_ms`
(semantic_seg.py:471-504, :507-557).

The reference runs the model on the frame and on five rescaled copies (`scales = [0.5, 0.75, 1.25, 1.5, 1.75]`,
semantic_seg.py:578; the copies come from the data loader, cityscapes_dataset.py:107-118), copies every log-prob
tensor to the host, resamples each of its planes to the frame size with PIL BILINEAR on CPU threads, sums them and
takes the argmax.  Here the log-probs stay on the device: `drnb200_ms_accumulate` resamples and sums (Pillow's exact
arithmetic: double accumulation per pass, float32 between passes), `drnb200_ms_argmax` produces the label map.
No CPU path: tensors must be CUDA tensors.
"""
import math
import os

import numpy as np
import torch

from . import ffi
from .evalops import ConfusionMeter

SCALES = [0.5, 0.75, 1.25, 1.5, 1.75]          # semantic_seg.py:578


def bilinear_coeffs(in_size, out_size):
    """Pillow's precompute_coeffs() (src/libImaging/Resample.c) for the BILINEAR (triangle, support 1) filter
    over a whole axis, in float64 like the C code: (xmin int32 [out], count int32 [out], k float64 [out, ksize])."""
    scale = float(in_size) / float(out_size)
    filterscale = max(scale, 1.0)
    support = filterscale                                        # filter support 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    center = (np.arange(out_size, dtype=np.float64) + 0.5) * scale
    lo = np.maximum(np.trunc(center - support + 0.5).astype(np.int64), 0)      # C (int) cast truncates
    hi = np.minimum(np.trunc(center + support + 0.5).astype(np.int64), in_size)
    cnt = hi - lo
    inv = 1.0 / filterscale
    kk = np.zeros((out_size, ksize), np.float64)
    ww = np.zeros(out_size, np.float64)
    for t in range(ksize):                                       # taps in order: ww is a sequential double sum
        a = np.abs((t + lo - center + 0.5) * inv)
        w = np.where((a < 1.0) & (t < cnt), 1.0 - a, 0.0)
        kk[:, t] = w
        ww = ww + w
    nz = ww != 0.0
    kk[nz] = kk[nz] / ww[nz, None]
    return lo.astype(np.int32), cnt.astype(np.int32), kk


_tables = {}


def _axis_tables(in_size, out_size, device):
    """device copies of the coefficient tables of one axis, or None when the axis keeps its size"""
    if in_size == out_size:
        return None
    key = (in_size, out_size, str(device))
    t = _tables.get(key)
    if t is None:
        lo, cnt, kk = bilinear_coeffs(in_size, out_size)
        t = _tables[key] = (torch.from_numpy(lo).to(device), torch.from_numpy(cnt).to(device),
                            torch.from_numpy(np.ascontiguousarray(kk)).to(device), kk.shape[1])
    return t


def resize_accumulate(src, acc, first):
    """acc[N,C,H,W] = (0 if first else acc) + resize_4d_tensor(src[N,C,Hs,Ws], W, H)   (semantic_seg.py:471-504, :540)"""
    for name, t in (("src", src), ("acc", acc)):
        if not (t.is_cuda and t.dtype == torch.float32 and t.dim() == 4 and t.is_contiguous()):
            raise ffi.Drnb200Error("%s must be a contiguous float32 NCHW CUDA tensor; there is no CPU path" % name)
    if src.shape[:2] != acc.shape[:2] or src.device != acc.device:
        raise ffi.Drnb200Error("src %s and acc %s disagree in N, C or device" % (tuple(src.shape), tuple(acc.shape)))
    N, Cc, Hs, Ws = src.shape
    H, W = acc.shape[2:]
    tx = _axis_tables(Ws, W, src.device)
    ty = _axis_tables(Hs, H, src.device)
    xa = (ffi.ptr(tx[0]), ffi.ptr(tx[1]), ffi.ptr(tx[2]), tx[3]) if tx else (None, None, None, 0)
    ya = (ffi.ptr(ty[0]), ffi.ptr(ty[1]), ffi.ptr(ty[2]), ty[3]) if ty else (None, None, None, 0)
    ffi.check(ffi.lib().drnb200_ms_accumulate(ffi.ptr(src), N, Cc, Hs, Ws, ffi.ptr(acc), H, W, *xa, *ya,
                                              1 if first else 0, ffi.stream_ptr()), "ms_accumulate")
    return acc


def argmax_labels(acc):
    """`final.argmax(axis=1)` (semantic_seg.py:543) -> uint8 [N,H,W]"""
    if not (acc.is_cuda and acc.dtype == torch.float32 and acc.dim() == 4 and acc.is_contiguous()):
        raise ffi.Drnb200Error("acc must be a contiguous float32 NCHW CUDA tensor; there is no CPU path")
    N, Cc, H, W = acc.shape
    labels = torch.empty((N, H, W), dtype=torch.uint8, device=acc.device)
    ffi.check(ffi.lib().drnb200_ms_argmax(ffi.ptr(acc), N, Cc, H, W, ffi.ptr(labels), ffi.stream_ptr()), "ms_argmax")
    return labels


def combine(outputs, height, width):
    """sum of the resized log-prob tensors and its argmax: (final float32 [N,C,H,W], pred uint8 [N,H,W])"""
    if not outputs:
        raise ffi.Drnb200Error("combine: no outputs")
    n, c = outputs[0].shape[:2]
    final = torch.empty((n, c, height, width), dtype=torch.float32, device=outputs[0].device)
    for i, out in enumerate(outputs):
        resize_accumulate(out.contiguous(), final, first=(i == 0))
    return final, argmax_labels(final)


def predict_ms(model, images):
    """the body of test_ms's loop (semantic_seg.py:531-543): `images[0]` is the frame batch at full size, the rest
    its rescaled copies.  Each scale's log-probs are folded into the accumulator as soon as they exist, so only one
    scale's [N,C,Hs,Ws] tensor is alive at a time (the reference keeps all six on the host)."""
    h, w = images[0].shape[2:4]
    final = None
    with torch.no_grad():
        for i, image in enumerate(images):
            out = model(image)[0]
            if final is None:
                final = torch.empty((out.shape[0], out.shape[1], h, w), dtype=torch.float32, device=out.device)
            resize_accumulate(out.contiguous(), final, first=(i == 0))
            del out
    return argmax_labels(final)


def test_ms(eval_data_loader, model, num_classes, scales, output_dir="pred", has_gt=True, save_vis=False,
            device=None):
    """signature and return value of semantic_seg.test_ms (semantic_seg.py:507-557): mIoU rounded to 2 decimals when
    `has_gt`.  Batches are `(image, label, name, *ms_images)` (or `(image, name, *ms_images)` without ground truth),
    as SegListMS yields them."""
    from .frameio import colorize
    model.eval()
    device = device if device is not None else next(model.parameters()).device
    meter = ConfusionMeter(num_classes, device)
    num_scales = len(scales)
    for input_data in eval_data_loader:
        name = input_data[2] if has_gt else input_data[1]
        images = [input_data[0]] + list(input_data[-num_scales:])
        pred = predict_ms(model, [im.to(device, non_blocking=True) for im in images])
        if save_vis:                                             # semantic_seg.py:84-112
            from PIL import Image
            color = colorize(pred).cpu().numpy()
            lab = pred.cpu().numpy()
            for ind in range(len(name)):
                for arr, root in ((lab[ind], output_dir), (color[ind], output_dir + "_color")):
                    fn = os.path.join(root, name[ind][:-4] + ".png")
                    os.makedirs(os.path.split(fn)[0], exist_ok=True)
                    Image.fromarray(arr).save(fn)
        if has_gt:
            meter.update(pred, input_data[1].to(device))
    if has_gt:
        return meter.miou()
