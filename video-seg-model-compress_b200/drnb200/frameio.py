"""Steps either side of the hot path in the reference's video caller, on the device:

* ingest: uint8 HWC frames go straight into `DRNSeg.predict` / `forward` after `DRNSeg.set_ingest(mean, std)`
  (the normalisation is fused into the stem kernel); `load_info()` reads the reference's `info.json`.
* `colorize(labels)` = `CITYSCAPE_PALETTE[pred]` (semantic_seg.py:52-72, :101-112), `overlay(labels, frames)` =
  the alpha=0.6 blend over the frame that seg_video.py:200-203 draws.
"""
import json

import numpy as np
import torch

from . import ffi

# the Cityscapes train-id colours (semantic_seg.py:52-72); row 19 (black) is used for the ignore label
CITYSCAPE_PALETTE = np.asarray([
    [128, 64, 128], [244, 35, 232], [70, 70, 70], [102, 102, 156], [190, 153, 153], [153, 153, 153],
    [250, 170, 30], [220, 220, 0], [107, 142, 35], [152, 251, 152], [70, 130, 180], [220, 20, 60], [255, 0, 0],
    [0, 0, 142], [0, 0, 70], [0, 60, 100], [0, 80, 100], [0, 0, 230], [119, 11, 32], [0, 0, 0]], dtype=np.uint8)


def load_info(path):
    """(mean, std) from the reference's info.json (`{"mean": [...], "std": [...]}`)"""
    with open(path) as fh:
        info = json.load(fh)
    return tuple(info["mean"]), tuple(info["std"])


_pal_cache = {}


def _device_palette(palette, device):
    pal = np.ascontiguousarray(np.asarray(palette, dtype=np.uint8))
    if pal.ndim != 2 or pal.shape[1] != 3 or not (1 <= pal.shape[0] <= 256):
        raise ffi.Drnb200Error("palette must be uint8 [n_colors<=256, 3] (got %s)" % (pal.shape,))
    key = (pal.tobytes(), str(device))
    t = _pal_cache.get(key)
    if t is None:
        t = _pal_cache[key] = torch.from_numpy(pal).to(device)
    return t


def _colorize(labels, frames, alpha, palette):
    if not (labels.is_cuda and labels.dtype == torch.uint8):
        raise ffi.Drnb200Error("labels must be a uint8 CUDA tensor (got %s on %s); there is no CPU path"
                               % (labels.dtype, labels.device))
    labels = labels.contiguous()
    if labels.numel() % 4:
        raise ffi.Drnb200Error("the pixel count must be a multiple of 4")
    if frames is not None:
        if not (frames.is_cuda and frames.dtype == torch.uint8 and tuple(frames.shape) == tuple(labels.shape) + (3,)):
            raise ffi.Drnb200Error("frames must be a uint8 CUDA tensor of shape labels.shape + (3,)")
        frames = frames.contiguous()
    pal = _device_palette(palette, labels.device)
    out = torch.empty(tuple(labels.shape) + (3,), dtype=torch.uint8, device=labels.device)
    if labels.numel() == 0:
        return out
    ffi.check(ffi.lib().drnb200_colorize(ffi.ptr(labels), labels.numel(), ffi.ptr(pal), pal.shape[0],
                                         ffi.ptr(frames), float(alpha), ffi.ptr(out), ffi.stream_ptr()), "colorize")
    return out


def colorize(labels, palette=CITYSCAPE_PALETTE):
    """uint8 labels [..., H, W] -> uint8 RGB [..., H, W, 3]"""
    return _colorize(labels, None, 0.0, palette)


def overlay(labels, frames, alpha=0.6, palette=CITYSCAPE_PALETTE):
    """rint(alpha * palette[labels] + (1 - alpha) * frames), uint8 RGB"""
    return _colorize(labels, frames, alpha, palette)


_resize_tables = {}


def _resize_axis_tables(in_size, out_size, device):
    """Pillow's 8-bit coefficient tables of one axis on the device (normalize_coeffs_8bpc), or None if the size is kept"""
    if in_size == out_size:
        return None
    key = (in_size, out_size, str(device))
    t = _resize_tables.get(key)
    if t is None:
        from .multiscale import bilinear_coeffs
        lo, cnt, kk = bilinear_coeffs(in_size, out_size)
        scaled = kk * float(1 << 22)
        ki = np.where(kk < 0, np.trunc(-0.5 + scaled), np.trunc(0.5 + scaled)).astype(np.int64).astype(np.int32)
        t = _resize_tables[key] = (torch.from_numpy(lo).to(device), torch.from_numpy(cnt).to(device),
                                   torch.from_numpy(np.ascontiguousarray(ki)).to(device), ki.shape[1])
    return t


def resize_frames(frames, size, out=None):
    """uint8 CUDA frames [N,Hs,Ws,3] -> uint8 [N,h,w,3], bit-identical to `T.Resize((h, w))` on every PIL frame
    (seg_video_old.py:125-128: Pillow's 8-bit BILINEAR resampler with its widened support when downscaling).
    `out`: optional destination (contiguous uint8 CUDA tensor [N,h,w,3] on the same device)."""
    if not (isinstance(frames, torch.Tensor) and frames.is_cuda and frames.dtype == torch.uint8 and frames.dim() == 4
            and frames.shape[3] == 3):
        raise ffi.Drnb200Error("resize_frames needs a uint8 CUDA tensor [N,H,W,3]; there is no CPU path")
    h, w = int(size[0]), int(size[1])
    frames = frames.contiguous()
    N, Hs, Ws, _ = frames.shape
    dev = frames.device
    with torch.cuda.device(dev):
        if out is None:
            out = torch.empty((N, h, w, 3), dtype=torch.uint8, device=dev)
        elif not (out.is_cuda and out.device == dev and out.dtype == torch.uint8 and out.is_contiguous()
                  and tuple(out.shape) == (N, h, w, 3)):
            raise ffi.Drnb200Error("resize_frames: `out` must be a contiguous uint8 tensor %s on %s" % ((N, h, w, 3), dev))
        tx, ty = _resize_axis_tables(Ws, w, dev), _resize_axis_tables(Hs, h, dev)
        tmp = torch.empty((N, Hs, w, 3), dtype=torch.uint8, device=dev) if tx is not None and ty is not None else None
        nul = (None, None, None, 0)
        xlo, xcnt, xk, kx = tx if tx is not None else nul
        ylo, ycnt, yk, ky = ty if ty is not None else nul
        ffi.check(ffi.lib().drnb200_resize_u8(ffi.ptr(frames), N, Hs, Ws, ffi.ptr(out), h, w, ffi.ptr(xlo), ffi.ptr(xcnt),
                                              ffi.ptr(xk), kx, ffi.ptr(ylo), ffi.ptr(ycnt), ffi.ptr(yk), ky, ffi.ptr(tmp),
                                              ffi.stream_ptr()), "resize_u8")
    return out


class HostBuffer:
    """Pinned host staging buffer for frames (H2D) or label maps (D2H), allocated by the library
    (`drnb200_host_alloc`): ``mode`` = "pinned" (cudaHostAlloc, what ``Tensor.pin_memory()`` gives), "wc"
    (write-combined: DMA reads skip CPU-cache snooping; never read it from the CPU — frames only) or "huge"
    (2 MiB-aligned MADV_HUGEPAGE mapping registered with CUDA: fewer IOMMU entries per copy).
    ``.tensor`` is a torch view of the memory; it must not outlive the buffer (``close()`` / garbage collection)."""

    MODES = {"pinned": ffi.HOST_PINNED, "wc": ffi.HOST_WC, "huge": ffi.HOST_HUGE}

    def __init__(self, shape, dtype=torch.uint8, mode="pinned"):
        import ctypes as C
        if mode not in self.MODES:
            raise ffi.Drnb200Error("HostBuffer mode must be one of %s" % sorted(self.MODES))
        self.mode = mode
        nbytes = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        if nbytes <= 0:
            raise ffi.Drnb200Error("HostBuffer needs a non-empty shape")
        p = C.c_void_p()
        ffi.check(ffi.lib().drnb200_host_alloc(C.byref(p), nbytes, self.MODES[mode]), "host_alloc(%s)" % mode)
        self.ptr, self.nbytes = p.value, nbytes
        raw = (C.c_uint8 * nbytes).from_address(self.ptr)
        self.tensor = torch.frombuffer(raw, dtype=torch.uint8).view(dtype).view(*shape)

    def close(self):
        if getattr(self, "ptr", None):
            self.tensor = None
            ffi.lib().drnb200_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
