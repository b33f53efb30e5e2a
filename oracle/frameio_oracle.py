"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference code either side of the hot path.
Pinned against fixtures produced by the real reference (tests/golden/gen_golden_io.py -> frameio.npz).

* ingest(): `ToTensorVideoImage` (data_transforms.py:256-281: HWC uint8 -> CHW float, `.float().div(255)`) followed
  by `Normalize` (data_transforms.py:109-125: per channel `t.sub_(m).div_(s)`, mean/std as FloatTensor), the
  transform seg_video_old.py:122-139 applies to every video frame.
* colorize(): `palettes[pred]` (semantic_seg.py:101-112 with CITYSCAPE_PALETTE :52-72; seg_video.py:168).
* overlay(): the alpha blend seg_video.py:200-203 draws (matplotlib imshow(alpha=0.6) over the frame); matplotlib is
  absent in this image, so the rule `rint(alpha*colour + (1-alpha)*frame)` in fp32 steps is this repo's definition
  (parity unpinned for the blend only; the palette lookup is pinned).
"""
import numpy as np
import torch

CITYSCAPE_PALETTE = np.asarray([
    [128, 64, 128], [244, 35, 232], [70, 70, 70], [102, 102, 156], [190, 153, 153], [153, 153, 153],
    [250, 170, 30], [220, 220, 0], [107, 142, 35], [152, 251, 152], [70, 130, 180], [220, 20, 60], [255, 0, 0],
    [0, 0, 142], [0, 0, 70], [0, 60, 100], [0, 80, 100], [0, 0, 230], [119, 11, 32], [0, 0, 0]], dtype=np.uint8)


def ingest(frames_u8_nhwc, mean, std, bgr=False):
    """uint8 [N,H,W,3] -> float32 [N,3,H,W]"""
    x = torch.from_numpy(np.ascontiguousarray(frames_u8_nhwc))
    if bgr:
        x = x.flip(-1)
    x = x.permute(0, 3, 1, 2).contiguous().float().div(255)
    m = torch.FloatTensor(list(mean))
    s = torch.FloatTensor(list(std))
    for c in range(3):
        x[:, c].sub_(m[c]).div_(s[c])
    return x


def ingest_table(mean, std):
    """the transform of every byte value: float32 [3,256]"""
    ramp = np.repeat(np.arange(256, dtype=np.uint8)[None, None, :, None], 3, axis=3)    # [1,1,256,3]
    return ingest(ramp, mean, std)[0, :, 0, :]


def colorize(pred, palette=CITYSCAPE_PALETTE):
    pred = np.asarray(pred)
    idx = np.where(pred < len(palette), pred, len(palette) - 1)
    return palette[idx]


def overlay(pred, frames_u8_nhwc, alpha=0.6, palette=CITYSCAPE_PALETTE):
    a = np.float32(alpha)
    b = np.float32(1.0) - a
    col = colorize(pred, palette).astype(np.float32)
    f = np.asarray(frames_u8_nhwc).astype(np.float32)
    return np.rint(a * col + b * f).astype(np.uint8)


# ------------------------------------------------------------------------------------------------ frame resize
# seg_video_old.py:125-128 resizes every decoded frame with torchvision `T.Resize((h, w))` on a PIL image, i.e.
# `Image.resize((w, h), BILINEAR)` — Pillow's 8-bit resampler (third-party, not in /root/reference; Pillow 12.2.0 here,
# src/libImaging/Resample.c, unchanged since 3.4): precompute_coeffs() in double (oracle/ms_oracle.bilinear_coeffs),
# then normalize_coeffs_8bpc(): k_int = (int)(k * 2^22 +- 0.5), horizontal pass over the source rows the vertical pass
# needs (ss = 2^21; ss += pixel * k_int; out = clip8(ss >> 22)) into a uint8 temporary, then the vertical pass likewise.
PRECISION_BITS = 32 - 8 - 2


def coeffs_8bpc(in_size, out_size):
    """-> (xmin int32 [out], count int32 [out], k int32 [out, ksize])"""
    from oracle import ms_oracle
    xmin, cnt, kk = ms_oracle.bilinear_coeffs(in_size, out_size)
    scaled = kk * float(1 << PRECISION_BITS)
    ki = np.where(kk < 0, np.trunc(-0.5 + scaled), np.trunc(0.5 + scaled)).astype(np.int64).astype(np.int32)
    return xmin, cnt, ki


def _pass_8bpc(a, out_size):
    """one 8-bit Pillow pass along axis -2 of a uint8 [..., L, C] array"""
    in_size = a.shape[-2]
    xmin, cnt, ki = coeffs_8bpc(in_size, out_size)
    ss = np.full(a.shape[:-2] + (out_size, a.shape[-1]), 1 << (PRECISION_BITS - 1), np.int64)
    for t in range(ki.shape[1]):
        live = (t < cnt)[:, None]
        idx = np.minimum(xmin + t, in_size - 1)
        ss = np.where(live, ss + a[..., idx, :].astype(np.int64) * ki[:, t][:, None], ss)
    return np.clip(ss >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resize_u8(frames, height, width):
    """uint8 [N,H,W,3] -> uint8 [N,height,width,3], = T.Resize((height, width)) on each PIL frame"""
    a = np.asarray(frames, np.uint8)
    if a.shape[2] != width:
        a = _pass_8bpc(a, width)                                   # horizontal first (axis -2 is W)
    if a.shape[1] != height:
        a = np.swapaxes(_pass_8bpc(np.swapaxes(a, 1, 2), height), 1, 2)
    return np.ascontiguousarray(a)
