"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's multi-scale test.
Pinned against fixtures produced by the real reference (tests/golden/gen_golden_ms.py -> multiscale.npz).

* resize_4d_tensor() — semantic_seg.py:471-504: every [i, j] plane of a float32 NCHW tensor goes through
  `Image.fromarray(plane).resize((width, height), Image.BILINEAR)` (mode "F"); a tensor that already has the target
  size is returned untouched.  The arithmetic lives in a third-party dependency that is NOT in /root/reference:
  **Pillow** (the reference pins no version; the build container has Pillow 12.2.0, whose resampler is unchanged
  since 3.4).  Its published algorithm (src/libImaging/Resample.c), restated here:
    - precompute_coeffs(): scale = in/out, filterscale = max(scale, 1), support = 1.0 * filterscale (triangle filter
      `1 - |x|` on |x| < 1), for every output index xx: center = (xx + 0.5) * scale,
      xmin = max(0, int(center - support + 0.5)), xmax = min(in, int(center + support + 0.5)) - xmin,
      k[x] = filter((x + xmin - center + 0.5) / filterscale), normalised by their sum; all in double.
    - ImagingResampleHorizontal_32bpc / Vertical_32bpc: `ss = 0.0; ss += pixel * k[x]` in double in tap order,
      result stored as float32.  The horizontal pass runs first (into a temporary float32 image), then the vertical
      pass; a pass whose size does not change is skipped.
* test_ms() loop body — semantic_seg.py:531-541: `final = sum([resize_4d_tensor(out, w, h) for out in outputs])`
  (Python's sum: ((0 + r0) + r1) + ... in float32), `pred = final.argmax(axis=1)` (first maximum).
"""
import math

import numpy as np


def bilinear_coeffs(in_size, out_size):
    """Pillow precompute_coeffs() for the BILINEAR filter over the full axis.
    -> (xmin int32 [out], count int32 [out], k float64 [out, ksize])"""
    scale = in_size / out_size
    filterscale = scale if scale >= 1.0 else 1.0
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xmin = np.zeros(out_size, np.int32)
    cnt = np.zeros(out_size, np.int32)
    kk = np.zeros((out_size, ksize), np.float64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        lo = int(center - support + 0.5)
        if lo < 0:
            lo = 0
        hi = int(center + support + 0.5)
        if hi > in_size:
            hi = in_size
        n = hi - lo
        ww = 0.0
        for x in range(n):
            a = abs((x + lo - center + 0.5) * ss)
            w = 1.0 - a if a < 1.0 else 0.0
            kk[xx, x] = w
            ww += w
        if ww != 0.0:
            for x in range(n):
                kk[xx, x] /= ww
        xmin[xx], cnt[xx] = lo, n
    return xmin, cnt, kk


def _resample_last_axis(a, out_size):
    """one Pillow pass along the last axis of a float32 array: double accumulation in tap order, float32 result"""
    in_size = a.shape[-1]
    xmin, cnt, kk = bilinear_coeffs(in_size, out_size)
    ss = np.zeros(a.shape[:-1] + (out_size,), np.float64)
    for t in range(kk.shape[1]):
        live = t < cnt                                           # taps beyond xmax are not visited by Pillow
        idx = np.minimum(xmin + t, in_size - 1)
        term = a[..., idx].astype(np.float64) * kk[:, t]
        ss = np.where(live, ss + term, ss)
    return ss.astype(np.float32)


def resize_bilinear_f32(a, width, height):
    """`Image.fromarray(plane).resize((width, height), Image.BILINEAR)` on every [..., H, W] float32 plane"""
    a = np.asarray(a, np.float32)
    if a.shape[-1] != width:
        a = _resample_last_axis(a, width)                         # horizontal pass first
    if a.shape[-2] != height:
        a = np.swapaxes(_resample_last_axis(np.swapaxes(a, -1, -2), height), -1, -2)
    return np.ascontiguousarray(a)


def resize_4d_tensor(t, width, height):
    """semantic_seg.py:471-504"""
    t = np.asarray(t, np.float32)
    if t.shape[2] == height and t.shape[3] == width:
        return t
    return resize_bilinear_f32(t, width, height)


def ms_combine(outputs, width, height):
    """semantic_seg.py:540-541 -> (summed float32 [N,C,H,W], pred int64 [N,H,W])"""
    final = sum([resize_4d_tensor(o, width, height) for o in outputs])
    return final, final.argmax(axis=1)
