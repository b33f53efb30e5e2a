"""Golden fixture for the frame resize of the video caller (seg_video_old.py:125-128): a crop of a real frame of the
reference's sample.mp4 through torchvision `T.Resize(size)` on the PIL image, exactly as FrameCapture does it, for
down-, up- and single-axis scaling.  Pillow is a third-party dependency of the reference (version here: see `pillow`
in the file); its 8-bit resampler is restated in oracle/frameio_oracle.py and on the device (csrc/frameio.cu).

Run in the build container only (needs /root/reference/sample.mp4, cv2, PIL, torchvision):
    python tests/golden/gen_golden_resize.py
"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SIZES = [(75, 75), (96, 160), (200, 300), (160, 100), (41, 284), (300, 300)]


def main():
    import cv2
    import PIL
    import torchvision.transforms as T
    from PIL import Image
    cap = cv2.VideoCapture("/root/reference/sample.mp4")
    ok, frame = cap.read()
    assert ok
    crop = frame[100:260, 300:584].copy()                       # 160 x 284 x 3, as cv2 delivers it
    out = {"src": crop, "sizes": np.asarray(SIZES), "pillow": PIL.__version__}
    for i, size in enumerate(SIZES):
        out["dst%d" % i] = np.asarray(T.Resize(size)(Image.fromarray(crop, "RGB")))    # seg_video_old.py:125-128
    np.savez_compressed(os.path.join(HERE, "frame_resize.npz"), **out)
    print("wrote frame_resize.npz", os.path.getsize(os.path.join(HERE, "frame_resize.npz")), "bytes, Pillow", PIL.__version__)


if __name__ == "__main__":
    main()
