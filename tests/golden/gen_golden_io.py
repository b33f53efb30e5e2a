"""Golden fixtures for the rows either side of the hot path: frame ingest and palette output,
produced by the REAL reference (imported unmodified from /root/reference).

Run in the build container only:   python tests/golden/gen_golden_io.py
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from gen_golden import import_reference, REF  # noqa: E402


def main():
    S = import_reference()
    import data_transforms as T                       # reference module
    info = json.load(open(os.path.join(REF, "info.json")))
    rng = np.random.RandomState(7)
    frame = rng.randint(0, 256, size=(24, 32, 3), dtype=np.uint8)
    frame[0, :4] = [[0, 0, 0], [255, 255, 255], [1, 128, 254], [17, 99, 200]]
    # seg_video_old.py:122-139: Compose([ToTensorVideoImage(), Normalize(mean, std)]) on one HWC frame
    tt = T.ToTensorVideoImage()
    norm = T.Normalize(mean=info["mean"], std=info["std"])
    from PIL import Image
    x = tt(Image.fromarray(frame, "RGB"))
    x = norm(x)[0]
    # every byte value through the same transform: the table the fused stem uses must equal this bit for bit
    ramp = np.repeat(np.arange(256, dtype=np.uint8)[None, :, None], 3, axis=2)       # [1,256,3]
    lut = norm(tt(Image.fromarray(ramp, "RGB")))[0][:, 0, :]                           # [3,256] fp32
    pred = rng.randint(0, 19, size=(2, 24, 32)).astype(np.int64)
    color = np.stack([S.CITYSCAPE_PALETTE[pred[i].squeeze()] for i in range(2)])       # semantic_seg.py:108
    np.savez_compressed(os.path.join(HERE, "frameio.npz"), frame=frame, mean=np.asarray(info["mean"], np.float64),
                        std=np.asarray(info["std"], np.float64), x=x.numpy(), lut=lut.numpy(), pred=pred,
                        color=color, palette=S.CITYSCAPE_PALETTE)
    print("wrote frameio.npz", x.shape, lut.shape, color.shape)


if __name__ == "__main__":
    main()
