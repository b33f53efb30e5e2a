"""Golden fixture of REAL video frames (SURVEY 8d "Inputs"): three frames of the reference's sample.mp4, decoded,
downscaled and normalised exactly as FrameCapture does (seg_video_old.py:110-139: cv2.VideoCapture.read() ->
Image.fromarray(image, 'RGB') -> T.Resize -> Compose([ToTensorVideoImage(), Normalize(info.json)])), by the REAL
reference's data_transforms module, plus the labels / low-res logits the REAL reference DRNSeg (semantic_seg.py)
produces on them with the seeded block-pruned DRN-D-22 weights of the parity tests.

Run in the build container only (needs /root/reference and cv2):   python tests/golden/gen_golden_frames.py
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
from gen_golden import import_reference, REF  # noqa: E402

SIZE = (256, 448)            # (H, W): 1138x640 source, aspect kept to within a pixel; W % 16 == 0 for the uint8 ingest
FRAMES = (0, 60, 150)
SEED = 71


def main():
    S = import_reference()
    import cv2
    import data_transforms as T                       # reference module
    import torchvision.transforms as TV
    from PIL import Image
    from helpers import gate_case_cpu
    info = json.load(open(os.path.join(REF, "info.json")))
    norm = T.Normalize(mean=info["mean"], std=info["std"])
    tt = T.ToTensorVideoImage()
    cap = cv2.VideoCapture(os.path.join(REF, "sample.mp4"))
    u8, xs, idx = [], [], 0
    while True:
        ok, image = cap.read()
        if not ok:
            break
        if idx in FRAMES:
            img = TV.Resize(SIZE)(Image.fromarray(image, "RGB"))       # seg_video_old.py:125-128
            u8.append(np.asarray(img).copy())
            xs.append(norm(tt(img))[0])
        idx += 1
    assert len(u8) == len(FRAMES), (len(u8), idx)
    u8 = np.stack(u8)
    x = torch.stack(xs)                                               # [3,3,H,W] float32, the reference's own transform
    # the real reference's DRNSeg on these frames, seeded block-pruned weights (same recipe as the GPU parity tests)
    _, sd, masks = gate_case_cpu("drn_d_22", True, SEED)
    ref = S.DRNSeg("drn_d_22", 19, pretrained_model=None, pretrained=False)
    missing = ref.load_state_dict(sd, strict=False)
    assert set(missing.missing_keys) <= {"up.weight"}, missing
    ref.eval()
    with torch.no_grad():
        final, seg = ref(x)
        _, pred = torch.max(final, 1)                                 # semantic_seg.py:444-445
    top2 = final.topk(2, dim=1)[0]
    np.savez_compressed(os.path.join(HERE, "real_frames.npz"), frames_u8=u8, frame_index=np.asarray(FRAMES),
                        x0=x[0].numpy(), labels=pred.numpy().astype(np.uint8), seg=seg.numpy().astype(np.float32),
                        margin_median=float((top2[:, 0] - top2[:, 1]).median()), seed=SEED,
                        mean=np.asarray(info["mean"], np.float64), std=np.asarray(info["std"], np.float64))
    print("wrote real_frames.npz", u8.shape, x.shape, "classes", len(pred.unique()),
          "bytes", os.path.getsize(os.path.join(HERE, "real_frames.npz")))


if __name__ == "__main__":
    main()
