#!/usr/bin/env python
"""Times the multi-scale kernels at BASELINE's frame size: one 1024x2048 frame, 19 classes, the
reference's scales.  Prints one JSON object: per-scale drnb200_ms_accumulate time and algorithmic GB/s
(source plane read once + accumulator read (not on the first scale) + written), drnb200_ms_argmax likewise."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-seg-model-compress_b200"))
from drnb200 import multiscale  # noqa: E402


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(iters):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / iters


def main():
    dev = torch.device("cuda:0")
    H, W, C = 1024, 2048, 19
    acc = torch.zeros(1, C, H, W, device=dev)
    out = {"frame": [H, W], "classes": C, "scales": []}
    total = 0.0
    for i, s in enumerate([1.0] + multiscale.SCALES):
        hs, ws = int(H * s), int(W * s)
        src = torch.log_softmax(torch.randn(1, C, hs, ws, device=dev), dim=1)
        first = i == 0
        ms = timed(lambda: multiscale.resize_accumulate(src, acc, first=first))
        nbytes = 4 * C * (hs * ws + H * W * (1 if first else 2))
        out["scales"].append({"scale": s, "ms": ms, "gbs": nbytes / ms / 1e6})
        total += ms
        del src
    ms = timed(lambda: multiscale.argmax_labels(acc))
    out["argmax"] = {"ms": ms, "gbs": (4 * C + 1) * H * W / ms / 1e6}
    out["combine_ms_per_frame"] = total + ms
    if "--model" in sys.argv:
        # whole multi-scale prediction of one frame (semantic_seg.test_ms loop body): six forward() calls of the pruned
        # DRN-D-22 (log-probs materialised per scale, as the reference does) + resize/sum/argmax, all on the device
        import types
        import torch.nn.functional as F
        sys.path.insert(0, ROOT)
        import bench
        args = types.SimpleNamespace(arch="drn_d_22", act="fp16", sparsity=0.75)
        model, _, _ = bench.build_model(args, dev)
        x = torch.randn(1, 3, H, W, device=dev)
        images = [x] + [F.interpolate(x, size=(int(H * s), int(W * s)), mode="bilinear") for s in multiscale.SCALES]
        single = timed(lambda: model.predict(x), iters=5)
        full = timed(lambda: multiscale.predict_ms(model, images), iters=5)
        fwd = timed(lambda: [model(im)[0] for im in images], iters=3)
        out["predict_ms"] = {"ms_per_frame": full, "six_forward_calls_ms": fwd, "single_scale_predict_ms": single,
                             "pixels_vs_single_scale": sum(s * s for s in [1.0] + multiscale.SCALES)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
