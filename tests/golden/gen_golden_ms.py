"""Golden fixtures for the multi-scale test, produced by the REAL reference: `resize_4d_tensor`
(semantic_seg.py:471-504, PIL BILINEAR on float planes) and the sum/argmax of `test_ms` (semantic_seg.py:540-541).

Run in the build container only:   python tests/golden/gen_golden_ms.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from gen_golden import import_reference  # noqa: E402

# (source H, W) -> target (H, W): the reference's scales 0.5/0.75/1.25/1.5/1.75 of a 32x56 frame, plus ragged ratios
TARGET = (32, 56)
SOURCES = [(32, 56), (16, 28), (24, 42), (40, 70), (48, 84), (56, 98), (23, 37), (77, 131)]


def make_sources(seed=11, n=1, c=4):
    """the log-prob-like source tensors (regenerated from the seed by the tests; not stored)"""
    g = torch.Generator().manual_seed(seed)
    return [torch.log_softmax(3.0 * torch.randn(n, c, h, w, generator=g), dim=1) for h, w in SOURCES]


def main():
    S = import_reference()
    import PIL
    out = {"target": np.asarray(TARGET), "sources": np.asarray(SOURCES), "pillow": np.asarray(PIL.__version__)}
    outputs = make_sources()
    for i, t in enumerate(outputs):
        out["dst%d" % i] = np.asarray(S.resize_4d_tensor(t, TARGET[1], TARGET[0]), np.float32)
    final = sum([S.resize_4d_tensor(o, TARGET[1], TARGET[0]) for o in outputs])      # semantic_seg.py:540
    out["final"] = final
    out["pred"] = final.argmax(axis=1)                                                # semantic_seg.py:543
    np.savez_compressed(os.path.join(HERE, "multiscale.npz"), **out)
    print("wrote multiscale.npz", final.shape, os.path.getsize(os.path.join(HERE, "multiscale.npz")))


if __name__ == "__main__":
    main()
