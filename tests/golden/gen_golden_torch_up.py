"""Golden fixture for `use_torch_up=True` (semantic_seg.py:144-145: nn.UpsamplingBilinear2d(scale_factor=8) instead of the
fixed ConvTranspose2d): the REAL reference DRNSeg on the seeded dense DRN-D-22 weights.

Run in the build container only:   python tests/golden/gen_golden_torch_up.py
"""
import collections
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from gen_golden import import_reference  # noqa: E402
from oracle import recipe  # noqa: E402

SEED, HW = 17, (40, 72)


def main():
    S = import_reference()
    model = S.DRNSeg("drn_d_22", 19, pretrained_model=None, pretrained=False, use_torch_up=True).eval()
    shapes = collections.OrderedDict((k, tuple(v.shape)) for k, v in model.state_dict().items())
    sd = recipe.make_state_dict(shapes, seed=SEED)
    missing = model.load_state_dict(sd, strict=False)
    assert not missing.missing_keys and not missing.unexpected_keys, missing
    x = recipe.make_frames(1, HW[0], HW[1], seed=1234 + SEED)
    with torch.no_grad():
        final, seg = model(x)
        _, pred = torch.max(final, 1)
    np.savez_compressed(os.path.join(HERE, "fwd_drn_d_22_40x72_torch_up.npz"), seed=SEED, hw=np.asarray(HW),
                        seg=seg.numpy(), labels=pred.numpy().astype(np.uint8), logprob=final.numpy().astype(np.float32))
    print("wrote fwd_drn_d_22_40x72_torch_up.npz", tuple(final.shape), "keys without up.weight:", "up.weight" not in shapes)


if __name__ == "__main__":
    main()
