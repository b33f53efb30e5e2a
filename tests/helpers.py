"""shared helpers for the test-suite (test infrastructure; may import oracle/)"""
import collections
import json
import os

import numpy as np
import torch

from conftest import GOLDEN
from oracle import recipe


def golden(name):
    return os.path.join(GOLDEN, name)


def load_keys(arch):
    with open(golden("state_dict_keys.json")) as fh:
        return collections.OrderedDict((k, tuple(s)) for k, s in json.load(fh)[arch])


def fixture_state_dict(arch, fx):
    """rebuild the exact state_dict the golden generator fed to the reference for fixture `fx`"""
    shapes = load_keys(arch)
    seed = int(fx["seed"])
    sd = recipe.make_state_dict(shapes, seed=seed)
    masks = collections.OrderedDict()
    for k in fx.files:
        if k.startswith("maskbits:"):
            key = k[len("maskbits:"):]
            masks[key] = torch.from_numpy(recipe.unpack_mask_bits(fx[k], shapes[key]))
    if masks:
        sd = recipe.sparse_reinit(sd, masks, seed=seed)
    return sd, masks


def fixture_frames(fx):
    h, w = (int(v) for v in fx["hw"])
    return recipe.make_frames(1, h, w, seed=1234 + int(fx["seed"]))


MS_SOURCES = [(32, 56), (16, 28), (24, 42), (40, 70), (48, 84), (56, 98), (23, 37), (77, 131)]


def ms_sources(seed=11, n=1, c=4):
    """the source tensors tests/golden/gen_golden_ms.py fed to the reference's resize_4d_tensor"""
    g = torch.Generator().manual_seed(seed)
    return [torch.log_softmax(3.0 * torch.randn(n, c, h, w, generator=g), dim=1) for h, w in MS_SOURCES]
