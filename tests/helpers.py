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


def gate_case_cpu(arch, pruned, seed, cfg=None):
    """(model on the CPU, state_dict, mask_dict) of a seeded random-init network, pruned by the host mirror of the
    reference's pruners when `pruned` (cfg = a pruner JSON dict, default BlockPruner 75 %)"""
    import contextlib
    import io
    import tempfile
    import drnb200
    shapes = load_keys(arch)
    sd = recipe.make_state_dict(shapes, seed=seed)
    model = drnb200.DRNSeg(arch, 19, pretrained=False)
    model.load_state_dict(sd, strict=False)
    masks = None
    if pruned:
        cfg = cfg if cfg is not None else recipe.block_pruner_config(shapes, 0.75)
        with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as fh:
            json.dump(cfg, fh)
        pruner = drnb200.pruners.make_pruner(fh.name, on_gpu=False)
        np.random.seed(seed)                       # srmbrep patterns draw from numpy's global RNG
        with contextlib.redirect_stdout(io.StringIO()):      # RmbPruner prints progress like the reference
            if cfg["pruner_type"] == "rmb":
                pruner.generate_masks(model)       # RmbPruner.generate_masks has no is_static (RmbPruner.py:111)
            else:
                pruner.generate_masks(model, is_static=False)
        os.unlink(fh.name)
        masks = pruner.mask_dict
        sd = recipe.sparse_reinit(sd, masks, seed=seed)
        model.load_state_dict(sd, strict=False)
    return model, sd, masks
