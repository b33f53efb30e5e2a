"""Hierarchical block pruning: successive BlockPruner levels on the residual (pruners/HbPruner.py:15-74).

Level masks are *summed*, so with overlapping static levels a mask value can exceed 1; consumers must
test ``mask != 0`` (the tile-list builder does)."""
import collections
import json

import numpy as np

from .BlockPruner import BlockPruner
from .Pruner import Pruner


class HbPrunerConfig:
    def __init__(self, block_configs):
        self.block_configs = block_configs


class HbPruner(Pruner):
    def __init__(self, config_fp, on_gpu=True):
        super(HbPruner, self).__init__(config_fp, on_gpu)

    def parse_config_file(self, config_fp):
        layer_configs = collections.OrderedDict()
        with open(config_fp) as fh:
            data = json.load(fh)
        for entry in data["configs"]:
            for layer in entry["layer_set"]:
                levels = [BlockPruner.generate_block_pruner_config(lv) for lv in entry["levels"]]
                layer_configs[layer] = HbPrunerConfig(levels)
        return layer_configs

    def generate_masks(self, model, is_static=False, verbose=False):
        sd = model.state_dict()
        for layer, cfg in self.layer_configs.items():
            if verbose:
                print("Generating mask for layer {}".format(layer))
            self._store(layer, HbPruner.generate_mask(sd[layer].cpu().numpy(), cfg, is_static))

    @staticmethod
    def generate_mask(tensor, pconfig, is_static=False):
        total = np.zeros(tensor.shape, dtype=tensor.dtype)
        for level in pconfig.block_configs:
            if is_static:
                mask = BlockPruner.generate_mask_by_construction(tensor, level)
            else:
                mask = BlockPruner.generate_mask_by_pruning(tensor, level)
            tensor = tensor - mask * tensor      # the next level sees only what is left (:69)
            total = total + mask
        return total
