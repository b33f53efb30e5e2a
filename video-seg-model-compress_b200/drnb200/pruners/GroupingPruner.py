"""Block-diagonal ("grouped convolution") masks (reference: pruners/GroupingPruner.py:13-61)."""
import collections
import json

import numpy as np

from .Pruner import Pruner


class GroupingPrunerConfig():
    def __init__(self, num_groups):
        self.num_groups = num_groups


class GroupingPruner(Pruner):
    def __init__(self, config_fp, on_gpu=True):
        super(GroupingPruner, self).__init__(config_fp, on_gpu)

    def parse_config_file(self, config_fp):
        layer_configs = collections.OrderedDict()
        with open(config_fp) as fh:
            data = json.load(fh)
        for entry in data["configs"]:
            for layer in entry["layer_set"]:
                layer_configs[layer] = GroupingPrunerConfig(entry["num_groups"])
        return layer_configs

    def generate_masks(self, model, is_static=True, verbose=False):
        sd = model.state_dict()
        for layer, cfg in self.layer_configs.items():
            if verbose:
                print("Generating mask for layer {}".format(layer))
            self._store(layer, GroupingPruner.construct_mask(sd[layer].cpu().numpy(), cfg))

    @staticmethod
    def construct_mask(tensor, config):
        """group g keeps out-channels [g*O/G,(g+1)*O/G) x in-channels [g*I/G,(g+1)*I/G)  (:52-61)"""
        g = config.num_groups
        mask = np.zeros(tensor.shape, dtype=tensor.dtype)
        so, si = tensor.shape[0] // g, tensor.shape[1] // g
        for gid in range(g):
            mask[gid * so:(gid + 1) * so, gid * si:(gid + 1) * si] = 1
        return mask
