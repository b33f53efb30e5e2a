"""Mask container shared by every pruner (reference: pruners/Pruner.py:6-27)."""
import collections

import numpy as np
import torch


class Pruner(object):
    """``mask_dict``: OrderedDict {state_dict key -> tensor shaped like the weight, 0 = pruned}.
    ``layer_configs``: OrderedDict {state_dict key -> per-type config}, from ``parse_config_file``."""

    def __init__(self, config_fp, on_gpu):
        super(Pruner, self).__init__()
        self.config_fp = config_fp
        self.on_gpu = on_gpu
        self.mask_dict = collections.OrderedDict()
        self.layer_configs = self.parse_config_file(config_fp)

    def parse_config_file(self, config_fp):
        raise NotImplementedError

    def _store(self, layer, mask):
        t = torch.from_numpy(np.ascontiguousarray(mask))
        self.mask_dict[layer] = t.cuda() if self.on_gpu else t

    def apply_masks(self, model):
        """in-place ``state_dict()[k] *= mask`` (pruners/Pruner.py:17-20)"""
        with torch.no_grad():
            sd = model.state_dict()
            for layer, mask in self.mask_dict.items():
                sd[layer] *= mask.to(sd[layer].device)

    def sparsity(self):
        """{layer: fraction of zero mask elements} — what print_stats prints (pruners/Pruner.py:22-27)"""
        out = collections.OrderedDict()
        for layer, mask in self.mask_dict.items():
            m = mask.cpu().numpy()
            out[layer] = 1.0 - np.count_nonzero(m) / m.size
        return out

    def print_stats(self):
        for layer, sp in self.sparsity().items():
            print(layer, "sparsity = {}".format(sp * 100))
