""""Regular multi cyclic-diagonal blocklet" masks (reference: pruners/RmcdbPruner.py:80-316).

Like RMB at the outer level; inside a surviving block the kept blocklets lie on cyclic diagonals
(blocklet-row r keeps blocklet-column (r + dia) % n_cols).  Pruning picks the ``count`` diagonals with the
largest sum |w| (``np.argsort(scores)[::-1][:count]``, :281); the static variant draws them with
``np.random.choice`` (:186).

Reference defects handled: ``construct_rmcdb_matrix`` uses an undefined ``rb`` when ``global_sp > 0`` (:167,
NameError) — the mirror applies the evident intent (independent draw per block-row); the pruning path's
"zeroing" statement double-slices rows (:293) and therefore clears whole rows of the working copy, which
only matters when several blocklet types are configured — reproduced literally so masks stay identical.
"""
import collections
import json

import numpy as np

from .Pruner import Pruner
from . import blocklet_export
from .RmbPruner import BlockletType, outer_block_mask, parse_blocklet_config
from .utils import get_meta_matrix


class RmcdbPrunerConfig(object):
    def __init__(self, bh, bw, spo, bl_types, bl_counts, collapse_tensor):
        self.bh, self.bw, self.spo = bh, bw, spo
        self.bl_types, self.bl_counts = bl_types, bl_counts
        self.collapse_tensor = collapse_tensor


class RmcdbPruner(Pruner):
    def __init__(self, config_fp, on_gpu=True):
        super(RmcdbPruner, self).__init__(config_fp, on_gpu)

    def parse_config_file(self, config_fp):
        layer_configs = collections.OrderedDict()
        with open(config_fp) as fh:
            data = json.load(fh)
        for entry in data["configs"]:
            bh, bw, sp, types, counts = parse_blocklet_config(entry)
            for layer in entry["layer_set"]:
                layer_configs[layer] = RmcdbPrunerConfig(bh, bw, sp, types, counts, entry["collapse_tensor"])
        return layer_configs

    def generate_masks(self, model, is_static=False, verbose=False):
        sd = model.state_dict()
        for layer, cfg in self.layer_configs.items():
            w = sd[layer].cpu().numpy()
            if verbose:
                print("Generating mask for layer {} using {} approach".format(
                    layer, "static" if is_static else "pruning"))
            self._store(layer, RmcdbPruner.construct_rmcdb_matrix(w, cfg) if is_static
                        else RmcdbPruner.prune_tensor_as_rmcdb(w, cfg))

    @staticmethod
    def _paint_diagonal(mask, rb, cb, bh, bw, btype, dia):
        n_r, n_c = bh // btype.bh, bw // btype.bw
        for br in range(n_r):
            bc = (br + dia) % n_c
            r0, c0 = rb * bh + br * btype.bh, cb * bw + bc * btype.bw
            mask[r0:r0 + btype.bh, c0:c0 + btype.bw] = 1

    @staticmethod
    def construct_rmcdb_matrix(tensor, config):
        rows, cols = tensor.shape[0], tensor.size // tensor.shape[0]
        bh, bw = config.bh, config.bw
        assert rows % bh == 0, "Block height should divide rows"
        assert cols % bw == 0, "Block width should divide columns"
        nrb, ncb = rows // bh, cols // bw
        mask = np.zeros((rows, cols), dtype=tensor.dtype)
        keep = np.ones((nrb, ncb), dtype=tensor.dtype)
        if config.spo > 0:
            n_zero = int(config.spo * ncb)
            for rb in range(nrb):
                keep[rb, np.random.choice(ncb, n_zero, replace=False)] = 0
        for rb in range(nrb):
            for cb in range(ncb):
                if keep[rb, cb] == 0:
                    continue
                for btype, count in zip(config.bl_types, config.bl_counts):
                    assert bh % btype.bh == 0, "Block height should divide rows in a blocklet"
                    assert bw % btype.bw == 0, "Block width should divide columns in a blocklet"
                    for dia in np.random.choice(bw // btype.bw, count, replace=False):
                        RmcdbPruner._paint_diagonal(mask, rb, cb, bh, bw, btype, dia)
        return mask.reshape(tensor.shape)

    @staticmethod
    def prune_tensor_as_rmcdb(tensor, config, dump_fpath=None):
        mat = tensor.reshape(tensor.shape[0], -1).copy()
        mask = np.zeros(mat.shape, dtype=mat.dtype)
        rows, cols = mat.shape
        bh, bw = config.bh, config.bw
        assert rows % bh == 0, "Block height should divide rows"
        assert cols % bw == 0, "Block width should divide columns"
        nrb, ncb = rows // bh, cols // bw
        keep = np.ones((nrb, ncb), dtype=mat.dtype)
        if config.spo > 0:
            keep = outer_block_mask(mat, bh, bw, config.spo, get_meta_matrix(mat, bh, bw))
        exported = []          # blocklets in creation order (only kept for the text export)
        for rb in range(nrb):
            for cb in range(ncb):
                if keep[rb, cb] == 0:
                    continue
                blk = mat[rb * bh:(rb + 1) * bh, cb * bw:(cb + 1) * bw]
                for btype, count in zip(config.bl_types, config.bl_counts):
                    assert bh % btype.bh == 0, "Block height should divide rows in a blocklet"
                    assert bw % btype.bw == 0, "Block width should divide columns in a blocklet"
                    n_r, n_c = bh // btype.bh, bw // btype.bw
                    meta = get_meta_matrix(blk, btype.bh, btype.bw)
                    r_idx = np.arange(n_r)
                    scores = np.zeros(n_c)
                    for dia in range(n_c):
                        scores[dia] = np.sum(meta[r_idx, (r_idx % n_c + dia) % n_c])
                    for dia in np.argsort(scores)[::-1][:count]:
                        RmcdbPruner._paint_diagonal(mask, rb, cb, bh, bw, btype, dia)
                        vals = np.zeros((bh, btype.bw), dtype=mat.dtype)
                        for br in range(n_r):
                            bc = (br + dia) % n_c
                            vals[br * btype.bh:(br + 1) * btype.bh] = \
                                blk[br * btype.bh:(br + 1) * btype.bh, bc * btype.bw:(bc + 1) * btype.bw]
                            # literal reproduction of pruners/RmcdbPruner.py:293 (row slice of a row slice)
                            blk[br * btype.bh:(br + 1) * btype.bh][bc * btype.bw:(bc + 1) * btype.bw] = 0
                            # the reference appends one record PER blocklet row, all sharing the diagonal's values
                            # (pruners/RmcdbPruner.py:303-304 sits inside the row loop)
                            exported.append(blocklet_export.Blocklet(bh, bw, rb, cb, btype.bh, btype.bw, vals, dia))
        if dump_fpath is not None:
            blocklet_export.write_rmcdb(dump_fpath, rows, cols, bh, bw, exported)
        return mask.reshape(tensor.shape)
