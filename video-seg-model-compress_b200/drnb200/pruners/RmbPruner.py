""""Regular multi-blocklet" masks (reference: pruners/RmbPruner.py:68-243).

Outer level: the matrix (O, I*kh*kw) is cut into ``global_bh x global_bw`` blocks; in every block-row the
``int(global_sp * n_col_blocks)`` lowest-magnitude blocks are dropped (score <= threshold, :155-163).
Inner level: each surviving block is cut into blocklets; for every blocklet type, ``count`` times, each
blocklet-row keeps its currently largest blocklet (argmax of sum |w|, first wins) and removes it from
further consideration (:190-231).

Note the reference's ``generate_masks(self, model, verbose)`` has no ``is_static`` parameter although the
drivers pass one (semantic_seg.py:849 -> TypeError); the mirror accepts and ignores it.
"""
import collections
import json

import numpy as np

from .Pruner import Pruner
from . import blocklet_export


class BlockletType(object):
    def __init__(self, bh, bw):
        self.bh = bh
        self.bw = bw

    def __str__(self):
        return "{}x{}".format(self.bh, self.bw)


class RmbPrunerConfig(object):
    def __init__(self, bh, bw, spo, bl_types, bl_counts):
        self.bh, self.bw, self.spo = bh, bw, spo
        self.bl_types, self.bl_counts = bl_types, bl_counts


def parse_blocklet_config(entry):
    types = [BlockletType(b["bh"], b["bw"]) for b in entry["blocklets"]]
    counts = [b["count"] for b in entry["blocklets"]]
    return entry["global_bh"], entry["global_bw"], entry["global_sp"], types, counts


def outer_block_mask(mat, bh, bw, spo, meta):
    """row-wise outer sparsity: keep[rb, cb] = 0 where score <= the (int(spo*ncb)-1)-th smallest of the row"""
    nrb, ncb = mat.shape[0] // bh, mat.shape[1] // bw
    keep = np.ones((nrb, ncb), dtype=mat.dtype)
    if spo > 0:
        cut = int(spo * meta.shape[1]) - 1
        if cut >= 0:
            for rb in range(nrb):
                thresh = np.sort(np.abs(meta[rb].flatten()))[cut]
                keep[rb][meta[rb] <= thresh] = 0
    return keep


class RmbPruner(Pruner):
    def __init__(self, config_fp, on_gpu=True):
        super(RmbPruner, self).__init__(config_fp, on_gpu)

    def parse_config_file(self, config_fp):
        layer_configs = collections.OrderedDict()
        with open(config_fp) as fh:
            data = json.load(fh)
        for entry in data["configs"]:
            bh, bw, sp, types, counts = parse_blocklet_config(entry)
            for layer in entry["layer_set"]:
                layer_configs[layer] = RmbPrunerConfig(bh, bw, sp, types, counts)
        return layer_configs

    def generate_masks(self, model, is_static=False, verbose=False):
        sd = model.state_dict()
        for layer, cfg in self.layer_configs.items():
            if verbose:
                print("Generating mask for layer {}".format(layer))
            self._store(layer, RmbPruner.prune_tensor_as_rmb(sd[layer].cpu().numpy(), cfg))

    @staticmethod
    def prune_tensor_as_rmb(tensor, config, dump_fpath=None):
        mat = tensor.reshape(tensor.shape[0], -1).copy()       # picked blocklets are zeroed in this copy
        mask = np.zeros(mat.shape, dtype=mat.dtype)
        rows, cols = mat.shape
        bh, bw = config.bh, config.bw
        assert rows % bh == 0, "Block height should divide rows"
        assert cols % bw == 0, "Block width should divide columns"
        nrb, ncb = rows // bh, cols // bw

        if config.spo > 0:
            if bh != 1 and bw != 1:
                meta = np.zeros((nrb, ncb))                    # float64 holder of float32 block sums (:148)
                for rb in range(nrb):
                    for cb in range(ncb):
                        meta[rb, cb] = np.sum(np.abs(mat[rb * bh:(rb + 1) * bh, cb * bw:(cb + 1) * bw]))
            else:
                meta = np.abs(mat)
            keep = outer_block_mask(mat, bh, bw, config.spo, meta)
        else:
            keep = np.ones((nrb, ncb), dtype=mat.dtype)

        exported = []          # blocklets in creation order (only kept for the text export)
        for rb in range(nrb):
            for cb in range(ncb):
                if keep[rb, cb] == 0:
                    continue
                blk = mat[rb * bh:(rb + 1) * bh, cb * bw:(cb + 1) * bw]      # view into `mat`
                for btype, count in zip(config.bl_types, config.bl_counts):
                    n_r, n_c = bh // btype.bh, bw // btype.bw
                    score = np.zeros(n_c)
                    for _ in range(count):
                        vals, picks = np.zeros((bh, btype.bw)), np.zeros(n_r)
                        exported.append(blocklet_export.Blocklet(bh, bw, rb, cb, btype.bh, btype.bw, vals, picks))
                        for br in range(n_r):
                            band = blk[br * btype.bh:(br + 1) * btype.bh]
                            for bc in range(n_c):
                                score[bc] = np.sum(np.abs(band[:, bc * btype.bw:(bc + 1) * btype.bw]))
                            pick = int(np.argmax(score))
                            vals[br * btype.bh:(br + 1) * btype.bh] = band[:, pick * btype.bw:(pick + 1) * btype.bw]
                            picks[br] = pick
                            band[:, pick * btype.bw:(pick + 1) * btype.bw] = 0
                            r0 = rb * bh + br * btype.bh
                            c0 = cb * bw + pick * btype.bw
                            mask[r0:r0 + btype.bh, c0:c0 + btype.bw] = 1.0
        if dump_fpath is not None:
            blocklet_export.write_rmb(dump_fpath, rows, cols, bh, bw, exported)
        return mask.reshape(tensor.shape)
